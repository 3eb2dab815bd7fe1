// Host-only check of the persistent kernel's work partition (ssimk::plan_slots + the PieceCursor both warps of a pair
// run in the kernel): for any shape the pieces of all slots must tile every band's rows of every frame exactly once, in
// order, the members of a team must walk the same rows of adjacent bands, shares must be balanced to +-1 unit, and no slot
// may touch more frames than the plan reserved partial-sum entries for.
#include <cstdint>
#include <cstdio>
#include <vector>
#include "ssim_kernels.h"

static int fails = 0;
static void expect(bool ok, const char* what, uint32_t w, uint32_t h, uint32_t f, uint32_t slots)
{
    if (!ok && fails < 20) std::printf("FAIL %s: %ux%u x%u on %u slots\n", what, w, h, f, slots);
    if (!ok) ++fails;
}

static void check(uint32_t maxSlots, uint32_t w, uint32_t h, uint32_t f, uint32_t minUnits)
{
    ssimk::SlotPlan plan;
    if (!ssimk::plan_slots(maxSlots, w, h, f, minUnits, &plan)) { expect(false, "plan_slots refused", w, h, f, maxSlots); return; }
    const ssimk::SlotGeo g = ssimk::make_slot_geo(plan, w);
    const uint32_t bands = (w + 63) / 64;
    const uint64_t cols = (uint64_t)bands * f;
    expect(plan.slots >= 1 && plan.slots <= maxSlots, "slot count in range", w, h, f, plan.slots);
    expect(plan.group >= 1 && plan.group <= 8 && plan.slots % plan.group == 0, "whole teams of 1..8 pairs", w, h, f, plan.slots);
    const uint32_t teams = plan.slots / plan.group;
    expect((uint64_t)teams * plan.shareQ + plan.shareR == (uint64_t)g.groupsPerFrame * f * plan.colUnits, "shares add up to all units", w, h, f, plan.slots);
    expect(plan.shareQ >= (teams > 1 ? minUnits : 1u), "no team thinner than minUnits", w, h, f, plan.slots);
    expect(plan.pad >= 10 && plan.colUnits == h + plan.pad, "columns are the rows plus at least 10 padding units", w, h, f, plan.slots);
    const bool oneColumn = (uint64_t)g.groupsPerFrame * f == 1;
    expect(plan.pad == (oneColumn ? 10u : 19u) || (plan.shareR == 0 && plan.colUnits % plan.shareQ == 0),
           "padding is the start-up rows (+ the crossing cost when there are several columns), or the columns are cut into equal parts", w, h, f, plan.slots);
    expect((uint64_t)g.groupsPerFrame * plan.group >= bands && (uint64_t)(g.groupsPerFrame - 1) * plan.group < bands, "groups cover the bands", w, h, f, plan.slots);
    expect((double)bands / ((double)g.groupsPerFrame * plan.group) >= 0.9 || bands < 8, "little ragged waste", w, h, f, plan.slots);
    {
        double bias, scale, inv;
        ssimk::acc_format(plan, &bias, &scale, &inv);
        expect(bias >= ((double)plan.shareQ + 1.0) * 64.0, "reduction bias covers a slot's pixels of one frame", w, h, f, plan.slots);
        expect(scale >= 1.0 && scale * inv == 1.0 && scale <= 1099511627776.0, "reduction scale is a power of two in [1, 2^40]", w, h, f, plan.slots);
        expect((double)plan.slots * 2.0 * bias * scale < 1125899906842624.0, "all slots of a frame together stay below 2^50 in the 52-bit field", w, h, f, plan.slots);
        expect(plan.slots <= ssimk::kMaxSlots, "slot count fits the 12-bit arrival counter", w, h, f, plan.slots);
    }
    std::vector<uint32_t> nextRow(cols, 0);          // rows [0, nextRow) of each column are covered so far
    uint64_t lastCol = 0;
    bool first = true;
    (void)lastCol; (void)first;
    for (uint32_t s = 0; s < plan.slots; ++s) {
        ssimk::PieceCursor c;
        ssimk::cursor_init(c, g, s);
        ssimk::Piece pc;
        uint32_t q0, q1;
        ssimk::slot_units(g, s, q0, q1);
        const uint32_t frameUnits = g.groupsPerFrame * g.colUnits;
        uint64_t prevCol = 0;
        bool any = false;
        while (ssimk::cursor_next(c, g, pc)) {
            const uint64_t col = (uint64_t)pc.frame * bands + pc.band;
            expect(pc.frame >= 0 && (uint32_t)pc.frame < f && pc.band >= 0 && (uint32_t)pc.band < bands, "piece inside the batch", w, h, f, plan.slots);
            if (pc.frame < 0 || (uint32_t)pc.frame >= f || pc.band < 0 || (uint32_t)pc.band >= bands) return;
            expect((uint32_t)pc.band % plan.group == s % plan.group, "member m of a team takes band group*k + m", w, h, f, plan.slots);
            expect(pc.nOut >= 1 && (uint32_t)(pc.r0 + pc.nOut) <= h, "piece inside its column", w, h, f, plan.slots);
            expect((uint32_t)pc.r0 == nextRow[col], "pieces of a band are contiguous and in order over the slots", w, h, f, plan.slots);
            expect(!any || col > prevCol, "a slot's pieces move forward", w, h, f, plan.slots);
            expect((uint32_t)pc.frame >= q0 / frameUnits && (uint32_t)pc.frame <= (q1 - 1) / frameUnits, "piece in a frame whose units the slot owns", w, h, f, plan.slots);
            nextRow[col] = (uint32_t)(pc.r0 + pc.nOut);
            prevCol = col; any = true;
        }
    }
    for (uint64_t c = 0; c < cols; ++c) expect(nextRow[c] == h, "every row of every column covered exactly once", w, h, f, plan.slots);
}

int main()
{
    const uint32_t slots = 148 * 8;              // B200: 148 SMs x one CTA of 8 warp pairs
    const uint32_t widths[] = {1, 63, 64, 65, 640, 1920, 3840, 16384, 100000};
    const uint32_t heights[] = {1, 2, 9, 10, 11, 23, 24, 25, 47, 141, 1080, 2058, 2160, 16384};
    const uint32_t frames[] = {1, 2, 3, 64, 512};
    for (uint32_t w : widths) for (uint32_t h : heights) for (uint32_t f : frames) {
        if ((uint64_t)((w + 63) / 64) * f * (h + 10) > 40000000ull) continue;     // keep the check fast
        check(slots, w, h, f, 24);
        check(slots / 2, w, h, f, 1);
        check(7, w, h, f, 100);
    }
    check(slots, 1920, 1080, 4096, 24);
    check(slots, 64, 64, 4096, 24);
    check(1, 3840, 2160, 2, 24);
    // the documented plans
    ssimk::SlotPlan p;
    ssimk::plan_slots(slots, 3840, 2160, 1, 24, &p);   expect(p.group == 6 && p.slots == 1182 && p.shareQ == 110 && p.pad == 19, "one 4K pair: 197 teams of 6 bands share 10 columns of 2160 + 19 units", 3840, 2160, 1, p.slots);
    ssimk::plan_slots(slots, 3840, 2160, 64, 24, &p);  expect(p.group == 6 && p.slots == 1182 && p.shareQ == 7078, "64 x 4K", 3840, 2160, 64, p.slots);
    ssimk::plan_slots(slots, 16384, 2058, 1, 24, &p);  expect(p.group == 8 && p.slots == 1184 && p.pad == 19, "a 16384-wide strip: one team of 8 bands per CTA, ranges cross columns", 16384, 2058, 1, p.slots);
    ssimk::plan_slots(slots, 1920, 1080, 1, 24, &p);   expect(p.group == 6 && p.slots == 1170 && p.shareQ == 28 && p.pad == 12, "1080p: 5 columns of 6 bands cut into 39 equal parts of 28 units", 1920, 1080, 1, p.slots);
    ssimk::plan_slots(slots, 333, 141, 1, 24, &p);     expect(p.group == 6 && p.slots == 36, "small image: fewer slots", 333, 141, 1, p.slots);
    ssimk::plan_slots(slots, 8, 8, 1, 24, &p);         expect(p.slots == 1 && p.shareQ == 18, "tiny image: one slot", 8, 8, 1, p.slots);
    expect(!ssimk::plan_slots(slots, 100000, 100000, 100, 24, &p), "more than 2^31 units is refused", 100000, 100000, 100, 0);
    std::printf(fails ? "plan_slots: %d failures\n" : "plan_slots ok\n", fails);
    return fails ? 1 : 0;
}
