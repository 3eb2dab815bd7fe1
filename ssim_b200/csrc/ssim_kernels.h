// ssim_kernels.h -- launch interface between the runtime (ssim_cuda.cu) and the sm_100a kernels.
#ifndef SSIM_B200_KERNELS_H
#define SSIM_B200_KERNELS_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssimk {

// ---- geometry of the fused kernel (see DESIGN.md "Kernel")
constexpr int kBandW       = 64;   // output columns per work column ("band"; a consumer lane owns columns lane and lane+32)
constexpr int kBlkRows     = 8;    // rows per TMA box = per horizontal-pass block = per ring unit
constexpr int kHalo        = 5;    // Gaussian radius (reference src/ssim.cpp:227)
constexpr int kBoxLeft     = 16;   // band column 0 sits at byte 16 of a box row: the innermost TMA coordinate (bx - 16) must be
                                   // a multiple of 16 bytes -- measured on B200: x = -16 works, x = -8 raises "illegal instruction"
                                   // (tools/dev/tma_probe.cu)
constexpr int kLoadRows    = 8;    // rows per TMA box (= one horizontal-pass block)
constexpr int kStages      = 2;    // TMA stages per warp pair
// ONE CTA of 8 warp pairs (512 threads) per SM, not two of 4.  Measured (tools/dev/slot_times.py, profiles/r02_cta_shapes.txt):
// with two resident CTAs the warp schedulers serve the CTA that arrived first on an SM with priority -- its pairs run 2.2x
// as fast as those of the second CTA (1.8 vs 3.9 us per 8 rows on 4K batches) -- so a static, equal partition ends 45% late
// on the second CTAs, and which CTA of a grid arrives first on an SM is not a function of blockIdx.  The warps of ONE CTA
// are served evenly: all 8 pairs of every CTA finish within 1% of each other.  (The unfair arrangement gets ~4% more
// instructions per cycle out of an SM while both CTAs run; cutting large inputs into many small CTAs so that the hardware
// block scheduler balances them -- what round 1 did -- measured 1% better than the single even wave on 64 x 4K and 7%
// worse on 16 x 4K, and was dropped.)
constexpr int kPairsPerCta = 8;    // warp pairs per CTA: each is one producer warp and one consumer warp
constexpr int kCtaThreads  = 2 * kPairsPerCta * 32;      // 512: warps 0-7 producers (TMA + horizontal pass), 8-15 consumers
// setmaxnreg budgets of the two warpgroups: 128*120 + 128*136 = 256*128.  Measured on 64 x 4K pairs (profiles/r02_variants.txt):
// 96/160 2062 us, 112/144 2056, 120/136 2038, 128/128 2074 -- the producer stops re-deriving per-lane constants in every block,
// the consumer's 88 accumulators + formula temporaries still fit without spills.
constexpr int kProducerRegs = 120;
constexpr int kConsumerRegs = 136;
constexpr unsigned kBackoffNs = 200; // default sleep between polls of the partner warp's mbarrier
constexpr int kRingPlaneBytes = kBandW * 8;              // 512: one row of packed {x, y} pairs
constexpr int kRingRowPad     = 32;                      // consecutive rows start 8 banks apart: see the ring layout in the kernel
constexpr int kRingRowBytes   = 2 * kRingPlaneBytes + kRingRowPad;   // 1056: {E[a'], E[b']} plane, {E[(a'-b')^2], E[a'b']} plane, pad
constexpr int kTaps           = 11;
constexpr int kRingRows       = 2 * kTaps;               // 22: two halves of 11 rows; the consumer's unrolled body is exactly one half
constexpr int kRingBytes      = kRingRows * kRingRowBytes; // 23232
constexpr int kBarsPerPair    = 8;                       // mbarriers per pair (64 bytes): tmaFull[2] stageEmpty[2] ringFull[2] ringEmpty[2]

// Per pixel type geometry of the TMA stage and the depth of the ring (everything after the widening is identical).
//   8-bit : box rows of 96 bytes = 16 left margin + 64 columns + 16 right margin; a stage (8 rows of A, 8 rows of B) is 1536
//           bytes.
//   16-bit: box rows of 80 elements = 160 bytes (band column 0 again at byte 16, x coordinate bx-8 elements = 16 bytes
//           aligned); a stage is 2560 bytes.
// The ring (two halves of 11 rows) is the same for both.
// Only the start of a box must be 128-byte aligned in shared memory (measured, tools/dev/tma_probe.cu): 8 x 96 = 768 and
// 8 x 160 = 1280 both are multiples of 128, so A and B boxes follow each other without padding.
template <bool kU16> struct PixGeo {
    static constexpr int kPixBytes      = kU16 ? 2 : 1;
    static constexpr int kBoxBytes      = kU16 ? 160 : 96;
    static constexpr int kBoxElems      = kBoxBytes / kPixBytes;            // 96 / 80
    static constexpr int kBoxLeftElems  = kBoxLeft / kPixBytes;             // 16 / 8
    static constexpr int kImgStageBytes = kBoxBytes * kLoadRows;            // 768 / 1280
    static constexpr int kStageBytes    = 2 * kImgStageBytes;               // 1536 / 2560
    static constexpr int kPairSmemBytes = (kStages * kStageBytes + kRingBytes + 127) / 128 * 128;   // 26368 / 28416
    static constexpr int kCtaSmemBytes  = kPairsPerCta * kPairSmemBytes;    // 210944 / 227328: one CTA per SM
};

// ---- work partition --------------------------------------------------------------------------------------------------
// The kernel is persistent: `slots` warp pairs (at most kPairsPerCta x numSMs, all resident at once) share the work evenly
// and statically.  Pairs work in TEAMS of `group` (1..8) pairs with consecutive slot numbers: a team walks
// down `group` ADJACENT 64-pixel bands side by side, member m taking band group*k + m, all members over the same rows.
// (Neighbouring bands share the 16-byte margins of their TMA boxes, i.e. sectors and DRAM atoms; pairs that reach the same
// rows at unrelated times each fetch them from DRAM again.  Measured with every pair on its own: 4.0x the algorithmic DRAM
// read traffic, 1.05 GB instead of 265 MB for 16 4K pairs, and 10% less throughput.)
// The work is laid out as one line per team member: a COLUMN is one group of bands of one frame (cols = frames x
// groupsPerFrame, frame-major), every column is `outRows` rows long and is preceded by `pad` "padding" units that stand
// for the cost of crossing into it: kPad = 2*kHalo start-up rows (the 10 extra input rows the vertical filter needs before
// its first output) + kCrossingUnits for the refill of the pair's pipeline.
// Team j owns the units [j*Q + min(j,R), ...) of that line (Q, R = quotient and remainder of units / teams), i.e. every team
// owns the same number of units +-1; the rows of a column that fall into a team's range form a PIECE per member, which is
// what a warp pair processes in one go (own halo above and below, replicated rows at the plane's edges).  A range that
// starts inside a column pays its own 10-row start-up, a range that crosses into the next column pays that column's
// padding units: either way cost = units + 10 for every slot, so all pairs finish together -- no tail, and no halo paid
// more often than once per slot and once per column.  When `group` does not divide the number of bands the last group of
// a frame is ragged: its surplus members have nothing to do there (plan_slots keeps that waste small).
constexpr int kPad = 2 * kHalo;
constexpr int kDbgWords = 32;
constexpr uint32_t kMaxSlots = 4095;     // the reduction counts the slots of a frame in 12 bits (see "reduction" in the kernel)

struct SlotPlan {
    uint32_t slots;          // warp pairs that get work = teams * group
    uint32_t group;          // pairs per team = adjacent bands walked side by side
    uint32_t shareQ, shareR; // units per team: team j owns shareQ + (j < shareR) units
    uint32_t colUnits;       // units per column = outRows + pad
    uint32_t pad;            // padding units in front of every column: kPad (one column), kPad + kCrossingUnits, or more when the columns are cut into equal parts
};
constexpr uint32_t kCrossingUnits = 9;   // what it costs a team to start a second piece (its range crosses into the next column),
                                         // in row units: measured, the crossing teams of a 4K pair finish 3 us after the others

// Pure host logic (unit-tested on the CPU: tests/clients/plan_check.cpp).  maxSlots: warp pairs resident at once (8 per SM).
// minUnits: do not spread the work thinner than this many units per slot (tiny images would otherwise pay 10 start-up rows
// for a handful of output rows per slot).
inline bool plan_slots(uint32_t maxSlots, uint32_t width, uint32_t outRows, uint32_t frames, uint32_t minUnits, SlotPlan* plan)
{
    const unsigned long long bands = ((unsigned long long)width + kBandW - 1) / kBandW;
    // the padding units in front of a column stand for what a team pays when its range crosses into that column: the 10
    // start-up rows of the new piece plus the refill of its pipeline (kCrossingUnits) -- with that in the line, teams whose
    // ranges cross a boundary get fewer rows and finish with the others
    // (a single column has no crossings)
    const unsigned long long maxColUnits = (unsigned long long)outRows + kPad + kCrossingUnits;
    if (maxSlots < 1 || width == 0 || outRows == 0 || frames == 0 || bands * frames > 0x7fffffffull || bands * frames * maxColUnits > 0x7fffffffull) return false;
    if (minUnits < 1) minUnits = 1;
    // team size: as many adjacent bands as possible side by side, as long as ragged last groups and slots that do not
    // fill a team waste less than the sharing is worth (an unshared band edge costs about a tenth of a band's time)
    unsigned long long bestG = 1;
    double bestCost = 1e30;
    for (unsigned long long g = 1; g <= 8 && g <= bands && g <= maxSlots; ++g) {
        const unsigned long long groups = (bands + g - 1) / g;
        const double ragged = 1.0 - (double)bands / (double)(groups * g);
        const double idle = (double)(maxSlots % g) / (double)maxSlots;
        const double cost = ragged + idle + 0.1 / (double)g;
        if (cost < bestCost - 1e-12) { bestCost = cost; bestG = g; }
    }
    const unsigned long long group = bestG, groupsPerFrame = (bands + group - 1) / group;
    const unsigned long long cols = groupsPerFrame * frames;
    const unsigned long long pad = cols > 1 ? kPad + kCrossingUnits : kPad;
    const unsigned long long colUnits = (unsigned long long)outRows + pad;
    unsigned long long units = cols * colUnits;                                    // per team member
    const unsigned long long maxTeams = maxSlots / group;
    unsigned long long teams = units / minUnits;
    if (teams > maxTeams) teams = maxTeams;
    if (teams < 1) teams = 1;
    plan->group = (uint32_t)group;
    plan->pad = (uint32_t)pad;
    plan->colUnits = (uint32_t)colUnits;
    // Few columns, many teams (a single image): cutting every column into k equal parts -- the column is padded up to
    // k * Q units -- leaves some teams idle but spares all others the second piece that a range crossing a column boundary
    // means.  Taken when it is the faster of the two by the model "time = units + start-up rows (+ a crossing)".
    if (cols > 1 && cols <= maxTeams) {
        const unsigned long long partUnits = (unsigned long long)outRows + kPad;   // no crossings: plain start-up rows
        unsigned long long k = maxTeams / cols;
        while (k > 1 && (partUnits + k - 1) / k < minUnits) --k;
        const unsigned long long partQ = (partUnits + k - 1) / k;
        const unsigned long long lineQ = (units + teams - 1) / teams;
        if (partQ >= minUnits && partQ <= lineQ && k * partQ * cols <= 0x7fffffffull) {
            teams = k * cols;
            plan->colUnits = (uint32_t)(k * partQ);
            plan->pad = (uint32_t)(k * partQ - outRows);
            units = cols * k * partQ;
        }
    }
    plan->slots = (uint32_t)(teams * group);
    plan->shareQ = (uint32_t)(units / teams);
    plan->shareR = (uint32_t)(units % teams);
    return true;
}

// Fixed-point format of the in-kernel reduction (see "reduction" in ssim_kernels.cu): every slot adds (its sum of a frame +
// bias) * scale to the 52-bit field of the frame's accumulator word.  bias >= the pixels a slot can hold of one frame (a slot's
// units there x 64 columns; SSIM > -1, so sum + bias >= 0); scale = 2^k, k <= 40 as large as keeps slots * 2 * bias * scale below
// 2^50.  Pure host logic, checked in tests/clients/plan_check.cpp.
inline void acc_format(const SlotPlan& plan, double* bias, double* scale, double* invScale)
{
    const double b = ((double)plan.shareQ + 1.0) * (double)kBandW;
    double s = 1.0;
    int k = 0;
    while (k < 40 && (double)plan.slots * 2.0 * b * (s * 2.0) * 4.0 < 4503599627370496.0) { s *= 2.0; ++k; }     // 2^52
    *bias = b; *scale = s; *invScale = 1.0 / s;
}

// One piece: rows [r0, r0 + nOut) of one band of one frame (relative to the first output row of the call).
struct Piece {
    int frame, band;
    int r0, nOut;
};

// Enumerates the pieces of a slot in order.  Used by both warps of a pair in the kernel (everything here is warp-uniform
// integer arithmetic) and by the CPU tests; divisions go through fast_div() constants.
struct PieceCursor {
    uint32_t q, qEnd;        // next unit / end of the slot's range
    uint32_t colBase;        // first unit of the current column
    int frame, band;
};

#if defined(__CUDACC__)
#define SSIMK_HD __host__ __device__ __forceinline__
#else
#define SSIMK_HD inline
#endif

SSIMK_HD uint32_t ssimk_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
SSIMK_HD uint32_t ssimk_div(uint32_t n, uint32_t mul, uint32_t shift) { return mul ? ssimk_mulhi(n, mul) >> shift : n; }

struct SlotGeo {             // the part of FusedParams the cursor needs (kept separate so that the CPU tests can build it)
    uint32_t slots, group, groupsPerFrame, shareQ, shareR, colUnits, pad, bands;
    uint32_t colMul, colShift;       // n / colUnits
    uint32_t gpfMul, gpfShift;       // n / groupsPerFrame
};

// units [q0, q1) of the team a slot belongs to: equal shares, the first shareR teams hold one unit more
SSIMK_HD void slot_units(const SlotGeo& g, uint32_t slot, uint32_t& q0, uint32_t& q1)
{
    const uint32_t team = slot / g.group;
    q0 = team * g.shareQ + (team < g.shareR ? team : g.shareR);
    q1 = q0 + g.shareQ + (team < g.shareR ? 1u : 0u);
}

// PieceCursor::band holds the band of THIS slot in the current group of bands: group index * group + member
SSIMK_HD void cursor_init(PieceCursor& c, const SlotGeo& g, uint32_t slot)
{
    slot_units(g, slot, c.q, c.qEnd);
    const uint32_t col = ssimk_div(c.q, g.colMul, g.colShift);
    c.colBase = col * g.colUnits;
    c.frame = (int)ssimk_div(col, g.gpfMul, g.gpfShift);
    c.band = (int)((col - (uint32_t)c.frame * g.groupsPerFrame) * g.group + slot % g.group);
}

// next piece with at least one output row; false when the slot's range is exhausted
SSIMK_HD bool cursor_next(PieceCursor& c, const SlotGeo& g, Piece& pc)
{
    while (c.q < c.qEnd) {
        const uint32_t ua = c.q - c.colBase;                                          // first unit inside the column
        const uint32_t ub = c.qEnd - c.colBase < g.colUnits ? c.qEnd - c.colBase : g.colUnits;
        const int r0 = (int)(ua > g.pad ? ua - g.pad : 0u);
        const int r1 = (int)(ub > g.pad ? ub - g.pad : 0u);
        pc.frame = c.frame; pc.band = c.band; pc.r0 = r0; pc.nOut = r1 - r0;
        const bool real = c.band < (int)g.bands;                                      // a ragged last group has members without a band
        c.colBase += g.colUnits;
        c.q = c.colBase;
        c.band += (int)g.group;
        if (c.band >= (int)(g.groupsPerFrame * g.group)) { c.band -= (int)(g.groupsPerFrame * g.group); ++c.frame; }
        if (r1 > r0 && real) return true;
    }
    return false;
}

struct FusedParams {
    int u16;                 // 0: 8-bit pixels, 1: 16-bit pixels (pitches and frame strides stay in BYTES)
    const uint8_t* a;        // raw planes (used only to fetch the per-piece centring pixel)
    const uint8_t* b;
    long long pitchA, frameStrideA, pitchB, frameStrideB;
    float*  map;             // NULL when no map is wanted
    long long mapPitch, mapFrameStride;   // floats
    long long mapStep;       // floats between horizontally adjacent map values (1 = dense rows)
    long long mapPitchBytes; // mapPitch in bytes.  The kernel shifts mapPitch itself and does not read this field, but it stays: where
                             // the fields behind it sit changes ptxas' register allocation in the consumer body (1125 instead of
                             // 1110 instructions per 11 rows without it; 64 x 4K 2043 instead of 2036 us, profiles/r02_variants.txt)
    int width, srcRows, outY0, outRows, frames;
    SlotGeo geo;
    // reduction workspace (per stream): one word per frame, zero before the launch and again after it: slot count in the
    // high 12 bits, fixed-point sum of (slot sum + accBias) * accScale in the low 52 (see "reduction" in the kernel)
    unsigned long long* frameAcc;
    double accBias, accScale, accInvScale;
    double*   sums;          // out, may be NULL: [frames] sum of the SSIM values of each frame
    float*    ssim;          // out, may be NULL: [frames] float(sum * invCount)
    double    invCount;      // 1 / double(uint32(width*outRows))
    alignas(16) float g[6];  // separable 11-tap weights: g[d] is the tap at distance d from the centre.  16-byte aligned: the hot loops
                             // re-load the taps from the constant bank (one LDCU.128 + one LDCU.64; three loads when the array
                             // sits at an odd multiple of 8 bytes: 2063 instead of 2036 us on 64 x 4K, profiles/r02_variants.txt)
    uint32_t magic;          // 0x4B000000 (float 2^23): kept opaque to ptxas, see the kernel
    uint32_t backoffNs;      // sleep between polls of the partner warp's mbarrier
    float eps2;              // 2*((sum of the 11x11 window) - 1): the reference window's normalisation bias, ~2.05e-8
    unsigned long long* dbgTimes;   // development aid (NULL in production): [slots][kDbgWords] globaltimer of each consumer warp: [0] start,
                                    // [1] end, [1+k] when it finished reading its k-th ring unit (k = 1 .. kDbgWords-2)
};

// Division of item indices by warp-uniform run-time divisors without the 64-bit division subroutine: keeps the whole item
// decode on the uniform datapath.  mul = ceil(2^(31+L) / d), shift = L - 1, L = ceil(log2 d); exact for n < 2^31.
inline void fast_div(uint32_t d, uint32_t* mul, uint32_t* shift)
{
    if (d <= 1) { *mul = 0; *shift = 0; return; }
    uint32_t L = 0;
    while ((1ull << L) < d) ++L;
    const unsigned long long k = 1ull << (31 + L);
    *mul = (uint32_t)((k + d - 1) / d);
    *shift = L - 1;
}

inline SlotGeo make_slot_geo(const SlotPlan& plan, uint32_t width)
{
    SlotGeo g;
    g.slots = plan.slots; g.group = plan.group; g.shareQ = plan.shareQ; g.shareR = plan.shareR; g.colUnits = plan.colUnits; g.pad = plan.pad;
    g.bands = (width + kBandW - 1) / kBandW;
    g.groupsPerFrame = (g.bands + g.group - 1) / g.group;
    fast_div(g.colUnits, &g.colMul, &g.colShift);
    fast_div(g.groupsPerFrame, &g.gpfMul, &g.gpfShift);
    return g;
}

// Cross-GPU sum fused into the kernel (strips of one image): every rank owns an exchange buffer of
// 2 x kMaxRanks slots; the consumer warp of rank r's kernel that completes the strip's sum stores it straight into
// slot [epoch & 1][r] of EVERY peer's buffer (NVLink peer stores: two 64-bit words, each 32 bits of the sum under a 32-bit epoch tag, so no ordering is needed), then
// waits until its own buffer holds all `world` slots of this epoch and adds them up in rank order (deterministic, identical
// on every rank).  world == 0: no exchange.
constexpr int kMaxRanks = 16;
struct ExchangeSlot { unsigned long long lo, hi; };   // (epoch tag << 32) | low / high 32 bits of the double: each word validates itself
struct ExchangeParams {
    ExchangeSlot* peers[kMaxRanks];   // device pointers to every rank's exchange buffer (own one at [rank])
    int world, rank;
    unsigned long long epoch;         // >= 1, same sequence on every rank; parity selects the half of the buffer
    unsigned long long timeoutNs;     // give up waiting after this long (status = 1, result NaN) instead of hanging the GPU
    double* sumAll;                   // out: sum over ranks
    float*  ssimAll;                  // out (may be NULL): float(sumAll * invCountAll)
    double invCountAll;
    int* status;                      // out: 0 ok, 1 timed out
};

// ONE launch per call: TMA loads, both filter passes, the formula, the map, the per-frame reduction and (xchg != NULL) the
// cross-GPU sum.  The grid is ceil(p.geo.slots / kPairsPerCta) CTAs, all resident.
cudaError_t launch_fused(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const FusedParams& p, const ExchangeParams* xchg);
// per-device preparation (sets the dynamic shared-memory limit on the CURRENT device) + kernel facts
cudaError_t fused_kernel_attributes(int* regsMap, int* regsNoMap, int* pairsPerSm);

// layout helpers
cudaError_t launch_pack_u8(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                           long long step, long long stride, int width, int height);
cudaError_t launch_pack_u16(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                            long long step, long long stride, int width, int height);   // step/stride in bytes
cudaError_t launch_pack_luma(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                             long long step, long long stride, int width, int height);
cudaError_t launch_deinterleave_u8(cudaStream_t stream, uint8_t* dst, long long dstPitch, long long planeStride, const uint8_t* src,
                                   long long srcPitch, int channels, int width, int height);
cudaError_t launch_scatter_map(cudaStream_t stream, float* dst, long long dstStep, long long dstStride,
                               const float* src, long long srcPitch, int width, int height);
cudaError_t launch_synth_fill(cudaStream_t stream, uint8_t* dA, long long pitchA, uint8_t* dB, long long pitchB,
                              int width, int rows, int y0, uint32_t frame, uint64_t seed);

}  // namespace ssimk
#endif
