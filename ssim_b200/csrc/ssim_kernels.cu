// ssim_kernels.cu -- hand-written sm_100a kernels for the SSIM hot path.
//
// ssim_fused_kernel replaces, in ONE launch, the reference's per-tile pipeline
//   retrieve_tile (src/ssim.cpp:515-583)  ->  TMA box loads of u8 rows (+5 px halo) into shared memory,
//                                             clamp-to-edge by coordinate clamping (rows) / patching (columns),
//                                             u8 -> f32 widening with PRMT + a mixed-precision add (centred on a per-item pixel)
//   multiply x3   (src/ssim.cpp:249-265)  ->  a'^2 + b'^2 and a'b' in registers (never materialised)
//   gaussian_blur x5 (src/ssim.cpp:321-489, src/ssim_fma.cpp:106-273)
//                                         ->  separable 11-tap horizontal pass (registers -> swizzled smem ring)
//                                             and vertical pass (11-deep shifted accumulators in registers), on FOUR
//                                             planes: E[a'], E[b'], E[(a'-b')^2], E[a'b'] (the reference's E[a^2] and
//                                             E[b^2] are only ever used as their sum, src/ssim.cpp:634-651, and
//                                             sigma_a^2+sigma_b^2 = 2 sigma_ab + var(a-b))
//   sum_tile      (src/ssim.cpp:590-704)  ->  per-pixel formula, coalesced map store, float->double partial sums
// and the reference's OpenMP tile distribution (src/ssim-openmp.c:26-37) by the grid: one warp per
// (frame, row segment, 64-column band) work item, no CTA-level synchronisation at all.
//
// All multiply-adds of the two filter passes are packed fma.rn.f32x2 (SASS FFMA2) whose tap operand is a
// uniform-register scalar: the FMA pipe is the binding resource (DESIGN.md "Roofline"), packed issue leaves the
// other half of the issue slots to LDS/STS/PRMT/address work.
#include "ssim_kernels.h"
#include "synth.h"
#include <string.h>

namespace ssimk {

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
    uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ u64 lds64(uint32_t addr) {
    u64 v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, u64 v) {
    float lo, hi; unpack2(v, lo, hi);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(addr), "f"(lo), "f"(hi) : "memory");
}

template <int kOff>
__device__ __forceinline__ void stg_f32(unsigned long long addr, float v) {
    asm volatile("st.global.f32 [%0+%1], %2;" :: "l"(addr), "n"(kOff), "f"(v) : "memory");
}

// predicated store, the predicate travelling as a register (flag != 0): one SETP + one predicated STG, no branch, and the
// compiler cannot re-derive the column test from the lane index in every row
template <int kOff>
__device__ __forceinline__ void stg_f32_if(unsigned long long addr, float v, uint32_t flag) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\n@q st.global.f32 [%0+%1], %2;\n}\n" :: "l"(addr), "n"(kOff), "f"(v), "r"(flag) : "memory");
}

// "if (pred) { store v; sum += v; }" as ONE predicate and two predicated instructions (no branch): the predicate travels as a
// register between rows; kStore = false leaves only the predicated add
template <int kOff, bool kStore>
__device__ __forceinline__ void store_and_sum_if(unsigned long long addr, float v, uint32_t pred, float& sum) {
    if (kStore)
        asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %4, 0;\n@q st.global.f32 [%1+%2], %3;\n@q add.f32 %0, %0, %3;\n}\n"
                     : "+f"(sum) : "l"(addr), "n"(kOff), "f"(v), "r"(pred) : "memory");
    else
        asm("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q add.f32 %0, %0, %1;\n}\n" : "+f"(sum) : "f"(v), "r"(pred));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// u8 -> f32 without the conversion pipe (I2F.U8 runs at 1/8 rate on B200, tools/microbench): PRMT puts pixel bytes under the
// exponent byte of a float (16-bit pixels: 2^23 + pixel as f32) or of two halves (8-bit pixels: 1024 + pixel as f16x2, one PRMT
// for two pixels); the add that follows removes the offset together with the centring pixel, exactly.  PRMT costs ~1.4 issue
// cycles under a packed-FMA stream (profiles/r01_issue_port_microbench.txt), so halving their number is worth 1.5% of the kernel.
// f16 + f32 -> f32 (mixed-precision add, sm_100: SASS FHADD) on the low / high half of a register
__device__ __forceinline__ float add_f16lo(uint32_t h2, float c) {
    float r; asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %1;\nadd.rn.f32.f16 %0, lo, %2;\n}\n" : "=f"(r) : "r"(h2), "f"(c)); return r;
}
__device__ __forceinline__ float add_f16hi(uint32_t h2, float c) {
    float r; asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %1;\nadd.rn.f32.f16 %0, hi, %2;\n}\n" : "=f"(r) : "r"(h2), "f"(c)); return r;
}

// ------------------------------------------------------------------------------------------------ fused kernel
// The kernel is PERSISTENT: the grid is one CTA of 8 warp pairs per SM (all resident at once) and every warp pair owns one
// SLOT of the work line (ssim_kernels.h, "work partition"): an equal share of the rows of all (frame, 64-column band)
// columns, cut into pieces at column boundaries.  A pair walks through its pieces without ever leaving the kernel:
//
//   producer warp (warps 0-7)   TMA loads of 8-row pixel boxes into a 2-stage ring (running two blocks ahead, across piece
//                               boundaries), clamp patching, u8/u16 -> f32, products, horizontal 11-tap pass; writes 8-row blocks
//                               of {E_h[a'], E_h[b']}, {E_h[(a'-b')^2], E_h[a'b']} into a shared-memory ring of two halves of 11
//                               rows, a full/empty mbarrier each
//   consumer warp (warps 8-15)  vertical 11-tap pass with eleven IN-PLACE accumulators per plane pair: its loop body is 11 rows
//                               = one ring half, fully unrolled, so accumulator slot s always owns the output rows == s (mod 11),
//                               every tap index and every ring offset is static and nothing has to be shifted or renamed across
//                               iterations; then the SSIM formula, map store and partial sums; when the slot leaves a frame it
//                               adds its sum to the frame's accumulator word with one atomic, and the slot whose atomic completes
//                               the frame writes the result and (strips across GPUs) exchanges it with the peers over NVLink
//
// The two roles overlap in time (the consumer's dependent formula chain hides behind the producer's FMAs and vice versa),
// setmaxnreg moves registers from the producers (120) to the consumers (136), and nothing is ever synchronised CTA-wide
// after the prologue.  Every hand-over (TMA stage full/empty, ring half full/empty) is an mbarrier on which each lane
// releases its own accesses and each lane acquires for itself: see the protocol table in DESIGN.md section 4.
//
// CODE LAYOUT MATTERS.  The two hot loops (consumer body ~1100 instructions, producer block loop ~750) together are ~30 KB
// of code, and the instruction cache level behind the per-scheduler L0s holds 32 KB: when the address range from the
// first instruction of the consumer body to the last of the producer loop exceeded it (36 KB, because the producer's
// prologue and per-piece code sat between the two loops) 8% of all issue-slot samples were stall_no_instruction and the
// kernel ran 9% slower (profiles/r02_code_layout.txt).  Hence: the first two pieces of a slot are looked up in the
// kernel's common prologue, the producer's per-piece code sits BEHIND its block loop, cold paths inside the loops are
// kept short, and tools/sass_summary.py reports the hot range.
struct PieceGeo {           // everything warp-uniform
    int frame, bx, oy0, nOut, inY0;
    int nRows;              // input rows the vertical pass consumes: nOut + 10
    int nBlk;               // 8-row blocks the producer makes: ceil(nRows / 8) (rows past nRows are filler the consumer skips)
};

__device__ __forceinline__ void piece_geo(const FusedParams& p, const Piece& pc, PieceGeo& g)
{
    g.frame = pc.frame;
    g.bx    = pc.band * kBandW;                                         // first output column of the band
    g.oy0   = p.outY0 + pc.r0;                                          // first output row (plane coordinates)
    g.nOut  = pc.nOut;
    g.inY0  = g.oy0 - kHalo;                                            // first input row needed (may be negative)
    g.nRows = pc.nOut + 2 * kHalo;
    g.nBlk  = (g.nRows + kBlkRows - 1) / kBlkRows;
}

// The per-piece centring pixels: moments are accumulated on (a - ca), (b - cb), which keeps the fp32 cancellation in
// E[x^2] - mu^2 small even on flat regions (DESIGN.md "Numerics").  Any integer works; both warps of a pair must of course
// use the same one.
template <bool kU16>
__device__ __forceinline__ void piece_centre_raw(const FusedParams& p, const PieceGeo& g, unsigned& ra, unsigned& rb)
{
    const int cx = min(g.bx + kBandW / 2, p.width - 1);
    const int cy = min(max(g.oy0, 0), p.srcRows - 1);
    const uint8_t* pa = p.a + (long long)g.frame * p.frameStrideA + (long long)cy * p.pitchA;
    const uint8_t* pb = p.b + (long long)g.frame * p.frameStrideB + (long long)cy * p.pitchB;
    if (kU16) { ra = __ldg((const uint16_t*)pa + cx); rb = __ldg((const uint16_t*)pb + cx); }
    else      { ra = __ldg(pa + cx);                  rb = __ldg(pb + cx); }
}
template <bool kU16>
__device__ __forceinline__ void piece_centre(const FusedParams& p, const PieceGeo& g, float& ca, float& cb)
{
    unsigned ra, rb;
    piece_centre_raw<kU16>(p, g, ra, rb);
    ca = (float)ra; cb = (float)rb;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// Waiting on an mbarrier.  Every lane polls (each lane needs its own acquire), but the decision to leave the loop is a
// warp vote, so the lanes of a warp can never leave a wait in different iterations.  (With warp-uniform barrier
// addresses ptxas emits no reconvergence point after these loops and turns __syncwarp() into a NOP, so lanes that got
// apart would stay apart.)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!__all_sync(0xffffffffu, mbar_test(bar, parity))) { }
}
// Waiting on the partner warp: same, with an explicit nanosleep between polls (try_wait's own suspend-time hint compiles
// to a NANOSLEEP.SYNCS loop that re-polls almost immediately).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned backoffNs) {
    while (!__all_sync(0xffffffffu, mbar_test(bar, parity))) __nanosleep(backoffNs);
}

// mbarriers of a pair (byte offsets from the pair's first barrier)
constexpr uint32_t kBarTmaFull = 0, kBarStageEmpty = 16, kBarRingFull = 32, kBarRingEmpty = 48;

// ---- producer: TMA + horizontal pass
template <bool kU16>
__device__ __forceinline__ void producer_warp(const CUtensorMap* tmA, const CUtensorMap* tmB, const FusedParams& p, int lane,
                                              uint32_t pairSmem, uint32_t barBase, PieceCursor cur, const PieceGeo& g0, unsigned ra0, unsigned rb0,
                                              bool have1, const PieceGeo& g1, unsigned ra1, unsigned rb1)
{
    typedef PixGeo<kU16> G;
    constexpr int kBoxW = G::kBoxBytes, kImgStageBytes = G::kImgStageBytes, kStageBytes = G::kStageBytes;
    const uint32_t ringBase = pairSmem + kStages * kStageBytes;
    const uint32_t barTma = barBase + kBarTmaFull, barStageEmpty = barBase + kBarStageEmpty;
    const uint32_t barFull = barBase + kBarRingFull, barEmpty = barBase + kBarRingEmpty;

    // One TMA load = an 8-row box x bytes [bx-16, bx+80) of both images (16-bit: elements [bx-8, bx+72)).  A block needs rows [y, y+8) with y = inY0 + 8*blk,
    // rows outside the plane replicating the nearest one (src/ssim.cpp:562-582): the distinct rows it needs always fit in
    // the 8-row box starting at clamp(y, 0, srcRows-8), so edge blocks load that box and each lane reads the row
    // clamp(y + hr) of it.  Columns outside the plane arrive as zeros and are patched after landing (src/ssim.cpp:541-554).
    const int lastBoxY = max(p.srcRows - kLoadRows, 0);

    // TMA issue runs exactly kStages (= 2) blocks ahead of the computation, across piece boundaries: the block that refills
    // the stage of block `blk` of the current piece is block blk + 2 of the same piece or, in its last two blocks, block 0
    // or 1 of the NEXT piece (every piece has >= 2 blocks), whose geometry and centring pixels are fetched one piece ahead.
    auto issue = [&](const PieceGeo& ge, int blkIdx, uint32_t stage, bool refill, uint32_t emptyParity, bool patched) {
        // one elected lane of the (converged) warp: with elect.sync ptxas emits the issue sequence once, behind one branch;
        // `lane == 0` made it wrap every UTMALDG in a loop that broadcasts the five operands lane by lane (-0.5% kernel time)
        uint32_t leader;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
        if (leader) {
            // TMA goes through the uniform datapath: one lane, warp-uniform operands (never issue it from divergent lanes).
            // The stage may be refilled once EVERY lane's loads of its previous contents have been performed: each lane
            // releases the stage through an mbarrier (count 32) and the elected lane acquires it before re-arming the TMA barrier.
            // Program order plus __syncwarp() is not enough here -- with the refill issued straight after the loads,
            // tools/dev/stress.py saw rare 8-row x 16-column blocks computed from the NEXT box's bytes.
            if (refill) {
                while (!mbar_test(barStageEmpty + 8 * stage, emptyParity)) { }
                if (patched) fence_proxy_async();           // the patch stores (generic proxy) precede the TMA write
            }
            const uint32_t bar = barTma + 8 * stage;
            const uint32_t dst = pairSmem + stage * kStageBytes;
            const int y0 = min(max(ge.inY0 + blkIdx * kLoadRows, 0), lastBoxY);
            mbar_arrive_expect_tx(bar, kStageBytes);
            tma_load_3d(dst, tmA, ge.bx - G::kBoxLeftElems, y0, ge.frame, bar);
            tma_load_3d(dst + kImgStageBytes, tmB, ge.bx - G::kBoxLeftElems, y0, ge.frame, bar);
        }
    };
    PieceGeo g = g0, gN = g1;                               // the slot's first two pieces: found by the kernel's common prologue
    bool haveN = have1;
    #pragma unroll 1
    for (int b = 0; b < kStages; ++b) issue(g, b, (uint32_t)b, false, 0, false);
    // the centring pixels were loaded by the common prologue and are first touched HERE, after the first TMA boxes are on
    // their way (converting them there put a DRAM round trip in front of the first TMA issue of every launch)
    // (volatile: keeps the conversions, i.e. the wait for the loads, behind the TMA issue above)
    float ca, cb, caN, cbN;
    asm volatile("cvt.rn.f32.u32 %0, %4;\n\tcvt.rn.f32.u32 %1, %5;\n\tcvt.rn.f32.u32 %2, %6;\n\tcvt.rn.f32.u32 %3, %7;"
                 : "=f"(ca), "=f"(cb), "=f"(caN), "=f"(cbN) : "r"(ra0), "r"(rb0), "r"(ra1), "r"(rb1));

    const uint32_t magic = p.magic;                    // 0x4B000000, passed as a parameter so that it lives in a register and
                                                       // PRMT takes the byte selector as its immediate (no per-PRMT selector MOV)
    u64 w2[6];
    #pragma unroll
    for (int d = 0; d < 6; ++d) w2[d] = pack2(p.g[d], p.g[d]);
    #define TAP(m) w2[(m) < 5 ? 5 - (m) : (m) - 5]

    // This lane's share of a block: row hr, 16 output columns starting at 16*hq.  Lane bits: [1:0] = hr & 3, [3:2] = hq,
    // [4] = hr >> 2.  With 96-byte box rows a quarter-warp (4 rows x 2 adjacent column groups) then reads eight distinct
    // 16-byte bank groups in each of its three 128-bit loads (row r starts 6r groups in, column group hq adds hq), and a
    // half-warp (4 rows x 4 groups) hits 16 distinct 8-byte slots of the ring with each store (see below).
    const int hr = (lane & 3) | ((lane >> 2) & 4), hq = (lane >> 2) & 3;
    // 8-bit: three 16-byte chunks holding columns 16hq-16 .. 16hq+31; 16-bit: 64-byte window holding columns 16hq-8 .. 16hq+23
    const uint32_t hColOff = kU16 ? hq * 32 : hq * 16;
    const uint32_t hSrcOff = hr * kBoxW + hColOff;
    // Ring: two halves of 11 rows (one half = one consumer body), a full/empty mbarrier each.  Input row i of a piece lives
    // in half (firstHalf + i / 11) & 1 at row t = i % 11 (every piece starts a new half).  Layout of a ring row (1056 bytes):
    // two planes of 64 packed pairs, {E_h[a'], E_h[b']} at +0 and {E_h[(a'-b')^2], E_h[a'b']} at +512, then 32 bytes of
    // padding; column c sits at 8*(c ^ (c>>4)).  Bank-conflict freedom without any per-access arithmetic: in the producer's
    // 8-byte stores a half-warp is 4 rows x 4 column groups writing the same j -- the XOR by the group index spreads the
    // groups over 4 adjacent slots and the 32-byte pad moves each following row by 4 slots (16 distinct slots); in the
    // consumer's 8-byte loads a half-warp reads 16 adjacent columns of one group (XOR by a constant).  Both sides address
    // "register + immediate": the store to column j goes to dst[j & 3] + 32*(j >> 2), the load of row t comes from
    // base + 1056*t.

    uint32_t gblk = 0;                                      // blocks computed so far
    uint32_t firstHalf = 0;                                 // ring halves (global count) used by the pieces before this one
    uint32_t acquired = 0;                                  // ring halves (global count) this warp may write
    uint32_t released = 0;                                  // ring halves (global count) handed to the consumer
    #pragma unroll 1
    for (;;) {
        // (a - ca, b - cb) from the bytes, exact: see the widening in the block loop
        const u64 negMagic = pack2(-(8388608.0f + ca), -(8388608.0f + cb));     // 16-bit pixels
        const float negCa = -(1024.0f + ca), negCb = -(1024.0f + cb);           // 8-bit pixels
        uint32_t hpa = 0, hpb = 0;                                              // the current pair of pixels of A / B as f16x2
        const float k2 = -0.5f * p.eps2 * (ca - cb) * (ca - cb);  // see the formula in consumer_warp()
        const bool patchLeft  = (g.bx == 0);
        const bool patchRight = (g.bx + kBandW + kHalo > p.width);
        const bool patched = patchLeft || patchRight;
        // input row i of the piece lives in ring row ((firstHalf & 1) * 11 + i) mod 22: every piece starts a new half
        const uint32_t ringLaneEnd = ringBase + hq * 128 + kRingRows * kRingRowBytes;
        uint32_t ringRowAddr = ringBase + hq * 128 + ((firstHalf & 1u) * kTaps + hr) * kRingRowBytes;

        #pragma unroll 1
        for (int blk = 0; blk < g.nBlk; ++blk) {
            const uint32_t stage = gblk & 1u;
            const uint32_t stageBase = pairSmem + stage * kStageBytes;
            mbar_wait(barTma + 8 * stage, (gblk >> 1) & 1u);

            if (patched) {                                       // warp-uniform; only the outermost bands
                if (lane < 2 * kLoadRows) {
                    const uint32_t row = stageBase + (lane >> 3) * kImgStageBytes + (lane & 7) * kBoxW;   // 2 images x 8 rows
                    const int lastCol = p.width - 1 - g.bx;      // last plane column, relative to the band (right edge only)
                    if (kU16) {
                        if (patchLeft) {
                            uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(row + kBoxLeft));
                            #pragma unroll
                            for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u16 [%0], %1;" :: "r"(row + kBoxLeft - 2 * k), "r"(v) : "memory");
                        }
                        if (patchRight) {
                            const uint32_t last = row + kBoxLeft + 2 * lastCol;
                            uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(last));
                            #pragma unroll
                            for (int k = 1; k <= kHalo; ++k)         // only columns the band reads (< 64 + 5): stays inside the box row
                                if (lastCol + k < kBandW + kHalo) asm volatile("st.shared.u16 [%0], %1;" :: "r"(last + 2 * k), "r"(v) : "memory");
                        }
                    } else {
                        if (patchLeft) {
                            uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row + kBoxLeft));
                            #pragma unroll
                            for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u8 [%0], %1;" :: "r"(row + kBoxLeft - k), "r"(v) : "memory");
                        }
                        if (patchRight) {
                            const uint32_t last = row + kBoxLeft + lastCol;
                            uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(last));
                            #pragma unroll
                            for (int k = 1; k <= kHalo; ++k)
                                if (lastCol + k < kBandW + kHalo) asm volatile("st.shared.u8 [%0], %1;" :: "r"(last + k), "r"(v) : "memory");
                        }
                    }
                }
                __syncwarp();
            }

            // this lane's row of the box: hr, except in edge blocks (warp-uniform test) where rows replicate
            uint32_t src = stageBase + hSrcOff;
            {
                const int y = g.inY0 + blk * kLoadRows;
                if (y < 0 || y + kLoadRows > p.srcRows) {
                    const int y0 = min(max(y, 0), lastBoxY);
                    src = stageBase + (uint32_t)(min(max(y + hr, 0), p.srcRows - 1) - y0) * kBoxW + hColOff;
                }
            }
            uint32_t wa[kU16 ? 16 : 12], wb[kU16 ? 16 : 12];
            #pragma unroll
            for (int q = 0; q < (kU16 ? 4 : 3); ++q) {
                const uint4 va = lds128(src + 16 * q), vb = lds128(src + kImgStageBytes + 16 * q);
                wa[4 * q] = va.x; wa[4 * q + 1] = va.y; wa[4 * q + 2] = va.z; wa[4 * q + 3] = va.w;
                wb[4 * q] = vb.x; wb[4 * q + 1] = vb.y; wb[4 * q + 2] = vb.z; wb[4 * q + 3] = vb.w;
            }
            // every lane releases the stage (the acquire + refill sit at ii == 10 below: by then the 32 arrivals have long
            // drained and the issuing lane never spins)
            mbar_arrive(barStageEmpty + 8 * stage);

            // ring position of this lane's row (see above)
            // the ring row of this lane's row, kept as an address that advances by 8 rows per block (mod 22 rows)
            const uint32_t dstRow = ringRowAddr;
            ringRowAddr += kBlkRows * kRingRowBytes;
            if (ringRowAddr >= ringLaneEnd) ringRowAddr -= kRingRows * kRingRowBytes;
            uint32_t dst4[4];
            #pragma unroll
            for (int m = 0; m < 4; ++m) dst4[m] = dstRow + ((uint32_t)(m ^ hq) << 3);

            u64 hab[16], hsp[16];
            #pragma unroll
            for (int ii = 0; ii < 26; ++ii) {                // input column 16hq - 5 + ii
                // 8-bit: byte 11 + ii of the 48-byte window; 16-bit: halfword 3 + ii of the 64-byte window
                const int byteIdx = kU16 ? ii + 3 : ii + 11;
                float a, b;                                             // (a - ca, b - cb), exact
                if (!kU16) {
                    // 8-bit: one PRMT makes TWO pixels of an image as f16 values 1024 + byte (bytes {pixel, 0x64, next pixel, 0x64});
                    // the mixed-precision add f16 + f32 -> f32 (SASS FHADD) takes 1024 + centre off again on the way to f32.
                    // Pairs start at even bytes (ii odd); the first and the last column of the window are singles.
                    if (ii == 0 || (ii & 1)) {
                        if ((byteIdx & 3) == 3)  { hpa = __byte_perm(wa[byteIdx >> 2], magic, 0x4343); hpb = __byte_perm(wb[byteIdx >> 2], magic, 0x4343); }
                        else if (byteIdx & 2)    { hpa = __byte_perm(wa[byteIdx >> 2], magic, 0x4342); hpb = __byte_perm(wb[byteIdx >> 2], magic, 0x4342); }
                        else                     { hpa = __byte_perm(wa[byteIdx >> 2], magic, 0x4140); hpb = __byte_perm(wb[byteIdx >> 2], magic, 0x4140); }
                        a = add_f16lo(hpa, negCa); b = add_f16lo(hpb, negCb);
                    } else {
                        a = add_f16hi(hpa, negCa); b = add_f16hi(hpb, negCb);
                    }
                } else {
                    // 16-bit: 2^23 + pixel as f32 -- the two pixel bytes under the two top bytes of the magic word (PRMT with an
                    // immediate selector) -- and one packed FADD that removes 2^23 + centre
                    float fa, fb;
                    if (byteIdx & 1) { fa = __uint_as_float(__byte_perm(wa[byteIdx >> 1], magic, 0x7632)); fb = __uint_as_float(__byte_perm(wb[byteIdx >> 1], magic, 0x7632)); }
                    else             { fa = __uint_as_float(__byte_perm(wa[byteIdx >> 1], magic, 0x7610)); fb = __uint_as_float(__byte_perm(wb[byteIdx >> 1], magic, 0x7610)); }
                    unpack2(add2(pack2(fa, fb), negMagic), a, b);
                }
                const u64 ab = pack2(a, b);
                const float d = a - b;
                const u64 sp = pack2(fmaf(d, d, k2), a * b);           // ((a'-b')^2 + k2, a'b')
                #pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = ii - j;                               // sample ii is tap k of output j
                    if (k == 0)                { hab[j] = mul2(ab, TAP(0)); hsp[j] = mul2(sp, TAP(0)); }
                    else if (k > 0 && k <= 10) { hab[j] = fma2(ab, TAP(k), hab[j]); hsp[j] = fma2(sp, TAP(k), hsp[j]); }
                }
                if (ii == 10) {
                    // refill this block's stage with the block two ahead (all lanes have released it above)
                    const bool own = blk + kStages < g.nBlk;         // one call site: the TMA issue sequence is not small
                    if (own || haveN) issue(own ? g : gN, own ? blk + kStages : blk + kStages - g.nBlk, stage, true, (gblk >> 1) & 1u, patched);
                    // first store of the block: the ring halves this block touches must have been drained by the consumer
                    // (waiting here, not at the top, lets the loads and the first 10 columns of math overlap the wait); parity
                    // of the half's previous use -- its first use passes at once on the fresh barrier.  A piece's last block
                    // may spill filler rows into the half after the piece's last one: that is the next piece's first half,
                    // acquired here already and simply overwritten by it.
                    const uint32_t lastHalf = firstHalf + ((uint32_t)blk * kBlkRows + kBlkRows - 1) / kTaps;
                    while (acquired <= lastHalf) {
                        mbar_wait_sleep(barEmpty + 8 * (acquired & 1u), ((acquired >> 1) & 1u) ^ 1u, p.backoffNs);
                        ++acquired;
                    }
                }
                if (ii >= 10) {                                         // output j = ii-10 is complete
                    const int j = ii - 10;
                    const uint32_t dst = dst4[j & 3] + 32 * (j >> 2);
                    sts64(dst, hab[j]);
                    sts64(dst + kRingPlaneBytes, hsp[j]);
                }
            }
            // every lane arrives (barrier count 32): each lane's release covers its own stores, no reliance on warp-level
            // cumulativity.  Halves completely written so far; the piece's last block also completes its last, partial half.
            const uint32_t nBodies = ((uint32_t)g.nRows + kTaps - 1) / kTaps;
            const uint32_t complete = firstHalf + (blk + 1 == g.nBlk ? nBodies : min(((uint32_t)blk * kBlkRows + kBlkRows) / kTaps, nBodies));
            // (a block completes at most two halves: 8 rows of its own plus, in a piece's last block, the partial last half)
            if (released < complete) { mbar_arrive(barFull + 8 * (released & 1u)); ++released; }
            if (released < complete) { mbar_arrive(barFull + 8 * (released & 1u)); ++released; }
            ++gblk;
        }
        firstHalf += ((uint32_t)g.nRows + kTaps - 1) / kTaps;
        if (!haveN) break;
        // The last block of a piece may have put filler rows into the half after the piece's last one, which is the next
        // piece's first half: other lanes of this warp are about to overwrite them.  Same warp, program order -- but two
        // lanes writing one address need a warp-level sync in between to be ordered (compute-sanitizer racecheck reports the
        // pair otherwise).
        __syncwarp();
        // the piece fetched one ahead becomes the current one; fetch the next.  (This code sits BEHIND the block loop in the
        // binary; its first two executions were moved into the kernel's common prologue so that no cold code lies between the
        // consumer's body and the block loop.)
        g = gN; ca = caN; cb = cbN;
        Piece pc;
        haveN = cursor_next(cur, p.geo, pc);
        if (haveN) { piece_geo(p, pc, gN); piece_centre<kU16>(p, gN, caN, cbN); }
    }
    #undef TAP
}

// ---- reduction: ONE 64-bit atomic per slot and frame, and nothing else.  Every frame owns one word of the workspace: the
// low 52 bits accumulate the slots' sums in fixed point, the high 12 bits count the slots that have delivered.  A slot adds
// (1 << 52) | fixed(sum + bias) with one atomicAdd; atomics on one address are totally ordered and return the old value, so
// the slot whose atomic returns count == expected - 1 holds the frame's total in (old + own) -- no fence, no second counter,
// nothing to poll -- and finishes the frame: result out, word back to zero for the next launch on the stream.  Integer
// addition is associative, so the total does not depend on the order in which the slots arrive: the result is deterministic
// (and, up to the rounding of each slot's sum to the fixed-point grid, independent of how the work was partitioned).
// bias: SSIM values may be negative (> -1); every slot adds p.accBias >= its number of pixels so that what goes into the
// field is non-negative, and the finisher takes expected * bias off again.  Scale: p.accScale = 2^k with k chosen by the
// host (acc_format() in ssim_kernels.h) so that the field cannot overflow (one 4K frame: k = 25, i.e. 1.5e-8 per slot; the rounding of all slots together
// moves the mean SSIM of a 4K frame by < 1e-12).
// (Protocols tried before: partial sums in memory + fence + atomic counter + the last arriver adds them up: 4.5 us behind
// the last row of a single image -- the fence waits for the slot's map stores, then the atomic's round trip, then 1184
// loads; per-slot entries polled by a designated reducer: 4 us, the polling rounds are two dependent groups of loads.)
// Strips across GPUs: the finishing slot also exchanges the strip sums with the peers.
constexpr unsigned long long kAccCountShift = 52, kAccSumMask = (1ull << kAccCountShift) - 1ull;

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ uint32_t share_of_unit(const SlotGeo& g, uint32_t q)
{
    const uint32_t big = g.shareR * (g.shareQ + 1u);
    return q < big ? q / (g.shareQ + 1u) : g.shareR + (q - big) / g.shareQ;
}

// Slot delivers v = its sum of frame f's values (0 when its units there held no rows).
__device__ __forceinline__ void frame_deliver(const FusedParams& p, const ExchangeParams& x, int f, double v, int lane)
{
    // the slots that deliver to this frame are the members of the teams whose units intersect the frame's: a contiguous run
    uint32_t expected = p.geo.slots;
    if (p.frames != 1) {
        const uint32_t frameUnits = p.geo.groupsPerFrame * p.geo.colUnits;
        const uint32_t u0 = (uint32_t)f * frameUnits;
        expected = (share_of_unit(p.geo, u0 + frameUnits - 1u) - share_of_unit(p.geo, u0) + 1u) * p.geo.group;
    }
    unsigned long long old = 0ull, mine = 0ull;
    if (lane == 0) {
        mine = (1ull << kAccCountShift) | __double2ull_rn((v + p.accBias) * p.accScale);
        old = atomicAdd(p.frameAcc + f, mine);
    }
    old = __shfl_sync(0xffffffffu, old, 0);
    mine = __shfl_sync(0xffffffffu, mine, 0);
    if (!__all_sync(0xffffffffu, (uint32_t)(old >> kAccCountShift) + 1u == expected)) return;     // not the last one (a vote: warp-uniform for ptxas)
    const double acc = (double)((old + mine) & kAccSumMask) * p.accInvScale - (double)expected * p.accBias;
    if (lane == 0) {
        p.frameAcc[f] = 0ull;                           // ready for the next launch on this stream
        if (p.sums) p.sums[f] = acc;
        if (p.ssim) p.ssim[f] = (float)(acc * p.invCount);
    }
    if (x.world > 0) {
        // strip sums of all ranks (frames == 1): one lane per peer stores the sum into the peer's buffer as two 64-bit words,
        // each carrying 32 bits of the double under a 32-bit epoch tag, then waits until both words of the peer's slot in its
        // own buffer carry this epoch; the sum runs in rank order on every rank.  Every word validates itself, so the stores
        // and loads are relaxed: no release / acquire at system scope (the release used to wait for this lane's map stores).
        const unsigned half = (unsigned)(x.epoch & 1ull) * kMaxRanks;
        const unsigned long long tag = ((x.epoch % 0xffffffffull) + 1ull) << 32;         // never 0: fresh buffers are zero
        double val = 0.0;
        int failed = 0;
        if (lane < x.world) {
            ExchangeSlot* dst = x.peers[lane] + half + x.rank;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(acc);
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(&dst->lo), "l"(tag | (bits & 0xffffffffull)) : "memory");
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(&dst->hi), "l"(tag | (bits >> 32)) : "memory");
            const ExchangeSlot* src = x.peers[x.rank] + half + lane;
            unsigned long long t0, now, lo, hi;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(lo) : "l"(&src->lo) : "memory");
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(hi) : "l"(&src->hi) : "memory");
                if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now - t0 > x.timeoutNs) { failed = 1; break; }
            }
            val = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
        }
        failed = __any_sync(0xffffffffu, failed);
        double total = 0.0;
        for (int r = 0; r < x.world; ++r) total += __shfl_sync(0xffffffffu, val, r);
        if (lane == 0) {
            if (failed) total = __longlong_as_double(0x7ff8000000000000ll);
            *x.sumAll = total;
            if (x.ssimAll) *x.ssimAll = (float)(total * x.invCountAll);
            if (x.status) *x.status = failed;
        }
    }
}

// ---- consumer: vertical pass + formula + outputs
// One input row of the vertical pass at position T of the 11-row body, for this lane's two columns: 44 FFMA2 into the
// in-place accumulators, then the SSIM value of the output row this input row completes.
template <bool kU16>
__device__ __forceinline__ void vertical_row(const int T, u64 (&qab0)[kTaps], u64 (&qsp0)[kTaps], u64 (&qab1)[kTaps], u64 (&qsp1)[kTaps], const u64 (&w2)[6],
                                             uint32_t addr0, uint32_t addr1, float ca, float cb, float eps2, float& sv0, float& sv1)
{
    #define TAP(m) w2[(m) < 5 ? 5 - (m) : (m) - 5]
    // (0.01*L)^2, (0.03*L)^2 as float, L = 255 (src/ssim.cpp:956-960) or 65535 (the 16-bit extension the reference's README names)
    constexpr float c1 = kU16 ? 429483.6225f : 6.5025f, c2 = kU16 ? 3865352.6025f : 58.5225f;
    const u64 hab0 = lds64(addr0), hsp0 = lds64(addr0 + kRingPlaneBytes);
    const u64 hab1 = lds64(addr1), hsp1 = lds64(addr1 + kRingPlaneBytes);
    #pragma unroll
    for (int s = 0; s < kTaps; ++s) {
        const int k = (T - s + kTaps) % kTaps;
        if (k == 0) {
            qab0[s] = mul2(hab0, TAP(0)); qsp0[s] = mul2(hsp0, TAP(0));
            qab1[s] = mul2(hab1, TAP(0)); qsp1[s] = mul2(hsp1, TAP(0));
        } else {
            qab0[s] = fma2(hab0, TAP(k), qab0[s]); qsp0[s] = fma2(hsp0, TAP(k), qsp0[s]);
            qab1[s] = fma2(hab1, TAP(k), qab1[s]); qsp1[s] = fma2(hsp1, TAP(k), qsp1[s]);
        }
    }
    #undef TAP
    // Output row completed by this input row.  The formula is evaluated unconditionally (the first 10 rows of a piece only
    // cost the pipeline fill); the store and the sum are predicated by the caller.
    const int done = (T + 1) % kTaps;
    float sv[2];
    #pragma unroll
    for (int c = 0; c < 2; ++c) {
        float ma, mb, D, P;
        unpack2(c == 0 ? qab0[done] : qab1[done], ma, mb);
        unpack2(c == 0 ? qsp0[done] : qsp1[done], D, P);
        // The reference formula (src/ssim.cpp:590-704) rearranged so that numerator and denominator share
        // their terms:  mu_a^2 + mu_b^2 = 2 mu_a mu_b + (mu_a - mu_b)^2  and
        // sigma_a^2 + sigma_b^2 = 2 sigma_ab + var(a - b),  var(a-b) = E[(a'-b')^2] - (E[a'] - E[b'])^2.
        // Identical images then give num == den bit for bit, hence exactly 1 like the reference.
        // The reference's window sums to 1+eps (see gaussian_taps() in ssim_cuda.cu), which on its RAW moments
        // shifts every covariance by -eps*mu_a*mu_b; centred moments only see -eps*ma*mb, so the difference
        // -eps*(mu_a mu_b - ma mb) is applied explicitly (the matching -eps*(ca-cb)^2 of var(a-b) is already
        // inside D: the producer added k2 to every (a'-b')^2 before the blur, turning an FMUL into an FFMA).
        const float mua = ma + ca, mub = mb + cb;
        const float tt  = mua * mub;
        const float n1  = fmaf(2.f, tt, c1);
        const float dmu = mua - mub;
        const float d1  = fmaf(dmu, dmu, n1);
        const float n2  = fmaf(-eps2, fmaf(-ma, mb, tt), fmaf(2.f, fmaf(-ma, mb, P), c2));
        const float dm  = ma - mb;
        const float d2  = n2 + fmaf(-dm, dm, D);
        const float num = n1 * n2, den = d1 * d2;
        // den >= c1*c2 > 0 on valid rows.  MUFU.RCP + one Newton step on the quotient: ~correctly rounded,
        // exact when num == den
        float rc; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));
        const float q = num * rc;
        sv[c] = fmaf(rc, fmaf(-q, den, num), q);
    }
    sv0 = sv[0]; sv1 = sv[1];
}

// kMap: 0 = no map, 1 = dense map rows (x step 1), 2 = map with a pixel step (interleaved maps: all channels of an image
// are written by one launch straight into the caller's layout, map[(y*pitch) + x*step + channel])
template <int kMap, bool kU16>
__device__ __forceinline__ void consumer_warp(const FusedParams& p, const ExchangeParams& x, uint32_t slot, int lane, uint32_t pairSmem, uint32_t barBase)
{
    typedef PixGeo<kU16> G;
    const uint32_t ringBase = pairSmem + kStages * G::kStageBytes;
    const uint32_t barFull = barBase + kBarRingFull, barEmpty = barBase + kBarRingEmpty;
    u64 w2[6];
    #pragma unroll
    for (int d = 0; d < 6; ++d) w2[d] = pack2(p.g[d], p.g[d]);
    const float eps2 = p.eps2;

    // this lane owns columns bx+lane and bx+32+lane (ring layout: see producer_warp); column c sits at 8*(c ^ (c>>4))
    const uint32_t vOff0 = ((uint32_t)(lane ^ (lane >> 4)) << 3);
    const uint32_t vOff1 = ((uint32_t)((32 + lane) ^ (2 + (lane >> 4))) << 3);

    // Eleven in-place accumulators per plane pair and column: slot s accumulates the output row whose first input row
    // is == s (mod 11).  At row t of a body slot s receives tap (t - s) mod 11; the slot receiving tap 0 is re-initialised,
    // the slot receiving tap 10 is complete.  All indices are compile-time constants.  They are cleared at the start of every
    // piece: not needed for the results (what they hold only reaches the 10 warm-up rows, which are never stored), but it
    // ends their live ranges at the piece boundary, so the per-frame delivery code between two pieces gets registers without
    // spilling 88 accumulator registers around it.
    u64 qab0[kTaps], qsp0[kTaps], qab1[kTaps], qsp1[kTaps];

    const unsigned long long mapPitchBytes = (unsigned long long)p.mapPitch * sizeof(float);
    const unsigned long long mapCol1Bytes = (unsigned long long)p.mapStep * 32 * sizeof(float);      // kMap == 2: this lane's second column
    // ring: this lane's two column addresses in half 0; the half of body number `gbody` (counted over all pieces of the slot) is
    // gbody & 1, rows of a body sit at +1056*t (static offsets: the unrolled body is exactly one half)
    const uint32_t col0Base = ringBase + vOff0, col1Base = ringBase + vOff1;
    uint32_t gbody = 0;

    // the frames this slot's unit range touches: it delivers a sum to every one of them (0 where its units held no rows)
    const uint32_t frameUnits = p.geo.groupsPerFrame * p.geo.colUnits;
    uint32_t q0, qEnd;
    slot_units(p.geo, slot, q0, qEnd);
    if (q0 >= qEnd) return;
    const int fFirst = (int)(q0 / frameUnits), fLast = (int)((qEnd - 1u) / frameUnits);
    int curFrame = fFirst;                                  // frames below this one have been delivered
    double total = 0.0;                                     // this lane's sum of curFrame's values so far

    if (p.dbgTimes && lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbgTimes[kDbgWords * slot] = t; }
    PieceCursor cur;
    cursor_init(cur, p.geo, slot);
    #pragma unroll 1
    for (;;) {
        Piece pc;
        const bool have = cursor_next(cur, p.geo, pc);
        const int f = have ? pc.frame : fLast + 1;
        if (f != curFrame) {
            // the slot moves on to another frame (or is done): deliver curFrame's sum, and zeros for frames whose units
            // here held no rows
            double v = warp_sum(total);
            #pragma unroll 1
            for (; curFrame < f; ++curFrame) {
                frame_deliver(p, x, curFrame, v, lane);     // (the one call site)
                v = 0.0;
            }
            total = 0.0;
        }
        if (!have) {
            if (p.dbgTimes && lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.dbgTimes[kDbgWords * slot + 1] = t; }
            break;
        }
        PieceGeo g;
        piece_geo(p, pc, g);
        #pragma unroll
        for (int m = 0; m < kTaps; ++m) qab0[m] = qsp0[m] = qab1[m] = qsp1[m] = 0ull;
        float ca, cb;
        piece_centre<kU16>(p, g, ca, cb);
        const bool colOk0 = g.bx + lane < p.width;
        const bool colOk1 = g.bx + 32 + lane < p.width;
        // column tests as opaque flags in registers: ptxas otherwise re-derives them from the lane index in every row
        uint32_t okFlag0 = colOk0 ? 1u : 0u, okFlag1 = colOk1 ? 1u : 0u, fullFlag = g.bx + kBandW <= p.width ? 1u : 0u;
        asm volatile("" : "+r"(okFlag0), "+r"(okFlag1), "+r"(fullFlag));
        // Map addressing: one 64-bit per-lane address that advances by the pitch per input row (starts 10 rows above the
        // piece, never dereferenced there); the second column is an immediate offset.  Stores are written in PTX so
        // that the address arithmetic stays these two adds per row.
        unsigned long long mapAddr = 0;
        if (kMap == 1) mapAddr = (unsigned long long)(p.map + (long long)g.frame * p.mapFrameStride + (long long)(pc.r0 - 2 * kHalo) * p.mapPitch + g.bx + lane);
        if (kMap == 2) mapAddr = (unsigned long long)(p.map + (long long)g.frame * p.mapFrameStride + (long long)(pc.r0 - 2 * kHalo) * p.mapPitch + (long long)(g.bx + lane) * p.mapStep);

        // The piece is consumed in bodies of 11 rows = one ring half each; the last body may run past the piece's rows: what
        // it then reads are stale ring rows, what it computes from them belongs to output rows that are never stored.
        const int nRows = g.nRows;                                      // input rows that complete a wanted output
        const int nBodies = (nRows + kTaps - 1) / kTaps;
        #pragma unroll 1
        for (int body = 0; body < nBodies; ++body, ++gbody) {
            const uint32_t halfOff = (gbody & 1u) * (kTaps * kRingRowBytes);
            const uint32_t col0 = col0Base + halfOff, col1 = col1Base + halfOff;
            mbar_wait_sleep(barFull + 8 * (gbody & 1u), (gbody >> 1) & 1u, p.backoffNs);
            const int iBase = body * kTaps;
            float bodySum0 = 0.f, bodySum1 = 0.f;
            #pragma unroll
            for (int t = 0; t < kTaps; ++t) {
                float sv0, sv1;
                vertical_row<kU16>(t, qab0, qsp0, qab1, qsp1, w2, col0 + t * kRingRowBytes, col1 + t * kRingRowBytes, ca, cb, eps2, sv0, sv1);
                if (t == kTaps - 1) mbar_arrive(barEmpty + 8 * (gbody & 1u));   // this lane is done reading the half (count 32)
                const int i = iBase + t;
                if (i >= 2 * kHalo && i < nRows) {                      // warp-uniform: the row completes a wanted output
                    if (kMap == 1) {
                        // two predicated stores; a separate unpredicated path for bands that lie fully inside the
                        // plane would double the store code of all 11 unrolled rows (hot code size matters: see the kernel)
                        if (fullFlag) { stg_f32<0>(mapAddr, sv0); stg_f32<128>(mapAddr, sv1); }
                        else { stg_f32_if<0>(mapAddr, sv0, okFlag0); stg_f32_if<128>(mapAddr, sv1, okFlag1); }
                    }
                    if (kMap == 2) {
                        if (colOk0) stg_f32<0>(mapAddr, sv0);
                        if (colOk1) stg_f32<0>(mapAddr + mapCol1Bytes, sv1);
                    }
                    bodySum0 += sv0; bodySum1 += sv1;
                }
                if (kMap) mapAddr += mapPitchBytes;
            }
            const float bodySum = (colOk0 ? bodySum0 : 0.f) + (colOk1 ? bodySum1 : 0.f);
            total += (double)bodySum;                                   // <= 22 values per float partial
        }
    }
}

template <int kMap, bool kU16>
__global__ void __launch_bounds__(kCtaThreads, 1)
ssim_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ FusedParams p, const __grid_constant__ ExchangeParams x)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kPairsPerCta][kBarsPerPair];  // per pair: tmaFull[2], stageEmpty[2], ringFull[4], ringEmpty[4] (+4 spare)

    // shuffled from lane 0 so that the compiler knows the warp index (and everything derived from it: slot, smem and
    // barrier addresses, TMA coordinates) is warp-uniform and keeps it on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int pair = warp & (kPairsPerCta - 1);
    const bool isConsumer = warp >= kPairsPerCta;

    if (threadIdx.x == 0) {
        for (int pr = 0; pr < kPairsPerCta; ++pr)
            for (int i = 0; i < kBarsPerPair; ++i) mbar_init(smem_u32(&bars[pr][i]), i < kStages ? 1 : 32);   // TMA barriers: 1 arrival; stage-empty, ring full/empty: all 32 lanes
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();                                                    // the only CTA-wide barrier
    // Programmatic dependent launch (launch_fused sets the stream-serialization attribute): the next launch on the stream may
    // be scheduled while this grid runs -- its CTAs move onto the SMs as ours leave and get as far as this point -- and
    // everything that touches memory waits here until the launch before this one has completed and flushed.  Takes the launch
    // latency out of calls queued back to back; without the attribute both instructions do nothing.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const uint32_t slot = blockIdx.x * kPairsPerCta + pair;
    const uint32_t pairSmem = __shfl_sync(0xffffffffu, smem_u32(smem), 0) + pair * PixGeo<kU16>::kPairSmemBytes;
    const uint32_t barBase  = __shfl_sync(0xffffffffu, smem_u32(&bars[0][0]), 0) + pair * (kBarsPerPair * 8);

    // Register hand-over between the two warpgroups: every warp of a warpgroup must execute its setmaxnreg (so it comes
    // before the early exit), and each role's code must follow its own setmaxnreg within the same branch -- ptxas budgets
    // registers per region, and any code shared by both roles would be held to the smaller budget.
    // The producer's first piece is looked up HERE, in code both roles run, not inside producer_warp: in the binary that
    // code would sit between the consumer's body and the producer's block loop, i.e. cold instructions in the middle of the
    // hot address range, which then no longer fits the 32 KB instruction cache behind the L0s (stall_no_instruction 8% of
    // the samples against 2% when the two hot loops are nearly adjacent).  The empty volatile asm pins the values here.
    PieceCursor cur0;
    PieceGeo g0 = PieceGeo(), g1 = PieceGeo();
    unsigned ra0 = 0, rb0 = 0, ra1 = 0, rb1 = 0;          // centring pixels of the two pieces, as loaded
    int have0 = 0, have1 = 0;
    if (slot < p.geo.slots) {
        cursor_init(cur0, p.geo, slot);
        Piece pc;
        have0 = cursor_next(cur0, p.geo, pc) ? 1 : 0;
        if (have0) { piece_geo(p, pc, g0); piece_centre_raw<kU16>(p, g0, ra0, rb0); }
        have1 = have0 && cursor_next(cur0, p.geo, pc) ? 1 : 0;
        if (have1) { piece_geo(p, pc, g1); piece_centre_raw<kU16>(p, g1, ra1, rb1); }
    } else {
        cur0.q = cur0.qEnd = cur0.colBase = 0; cur0.frame = cur0.band = 0;
    }
    asm volatile("" : "+r"(cur0.q), "+r"(cur0.qEnd), "+r"(cur0.colBase), "+r"(cur0.frame), "+r"(cur0.band), "+r"(have0), "+r"(have1));
    // (the centring pixels are NOT pinned and not converted here: their loads stay in flight across the role branch)
    asm volatile("" : "+r"(g0.frame), "+r"(g0.bx), "+r"(g0.oy0), "+r"(g0.nOut), "+r"(g0.inY0), "+r"(g0.nRows), "+r"(g0.nBlk));
    asm volatile("" : "+r"(g1.frame), "+r"(g1.bx), "+r"(g1.oy0), "+r"(g1.nOut), "+r"(g1.inY0), "+r"(g1.nRows), "+r"(g1.nBlk));

    if (isConsumer) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kConsumerRegs));
        if (slot >= p.geo.slots) return;
        consumer_warp<kMap, kU16>(p, x, slot, lane, pairSmem, barBase);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kProducerRegs));
        if (slot >= p.geo.slots || !have0) return;          // a slot without any output row
        producer_warp<kU16>(&tmA, &tmB, p, lane, pairSmem, barBase, cur0, g0, ra0, rb0, have1 != 0, g1, ra1, rb1);
    }
}

// ------------------------------------------------------------------------------------------------ layout helpers
// gathers one channel of an arbitrarily strided u8 image into a dense pitched plane (the canonical input of the
// fused kernel); replaces the addressing part of retrieve_tile (src/ssim.cpp:531-548) for step != 1 / negative strides.
__global__ void pack_u8_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                               long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < height; y += gridDim.y * blockDim.y)     // grid-stride: any height
        dst[(long long)y * dstPitch + x] = src[(long long)x * step + (long long)y * stride];
}

// same for 16-bit pixels: src is a byte pointer, step/stride are BYTE distances (multiples of 2)
__global__ void pack_u16_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                                long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < height; y += gridDim.y * blockDim.y)
        *(uint16_t*)(dst + (long long)y * dstPitch + 2 * x) = *(const uint16_t*)(src + (long long)x * step + (long long)y * stride);
}

// BT.601 luma of an interleaved RGB(A) image, integer arithmetic identical to the reference CLI's CPU loop
// (src/ssim-cli.cpp:158-186: (19595 R + 38470 G + 7471 B + 32768) >> 16), written as a dense pitched plane.
__global__ void pack_luma_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                                 long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < height; y += gridDim.y * blockDim.y) {
        const uint8_t* px = src + (long long)x * step + (long long)y * stride;
        const unsigned r = px[0], g = px[1], b = px[2];
        dst[(long long)y * dstPitch + x] = (uint8_t)((r * 19595u + g * 38470u + b * 7471u + 32768u) >> 16);
    }
}

// splits an interleaved C-channel image into C dense pitched planes (plane c at dst + c*planeStride) in one read of the
// bytes.  C <= 4: a thread handles 4 neighbouring pixels, i.e. C aligned 32-bit loads and one 32-bit store per plane (rows
// start 16-byte aligned on both sides); other channel counts go byte by byte.
template <int kC>
__global__ void deinterleave4_u8_kernel(uint8_t* __restrict__ dst, long long dstPitch, long long planeStride,
                                        const uint8_t* __restrict__ src, long long srcPitch, int width, int height)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < height; y += gridDim.y * blockDim.y) {
        const uint32_t* in = (const uint32_t*)(src + (long long)y * srcPitch + (long long)x4 * kC);      // 4*kC bytes = kC words
        uint32_t w[kC];
        #pragma unroll
        for (int i = 0; i < kC; ++i) w[i] = __ldg(in + i);            // reads past `width` stay inside the 16-byte padded row
        #pragma unroll
        for (int c = 0; c < kC; ++c) {
            uint32_t out = 0;
            #pragma unroll
            for (int k = 0; k < 4; ++k) {                             // pixel k, channel c = byte k*kC + c of the group
                const int byte = k * kC + c;
                out |= ((w[byte >> 2] >> (8 * (byte & 3))) & 0xffu) << (8 * k);
            }
            *(uint32_t*)(dst + c * planeStride + (long long)y * dstPitch + x4) = out;     // plane rows are padded to 16 bytes too
        }
    }
}

__global__ void deinterleave_u8_kernel(uint8_t* __restrict__ dst, long long dstPitch, long long planeStride,
                                       const uint8_t* __restrict__ src, long long srcPitch, int channels, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < height; y += gridDim.y * blockDim.y) {
        const uint8_t* px = src + (long long)y * srcPitch + (long long)x * channels;
        for (int c = 0; c < channels; ++c) dst[c * planeStride + (long long)y * dstPitch + x] = px[c];
    }
}

// scatters a dense float map into an arbitrarily strided one (ssimStep != 1, negative ssimStride; src/ssim.cpp:661-667)
__global__ void scatter_map_kernel(float* __restrict__ dst, long long dstStep, long long dstStride,
                                   const float* __restrict__ src, long long srcPitch, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < height; y += gridDim.y * blockDim.y)
        dst[(long long)x * dstStep + (long long)y * dstStride] = src[(long long)y * srcPitch + x];
}

__global__ void synth_fill_kernel(uint8_t* __restrict__ dA, long long pitchA, uint8_t* __restrict__ dB, long long pitchB,
                                  int width, int rows, int y0, uint32_t frame, uint64_t seed)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < rows; y += gridDim.y * blockDim.y) {
        uint8_t a, b;
        ssim_synth_pixel(seed, frame, (uint32_t)x, (uint32_t)(y0 + y), &a, &b);
        dA[(long long)y * pitchA + x] = a;
        dB[(long long)y * pitchB + x] = b;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
// cudaFuncSetAttribute is per DEVICE: called from every device context's initialisation (current device = that device)
template <int kMap, bool kU16>
static cudaError_t set_smem_attr_one()
{
    return cudaFuncSetAttribute(ssim_fused_kernel<kMap, kU16>, cudaFuncAttributeMaxDynamicSharedMemorySize, PixGeo<kU16>::kCtaSmemBytes);
}
static cudaError_t set_smem_attr()
{
    cudaError_t e = set_smem_attr_one<0, false>();
    if (e == cudaSuccess) e = set_smem_attr_one<1, false>();
    if (e == cudaSuccess) e = set_smem_attr_one<2, false>();
    if (e == cudaSuccess) e = set_smem_attr_one<0, true>();
    if (e == cudaSuccess) e = set_smem_attr_one<1, true>();
    return e;
}

cudaError_t launch_fused(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const FusedParams& p, const ExchangeParams* xchg)
{
    const unsigned ctas = (p.geo.slots + kPairsPerCta - 1) / kPairsPerCta;
    if (ctas == 0) return cudaErrorInvalidValue;
    ExchangeParams none;
    if (!xchg) { memset(&none, 0, sizeof(none)); xchg = &none; }
    const int mapKind = !p.map ? 0 : p.mapStep == 1 ? 1 : 2;
    if (p.u16 && mapKind == 2) return cudaErrorInvalidValue;         // 16-bit pixels: dense maps only
    typedef void (*Kernel)(const CUtensorMap, const CUtensorMap, const FusedParams, const ExchangeParams);
    const Kernel kernel = p.u16 ? (mapKind ? ssim_fused_kernel<1, true> : ssim_fused_kernel<0, true>)
                                : (mapKind == 2 ? ssim_fused_kernel<2, false> : mapKind == 1 ? ssim_fused_kernel<1, false> : ssim_fused_kernel<0, false>);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(kCtaThreads);
    cfg.dynamicSmemBytes = p.u16 ? PixGeo<true>::kCtaSmemBytes : PixGeo<false>::kCtaSmemBytes;
    cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, p, *xchg);
}

cudaError_t fused_kernel_attributes(int* regsMap, int* regsNoMap, int* pairsPerSm)
{
    cudaError_t e = set_smem_attr();
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, ssim_fused_kernel<1, false>)) != cudaSuccess) return e;
    if (regsMap) *regsMap = fa.numRegs;
    if ((e = cudaFuncGetAttributes(&fa, ssim_fused_kernel<0, false>)) != cudaSuccess) return e;
    if (regsNoMap) *regsNoMap = fa.numRegs;
    // warp pairs resident per SM (one CTA of kPairsPerCta for every variant, by construction)
    int n = 0, n16 = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ssim_fused_kernel<1, false>, kCtaThreads, PixGeo<false>::kCtaSmemBytes);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n16, ssim_fused_kernel<1, true>, kCtaThreads, PixGeo<true>::kCtaSmemBytes);
    if (n16 < n) n = n16;
    if (pairsPerSm) *pairsPerSm = n * kPairsPerCta;
    return e;
}

// rows go on gridDim.y, which is limited to 65535 blocks: the kernels stride over y, so any height works
static dim3 grid2d(int width, int height, dim3 block)
{
    const unsigned gy = (unsigned)((height + block.y - 1) / block.y);
    return dim3((width + block.x - 1) / block.x, gy < 65535u ? gy : 65535u);
}

cudaError_t launch_pack_u8(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                           long long step, long long stride, int width, int height)
{
    const dim3 block(64, 4);
    pack_u8_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstPitch, src, step, stride, width, height);
    return cudaGetLastError();
}

cudaError_t launch_pack_u16(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                            long long step, long long stride, int width, int height)
{
    const dim3 block(64, 4);
    pack_u16_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstPitch, src, step, stride, width, height);
    return cudaGetLastError();
}

cudaError_t launch_pack_luma(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                             long long step, long long stride, int width, int height)
{
    const dim3 block(64, 4);
    pack_luma_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstPitch, src, step, stride, width, height);
    return cudaGetLastError();
}

cudaError_t launch_deinterleave_u8(cudaStream_t stream, uint8_t* dst, long long dstPitch, long long planeStride, const uint8_t* src,
                                   long long srcPitch, int channels, int width, int height)
{
    const dim3 block(64, 4);
    // the 4-pixel path reads and writes whole words up to the padded end of a row: both pitches must cover ceil(width/4)*4 pixels
    const int w4 = (width + 3) / 4;
    const bool vec = channels <= 4 && (srcPitch & 3) == 0 && (dstPitch & 3) == 0 && srcPitch >= (long long)w4 * 4 * channels && dstPitch >= (long long)w4 * 4 &&
                     (((uintptr_t)src | (uintptr_t)dst | (uintptr_t)planeStride) & 3) == 0;
    if (vec && channels == 1)      deinterleave4_u8_kernel<1><<<grid2d(w4, height, block), block, 0, stream>>>(dst, dstPitch, planeStride, src, srcPitch, width, height);
    else if (vec && channels == 2) deinterleave4_u8_kernel<2><<<grid2d(w4, height, block), block, 0, stream>>>(dst, dstPitch, planeStride, src, srcPitch, width, height);
    else if (vec && channels == 3) deinterleave4_u8_kernel<3><<<grid2d(w4, height, block), block, 0, stream>>>(dst, dstPitch, planeStride, src, srcPitch, width, height);
    else if (vec && channels == 4) deinterleave4_u8_kernel<4><<<grid2d(w4, height, block), block, 0, stream>>>(dst, dstPitch, planeStride, src, srcPitch, width, height);
    else deinterleave_u8_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstPitch, planeStride, src, srcPitch, channels, width, height);
    return cudaGetLastError();
}

cudaError_t launch_scatter_map(cudaStream_t stream, float* dst, long long dstStep, long long dstStride,
                               const float* src, long long srcPitch, int width, int height)
{
    const dim3 block(64, 4);
    scatter_map_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstStep, dstStride, src, srcPitch, width, height);
    return cudaGetLastError();
}

cudaError_t launch_synth_fill(cudaStream_t stream, uint8_t* dA, long long pitchA, uint8_t* dB, long long pitchB,
                              int width, int rows, int y0, uint32_t frame, uint64_t seed)
{
    const dim3 block(64, 4);
    synth_fill_kernel<<<grid2d(width, rows, block), block, 0, stream>>>(dA, pitchA, dB, pitchB, width, rows, y0, frame, seed);
    return cudaGetLastError();
}

}  // namespace ssimk
