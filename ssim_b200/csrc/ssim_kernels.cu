// ssim_kernels.cu -- hand-written sm_100a kernels for the SSIM hot path.
//
// ssim_fused_kernel replaces, in ONE launch, the reference's per-tile pipeline
//   retrieve_tile (src/ssim.cpp:515-583)  ->  TMA box loads of u8 rows (+5 px halo) into shared memory,
//                                             clamp-to-edge by coordinate clamping (rows) / patching (columns),
//                                             u8 -> f32 widening with PRMT + one packed FADD (centred on a per-item pixel)
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
    asm volatile("st.shared.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}

template <int kOff>
__device__ __forceinline__ void stg_f32(unsigned long long addr, float v) {
    asm volatile("st.global.f32 [%0+%1], %2;" :: "l"(addr), "n"(kOff), "f"(v) : "memory");
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

// u8 -> f32 without the conversion pipe: PRMT builds the float 2^23 + byte, the packed FADD that follows removes
// 2^23 + centre (exact).  I2F.U8 runs at 1/8 rate on B200 (tools/microbench), PRMT is an ALU-pipe op.
template <int kByte>
__device__ __forceinline__ float magic_byte(uint32_t word, uint32_t magic) { return __uint_as_float(__byte_perm(word, magic, 0x7440 + kByte)); }

// ------------------------------------------------------------------------------------------------ fused kernel
// Work item = (frame, row segment, 64-column band).  Each item is served by a PAIR of warps of the same CTA:
//
//   producer warp (warps 0-3)   TMA loads of 8-row pixel boxes into a 2-stage ring, clamp patching, u8/u16 -> f32, products,
//                               horizontal 11-tap pass; writes 8-row blocks of {E_h[a'], E_h[b']}, {E_h[(a'-b')^2], E_h[a'b']}
//                               into a 22-row shared-memory ring = two halves of 11 rows (full/empty mbarrier per half)
//   consumer warp (warps 4-7)   vertical 11-tap pass with eleven IN-PLACE accumulators per plane pair: its loop body is
//                               exactly half the ring = 11 rows, fully unrolled, so accumulator slot s always owns the
//                               output rows == s (mod 11), every tap index is static and nothing has to be shifted or
//                               renamed across iterations; then the SSIM formula, map store and partial sums
//
// The two roles overlap in time (the consumer's dependent formula chain hides behind the producer's FMAs and vice versa),
// setmaxnreg moves registers from the producers (96) to the consumers (160), and nothing is ever synchronised CTA-wide
// after the prologue.  Every hand-over (TMA stage full/empty, ring half full/empty) is an mbarrier on which each lane
// releases its own accesses and each lane acquires for itself: see the protocol table in DESIGN.md section 4.
struct ItemCoords {
    int frame, bx, oy0, nOut, inY0;
    int nBodies;    // 11-row bodies the consumer runs: ceil((nOut + 10) / 11)
    int nBlk;       // 8-row blocks the producer makes: ceil(11 * nBodies / 8) (rows past the segment are clamp-loaded filler)
};

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

// (frame, segment, band) of a work item, band fastest so that a CTA covers 4 adjacent bands; plus the per-item centring
// pixels: moments are accumulated on (a - ca), (b - cb), which keeps the fp32 cancellation in E[x^2] - mu^2 small even on
// flat regions (DESIGN.md "Numerics").  Any integer works; both warps of a pair must of course use the same one.
__device__ __forceinline__ uint32_t div_u32(uint32_t n, uint32_t mul, uint32_t shift) { return mul ? __umulhi(n, mul) >> shift : n; }

template <bool kU16>
__device__ __forceinline__ void decode_item(const FusedParams& p, uint32_t item, ItemCoords& it, float& ca, float& cb)
{
    const uint32_t q1 = div_u32(item, p.bandsMul, p.bandsShift);       // item / bands          (fast_div() constants)
    const uint32_t q2 = div_u32(q1, p.segsMul, p.segsShift);           // item / (bands * segs)
    const int band = (int)(item - q1 * (uint32_t)p.bands);
    const int seg  = (int)(q1 - q2 * (uint32_t)p.segs);
    it.frame = (int)q2;
    it.bx    = band * kBandW;                                           // first output column of the band
    it.oy0   = p.outY0 + seg * p.segRows;                               // first output row (plane coordinates)
    it.nOut  = min(p.segRows, p.outY0 + p.outRows - it.oy0);            // output rows of this item
    it.inY0  = it.oy0 - kHalo;                                          // first input row needed (may be negative)
    it.nBodies = (it.nOut + 2 * kHalo + kTaps - 1) / kTaps;
    it.nBlk    = (it.nBodies * kTaps + kBlkRows - 1) / kBlkRows;
    const int cx = min(it.bx + kBandW / 2, p.width - 1);
    const int cy = min(max(it.oy0, 0), p.srcRows - 1);
    const uint8_t* pa = p.a + (long long)it.frame * p.frameStrideA + (long long)cy * p.pitchA;
    const uint8_t* pb = p.b + (long long)it.frame * p.frameStrideB + (long long)cy * p.pitchB;
    if (kU16) { ca = (float)__ldg((const uint16_t*)pa + cx); cb = (float)__ldg((const uint16_t*)pb + cx); }
    else      { ca = (float)__ldg(pa + cx);                  cb = (float)__ldg(pb + cx); }
}

// ---- producer: TMA + horizontal pass
template <bool kU16>
__device__ __forceinline__ void producer_warp(const CUtensorMap* tmA, const CUtensorMap* tmB, const FusedParams& p, const ItemCoords& it, int lane,
                                              uint32_t pairSmem, uint32_t barTma, uint32_t barFull, uint32_t barEmpty, float ca, float cb)
{
    typedef PixGeo<kU16> G;
    constexpr int kBoxW = G::kBoxBytes, kImgStageBytes = G::kImgStageBytes, kStageBytes = G::kStageBytes;   // shadow the 8-bit constants
    const uint32_t ringBase = pairSmem + kStages * kStageBytes;
    const uint32_t barStageEmpty = barTma + 8 * kStages;      // per stage: "all 32 lanes have read it" (count 32)

    // One TMA load = an 8-row box x bytes [bx-16, bx+112) of both images (16-bit: elements [bx-8, bx+72)).  A block needs rows [y, y+8) with y = inY0 + 8*blk,
    // rows outside the plane replicating the nearest one (src/ssim.cpp:562-582): the distinct rows it needs always fit in
    // the 8-row box starting at clamp(y, 0, srcRows-8), so edge blocks load that box and each lane reads the row
    // clamp(y + hr) of it.  Columns outside the plane arrive as zeros and are patched after landing (src/ssim.cpp:541-554).
    const int lastBoxY = max(p.srcRows - kLoadRows, 0);
    auto issue_load = [&](int blk) {
        if (lane == 0) {
            // TMA goes through the uniform datapath: one lane, warp-uniform operands (never issue it from divergent lanes)
            const int stage = blk % kStages;
            const uint32_t bar = barTma + 8 * stage;
            const uint32_t dst = pairSmem + stage * kStageBytes;
            const int y0 = min(max(it.inY0 + blk * kLoadRows, 0), lastBoxY);
            mbar_arrive_expect_tx(bar, kStageBytes);
            tma_load_3d(dst, tmA, it.bx - G::kBoxLeftElems, y0, it.frame, bar);
            tma_load_3d(dst + kImgStageBytes, tmB, it.bx - G::kBoxLeftElems, y0, it.frame, bar);
        }
    };
    #pragma unroll
    for (int s = 0; s < kStages; ++s)
        if (s < it.nBlk) issue_load(s);

    // (a - ca, b - cb) from the bytes: PRMT builds 2^23 + byte, one packed FADD removes 2^23 + centre (exact)
    const u64 negMagic = pack2(-(8388608.0f + ca), -(8388608.0f + cb));
    const uint32_t magic = p.magic;                    // 0x4B000000, passed as a parameter so that it lives in a register and
                                                       // PRMT takes the byte selector as its immediate (no per-PRMT selector MOV)
    const float k2 = -0.5f * p.eps2 * (ca - cb) * (ca - cb);  // see the formula in consumer_warp()

    u64 w2[6];
    #pragma unroll
    for (int d = 0; d < 6; ++d) w2[d] = pack2(p.g[d], p.g[d]);
    #define TAP(m) w2[(m) < 5 ? 5 - (m) : (m) - 5]

    // this lane's share of a block: row hr, 16 output columns starting at 16*hq
    const int hr = lane >> 2, hq = lane & 3;
    // 8-bit: 32-byte window holding columns 16hq-8 .. 16hq+23; 16-bit: 64-byte window holding columns 16hq-8 .. 16hq+23
    const uint32_t hColOff  = kU16 ? hq * 32 : hq * 16 + (kBoxLeft - 8);
    const uint32_t hSrcOff  = hr * kBoxW + hColOff;

    const bool patchLeft  = (it.bx == 0);
    const bool patchRight = (it.bx + kBandW + kHalo > p.width);

    int acquired = 0;      // ring halves (global count) this warp may write
    int released = 0;      // ring halves (global count) handed to the consumer

    #pragma unroll 1
    for (int blk = 0; blk < it.nBlk; ++blk) {
        const int stage = blk % kStages;
        const uint32_t stageBase = pairSmem + stage * kStageBytes;
        mbar_wait(barTma + 8 * stage, (uint32_t)(blk / kStages) & 1u);

        if (patchLeft || patchRight) {                       // warp-uniform; only the outermost bands
            if (lane < 2 * kLoadRows) {
                const uint32_t row = stageBase + (lane >> 3) * kImgStageBytes + (lane & 7) * kBoxW;   // 2 images x 8 rows
                if (kU16) {
                    if (patchLeft) {
                        uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(row + kBoxLeft));
                        #pragma unroll
                        for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u16 [%0], %1;" :: "r"(row + kBoxLeft - 2 * k), "r"(v) : "memory");
                    }
                    if (patchRight) {
                        const uint32_t last = row + kBoxLeft + 2 * (p.width - 1 - it.bx);
                        uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(last));
                        #pragma unroll
                        for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u16 [%0], %1;" :: "r"(last + 2 * k), "r"(v) : "memory");
                    }
                } else {
                    if (patchLeft) {
                        uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row + kBoxLeft));
                        #pragma unroll
                        for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u8 [%0], %1;" :: "r"(row + kBoxLeft - k), "r"(v) : "memory");
                    }
                    if (patchRight) {
                        const uint32_t last = row + kBoxLeft + (p.width - 1 - it.bx);
                        uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(last));
                        #pragma unroll
                        for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u8 [%0], %1;" :: "r"(last + k), "r"(v) : "memory");
                    }
                }
            }
            __syncwarp();
        }

        // this lane's row of the box: hr, except in edge blocks (warp-uniform test) where rows replicate
        uint32_t src = stageBase + hSrcOff;
        {
            const int y = it.inY0 + blk * kLoadRows;
            if (y < 0 || y > lastBoxY) {
                const int y0 = min(max(y, 0), lastBoxY);
                src = stageBase + (uint32_t)(min(max(y + hr, 0), p.srcRows - 1) - y0) * kBoxW + hColOff;
            }
        }
        uint32_t wa[kU16 ? 16 : 8], wb[kU16 ? 16 : 8];
        if (kU16) {
            #pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 va = lds128(src + 16 * q), vb = lds128(src + kImgStageBytes + 16 * q);
                wa[4 * q] = va.x; wa[4 * q + 1] = va.y; wa[4 * q + 2] = va.z; wa[4 * q + 3] = va.w;
                wb[4 * q] = vb.x; wb[4 * q + 1] = vb.y; wb[4 * q + 2] = vb.z; wb[4 * q + 3] = vb.w;
            }
        } else {
            // the window starts 8 bytes into a 16-byte chunk: 8 + 16 + 8 byte loads
            const uint2 a0 = lds64u(src), a2 = lds64u(src + 24);
            const uint4 a1 = lds128(src + 8);
            const uint2 b0 = lds64u(src + kImgStageBytes), b2 = lds64u(src + kImgStageBytes + 24);
            const uint4 b1 = lds128(src + kImgStageBytes + 8);
            wa[0] = a0.x; wa[1] = a0.y; wa[2] = a1.x; wa[3] = a1.y; wa[4] = a1.z; wa[5] = a1.w; wa[6] = a2.x; wa[7] = a2.y;
            wb[0] = b0.x; wb[1] = b0.y; wb[2] = b1.x; wb[3] = b1.y; wb[4] = b1.z; wb[5] = b1.w; wb[6] = b2.x; wb[7] = b2.y;
        }
        // The stage may be refilled once EVERY lane's loads have been performed: each lane releases the stage through an
        // mbarrier (count 32) and lane 0 acquires it before re-arming the TMA barrier.  Program order plus __syncwarp() is
        // not enough here -- with the refill issued straight after the loads, tools/dev/stress.py saw rare 8-row x 16-column
        // blocks computed from the NEXT box's bytes (1080p, first block of an item).
        // (the acquire + refill sit at ii == 10 below: by then the 32 arrivals have long drained and lane 0 never spins)
        mbar_arrive(barStageEmpty + 8 * stage);

        // Ring position of this lane's row: input row i = 8*blk + hr lives in half (i / 11) & 1 at row t = i % 11.
        // Layout of a ring row (1056 bytes): two planes of 64 packed pairs, {E_h[a'], E_h[b']} at +0 and
        // {E_h[(a'-b')^2], E_h[a'b']} at +512, then 32 bytes of padding; column c sits at 8*(c ^ (c>>4)).
        // Bank-conflict freedom without any per-access arithmetic: in the producer's 8-byte stores a half-warp is 4 rows x
        // 4 column groups writing the same j -- the XOR by the group index spreads the groups over 4 adjacent slots and the
        // 32-byte pad moves each following row by 4 slots (16 distinct slots); in the consumer's 8-byte loads a half-warp
        // reads 16 adjacent columns of one group (XOR by a constant).  Both sides address "register + immediate": the
        // store to column j goes to dst[j & 3] + 32*(j >> 2), the load of row t comes from base + 1056*t.
        const int i  = blk * kBlkRows + hr;
        const int ih = i / kTaps, t = i - ih * kTaps;
        const uint32_t dstRow = ringBase + (uint32_t)((ih & 1) * kTaps + t) * kRingRowBytes + hq * 128;
        uint32_t dst4[4];
        #pragma unroll
        for (int m = 0; m < 4; ++m) dst4[m] = dstRow + ((uint32_t)(m ^ hq) << 3);

        u64 hab[16], hsp[16];
        #pragma unroll
        for (int ii = 0; ii < 26; ++ii) {                // input column 16hq - 5 + ii  = byte 3 + ii of the 32-byte window
            const int byteIdx = ii + 3;                     // 16-bit: halfword index
            float fa, fb;
            if (kU16) {
                // 2^23 + pixel: the two pixel bytes under the two top bytes of the magic word (PRMT with an immediate selector)
                if (byteIdx & 1) { fa = __uint_as_float(__byte_perm(wa[byteIdx >> 1], magic, 0x7632)); fb = __uint_as_float(__byte_perm(wb[byteIdx >> 1], magic, 0x7632)); }
                else             { fa = __uint_as_float(__byte_perm(wa[byteIdx >> 1], magic, 0x7610)); fb = __uint_as_float(__byte_perm(wb[byteIdx >> 1], magic, 0x7610)); }
            } else {
                switch (byteIdx & 3) {
                    case 0:  fa = magic_byte<0>(wa[byteIdx >> 2], magic); fb = magic_byte<0>(wb[byteIdx >> 2], magic); break;
                    case 1:  fa = magic_byte<1>(wa[byteIdx >> 2], magic); fb = magic_byte<1>(wb[byteIdx >> 2], magic); break;
                    case 2:  fa = magic_byte<2>(wa[byteIdx >> 2], magic); fb = magic_byte<2>(wb[byteIdx >> 2], magic); break;
                    default: fa = magic_byte<3>(wa[byteIdx >> 2], magic); fb = magic_byte<3>(wb[byteIdx >> 2], magic); break;
                }
            }
            const u64 ab = add2(pack2(fa, fb), negMagic);          // (a - ca, b - cb), exact
            float a, b; unpack2(ab, a, b);
            const float d = a - b;
            const u64 sp = pack2(fmaf(d, d, k2), a * b);           // ((a'-b')^2 + k2, a'b')
            #pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = ii - j;                               // sample ii is tap k of output j
                if (k == 0)                { hab[j] = mul2(ab, TAP(0)); hsp[j] = mul2(sp, TAP(0)); }
                else if (k > 0 && k <= 10) { hab[j] = fma2(ab, TAP(k), hab[j]); hsp[j] = fma2(sp, TAP(k), hsp[j]); }
            }
            if (ii == 10) {
                if (blk + kStages < it.nBlk) {                      // refill this block's stage (see above)
                    if (lane == 0) while (!mbar_test(barStageEmpty + 8 * stage, (uint32_t)(blk / kStages) & 1u)) { }
                    if (patchLeft || patchRight) fence_proxy_async();
                    issue_load(blk + kStages);
                }
                // first store of the block: the ring halves this block touches must have been drained by the consumer
                // (waiting here, not at the top, lets the loads and the first 10 columns of math overlap the wait)
                const int lastHalf = (blk * kBlkRows + kBlkRows - 1) / kTaps;
                while (acquired <= lastHalf) {
                    mbar_wait_sleep(barEmpty + 8 * (acquired & 1), ((uint32_t)(acquired >> 1) & 1u) ^ 1u, p.backoffNs);
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
        // every lane arrives (barrier count 32): each lane's release covers its own stores, no reliance on warp-level cumulativity
        const int complete = (blk * kBlkRows + kBlkRows) / kTaps;   // ring halves fully written so far
        for (int hdone = released; hdone < complete; ++hdone) mbar_arrive(barFull + 8 * (hdone & 1));
        released = max(released, complete);
    }
    #undef TAP
}

// ---- consumer: vertical pass + formula + outputs
template <bool kMap, bool kU16>
__device__ __forceinline__ void consumer_warp(const FusedParams& p, const ItemCoords& it, int lane, uint32_t item, uint32_t pairSmem,
                                              uint32_t barFull, uint32_t barEmpty, float ca, float cb)
{
    const uint32_t ringBase = pairSmem + kStages * PixGeo<kU16>::kStageBytes;
    u64 w2[6];
    #pragma unroll
    for (int d = 0; d < 6; ++d) w2[d] = pack2(p.g[d], p.g[d]);
    #define TAP(m) w2[(m) < 5 ? 5 - (m) : (m) - 5]
    // (0.01*L)^2, (0.03*L)^2 as float, L = 255 (src/ssim.cpp:956-960) or 65535 (the 16-bit extension the reference's README names)
    constexpr float c1 = kU16 ? 429483.6225f : 6.5025f, c2 = kU16 ? 3865352.6025f : 58.5225f;

    // this lane owns columns bx+lane and bx+32+lane (ring layout: see producer_warp); column c sits at 8*(c ^ (c>>4))
    const uint32_t vBase0 = ringBase + ((uint32_t)(lane ^ (lane >> 4)) << 3);
    const uint32_t vBase1 = ringBase + ((uint32_t)((32 + lane) ^ (2 + (lane >> 4))) << 3);
    const bool colOk0 = it.bx + lane < p.width;
    const bool colOk1 = it.bx + 32 + lane < p.width;

    // Eleven in-place accumulators per plane pair and column: slot s accumulates the output row whose first input row
    // is == s (mod 11).  At row t of a body slot s receives tap (t - s) mod 11; the slot receiving tap 0 is re-initialised,
    // the slot receiving tap 10 is complete.  All indices are compile-time constants.
    u64 qab0[kTaps], qsp0[kTaps], qab1[kTaps], qsp1[kTaps];
    #pragma unroll
    for (int m = 0; m < kTaps; ++m) qab0[m] = qsp0[m] = qab1[m] = qsp1[m] = 0ull;

    // Map addressing: one 64-bit per-lane address that advances by the pitch per input row (starts 10 rows above the
    // segment, never dereferenced there); the second column is an immediate offset.  Stores are written in PTX so
    // that the address arithmetic stays these two adds per row.
    unsigned long long mapAddr = 0;
    const unsigned long long mapPitchBytes = (unsigned long long)p.mapPitch * sizeof(float);
    if (kMap) mapAddr = (unsigned long long)(p.map + (long long)it.frame * p.mapFrameStride + (long long)(it.oy0 - p.outY0 - 2 * kHalo) * p.mapPitch + it.bx + lane);
    const bool fullBand = it.bx + kBandW <= p.width;                    // warp-uniform: no column predicates needed

    const int nRows = it.nOut + 2 * kHalo;                              // input rows that complete a wanted output
    double total = 0.0;

    #pragma unroll 1
    for (int body = 0; body < it.nBodies; ++body) {
        const uint32_t halfOff = (uint32_t)(body & 1) * (kTaps * kRingRowBytes);
        const uint32_t col0 = vBase0 + halfOff, col1 = vBase1 + halfOff;
        mbar_wait_sleep(barFull + 8 * (body & 1), (uint32_t)(body >> 1) & 1u, p.backoffNs);
        const int iBase = body * kTaps;
        float bodySum0 = 0.f, bodySum1 = 0.f;
        #pragma unroll
        for (int t = 0; t < kTaps; ++t) {
            const u64 hab0 = lds64(col0 + t * kRingRowBytes), hsp0 = lds64(col0 + t * kRingRowBytes + kRingPlaneBytes);
            const u64 hab1 = lds64(col1 + t * kRingRowBytes), hsp1 = lds64(col1 + t * kRingRowBytes + kRingPlaneBytes);
            #pragma unroll
            for (int s = 0; s < kTaps; ++s) {
                const int k = (t - s + kTaps) % kTaps;
                if (k == 0) {
                    qab0[s] = mul2(hab0, TAP(0)); qsp0[s] = mul2(hsp0, TAP(0));
                    qab1[s] = mul2(hab1, TAP(0)); qsp1[s] = mul2(hsp1, TAP(0));
                } else {
                    qab0[s] = fma2(hab0, TAP(k), qab0[s]); qsp0[s] = fma2(hsp0, TAP(k), qsp0[s]);
                    qab1[s] = fma2(hab1, TAP(k), qab1[s]); qsp1[s] = fma2(hsp1, TAP(k), qsp1[s]);
                }
            }
            if (t == kTaps - 1) mbar_arrive(barEmpty + 8 * (body & 1));   // this lane is done reading the half (count 32)

            // Output row completed by this input row.  The formula is evaluated unconditionally (the first 10 rows of a
            // segment and the filler rows at its end only cost the pipeline fill); the store and the sum are predicated.
            const int done = (t + 1) % kTaps;
            const int i = iBase + t;
            const bool rowOk = (i >= 2 * kHalo) && (i < nRows);
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
                const float n2  = fmaf(-p.eps2, fmaf(-ma, mb, tt), fmaf(2.f, fmaf(-ma, mb, P), c2));
                const float dm  = ma - mb;
                const float d2  = n2 + fmaf(-dm, dm, D);
                const float num = n1 * n2, den = d1 * d2;
                // den >= c1*c2 > 0 on valid rows.  MUFU.RCP + one Newton step on the quotient: ~correctly rounded,
                // exact when num == den
                float rc; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));
                const float q = num * rc;
                sv[c] = fmaf(rc, fmaf(-q, den, num), q);
            }
            if (rowOk) {                                                // warp-uniform
                if (kMap) {
                    if (fullBand) { stg_f32<0>(mapAddr, sv[0]); stg_f32<128>(mapAddr, sv[1]); }
                    else {
                        if (colOk0) stg_f32<0>(mapAddr, sv[0]);
                        if (colOk1) stg_f32<128>(mapAddr, sv[1]);
                    }
                }
                bodySum0 += sv[0]; bodySum1 += sv[1];
            }
            if (kMap) mapAddr += mapPitchBytes;
        }
        const float bodySum = (colOk0 ? bodySum0 : 0.f) + (colOk1 ? bodySum1 : 0.f);
        total += (double)bodySum;                                       // <= 22 values per float partial
    }
    #undef TAP

    // ---- warp-level double reduction, fixed order => deterministic
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
    if (lane == 0) p.partials[item] = total;
}

template <bool kMap, bool kU16>
__global__ void __launch_bounds__(kCtaThreads, 2)
ssim_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kPairsPerCta][8];            // per pair: tmaFull[kStages], stageEmpty[kStages], ringFull[2], ringEmpty[2]

    // shuffled from lane 0 so that the compiler knows the warp index (and everything derived from it: item, smem and
    // barrier addresses, TMA coordinates) is warp-uniform and keeps it on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int pair = warp & (kPairsPerCta - 1);
    const bool isConsumer = warp >= kPairsPerCta;

    if (threadIdx.x == 0) {
        for (int pr = 0; pr < kPairsPerCta; ++pr)
            for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[pr][i]), i < kStages ? 1 : 32);   // TMA barriers: 1 arrival; stage-empty, ring full/empty: all 32 lanes
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();                                                    // the only CTA-wide barrier
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");       // the reduction grid may be set up from now on (it waits for our completion)

    const uint32_t item = blockIdx.x * kPairsPerCta + pair;           // < 2^31, checked by the host
    const uint32_t pairSmem = __shfl_sync(0xffffffffu, smem_u32(smem), 0) + pair * PixGeo<kU16>::kPairSmemBytes;
    const uint32_t barBase  = __shfl_sync(0xffffffffu, smem_u32(&bars[0][0]), 0) + pair * 64;

    // Register hand-over between the two warpgroups: every warp of a warpgroup must execute its setmaxnreg (so it comes
    // before the early exit), and each role's code must follow its own setmaxnreg within the same branch -- ptxas budgets
    // registers per region, and any code shared by both roles would be held to the smaller budget.
    if (isConsumer) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kConsumerRegs));
        if (item >= p.items) return;
        ItemCoords it; float ca, cb;
        decode_item<kU16>(p, item, it, ca, cb);
        consumer_warp<kMap, kU16>(p, it, lane, item, pairSmem, barBase + 32, barBase + 48, ca, cb);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kProducerRegs));
        if (item >= p.items) return;
        ItemCoords it; float ca, cb;
        decode_item<kU16>(p, item, it, ca, cb);
        producer_warp<kU16>(&tmA, &tmB, p, it, lane, pairSmem, barBase, barBase + 32, barBase + 48, ca, cb);
    }
}

// Sums the per-item partials of each frame in a fixed order (deterministic), writes the double sum and
// float(sum / double(uint32(width*height))) -- the reference's final step, src/ssim.cpp:1091-1103.
__global__ void __launch_bounds__(256) ssim_finalize_kernel(const FinalizeParams p)
{
    __shared__ double sh[256];
    // launched with programmatic stream serialization: this grid may start while the fused kernel is still draining; it
    // must not read the partial sums before that kernel has completed and flushed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int frame = blockIdx.x;
    const double* src = p.partials + (long long)frame * p.itemsPerFrame;
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.itemsPerFrame; i += 256) acc += src[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    #pragma unroll
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (p.sums) p.sums[frame] = sh[0];
        if (p.ssim) p.ssim[frame] = (float)(sh[0] * p.invCount);
    }
}

// ------------------------------------------------------------------------------------------------ layout helpers
// Reduction of ONE strip + all-reduce of the strip sums over peer memory, in one kernel (see ExchangeParams).
__global__ void __launch_bounds__(256) ssim_finalize_allreduce_kernel(const FinalizeParams p, const ExchangeParams x)
{
    __shared__ double sh[256];
    __shared__ double vals[kMaxRanks];
    __shared__ int failed;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // see ssim_finalize_kernel
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.itemsPerFrame; i += 256) acc += p.partials[i];
    sh[threadIdx.x] = acc;
    if (threadIdx.x == 0) failed = 0;
    __syncthreads();
    #pragma unroll
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    const unsigned half = (unsigned)(x.epoch & 1ull) * kMaxRanks;
    if ((int)threadIdx.x < x.world) {
        // one thread per peer: value, then the epoch with release semantics at system scope
        ExchangeSlot* dst = x.peers[threadIdx.x] + half + x.rank;
        dst->value = sh[0];
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(&dst->epoch), "l"(x.epoch) : "memory");
    }
    if (threadIdx.x == 0) {
        if (p.sums) p.sums[0] = sh[0];
        if (p.ssim) p.ssim[0] = (float)(sh[0] * p.invCount);
    }
    if ((int)threadIdx.x < x.world) {
        const ExchangeSlot* src = x.peers[x.rank] + half + threadIdx.x;
        unsigned long long t0, now, seen;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(&src->epoch) : "memory");
            if (seen == x.epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > x.timeoutNs) { failed = 1; break; }
        }
        vals[threadIdx.x] = src->value;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int r = 0; r < x.world; ++r) total += vals[r];
        if (failed) total = __longlong_as_double(0x7ff8000000000000ll);
        *x.sumAll = total;
        if (x.ssimAll) *x.ssimAll = (float)(total * x.invCountAll);
        if (x.status) *x.status = failed;
    }
}

// gathers one channel of an arbitrarily strided u8 image into a dense pitched plane (the canonical input of the
// fused kernel); replaces the addressing part of retrieve_tile (src/ssim.cpp:531-548) for step != 1 / negative strides.
__global__ void pack_u8_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                               long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height) dst[(long long)y * dstPitch + x] = src[(long long)x * step + (long long)y * stride];
}

// same for 16-bit pixels: src is a byte pointer, step/stride are BYTE distances (multiples of 2)
__global__ void pack_u16_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                                long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height)
        *(uint16_t*)(dst + (long long)y * dstPitch + 2 * x) = *(const uint16_t*)(src + (long long)x * step + (long long)y * stride);
}

// BT.601 luma of an interleaved RGB(A) image, integer arithmetic identical to the reference CLI's CPU loop
// (src/ssim-cli.cpp:158-186: (19595 R + 38470 G + 7471 B + 32768) >> 16), written as a dense pitched plane.
__global__ void pack_luma_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                                 long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height) {
        const uint8_t* px = src + (long long)x * step + (long long)y * stride;
        const unsigned r = px[0], g = px[1], b = px[2];
        dst[(long long)y * dstPitch + x] = (uint8_t)((r * 19595u + g * 38470u + b * 7471u + 32768u) >> 16);
    }
}

// splits an interleaved C-channel image into C dense pitched planes (plane c at dst + c*planeStride) in one read of the bytes
__global__ void deinterleave_u8_kernel(uint8_t* __restrict__ dst, long long dstPitch, long long planeStride,
                                       const uint8_t* __restrict__ src, long long srcPitch, int channels, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height) {
        const uint8_t* px = src + (long long)y * srcPitch + (long long)x * channels;
        for (int c = 0; c < channels; ++c) dst[c * planeStride + (long long)y * dstPitch + x] = px[c];
    }
}

// merges C dense float maps (map c at src + c*planeStride) into one interleaved map: dst[y*dstPitch + x*C + c]
__global__ void interleave_map_kernel(float* __restrict__ dst, long long dstPitch, const float* __restrict__ src, long long srcPitch,
                                      long long planeStride, int channels, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height)
        for (int c = 0; c < channels; ++c) dst[(long long)y * dstPitch + (long long)x * channels + c] = src[c * planeStride + (long long)y * srcPitch + x];
}

// scatters a dense float map into an arbitrarily strided one (ssimStep != 1, negative ssimStride; src/ssim.cpp:661-667)
__global__ void scatter_map_kernel(float* __restrict__ dst, long long dstStep, long long dstStride,
                                   const float* __restrict__ src, long long srcPitch, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height) dst[(long long)x * dstStep + (long long)y * dstStride] = src[(long long)y * srcPitch + x];
}

__global__ void synth_fill_kernel(uint8_t* __restrict__ dA, long long pitchA, uint8_t* __restrict__ dB, long long pitchB,
                                  int width, int rows, int y0, uint32_t frame, uint64_t seed)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < rows) {
        uint8_t a, b;
        ssim_synth_pixel(seed, frame, (uint32_t)x, (uint32_t)(y0 + y), &a, &b);
        dA[(long long)y * pitchA + x] = a;
        dB[(long long)y * pitchB + x] = b;
    }
}

// ------------------------------------------------------------------------------------------------ launchers
// cudaFuncSetAttribute is per DEVICE: called from every device context's initialisation (current device = that device)
static cudaError_t set_smem_attr()
{
    cudaError_t e = cudaFuncSetAttribute(ssim_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PixGeo<false>::kCtaSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ssim_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PixGeo<false>::kCtaSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ssim_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PixGeo<true>::kCtaSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ssim_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PixGeo<true>::kCtaSmemBytes);
    return e;
}

cudaError_t launch_fused(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const FusedParams& p)
{
    const long long ctas = (p.items + kPairsPerCta - 1) / kPairsPerCta;
    if (ctas <= 0 || ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (p.u16) {
        if (p.map) ssim_fused_kernel<true, true><<<(unsigned)ctas, kCtaThreads, PixGeo<true>::kCtaSmemBytes, stream>>>(tmA, tmB, p);
        else       ssim_fused_kernel<false, true><<<(unsigned)ctas, kCtaThreads, PixGeo<true>::kCtaSmemBytes, stream>>>(tmA, tmB, p);
    } else {
        if (p.map) ssim_fused_kernel<true, false><<<(unsigned)ctas, kCtaThreads, PixGeo<false>::kCtaSmemBytes, stream>>>(tmA, tmB, p);
        else       ssim_fused_kernel<false, false><<<(unsigned)ctas, kCtaThreads, PixGeo<false>::kCtaSmemBytes, stream>>>(tmA, tmB, p);
    }
    return cudaGetLastError();
}

cudaError_t launch_finalize(cudaStream_t stream, const FinalizeParams& p, int frames)
{
    // programmatic dependent launch: the reduction grid is set up while the fused kernel's last CTAs are still running
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(frames); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ssim_finalize_kernel, p);
}

cudaError_t launch_finalize_allreduce(cudaStream_t stream, const FinalizeParams& p, const ExchangeParams& x)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ssim_finalize_allreduce_kernel, p, x);
}

cudaError_t fused_kernel_attributes(int* regsMap, int* regsNoMap, int* ctasPerSm)
{
    cudaError_t e = set_smem_attr();
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, ssim_fused_kernel<true, false>)) != cudaSuccess) return e;
    if (regsMap) *regsMap = fa.numRegs;
    if ((e = cudaFuncGetAttributes(&fa, ssim_fused_kernel<false, false>)) != cudaSuccess) return e;
    if (regsNoMap) *regsNoMap = fa.numRegs;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ssim_fused_kernel<true, false>, kCtaThreads, PixGeo<false>::kCtaSmemBytes);
    if (e == cudaSuccess) {
        int n16 = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n16, ssim_fused_kernel<true, true>, kCtaThreads, PixGeo<true>::kCtaSmemBytes);
        if (n16 < n) n = n16;
    }
    if (ctasPerSm) *ctasPerSm = n;
    return e;
}

static dim3 grid2d(int width, int height, dim3 block) { return dim3((width + block.x - 1) / block.x, (height + block.y - 1) / block.y); }

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
    deinterleave_u8_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstPitch, planeStride, src, srcPitch, channels, width, height);
    return cudaGetLastError();
}

cudaError_t launch_interleave_map(cudaStream_t stream, float* dst, long long dstPitch, const float* src, long long srcPitch,
                                  long long planeStride, int channels, int width, int height)
{
    const dim3 block(64, 4);
    interleave_map_kernel<<<grid2d(width, height, block), block, 0, stream>>>(dst, dstPitch, src, srcPitch, planeStride, channels, width, height);
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
