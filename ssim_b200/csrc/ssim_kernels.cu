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

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
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
template <bool kMap>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
ssim_fused_kernel(const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB8, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta][kStages];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (item >= p.items) return;                       // warps are independent: no CTA-wide barrier anywhere below

    // ---- decode the work item: (frame, segment, band), band fastest so that a CTA covers 4 adjacent bands
    const int band  = (int)(item % p.bands);
    const int seg   = (int)((item / p.bands) % p.segs);
    const int frame = (int)(item / ((long long)p.bands * p.segs));
    const int bx    = band * kBandW;                                 // first output column of the band
    const int oy0   = p.outY0 + seg * p.segRows;                     // first output row (plane coordinates)
    const int nOut  = min(p.segRows, p.outY0 + p.outRows - oy0);     // output rows of this item
    const int inY0  = oy0 - kHalo;                                   // first input row needed (may be negative)
    const int nBlk  = (nOut + 2 * kHalo + kBlkRows - 1) / kBlkRows;

    const uint32_t warpSmem = smem_u32(smem) + warp * kWarpSmemBytes;
    const uint32_t ringBase = warpSmem + kStages * kStageBytes;
    const uint32_t barBase  = smem_u32(&bars[warp][0]);

    if (lane == 0) {
        #pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(barBase + 8 * s, 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncwarp();

    // One TMA load = rows [inY0 + 16*ld, +16) x bytes [bx-16, bx+112) of both images = two 8-row blocks.  Rows outside
    // the plane are clamped by loading single-row boxes at clamped coordinates (replicates the nearest row,
    // src/ssim.cpp:562-582); columns outside the plane arrive as zeros and are patched after landing (src/ssim.cpp:541-554).
    const int nLoads = (nBlk + 1) >> 1;
    auto issue_load = [&](int ld) {
        const int stage = ld & (kStages - 1);
        const uint32_t bar = barBase + 8 * stage;
        const uint32_t dst = warpSmem + stage * kStageBytes;
        const int y = inY0 + ld * kLoadRows;
        if (lane == 0) {
            // TMA goes through the uniform datapath: one lane, warp-uniform operands (never issue it from divergent lanes)
            mbar_arrive_expect_tx(bar, kStageBytes);
            if (y >= 0 && y + kLoadRows <= p.srcRows) {
                tma_load_3d(dst, &tmA8, bx - kBoxLeft, y, frame, bar);
                tma_load_3d(dst + kImgStageBytes, &tmB8, bx - kBoxLeft, y, frame, bar);
            } else {
                #pragma unroll 1
                for (int r = 0; r < kLoadRows; ++r) {
                    const int yy = min(max(y + r, 0), p.srcRows - 1);
                    tma_load_3d(dst + r * kBoxW, &tmA1, bx - kBoxLeft, yy, frame, bar);
                    tma_load_3d(dst + kImgStageBytes + r * kBoxW, &tmB1, bx - kBoxLeft, yy, frame, bar);
                }
            }
        }
    };

    #pragma unroll
    for (int s = 0; s < kStages; ++s)
        if (s < nLoads) issue_load(s);

    // ---- per-item centring pixel: moments are accumulated on (a - ca), (b - cb), which keeps the fp32
    // cancellation in E[x^2] - mu^2 small even on flat regions (DESIGN.md "Numerics").  Any integer works.
    const int cx = min(bx + kBandW / 2, p.width - 1);
    const int cy = min(max(oy0, 0), p.srcRows - 1);
    const float ca = (float)__ldg(p.a + (long long)frame * p.frameStrideA + (long long)cy * p.pitchA + cx);
    const float cb = (float)__ldg(p.b + (long long)frame * p.frameStrideB + (long long)cy * p.pitchB + cx);
    const u64 negMagic = pack2(-(8388608.0f + ca), -(8388608.0f + cb));
    const uint32_t magic = p.magic;                    // 0x4B000000, passed as a parameter so that it lives in a register and
                                                       // PRMT takes the byte selector as its immediate (no per-PRMT selector MOV)
    const float k2     = -0.5f * p.eps2 * (ca - cb) * (ca - cb);                 // see the formula below

    // taps: w[m] multiplies the sample at offset m of an 11-sample window, w[m] = g[|m-5|]
    u64 w2[6];
    #pragma unroll
    for (int d = 0; d < 6; ++d) w2[d] = pack2(p.g[d], p.g[d]);
    #define TAP(m) w2[(m) < 5 ? 5 - (m) : (m) - 5]

    // ---- horizontal-pass role of this lane: row hr of the block, 16 output columns starting at 16*hq
    const int hr = lane >> 2, hq = lane & 3;
    const uint32_t hSrcOff  = hr * kBoxW + hq * 16 + (kBoxLeft - 8);              // 32-byte window holding columns 16hq-8 .. 16hq+23
    const uint32_t hSwz     = (uint32_t)(hq | ((hr & 3) << 2)) << 3;              // ring swizzle (see ring layout below)
    const uint32_t hDstBase = ringBase + hr * kRingRowBytes + hq * 128;

    // ---- vertical-pass role: columns bx+lane and bx+32+lane.  Ring layout: row r holds two planes of 64 packed
    // pairs, {mu_a', mu_b'} at +0 and {D, P} at +512; column c sits at 8*(c ^ ((c>>4) | ((r&3)<<2))).  The XOR makes
    // both the 8-byte stores of the horizontal pass (lanes = 4 rows x 4 column groups per half-warp) and the 8-byte
    // loads of the vertical pass (lanes = 16 adjacent columns per half-warp) bank-conflict free.
    const uint32_t vBase0 = (uint32_t)(lane ^ (lane >> 4)) << 3;
    const uint32_t vBase1 = (uint32_t)((32 + lane) ^ (2 + (lane >> 4))) << 3;
    const bool colOk0 = bx + lane < p.width;
    const bool colOk1 = bx + 32 + lane < p.width;

    // 11-deep shifted accumulators of the vertical pass: q[m] holds the partial sum of the output row that will
    // complete m rows from now; per input row q[m] = fma(h, w[m], q[m+1]) and q[0] is a finished output.
    u64 qab0[11], qsp0[11], qab1[11], qsp1[11];
    #pragma unroll
    for (int m = 0; m < 11; ++m) qab0[m] = qsp0[m] = qab1[m] = qsp1[m] = 0ull;

    const bool patchLeft  = (bx == 0);
    const bool patchRight = (bx + kBandW + kHalo > p.width);
    // running map pointer: row of the output completed by the current input row (starts 10 rows above the segment;
    // never dereferenced there)
    float* mapPtr = nullptr;
    if (kMap) mapPtr = p.map + (long long)frame * p.mapFrameStride + (long long)(oy0 - p.outY0 - 2 * kHalo) * p.mapPitch + bx + lane;

    double total = 0.0;

    #pragma unroll 1
    for (int blk = 0; blk < nBlk; ++blk) {
        const int ld = blk >> 1, half = blk & 1;                        // TMA load and which 8 of its 16 rows
        const uint32_t loadBase  = warpSmem + (ld & (kStages - 1)) * kStageBytes;
        const uint32_t stageBase = loadBase + half * (kBlkRows * kBoxW);

        if (half == 0) {
            mbar_wait(barBase + 8 * (ld & (kStages - 1)), (uint32_t)(ld / kStages) & 1u);
            if (patchLeft || patchRight) {                   // warp-uniform; only the outermost bands
                const uint32_t row = loadBase + (lane >> 4) * kImgStageBytes + (lane & 15) * kBoxW;   // 2 images x 16 rows
                if (patchLeft) {
                    uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(row + kBoxLeft));
                    #pragma unroll
                    for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u8 [%0], %1;" :: "r"(row + kBoxLeft - k), "r"(v) : "memory");
                }
                if (patchRight) {
                    const uint32_t last = row + kBoxLeft + (p.width - 1 - bx);
                    uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(last));
                    #pragma unroll
                    for (int k = 1; k <= kHalo; ++k) asm volatile("st.shared.u8 [%0], %1;" :: "r"(last + k), "r"(v) : "memory");
                }
                __syncwarp();
            }
        }

        // ================================================================ horizontal pass: 8 rows x 64 columns
        {
            // the window starts 8 bytes into a 16-byte chunk: 8 + 16 + 8 byte loads
            const uint2 a0 = lds64u(stageBase + hSrcOff), a2 = lds64u(stageBase + hSrcOff + 24);
            const uint4 a1 = lds128(stageBase + hSrcOff + 8);
            const uint2 b0 = lds64u(stageBase + kImgStageBytes + hSrcOff), b2 = lds64u(stageBase + kImgStageBytes + hSrcOff + 24);
            const uint4 b1 = lds128(stageBase + kImgStageBytes + hSrcOff + 8);
            const uint32_t wa[8] = {a0.x, a0.y, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y};
            const uint32_t wb[8] = {b0.x, b0.y, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y};
            // every lane has read its inputs once the warp reconverges: after the second half the stage can be refilled
            __syncwarp();
            if ((half == 1 || blk + 1 >= nBlk) && ld + kStages < nLoads) {
                if (patchLeft || patchRight) fence_proxy_async();
                issue_load(ld + kStages);
            }

            u64 hab[16], hsp[16];
            #pragma unroll
            for (int i = 0; i < 26; ++i) {                   // input column 16hq - 5 + i  = byte 3 + i of the 32-byte window
                const int byteIdx = i + 3;
                float fa, fb;
                switch (byteIdx & 3) {
                    case 0:  fa = magic_byte<0>(wa[byteIdx >> 2], magic); fb = magic_byte<0>(wb[byteIdx >> 2], magic); break;
                    case 1:  fa = magic_byte<1>(wa[byteIdx >> 2], magic); fb = magic_byte<1>(wb[byteIdx >> 2], magic); break;
                    case 2:  fa = magic_byte<2>(wa[byteIdx >> 2], magic); fb = magic_byte<2>(wb[byteIdx >> 2], magic); break;
                    default: fa = magic_byte<3>(wa[byteIdx >> 2], magic); fb = magic_byte<3>(wb[byteIdx >> 2], magic); break;
                }
                const u64 ab = add2(pack2(fa, fb), negMagic);          // (a - ca, b - cb), exact
                float a, b; unpack2(ab, a, b);
                const float d = a - b;
                const u64 sp = pack2(fmaf(d, d, k2), a * b);           // ((a'-b')^2 + k2, a'b'); k2: see the formula below
                #pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = i - j;                                // sample i is tap k of output j
                    if (k == 0)                { hab[j] = mul2(ab, TAP(0)); hsp[j] = mul2(sp, TAP(0)); }
                    else if (k > 0 && k <= 10) { hab[j] = fma2(ab, TAP(k), hab[j]); hsp[j] = fma2(sp, TAP(k), hsp[j]); }
                }
                if (i >= 10) {                                          // output j = i-10 is complete
                    const int j = i - 10;
                    const uint32_t dst = hDstBase + ((uint32_t)(j << 3) ^ hSwz);
                    sts64(dst, hab[j]);
                    sts64(dst + kRingPlaneBytes, hsp[j]);
                }
            }
        }
        __syncwarp();

        // ================================================================ vertical pass + formula: 8 rows x 2 columns per lane
        float blockSum = 0.f;
        #pragma unroll
        for (int r = 0; r < kBlkRows; ++r) {
            const uint32_t row0 = ringBase + r * kRingRowBytes + (vBase0 ^ ((r & 3) << 5));
            const uint32_t row1 = ringBase + r * kRingRowBytes + (vBase1 ^ ((r & 3) << 5));
            const u64 hab0 = lds64(row0), hsp0 = lds64(row0 + kRingPlaneBytes);
            const u64 hab1 = lds64(row1), hsp1 = lds64(row1 + kRingPlaneBytes);
            #pragma unroll
            for (int m = 0; m < 10; ++m) {
                qab0[m] = fma2(hab0, TAP(m), qab0[m + 1]);
                qsp0[m] = fma2(hsp0, TAP(m), qsp0[m + 1]);
                qab1[m] = fma2(hab1, TAP(m), qab1[m + 1]);
                qsp1[m] = fma2(hsp1, TAP(m), qsp1[m + 1]);
            }
            qab0[10] = mul2(hab0, TAP(10)); qsp0[10] = mul2(hsp0, TAP(10));
            qab1[10] = mul2(hab1, TAP(10)); qsp1[10] = mul2(hsp1, TAP(10));

            // Output row (segment-relative) completed by this input row.  The formula is evaluated unconditionally (rows
            // outside [0,nOut) only cost the pipeline fill) so that its dependent chain overlaps the next row's FMAs
            // instead of sitting behind a branch; only the store and the sum are predicated.
            const int o = blk * kBlkRows + r - 2 * kHalo;
            const bool rowOk = (o >= 0) && (o < nOut);
            float s[2];
            #pragma unroll
            for (int c = 0; c < 2; ++c) {
                float ma, mb, D, P;
                unpack2(c == 0 ? qab0[0] : qab1[0], ma, mb);
                unpack2(c == 0 ? qsp0[0] : qsp1[0], D, P);
                // The reference formula (src/ssim.cpp:590-704) rearranged so that numerator and denominator share
                // their terms:  mu_a^2 + mu_b^2 = 2 mu_a mu_b + (mu_a - mu_b)^2  and
                // sigma_a^2 + sigma_b^2 = 2 sigma_ab + var(a - b),  var(a-b) = E[(a'-b')^2] - (E[a'] - E[b'])^2.
                // Identical images then give num == den bit for bit, hence exactly 1 like the reference.
                // The reference's window sums to 1+eps (see gaussian_taps() in ssim_cuda.cu), which on its RAW moments
                // shifts every covariance by -eps*mu_a*mu_b; centred moments only see -eps*ma*mb, so the difference
                // -eps*(mu_a mu_b - ma mb) is applied explicitly (the matching -eps*(ca-cb)^2 of var(a-b) is already
                // inside D: k2 was added to every (a'-b')^2 before the blur, for free, by turning an FMUL into an FFMA).
                const float mua = ma + ca, mub = mb + cb;
                const float t   = mua * mub;
                const float n1  = fmaf(2.f, t, p.c1);
                const float dmu = mua - mub;
                const float d1  = fmaf(dmu, dmu, n1);
                const float n2  = fmaf(-p.eps2, fmaf(-ma, mb, t), fmaf(2.f, fmaf(-ma, mb, P), p.c2));
                const float dm  = ma - mb;
                const float d2  = n2 + fmaf(-dm, dm, D);
                const float num = n1 * n2, den = d1 * d2;
                // den >= c1*c2 > 0 on valid rows.  MUFU.RCP + one Newton step on the quotient: ~correctly rounded, exact
                // when num == den
                float rc; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));
                const float q = num * rc;
                s[c] = fmaf(rc, fmaf(-q, den, num), q);
            }
            const bool ok0 = rowOk && colOk0, ok1 = rowOk && colOk1;
            if (kMap) {
                if (ok0) mapPtr[0]  = s[0];
                if (ok1) mapPtr[32] = s[1];
                mapPtr += p.mapPitch;
            }
            blockSum += (ok0 ? s[0] : 0.f) + (ok1 ? s[1] : 0.f);
        }
        total += (double)blockSum;                                      // <= 16 values per float partial
        __syncwarp();                                                   // ring is rewritten by the next horizontal pass
    }
    #undef TAP

    // ---- warp-level double reduction, fixed order => deterministic
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
    if (lane == 0) p.partials[item] = total;
}

// Sums the per-item partials of each frame in a fixed order (deterministic), writes the double sum and
// float(sum / double(uint32(width*height))) -- the reference's final step, src/ssim.cpp:1091-1103.
__global__ void __launch_bounds__(256) ssim_finalize_kernel(const FinalizeParams p)
{
    __shared__ double sh[256];
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
// gathers one channel of an arbitrarily strided u8 image into a dense pitched plane (the canonical input of the
// fused kernel); replaces the addressing part of retrieve_tile (src/ssim.cpp:531-548) for step != 1 / negative strides.
__global__ void pack_u8_kernel(uint8_t* __restrict__ dst, long long dstPitch, const uint8_t* __restrict__ src,
                               long long step, long long stride, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < width && y < height) dst[(long long)y * dstPitch + x] = src[(long long)x * step + (long long)y * stride];
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
    cudaError_t e = cudaFuncSetAttribute(ssim_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtaSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ssim_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtaSmemBytes);
}

cudaError_t launch_fused(cudaStream_t stream, const CUtensorMap& tmA8, const CUtensorMap& tmA1,
                         const CUtensorMap& tmB8, const CUtensorMap& tmB1, const FusedParams& p)
{
    const long long ctas = (p.items + kWarpsPerCta - 1) / kWarpsPerCta;
    if (ctas <= 0 || ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (p.map) ssim_fused_kernel<true><<<(unsigned)ctas, kWarpsPerCta * 32, kCtaSmemBytes, stream>>>(tmA8, tmA1, tmB8, tmB1, p);
    else       ssim_fused_kernel<false><<<(unsigned)ctas, kWarpsPerCta * 32, kCtaSmemBytes, stream>>>(tmA8, tmA1, tmB8, tmB1, p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(cudaStream_t stream, const FinalizeParams& p, int frames)
{
    ssim_finalize_kernel<<<frames, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t fused_kernel_attributes(int* regsMap, int* regsNoMap, int* ctasPerSm)
{
    cudaError_t e = set_smem_attr();
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, ssim_fused_kernel<true>)) != cudaSuccess) return e;
    if (regsMap) *regsMap = fa.numRegs;
    if ((e = cudaFuncGetAttributes(&fa, ssim_fused_kernel<false>)) != cudaSuccess) return e;
    if (regsNoMap) *regsNoMap = fa.numRegs;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ssim_fused_kernel<true>, kWarpsPerCta * 32, kCtaSmemBytes);
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
