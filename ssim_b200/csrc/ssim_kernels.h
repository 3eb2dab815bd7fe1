// ssim_kernels.h -- launch interface between the runtime (ssim_cuda.cu) and the sm_100a kernels.
#ifndef SSIM_B200_KERNELS_H
#define SSIM_B200_KERNELS_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssimk {

// ---- geometry of the fused kernel (see DESIGN.md "Kernel")
constexpr int kBandW       = 64;   // output columns per warp work item (lane owns columns lane and lane+32)
constexpr int kBlkRows     = 8;    // rows per TMA block / horizontal-pass block
constexpr int kHalo        = 5;    // Gaussian radius (reference src/ssim.cpp:227)
constexpr int kBoxW        = 128;  // TMA box width in bytes: 16 left margin + 64 columns + right margin; 128 so that every box
                                   // row lands on a 128-byte aligned shared-memory address (TMA needs that of the box start)
constexpr int kBoxLeft     = 16;   // band column 0 sits at byte 16 of a box row: the innermost TMA coordinate (bx - 16) must be
                                   // a multiple of 16 bytes -- measured on B200: x = -16 works, x = -8 raises "illegal instruction"
                                   // (tools/dev/tma_probe.cu)
constexpr int kLoadRows    = 8;    // rows per TMA box (= one horizontal-pass block)
constexpr int kStages      = 2;    // TMA ring depth per work item (the producer has slack; 3 stages would not fit 2 CTAs/SM)
constexpr int kPairsPerCta = 4;    // work items per CTA: each is served by one producer warp and one consumer warp
constexpr int kCtaThreads  = 2 * kPairsPerCta * 32;      // 256: warps 0-3 producers (TMA + horizontal pass), 4-7 consumers
constexpr int kProducerRegs = 96;  // setmaxnreg budgets of the two warpgroups: 128*96 + 128*160 = 256*128
constexpr int kConsumerRegs = 160;
constexpr unsigned kBackoffNs = 200; // default sleep between polls of the partner warp's mbarrier
constexpr int kImgStageBytes = kBoxW * kLoadRows;        // 1024
constexpr int kStageBytes    = 2 * kImgStageBytes;       // 2048 (A then B)
constexpr int kRingPlaneBytes = kBandW * 8;              // 512: one row of packed {x, y} pairs
constexpr int kRingRowPad     = 32;                      // consecutive rows start 8 banks apart: see the ring layout in the kernel
constexpr int kRingRowBytes   = 2 * kRingPlaneBytes + kRingRowPad;   // 1056: {E[a'], E[b']} plane, {E[(a'-b')^2], E[a'b']} plane, pad
constexpr int kTaps           = 11;
constexpr int kRingRows       = 2 * kTaps;               // 22: two halves of 11 rows; the consumer's unrolled body is 11 rows
constexpr int kRingBytes      = kRingRows * kRingRowBytes; // 23232
constexpr int kPairSmemBytes = (kStages * kStageBytes + kRingBytes + 127) / 128 * 128;  // 27392
constexpr int kCtaSmemBytes  = kPairsPerCta * kPairSmemBytes;       // 109568 -> 2 CTAs (16 warps) per SM

// Per pixel type geometry of the TMA stage (the ring and everything after the widening are identical).  16-bit pixels
// (the extension the reference's README names): box rows of 80 elements = 160 bytes -- band column 0 sits
// at byte 16 in both layouts, the innermost TMA coordinate bx-8 elements is again a multiple of 16 bytes, and since a
// stage is one box only its start has to be 128-byte aligned.
template <bool kU16> struct PixGeo {
    static constexpr int kPixBytes      = kU16 ? 2 : 1;
    static constexpr int kBoxBytes      = kU16 ? 160 : kBoxW;
    static constexpr int kBoxElems      = kBoxBytes / kPixBytes;            // 128 / 80
    static constexpr int kBoxLeftElems  = kBoxLeft / kPixBytes;             // 16 / 8
    static constexpr int kImgStageBytes = kBoxBytes * kLoadRows;            // 1024 / 1280
    static constexpr int kStageBytes    = 2 * kImgStageBytes;               // 2048 / 2560
    static constexpr int kPairSmemBytes = (kStages * kStageBytes + kRingBytes + 127) / 128 * 128;   // 27392 / 28416
    static constexpr int kCtaSmemBytes  = kPairsPerCta * kPairSmemBytes;    // 109568 / 113664 (2 CTAs/SM either way)
};

struct FusedParams {
    int u16;                 // 0: 8-bit pixels, 1: 16-bit pixels (pitches and frame strides stay in BYTES)
    const uint8_t* a;        // raw planes (used only to fetch the per-item centring pixel)
    const uint8_t* b;
    long long pitchA, frameStrideA, pitchB, frameStrideB;
    float*  map;             // NULL when no map is wanted
    long long mapPitch, mapFrameStride;   // floats
    double* partials;        // [items] per-warp-item partial sums
    int width, srcRows, outY0, outRows, frames;
    int bands, segs, segRows;
    long long items;         // < 2^31 (checked by the host)
    uint32_t bandsMul, bandsShift, segsMul, segsShift;   // n / d == umulhi(n, mul) >> shift for n < 2^31 (d == 1: mul == 0); see fast_div()
    float g[6];              // separable 11-tap weights: g[d] is the tap at distance d from the centre
    float c1, c2;
    uint32_t magic;          // 0x4B000000 (float 2^23): kept opaque to ptxas, see the kernel
    uint32_t backoffNs;      // sleep between polls of the partner warp's mbarrier
    float eps2;              // 2*((sum of the 11x11 window) - 1): the reference window's normalisation bias, ~2.05e-8
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

// Work partition (pure host logic, unit-tested on the CPU: tests/clients/plan_check.cpp).  A work item is (frame, row
// segment, 64-column band).  The number of segments per frame decides both the halo overhead (10 extra input rows per
// segment) and how full the last wave of CTAs is.  Measured on B200 (tools/dev/batch_sweep.py, strip_sweep.py): 64 x 4K
// with 720-row segments = 9.73 waves 240.8k Mpix/s, 540-row = 12.97 waves 248.4k; a 16384 x 2058 strip with 515-row
// segments (0.86 waves) 187 us, 229-row (1.95 waves) 169 us.  Pick the count that minimises
//     waves x (rows + halo + per-item set-up),
// a last wave that fills at most half the CTA slots counting 0.4-0.7 (its CTAs have an SM to themselves and run faster);
// the candidates stop where a segment would drop below 24 rows.  overrideRows > 0 forces that many rows per segment.
inline void plan_segments(long long ctaSlots, uint32_t width, uint32_t outRows, uint32_t frames, int overrideRows, int* segRows, int* segs)
{
    const long long bands = ((long long)width + kBandW - 1) / kBandW;
    const long long units = bands * frames;
    long long s = 1;
    if (overrideRows > 0) {
        s = ((long long)outRows + overrideRows - 1) / overrideRows;
    } else {
        long long maxSegs = outRows / 24;
        if (maxSegs > 256) maxSegs = 256;
        if (maxSegs < 1) maxSegs = 1;
        if (ctaSlots < 1) ctaSlots = 1;
        double best = 0;
        for (long long cand = 1; cand <= maxSegs; ++cand) {
            const long long rows = ((long long)outRows + cand - 1) / cand;
            const long long nseg = ((long long)outRows + rows - 1) / rows;
            if (nseg != cand && cand != 1) continue;                    // same partition as a smaller candidate
            const long long ctas = (units * nseg + kPairsPerCta - 1) / kPairsPerCta;
            const long long full = ctas / ctaSlots, rem = ctas % ctaSlots;
            double tail = 0.0;
            if (rem > 0) tail = 2 * rem > ctaSlots ? 1.0 : 0.4 + 0.3 * (double)(2 * rem) / (double)ctaSlots;
            const double cost = ((double)full + tail) * (double)(rows + 2 * kHalo + 12);
            if (cand == 1 || cost < best) { best = cost; s = cand; }
        }
    }
    if (s > (long long)outRows) s = outRows;
    if (s < 1) s = 1;
    const int rows = (int)(((long long)outRows + s - 1) / s);
    *segRows = rows;
    *segs = (int)(((long long)outRows + rows - 1) / rows);
}

struct FinalizeParams {
    const double* partials;
    double* sums;            // may be NULL
    float*  ssim;            // may be NULL
    int itemsPerFrame;
    double invCount;         // 1 / double(uint32(width*outRows))
};

// Cross-GPU sum fused into the reduction kernel (strips of one image): every rank owns an exchange buffer of
// 2 x kMaxRanks slots; the reduction kernel of rank r stores its partial sum straight into slot [epoch & 1][r] of EVERY
// peer's buffer (NVLink peer stores, value then epoch with release semantics at system scope), then waits until its own
// buffer holds all `world` slots of this epoch and adds them up in rank order (deterministic, identical on every rank).
constexpr int kMaxRanks = 16;
struct ExchangeSlot { double value; unsigned long long epoch; };
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
cudaError_t launch_finalize_allreduce(cudaStream_t stream, const FinalizeParams& p, const ExchangeParams& x);

cudaError_t launch_fused(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const FusedParams& p);
cudaError_t launch_finalize(cudaStream_t stream, const FinalizeParams& p, int frames);
// per-device preparation (sets the dynamic shared-memory limit on the CURRENT device) + kernel facts
cudaError_t fused_kernel_attributes(int* regsMap, int* regsNoMap, int* ctasPerSm);

// layout helpers
cudaError_t launch_pack_u8(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                           long long step, long long stride, int width, int height);
cudaError_t launch_pack_u16(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                            long long step, long long stride, int width, int height);   // step/stride in bytes
cudaError_t launch_pack_luma(cudaStream_t stream, uint8_t* dst, long long dstPitch, const uint8_t* src,
                             long long step, long long stride, int width, int height);
cudaError_t launch_deinterleave_u8(cudaStream_t stream, uint8_t* dst, long long dstPitch, long long planeStride, const uint8_t* src,
                                   long long srcPitch, int channels, int width, int height);
cudaError_t launch_interleave_map(cudaStream_t stream, float* dst, long long dstPitch, const float* src, long long srcPitch,
                                  long long planeStride, int channels, int width, int height);
cudaError_t launch_scatter_map(cudaStream_t stream, float* dst, long long dstStep, long long dstStride,
                               const float* src, long long srcPitch, int width, int height);
cudaError_t launch_synth_fill(cudaStream_t stream, uint8_t* dA, long long pitchA, uint8_t* dB, long long pitchB,
                              int width, int rows, int y0, uint32_t frame, uint64_t seed);

}  // namespace ssimk
#endif
