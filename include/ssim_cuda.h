/*
 * ssim_cuda.h -- C ABI of libssim_cuda.so, the CUDA (sm_100a) engine under rmgr::ssim.
 *
 * Plain C types only (pointers, sizes, errno-style int returns); no C++/CUDA/torch types cross this
 * boundary, so it can be bound from C, C++, ctypes, cgo, JNI ...  librmgr-ssim.so (the reference's
 * API, include/rmgr/ssim.h) is a thin C++ layer over these calls.
 *
 * What each entry point replaces in the reference:
 *   ssim_cuda_compute()          the body of rmgr::ssim::compute_ssim() after parameter validation:
 *                                Gaussian set-up, the 256x64 tile loop / thread-pool dispatch and the
 *                                final mean (reference src/ssim.cpp:990-1103), i.e. process_tile()
 *                                (:747-783) = retrieve_tile (:515-583) + multiply x3 (:249-265) +
 *                                gaussian_blur x5 (:321-489, ISA variants src/ssim_{sse,avx,fma,neon}.cpp)
 *                                + sum_tile (:590-704), and run_in_openmp (src/ssim-openmp.c:26-37)
 *   ssim_cuda_compute_device()   the same, for callers that already hold the planes in device memory
 *                                (frames / strips; BASELINE.json configs 2-5).  No reference equivalent:
 *                                the reference has no device or batch notion; this is the entry the
 *                                bench and multi-GPU drivers use.
 *   ssim_cuda_compute_strips()   one large image split into row strips across several GPUs of one
 *                                process, the per-GPU sums exchanged over NVLink peer memory inside the
 *                                kernels (NCCL all-reduce on request); replaces the OpenMP tile
 *                                distribution for gigapixel inputs (src/ssim-openmp.c:26-47)
 *
 * Return values: 0, EINVAL, ENOMEM, ENODEV (no usable device), EIO (CUDA/NCCL runtime failure; text in
 * ssim_cuda_last_error_string()).  All functions are thread-safe; a call never retains caller pointers.
 */
#ifndef SSIM_CUDA_H
#define SSIM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSIM_CUDA_ABI_VERSION 2

int         ssim_cuda_abi_version(void);
int         ssim_cuda_device_count(void);            /* 0 when there is no driver / device */
int         ssim_cuda_init(int device);              /* optional: creates the per-device context eagerly */
void        ssim_cuda_shutdown(void);                /* releases every context; safe to call twice */
const char* ssim_cuda_last_error_string(void);       /* thread-local, never NULL */

/* Pinned host memory for callers that want the fastest host path (optional; any host memory works). */
void* ssim_cuda_host_alloc(size_t bytes);
void  ssim_cuda_host_free(void* p);

/*
 * One image pair, any layout.  Pixel (x,y) of image A is a[x*stepA + y*strideA] (bytes, signed);
 * map pixel (x,y) is map[x*mapStep + y*mapStride] (floats, signed).  `a`, `b`, `map` may each be host
 * or device pointers.  `ssim` (host) and `map` may be NULL, but not both.  Blocking.
 * Semantics: 11x11 Gaussian window sigma 1.5, C1=(0.01*255)^2, C2=(0.03*255)^2, clamp-to-edge borders,
 * float per-pixel values, double global sum, *ssim = float(sum / double(uint32(width*height))).
 */
int ssim_cuda_compute(int device, uint32_t width, uint32_t height,
                      const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                      const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                      float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                      float* ssim);

/*
 * SSIM of the BT.601 luma of two interleaved RGB(A) images: rgb points at the R byte of pixel (0,0), step >= 3 is the
 * distance between pixels.  The luma planes ((19595 R + 38470 G + 7471 B + 32768) >> 16, the integer formula of the
 * reference CLI, src/ssim-cli.cpp:158-186) are produced on the GPU.  Otherwise like ssim_cuda_compute().  Blocking.
 */
int ssim_cuda_compute_luma(int device, uint32_t width, uint32_t height,
                           const uint8_t* rgbA, ptrdiff_t stepA, ptrdiff_t strideA,
                           const uint8_t* rgbB, ptrdiff_t stepB, ptrdiff_t strideB,
                           float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                           float* ssim);

/*
 * All channels of two interleaved host images in one pass: channel c of pixel (x,y) is a[y*strideA + x*channels + c].
 * Replaces the per-channel loop of the reference's CLI and tests (src/ssim-cli.cpp:199-209, tests/rmgr-ssim-tests.cpp:273-291),
 * which call compute_ssim() once per channel with step = channels on the same bytes: here the bytes are uploaded once, split
 * on the GPU and all channels go through ONE fused launch.  ssim (or NULL) receives `channels` floats; map (or NULL)
 * receives channels floats per pixel: map[y*mapStride + x*channels + c].  Strides must be positive.  Blocking.
 */
int ssim_cuda_compute_channels(int device, uint32_t width, uint32_t height, uint32_t channels,
                               const uint8_t* a, ptrdiff_t strideA, const uint8_t* b, ptrdiff_t strideB,
                               float* map, ptrdiff_t mapStride, float* ssim);

/*
 * Device-resident planes, asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 * stream).  `frames` independent pairs are processed by ONE kernel launch, reduction included.
 *
 *   dA/dB      8-bit planes, 1 byte per pixel; pixel (x,y) of frame f at d[f*frameStride + y*pitch + x].
 *              Base address, pitch and frameStride must be multiples of 16 bytes (TMA requirement).
 *   srcRows    rows present in each plane; rows outside [0,srcRows) replicate the nearest one.
 *   outY0,outRows  the rows for which SSIM is produced: a full image uses (0, srcRows); a strip that
 *              carries 5 halo rows on an interior edge uses outY0 = 5 there.  Columns always clamp at
 *              [0,width).
 *   dMap       NULL, or floats: value of (x, outY0+r) of frame f at dMap[f*mapFrameStride + r*mapPitch + x]
 *   dSums      NULL or [frames] doubles: sum of the SSIM values of the outRows x width outputs
 *   dSsim      NULL or [frames] floats : float(sum / double(uint32(width*outRows)))
 */
int ssim_cuda_compute_device(int device, void* stream,
                             uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows, uint32_t frames,
                             const uint8_t* dA, size_t pitchA, size_t frameStrideA,
                             const uint8_t* dB, size_t pitchB, size_t frameStrideB,
                             float* dMap, size_t mapPitch, size_t mapFrameStride,
                             double* dSums, float* dSsim);

/*
 * 16-bit pixels (dynamic range L = 65535: C1 = (0.01*65535)^2, C2 = (0.03*65535)^2), the extension the reference names but
 * does not implement (reference README.md:107-111; L is hard-wired to 255 at src/ssim.cpp:958, retrieve_tile reads bytes at
 * src/ssim.cpp:515-516).  Same kernel, same window, same border rule.  Parity is pinned against the reference's OWN
 * template implementation: tests/ssim_naive.h instantiated as compute_ssim<double, uint16_t> (it takes its dynamic
 * range from the pixel type) is compiled from the reference tree by oracle/Makefile (oracle/_ref/libnaive.so) and gates both
 * the 16-bit oracle and the GPU path at |delta| <= 2e-6 global / 1e-3 per pixel (tests/test_oracle.py, tests/test_u16_gpu.py,
 * committed vectors tests/golden/golden.json "u16_naive"); the scale invariance SSIM_16(257*a, 257*b) == SSIM_8(a, b) is a
 * second, independent check.
 *
 * ssim_cuda_compute_u16():        like ssim_cuda_compute(); stepA/strideA/stepB/strideB are in uint16 ELEMENTS (signed).
 * ssim_cuda_compute_device_u16(): like ssim_cuda_compute_device(); pitches and frame strides in BYTES, multiples of 16.
 */
int ssim_cuda_compute_u16(int device, uint32_t width, uint32_t height,
                          const uint16_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                          const uint16_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                          float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                          float* ssim);
int ssim_cuda_compute_device_u16(int device, void* stream,
                                 uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows, uint32_t frames,
                                 const uint16_t* dA, size_t pitchA, size_t frameStrideA,
                                 const uint16_t* dB, size_t pitchB, size_t frameStrideB,
                                 float* dMap, size_t mapPitch, size_t mapFrameStride,
                                 double* dSums, float* dSsim);

/* Number of kernels the previous ssim_cuda_compute_device() call on this thread launched (for bench.py): 1 */
int ssim_cuda_last_launch_count(void);

/*
 * One host image pair split into horizontal strips (with 5 halo rows on interior edges) across
 * `nDevices` GPUs of this process.  The per-GPU double sums are combined INSIDE the kernels: the warp that completes a
 * GPU's strip sum stores it into the peers' exchange buffers over NVLink and adds up what lands in its own
 * (see ssim_cuda_compute_strip_allreduce below); SSIM_CUDA_STRIPS_NCCL=1 selects one ncclAllReduce of a double per GPU
 * instead.  Same argument meaning as ssim_cuda_compute(); pointers must be host pointers.
 */
int ssim_cuda_compute_strips(int nDevices, const int* devices, uint32_t width, uint32_t height,
                             const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                             const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                             float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                             float* ssim);

/*
 * Strips of one image across GPUs with the cross-GPU sum fused into the kernel (the one launch per rank): instead of an NCCL
 * all-reduce after the launch (torch.distributed in a one-process-per-GPU driver), the warp that completes the strip's sum on
 * a rank stores it straight into every peer's exchange buffer over NVLink (relaxed peer stores of self-validating words at
 * system scope), waits for the other ranks' sums to land in its own buffer and adds them in rank order: same result bits on
 * every rank, no host round trip, no collective launch.  Replaces the same reference code as ssim_cuda_compute_strips().
 *
 *   exchange_create   allocates this rank's exchange buffer; ipcHandle64 (64 bytes, may be NULL) receives a handle that other
 *                     PROCESSES pass to exchange_open(); threads of one process may use the returned pointer directly after
 *                     exchange_enable_peer(device, ownerDevice).
 *   compute_strip_allreduce   like ssim_cuda_compute_device() for ONE strip (frames = 1) of an image of `imageRows` rows:
 *                     peerBufs[r] = rank r's exchange buffer as seen from this device (own buffer at [rank]); epoch = 1, 2, 3...
 *                     the same on every rank for the same image; dSumAll/dSsimAll receive the sum over all strips and
 *                     float(sum / double(uint32(width*imageRows))); *dStatus = 1 (and the results NaN) if a peer did not
 *                     show up within SSIM_CUDA_EXCHANGE_TIMEOUT_MS (default 2000) -- the kernel never spins forever.
 */
int ssim_cuda_exchange_create(int device, void** dBuf, void* ipcHandle64);
int ssim_cuda_exchange_open(int device, const void* ipcHandle64, void** dPeerBuf);
int ssim_cuda_exchange_close(int device, void* dPeerBuf);
int ssim_cuda_exchange_destroy(int device, void* dBuf);
int ssim_cuda_exchange_enable_peer(int device, int peerDevice);
int ssim_cuda_compute_strip_allreduce(int device, void* stream,
                                      uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows, uint32_t imageRows,
                                      const uint8_t* dA, size_t pitchA, const uint8_t* dB, size_t pitchB,
                                      float* dMap, size_t mapPitch,
                                      void* const* peerBufs, int world, int rank, uint64_t epoch,
                                      double* dSumAll, float* dSsimAll, int* dStatus);

/* Fills device planes with rows y0..y0+rows-1 of synthetic frame `frame` (ssim_b200/csrc/synth.h). */
int ssim_cuda_synth_fill(int device, void* stream, uint8_t* dA, size_t pitchA, uint8_t* dB, size_t pitchB,
                         uint32_t width, uint32_t rows, uint32_t y0, uint32_t frame, uint64_t seed);

/*
 * Tuning knobs of the persistent kernel's work partition, for experiments/bench only (0 = automatic for either):
 * maxPairsPerSm caps the warp pairs per SM the work is spread over (1..8), minSlotRows is the smallest share of rows (incl.
 * the 10 start-up rows per column) a warp pair is given before fewer pairs are used.  Process-wide.
 */
void ssim_cuda_set_tuning(int maxPairsPerSm, int minSlotRows);

/*
 * Development aid: while dTimes (device memory, 32 x 64-bit words per warp pair of the persistent grid, i.e. at least
 * 32 * 8 * numSMs words) is non-NULL, every launch records the %globaltimer at which each consumer warp started [0], finished
 * its share [1] and finished its k-th 8-row ring unit [1+k]; NULL switches it off.  Used by tools/dev/slot_times.py to look at load balance.
 */
void ssim_cuda_debug_slot_times(unsigned long long* dTimes);

/*
 * Development aid for the tests: the entry (0..511) of the per-thread tensor-map descriptor cache that a plane with this base
 * address and geometry uses.  Two planes of one call may share an entry; tests/test_device_api_gpu.py builds such a pair.
 */
int ssim_cuda_debug_map_cache_entry(const void* base, uint32_t width, uint32_t rows, uint32_t frames, size_t pitch, int elemBytes);

#ifdef __cplusplus
}
#endif
#endif /* SSIM_CUDA_H */
