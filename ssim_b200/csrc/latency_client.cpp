// latency_client.cpp -- what one ssim_cuda_compute_device() call costs a C/C++ caller (no Python in the timed path).
// Usage: latency_client [width height map(0|1) reps]; prints one JSON object.
//   one_call_us  : cudaEvents around ONE call on an idle stream (event, call, event, synchronize), median / min of reps
//   queued_us    : reps calls queued back to back, per call
//   host_call_us : CPU time the call itself takes (launch included), median
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "ssim_cuda.h"

#define CHECK(x) do { if ((x) != cudaSuccess) { std::fprintf(stderr, "CUDA error at %s\n", #x); return 2; } } while (0)

int main(int argc, char** argv)
{
    const unsigned W = argc > 1 ? std::atoi(argv[1]) : 3840, H = argc > 2 ? std::atoi(argv[2]) : 2160;
    const int withMap = argc > 3 ? std::atoi(argv[3]) : 1, reps = argc > 4 ? std::atoi(argv[4]) : 50;
    const size_t pitch = (W + 15) / 16 * 16;
    unsigned char *a, *b; float* m = nullptr; double* sum; float* val;
    CHECK(cudaMalloc(&a, pitch * H)); CHECK(cudaMalloc(&b, pitch * H));
    if (withMap) CHECK(cudaMalloc(&m, sizeof(float) * W * H));
    CHECK(cudaMalloc(&sum, sizeof(double))); CHECK(cudaMalloc(&val, sizeof(float)));
    cudaStream_t st; CHECK(cudaStreamCreate(&st));
    if (ssim_cuda_synth_fill(0, st, a, pitch, b, pitch, W, H, 0, 0, 0x5517)) { std::fprintf(stderr, "%s\n", ssim_cuda_last_error_string()); return 3; }
    auto call = [&] { return ssim_cuda_compute_device(0, st, W, H, 0, H, 1, a, pitch, 0, b, pitch, 0, m, W, 0, sum, val); };
    for (int i = 0; i < 5; ++i) if (call()) { std::fprintf(stderr, "%s\n", ssim_cuda_last_error_string()); return 3; }
    CHECK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    std::vector<float> one, host;
    for (int i = 0; i < reps; ++i) {
        CHECK(cudaStreamSynchronize(st));
        CHECK(cudaEventRecord(e0, st));
        const auto t0 = std::chrono::steady_clock::now();
        call();
        const auto t1 = std::chrono::steady_clock::now();
        CHECK(cudaEventRecord(e1, st));
        CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
        one.push_back(ms * 1e3f);
        host.push_back(std::chrono::duration<float, std::micro>(t1 - t0).count());
    }
    CHECK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) call();
    CHECK(cudaEventRecord(e1, st));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    float ssim = 0; CHECK(cudaMemcpy(&ssim, val, sizeof(float), cudaMemcpyDeviceToHost));
    std::sort(one.begin(), one.end()); std::sort(host.begin(), host.end());
    std::printf("{\"width\": %u, \"height\": %u, \"map\": %d, \"reps\": %d, \"one_call_us\": {\"median\": %.2f, \"min\": %.2f}, \"queued_us\": %.2f, "
                "\"host_call_us\": %.2f, \"ssim\": %.7f}\n", W, H, withMap, reps, one[one.size() / 2], one[0], ms * 1e3f / reps, host[host.size() / 2], ssim);
    return 0;
}
