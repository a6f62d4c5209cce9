// Pipe-rate microbenchmarks for bench.py's roofline denominators ("measure FP32 / FP64 / POPC peaks with
// microbenchmarks first and use MEASURED peaks").  MEASURED_PEAKS.json only carries HBM copy bandwidth and cuBLAS bf16; the
// pair stage of this path runs on the FP64 pipe without FMA contraction (k_ransac, k_cheirality: cv2's arithmetic is unfused),
// the POPC matcher on the XU pipe and the default matcher on the int8 tensor pipe -- their rates are measured here, on the
// device the benchmark runs on, with the same clocks.  Each kernel is a register-only loop with enough independent chains
// to cover the pipe latency; the result is the best of a few launches timed with CUDA events.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dvo.h"

namespace dvo {
double nn_tensor_peak_tops(int numSms, int iters);      // nn_tensor.cu: back-to-back tcgen05.mma.kind::i8 from resident tiles

namespace {

constexpr int kChains = 8;

template <int kKind>
__global__ void __launch_bounds__(256) k_pipe_rate(int iters, double seed, double* sink) {
    // kKind: 0 FFMA, 1 FMUL+FADD (no contraction), 2 DFMA, 3 DMUL+DADD, 4 POPC
    if (kKind <= 1) {
        float a[kChains], m = 1.0000001f + (float)seed, c = 1e-7f;
#pragma unroll
        for (int k = 0; k < kChains; ++k) a[k] = (float)(threadIdx.x + k) * 1e-3f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) {
                if (kKind == 0) a[k] = __fmaf_rn(a[k], m, c);
                else a[k] = __fadd_rn(__fmul_rn(a[k], m), c);
            }
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < kChains; ++k) s += a[k];
        if (s == 12345.678f) *sink = s;
    } else if (kKind <= 3) {
        double a[kChains], m = 1.0000000001 + seed, c = 1e-9;
#pragma unroll
        for (int k = 0; k < kChains; ++k) a[k] = (double)(threadIdx.x + k) * 1e-3;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) {
                if (kKind == 2) a[k] = __fma_rn(a[k], m, c);
                else a[k] = __dadd_rn(__dmul_rn(a[k], m), c);
            }
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < kChains; ++k) s += a[k];
        if (s == 12345.678) *sink = s;
    } else {
        uint32_t a[kChains], acc = 0;
#pragma unroll
        for (int k = 0; k < kChains; ++k) a[k] = threadIdx.x * 2654435761u + k + (uint32_t)seed;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < kChains; ++k) a[k] = __popc(a[k]) + a[k];      // POPC on the XU pipe, the add on the ALU pipe
        }
#pragma unroll
        for (int k = 0; k < kChains; ++k) acc += a[k];
        if (acc == 0x12345678u) *sink = (double)acc;
    }
}

template <int kKind>
double run_rate(int numSms, double* d_sink) {
    const int iters = kKind >= 2 && kKind <= 3 ? 2048 : 8192;
    const int grid = numSms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_pipe_rate<kKind><<<grid, 256>>>(iters, 0.0, d_sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * 256 * iters * kChains * (kKind == 4 ? 1.0 : 2.0);   // a fused or unfused multiply-add = 2 flop
        if (rep > 0 && ms > 0) best = best > ops / (ms * 1e-3) ? best : ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

}  // namespace
}  // namespace dvo

extern "C" int dvo_measure_peaks(int device, double* out, int n) {
    if (!out || n < DVO_PEAK_COUNT) return DVO_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return DVO_E_NODEVICE;
    int numSms = 148;
    cudaDeviceGetAttribute(&numSms, cudaDevAttrMultiProcessorCount, device);
    double* d_sink = nullptr;
    if (cudaMalloc(&d_sink, 8) != cudaSuccess) return DVO_E_CUDA;
    out[DVO_PEAK_FP32_FMA] = dvo::run_rate<0>(numSms, d_sink);
    out[DVO_PEAK_FP32_MUL_ADD] = dvo::run_rate<1>(numSms, d_sink);
    out[DVO_PEAK_FP64_FMA] = dvo::run_rate<2>(numSms, d_sink);
    out[DVO_PEAK_FP64_MUL_ADD] = dvo::run_rate<3>(numSms, d_sink);
    out[DVO_PEAK_POPC] = dvo::run_rate<4>(numSms, d_sink);
    out[DVO_PEAK_INT8_TENSOR] = dvo::nn_tensor_peak_tops(numSms, 4096);
    cudaFree(d_sink);
    return cudaGetLastError() == cudaSuccess ? DVO_OK : DVO_E_CUDA;
}
