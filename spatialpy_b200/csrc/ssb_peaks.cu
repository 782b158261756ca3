// ssb_peaks.cu — measurement-only: the fp64 FMA peak of the device (the neighbour sweeps are fp64-issue / gather
// bound, so the bench reports their flop rate against a MEASURED fp64 peak next to the HBM roofline).  Not part of the engine:
// built into its own library (libssb_peaks.so) and run by bench.py in a child process.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ssb_peaks.h"

// 8 independent DFMA chains per thread, 2048 resident threads per SM: issue bound on the fp64 pipe, no memory traffic
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + 1.0e-9 * (double) (threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = fma(a[k], b, c);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int ssb_fp64_peak(int device, double *tflops, double *best_ms) {
    if (!tflops) return 4;
    if (cudaSetDevice(device) != cudaSuccess) return 3;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 3;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *d_out = nullptr;
    if (cudaMalloc((void **) &d_out, sizeof(double) * (size_t) blocks * threads) != cudaSuccess) return 3;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1.0e30f;
    int rc = 0;
    for (int rep = 0; rep < 5 && !rc; rep++) {          // rep 0 is the warm-up
        cudaEventRecord(e0, 0);
        k_dfma<<<blocks, threads>>>(d_out, iters, 1.0000001, 1.0e-9);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = 3; break; }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    if (rc) return rc;
    const double flop = 2.0 * 8.0 * (double) iters * (double) blocks * (double) threads;
    *tflops = flop / ((double) best * 1.0e-3) / 1.0e12;
    if (best_ms) *best_ms = (double) best;
    return 0;
}
