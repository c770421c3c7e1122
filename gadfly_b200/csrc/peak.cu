// FP64 FMA throughput microbenchmark: the denominator of the scan's roofline.
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the O(N J^2) scan is bound by
// the FP64 FMA pipe, so the library measures that pipe itself on the device it runs on:
// independent DFMA chains, enough warps to cover the pipe latency, flops = 2 per FMA.
#include "common.cuh"

namespace gf {

namespace {

constexpr int PK_THREADS = 256;
constexpr int PK_CHAINS = 8;
constexpr int PK_INNER = 64;

__global__ void __launch_bounds__(PK_THREADS) dfma_kernel(double *out, int iters, double a, double b)
{
    double x[PK_CHAINS];
#pragma unroll
    for (int c = 0; c < PK_CHAINS; ++c) x[c] = 1e-3 * (threadIdx.x + c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < PK_INNER; ++k) {
#pragma unroll
            for (int c = 0; c < PK_CHAINS; ++c) x[c] = fma(x[c], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < PK_CHAINS; ++c) s += x[c];
    if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace

cudaError_t measure_fp64_peak(int sm_count, cudaStream_t stream, double *flops)
{
    double *out = nullptr;
    cudaError_t e = cudaMalloc(&out, sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sm_count * 4;
    const int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, stream);
        dfma_kernel<<<grid, PK_THREADS, 0, stream>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * (double)grid * PK_THREADS * PK_CHAINS * PK_INNER * (double)iters;
        if (rep > 0 && ms > 0.0f) best = fmax(best, fl / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e == cudaSuccess) e = cudaGetLastError();
    *flops = best;
    return e;
}

}  // namespace gf
