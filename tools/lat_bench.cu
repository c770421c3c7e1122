// Latency / issue microbenchmarks that the scan kernel's design leans on (FP64 pipe, shuffles,
// shared memory, named barriers) -- B200, sm_100a.  Build and run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench tools/lat_bench.cu && ./lat_bench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

// dependent DFMA chain with ILP independent accumulators, executed by `nw` warps of the block that
// sit on the same SM sub-partition (warp id % 4 == 0)
template <int ILP>
__global__ void k_dfma(double *out, long long *cyc, int nw)
{
    const int warp = threadIdx.x >> 5;
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, b = 1e-9;
    __syncthreads();
    const bool on = (warp & 3) == 0 && (warp >> 2) < nw;
    long long t0 = clk();
    if (on) {
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, b);
        }
    }
    long long t1 = clk();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// one "victim" warp runs a dependent DADD chain while `nbg` background warps on the same
// sub-partition saturate the FP64 pipe with independent DFMAs
__global__ void k_contend(double *out, long long *cyc, int nbg)
{
    const int warp = threadIdx.x >> 5;
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, b = 1e-9;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        double x = a[0];
        long long t0 = clk();
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) x = x + b;
        long long t1 = clk();
        if (threadIdx.x == 0) { *cyc = t1 - t0; stop = 1; }
        a[0] = x;
    } else if ((warp & 3) == 0 && (warp >> 2) <= nbg) {
        while (!stop) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[threadIdx.x] = s;
}

__global__ void k_shfl(double *out, long long *cyc)
{
    double x = threadIdx.x;
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) x += __shfl_xor_sync(0xffffffffu, x, 1 + (it & 15));
    long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_lds(double *out, long long *cyc)
{
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) next[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) p = next[p];
    long long t1 = clk();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// store -> named barrier -> load round trip between two warps (ping-pong)
__global__ void k_bar(double *out, long long *cyc)
{
    __shared__ double box[2];
    const int warp = threadIdx.x >> 5;
    double x = threadIdx.x;
    if (threadIdx.x < 2) box[threadIdx.x] = 0;
    __syncthreads();
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (warp == 0) {
            if (threadIdx.x == 0) box[0] = x;
            asm volatile("bar.arrive 1, 64;" ::: "memory");
            asm volatile("bar.sync 2, 64;" ::: "memory");
            x += box[1];
        } else {
            asm volatile("bar.sync 1, 64;" ::: "memory");
            x += box[0];
            if (threadIdx.x == 32) box[1] = x;
            asm volatile("bar.arrive 2, 64;" ::: "memory");
        }
    }
    long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_rcp(double *out, long long *cyc)
{
    double x = 1.5 + threadIdx.x;
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        x = r + 1.25;
    }
    long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_sts_lds(double *out, long long *cyc)
{
    __shared__ double buf[256];
    double x = threadIdx.x;
    long long t0 = clk();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        buf[threadIdx.x] = x;
        __syncwarp();
        x += buf[threadIdx.x ^ 1];
        __syncwarp();
    }
    long long t1 = clk();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 4096 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
#define RUN(name, per, ...)                                                                  \
    do {                                                                                     \
        __VA_ARGS__;                                                                         \
        cudaDeviceSynchronize();                                                             \
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                              \
        printf("%-46s %8.2f cycles per %s\n", name, (double)h / ITERS, per);                 \
    } while (0)
    RUN("DFMA dependent, 1 warp ILP1", "DFMA", k_dfma<1><<<1, 512>>>(out, cyc, 1));
    RUN("DFMA 1 warp ILP2", "2 DFMA", k_dfma<2><<<1, 512>>>(out, cyc, 1));
    RUN("DFMA 1 warp ILP4", "4 DFMA", k_dfma<4><<<1, 512>>>(out, cyc, 1));
    RUN("DFMA 1 warp ILP8", "8 DFMA", k_dfma<8><<<1, 512>>>(out, cyc, 1));
    RUN("DFMA 2 warps/SMSP ILP1", "DFMA each", k_dfma<1><<<1, 512>>>(out, cyc, 2));
    RUN("DFMA 2 warps/SMSP ILP4", "4 DFMA each", k_dfma<4><<<1, 512>>>(out, cyc, 2));
    RUN("DFMA 2 warps/SMSP ILP8", "8 DFMA each", k_dfma<8><<<1, 512>>>(out, cyc, 2));
    RUN("DFMA 3 warps/SMSP ILP8", "8 DFMA each", k_dfma<8><<<1, 512>>>(out, cyc, 3));
    RUN("DFMA 1 warp ILP16", "16 DFMA", k_dfma<16><<<1, 512>>>(out, cyc, 1));
    RUN("DFMA 1 warp ILP32", "32 DFMA", k_dfma<32><<<1, 512>>>(out, cyc, 1));
    RUN("DFMA 2 warps/SMSP ILP16", "16 DFMA each", k_dfma<16><<<1, 512>>>(out, cyc, 2));
    RUN("DFMA 2 warps/SMSP ILP32", "32 DFMA each", k_dfma<32><<<1, 512>>>(out, cyc, 2));
    RUN("DFMA 3 warps/SMSP ILP16", "16 DFMA each", k_dfma<16><<<1, 512>>>(out, cyc, 3));
    RUN("DFMA 4 warps/SMSP ILP16", "16 DFMA each", k_dfma<16><<<1, 512>>>(out, cyc, 4));
    RUN("DADD dependent, idle SMSP", "DADD", k_contend<<<1, 512>>>(out, cyc, 0));
    RUN("DADD dependent vs 1 saturating warp", "DADD", k_contend<<<1, 512>>>(out, cyc, 1));
    RUN("DADD dependent vs 2 saturating warps", "DADD", k_contend<<<1, 512>>>(out, cyc, 2));
    RUN("DADD dependent vs 3 saturating warps", "DADD", k_contend<<<1, 512>>>(out, cyc, 3));
    RUN("SHFL.64 + DADD dependent", "pair", k_shfl<<<1, 32>>>(out, cyc));
    RUN("LDS pointer chase", "LDS", k_lds<<<1, 32>>>(out, cyc));
    RUN("STS + syncwarp + LDS + DADD", "round", k_sts_lds<<<1, 32>>>(out, cyc));
    RUN("bar ping-pong (2 x STS+bar+LDS+DADD)", "round trip", k_bar<<<1, 64>>>(out, cyc));
    RUN("MUFU.RCP64H + DADD dependent", "pair", k_rcp<<<1, 32>>>(out, cyc));
    return 0;
}
