// Does a host-to-device copy submitted while a long kernel runs start before that kernel ends?
// Variants of the kernel: plain; with > 48 KB of opted-in dynamic shared memory; with setmaxnreg;
// 148 or 100 CTAs.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o variants/overlap_probe tools/overlap_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <thread>

__global__ void spin_kernel(long long cycles, int *sink)
{
    extern __shared__ int sm[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { if (threadIdx.x == 999999) sm[0] = 1; }
    if (threadIdx.x == 0 && sm[0] == 123456) sink[0] = 1;
}

__global__ void __launch_bounds__(384, 1) spin_setmaxnreg_kernel(long long cycles, int *sink)
{
    extern __shared__ int sm[];
    if (threadIdx.x < 256) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { if (threadIdx.x == 999999) sm[0] = 1; }
    if (threadIdx.x == 0 && sm[0] == 123456) sink[0] = 1;
}

static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main()
{
    const size_t n = 1u << 30;
    void *h, *d; int *sink;
    cudaMallocHost(&h, n); cudaMalloc(&d, n); cudaMalloc(&sink, 4);
    cudaStream_t sk, sc;
    cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
    const long long cycles = 1000000000LL;   // ~0.5 s
    cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(spin_setmaxnreg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Case { const char *name; int grid, threads; size_t smem; int kind; };
    Case cases[] = {{"plain, 148 x 384, 1 KB dynamic smem", 148, 384, 1024, 0},
                    {"148 x 384, 160 KB dynamic smem", 148, 384, 160 * 1024, 0},
                    {"100 x 384, 160 KB dynamic smem", 100, 384, 160 * 1024, 0},
                    {"148 x 384, 160 KB + setmaxnreg 208/88", 148, 384, 160 * 1024, 1},
                    {"148 x 384, 40 KB + setmaxnreg 208/88", 148, 384, 40 * 1024, 1}};
    for (const Case &c : cases) {
        for (int dir = 0; dir < 2; ++dir) {
            cudaDeviceSynchronize();
            const double t0 = now();
            if (c.kind == 0) spin_kernel<<<c.grid, c.threads, c.smem, sk>>>(cycles, sink);
            else spin_setmaxnreg_kernel<<<c.grid, c.threads, c.smem, sk>>>(cycles, sink);
            cudaStreamQuery(sk);     // flush the launch
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
            if (dir == 0) cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, sc);
            else cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, sc);
            cudaStreamSynchronize(sc);
            const double t1 = now();
            cudaStreamSynchronize(sk);
            const double t2 = now();
            printf("%-44s %s issued at 100 ms: copy done at %6.0f ms, kernel done at %6.0f ms  (%s)\n", c.name,
                   dir == 0 ? "H2D" : "D2H", t1 - t0, t2 - t0, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
