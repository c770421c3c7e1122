// How fast can the MATRIX side of a k-step blocked scan run?  ("k-step blocked
// recurrence".)  A stand-in for the matrix warps of scan_fast.cu with the same resources -- 148 CTAs x
// (256 matrix + 128 helper) threads, setmaxnreg 208 / 88, one 8x8 register tile of the symmetric
// J = 176 state per matrix thread, operands from shared memory, 2x2 exchange of the partial sums and
// 16-byte stores -- but per hand-over it applies a rank-K update and forms K matrix-vector products
// against the same state (K = 1 is the shape of today's phase without the quadratic-form butterfly).
// The helpers only take part in the barrier; operands are constant, results are checksummed.
// Prints cycles per time step.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o variants/blk_matrix_bench tools/blk_matrix_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int TILE = 8, NB = 22, J = NB * TILE, NTILE = NB * (NB + 1) / 2;   // 253 tiles
constexpr int MAT = 256, THREADS = 384, NBP = NB + 1;

template <int K>
struct Smem {
    // element-major operand layout [pair of elements q][block], pitch NBP: the lanes of a warp (consecutive
    // blocks) read consecutive 16-byte words -- conflict-free, as the A / C arrays of scan_fast.cu
    double2 U[K][4][NBP];       // rows u~_i of this block
    double2 T[K][4][NBP];       // t~_m of the previous block (rank-K update S += sum t~ w~^T)
    double2 W[K][4][NBP];
    double2 P[K][4][MAT];       // partial sums after the 2x2 exchange: 4 x 16 B per thread and row
};

__device__ __forceinline__ double shfl_xor_d(double x, int m) { return __shfl_xor_sync(0xffffffffu, x, m); }

template <int K, int PAIR, int NOLOAD, int NOSTORE>
__global__ void __launch_bounds__(THREADS, 1) blk_kernel(int blocks, double *out, long long *cycles)
{
    extern __shared__ __align__(16) unsigned char raw[];
    Smem<K> &sm = *reinterpret_cast<Smem<K> *>(raw);
    const int tid = threadIdx.x;
    for (int i = tid; i < K * 4 * NBP * 2; i += THREADS) {
        reinterpret_cast<double *>(sm.U)[i] = 1e-3 * ((i * 37) % 101 - 50);
        reinterpret_cast<double *>(sm.T)[i] = 1e-4 * ((i * 53) % 97 - 48);
        reinterpret_cast<double *>(sm.W)[i] = 1e-4 * ((i * 29) % 89 - 44);
    }
    __syncthreads();
    if (tid >= MAT) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        for (int b = 0; b < blocks; ++b) asm volatile("bar.sync 1, 384;" ::: "memory");
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    int bi = 0, bj = 0;
    {
        int rem = tid < NTILE ? tid : 0, row = 0;
        while (rem >= NB - row) { rem -= NB - row; ++row; }
        bi = row; bj = row + rem;
    }
    const double2 *ub = &sm.U[0][0][bi], *uc = &sm.U[0][0][bj];
    const double2 *tb = &sm.T[0][0][bi], *wc = &sm.W[0][0][bj];
    double S[TILE][TILE];
#pragma unroll
    for (int i = 0; i < TILE; ++i)
#pragma unroll
        for (int j = 0; j < TILE; ++j) S[i][j] = 1e-2 * (i - j);
    const bool hr = (tid & 2) != 0, hc = (tid & 1) != 0;
    // NOLOAD: operands live in registers (loaded once), to separate the FP64 / register-file side from
    // the shared-memory side; NOSTORE: no exchange / stores, the sums are folded into a checksum
    double kt[TILE], kw[TILE], kr[TILE], kc[TILE], fold = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        kt[2 * q] = tb[q * NBP].x; kt[2 * q + 1] = tb[q * NBP].y; kw[2 * q] = wc[q * NBP].x; kw[2 * q + 1] = wc[q * NBP].y;
        kr[2 * q] = ub[q * NBP].x; kr[2 * q + 1] = ub[q * NBP].y; kc[2 * q] = uc[q * NBP].x; kc[2 * q + 1] = uc[q * NBP].y;
    }
    const long long t0 = clock64();
    for (int b = 0; b < blocks; ++b) {
        // rank-K update
#pragma unroll
        for (int m = 0; m < K; ++m) {
            double t[TILE], w[TILE];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (NOLOAD) {
                    t[2 * q] = kt[2 * q]; t[2 * q + 1] = kt[2 * q + 1]; w[2 * q] = kw[2 * q]; w[2 * q + 1] = kw[2 * q + 1];
                } else {
                    const double2 a = tb[(m * 4 + q) * NBP], c = wc[(m * 4 + q) * NBP];
                    t[2 * q] = a.x; t[2 * q + 1] = a.y; w[2 * q] = c.x; w[2 * q + 1] = c.y;
                }
            }
#pragma unroll
            for (int i = 0; i < TILE; ++i)
#pragma unroll
                for (int j = 0; j < TILE; ++j) S[i][j] = fma(t[i], w[j], S[i][j]);
        }
        // K matrix-vector products against the same tile, PAIR rows at a time
#pragma unroll
        for (int i0 = 0; i0 < K; i0 += PAIR) {
            double rowp[PAIR][TILE], colp[PAIR][TILE];
#pragma unroll
            for (int p = 0; p < PAIR; ++p) {
                double ur[TILE], uj[TILE];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (NOLOAD) {
                        ur[2 * q] = kr[2 * q]; ur[2 * q + 1] = kr[2 * q + 1]; uj[2 * q] = kc[2 * q]; uj[2 * q + 1] = kc[2 * q + 1];
                    } else {
                        const double2 a = ub[((i0 + p) * 4 + q) * NBP], c = uc[((i0 + p) * 4 + q) * NBP];
                        ur[2 * q] = a.x; ur[2 * q + 1] = a.y; uj[2 * q] = c.x; uj[2 * q + 1] = c.y;
                    }
                }
#pragma unroll
                for (int j = 0; j < TILE; ++j) colp[p][j] = ur[0] * S[0][j];
#pragma unroll
                for (int i = 1; i < TILE; ++i)
#pragma unroll
                    for (int j = 0; j < TILE; ++j) colp[p][j] = fma(ur[i], S[i][j], colp[p][j]);
#pragma unroll
                for (int i = 0; i < TILE; ++i) rowp[p][i] = S[i][0] * uj[0];
#pragma unroll
                for (int j = 1; j < TILE; ++j)
#pragma unroll
                    for (int i = 0; i < TILE; ++i) rowp[p][i] = fma(S[i][j], uj[j], rowp[p][i]);
            }
#pragma unroll
            if (NOSTORE) {
#pragma unroll
                for (int p = 0; p < PAIR; ++p)
#pragma unroll
                    for (int q = 0; q < TILE; ++q) fold += rowp[p][q] + colp[p][q];
                // keep the rank-K operands changing so that nothing is hoisted out of the loop
                kt[0] = fold * 1e-300;
            } else
#pragma unroll
            for (int p = 0; p < PAIR; ++p) {
                double rs[4], cs[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double send_r = hr ? rowp[p][q] : rowp[p][4 + q];
                    const double keep_r = hr ? rowp[p][4 + q] : rowp[p][q];
                    rs[q] = keep_r + shfl_xor_d(send_r, 1);
                    const double send_c = hc ? colp[p][q] : colp[p][4 + q];
                    const double keep_c = hc ? colp[p][4 + q] : colp[p][q];
                    cs[q] = keep_c + shfl_xor_d(send_c, 2);
                }
                sm.P[i0 + p][0][tid] = make_double2(rs[0], rs[1]);
                sm.P[i0 + p][1][tid] = make_double2(rs[2], rs[3]);
                sm.P[i0 + p][2][tid] = make_double2(cs[0], cs[1]);
                sm.P[i0 + p][3][tid] = make_double2(cs[2], cs[3]);
            }
        }
        asm volatile("bar.sync 1, 384;" ::: "memory");
    }
    const long long t1 = clock64();
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < TILE; ++i)
#pragma unroll
        for (int j = 0; j < TILE; ++j) acc += S[i][j];
    acc += sm.P[0][0][tid].x + fold;
    out[(size_t)blockIdx.x * MAT + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K, int PAIR, int NOLOAD = 0, int NOSTORE = 0>
void run(int steps)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * MAT * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaFuncSetAttribute(blk_kernel<K, PAIR, NOLOAD, NOSTORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<K>));
    const int blocks = steps / K;
    for (int rep = 0; rep < 2; ++rep) blk_kernel<K, PAIR, NOLOAD, NOSTORE><<<148, THREADS, sizeof(Smem<K>)>>>(blocks, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long v : h) mx = v > mx ? v : mx;
    printf("K = %d, %d row(s) at a time%s%s: %7.1f cycles per time step (%s); FP64 issue floor 2 warps x 192 DFMA x 2.07 = 795\n",
           K, PAIR, NOLOAD ? ", operands in registers" : "", NOSTORE ? ", no exchange / stores" : "",
           (double)mx / ((double)blocks * K), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    const int steps = argc > 1 ? atoi(argv[1]) : 32768;
    run<1, 1>(steps);
    run<2, 1>(steps);
    run<2, 2>(steps);
    run<4, 1>(steps);
    run<4, 2>(steps);
    run<8, 1>(steps);
    run<8, 2>(steps);
    run<1, 1, 1, 0>(steps);
    run<1, 1, 0, 1>(steps);
    run<1, 1, 1, 1>(steps);
    run<4, 1, 0, 1>(steps);
    run<8, 1, 0, 1>(steps);
    return 0;
}
