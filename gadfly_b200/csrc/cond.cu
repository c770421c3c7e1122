// Conditional mean at new times (celerite2 general_matmul_lower + general_matmul_upper):
//
//   mu[i] = sum_m k(|ts[i] - t[m]|) alpha[m],   k(tau) = sum_j exp(-c_j tau) (a_j cos d_j tau + b_j sin d_j tau)
//
// in O((N + M) J): forward state F = sum_{t_m <= t*} exp(-c (t* - t_m)) (cos, sin)(d t_m) alpha_m
// read through U(t*), backward state B = sum_{t_m > t*} exp(-c (t_m - t*)) U(t_m) alpha_m read through
// V(t*) (SURVEY.md A.2 generators; reference use: gadfly/gp.py:243-306 predict at new times).
// The recurrences are linear in the state, so every complex term is independent: one thread per
// term walks the merged, sorted time axes; the J/2 contributions of a new point are summed over
// the block.  Block 0 does the forward pass, block 1 the backward pass; out = [forward[M] | backward[M]].
#include "common.cuh"

namespace gf {

namespace {

constexpr int COND_THREADS = 128;

__device__ __forceinline__ double block_sum(double v, double *red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int warp = threadIdx.x >> 5;
    __syncthreads();                       // red is free again
    if ((threadIdx.x & 31) == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < COND_THREADS / 32; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(COND_THREADS)
cond_mean_kernel(const int64_t N, const double *__restrict__ t, const int64_t M,
                 const double *__restrict__ ts, const int Jc, const double *__restrict__ coef,
                 const double *__restrict__ alpha, double *__restrict__ out)
{
    __shared__ double red[COND_THREADS / 32];
    const bool backward = blockIdx.x == 1;
    double *o = out + (backward ? M : 0);
    // more terms than threads: several passes over the axes, accumulating into o
    const int passes = (Jc + COND_THREADS - 1) / COND_THREADS;
    for (int pass = 0; pass < (passes > 0 ? passes : 1); ++pass) {
        const int j0 = pass * COND_THREADS;
        const int j = j0 + threadIdx.x;
        const bool on = j < Jc;
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
        if (on) { a = coef[4 * j]; b = coef[4 * j + 1]; c = coef[4 * j + 2]; d = coef[4 * j + 3]; }
        double s0 = 0.0, s1 = 0.0, t_last = 0.0;
        bool any = false;
        if (!backward) {
            int64_t m = 0;
            for (int64_t i = 0; i < M; ++i) {
                const double tq = ts[i];
                while (m < N && t[m] <= tq) {
                    const double tm = t[m], am = alpha[m];
                    if (any) { const double p = exp(-c * (tm - t_last)); s0 *= p; s1 *= p; }
                    double sn, cs;
                    sincos_cw(__dmul_rn(d, tm), &sn, &cs);
                    s0 = fma(cs, am, s0); s1 = fma(sn, am, s1);
                    t_last = tm; any = true; ++m;
                }
                double contrib = 0.0;
                if (any && on) {
                    const double p = exp(-c * (tq - t_last));
                    double sn, cs;
                    sincos_cw(__dmul_rn(d, tq), &sn, &cs);
                    contrib = p * ((a * cs + b * sn) * s0 + (a * sn - b * cs) * s1);
                }
                const double tot = block_sum(contrib, red);
                if (threadIdx.x == 0) o[i] = (j0 == 0 ? 0.0 : o[i]) + tot;
            }
        } else {
            int64_t m = N - 1;
            for (int64_t i = M - 1; i >= 0; --i) {
                const double tq = ts[i];
                while (m >= 0 && t[m] > tq) {
                    const double tm = t[m], am = alpha[m];
                    if (any) { const double p = exp(-c * (t_last - tm)); s0 *= p; s1 *= p; }
                    double sn, cs;
                    sincos_cw(__dmul_rn(d, tm), &sn, &cs);
                    s0 = fma(a * cs + b * sn, am, s0); s1 = fma(a * sn - b * cs, am, s1);
                    t_last = tm; any = true; --m;
                }
                double contrib = 0.0;
                if (any && on) {
                    const double p = exp(-c * (t_last - tq));
                    double sn, cs;
                    sincos_cw(__dmul_rn(d, tq), &sn, &cs);
                    contrib = p * (cs * s0 + sn * s1);
                }
                const double tot = block_sum(contrib, red);
                if (threadIdx.x == 0) o[i] = (j0 == 0 ? 0.0 : o[i]) + tot;
            }
        }
    }
}

__global__ void add_halves_kernel(const int64_t M, const double *__restrict__ two, double *__restrict__ mu)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) mu[i] = two[i] + two[M + i];
}

}  // namespace

// scratch: 2 M doubles
cudaError_t launch_cond_mean(int64_t N, const double *t, int64_t M, const double *ts, int Jc,
                             const double *coef, const double *alpha, double *scratch, double *mu,
                             cudaStream_t stream)
{
    if (M == 0) return cudaSuccess;
    cond_mean_kernel<<<2, COND_THREADS, 0, stream>>>(N, t, M, ts, Jc, coef, alpha, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    add_halves_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, scratch, mu);
    return cudaGetLastError();
}

}  // namespace gf
