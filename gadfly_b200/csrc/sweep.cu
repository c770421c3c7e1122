// K4: O(N J) sweeps over a stored factor W (celerite2 driver.solve_lower / matmul_lower /
// solve_upper / matmul_upper; reference call sites gadfly/gp.py:327,350,370).
//
//   lower, n = 0..N-1:   F <- p_n o (F + W_{n-1} prev),   z_n = y_n -/+ U_n . F
//   upper, n = N-1..0:   F <- p_n o (F + U_{n+1} prev),   z_n = y_n -/+ W_n . F
// with prev = z (solve) or y (matmul) of the neighbouring step and p_n the decay over the
// step.  U rows are regenerated from (t, coef); W is read once, coalesced.
//
// One CTA per sequence, one thread per complex term (cos and sin columns).  The sweep is a
// length-N dependency chain, so the kernel works in chunks of CH steps: the chunk's rows
// (sincos, exp) and the W / t / y loads of the NEXT chunk are independent of the chain and
// overlap with it; per step the chain is one FMA pair, a warp-shuffle sum and one barrier.
#include "common.cuh"

namespace gf {

namespace {

constexpr int SW_THREADS = 96;   // >= GF_MAX_J / 2 complex terms
constexpr int SW_WARPS = SW_THREADS / 32;
constexpr int CH = 8;

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

template <bool UPPER, bool SOLVE>
__global__ void __launch_bounds__(SW_THREADS)
sweep_kernel(int64_t B, const int64_t *__restrict__ n_off, const int64_t *__restrict__ t_off,
             const int64_t *__restrict__ j_off, const int64_t *__restrict__ w_off,
             const double *__restrict__ t_all, const double *__restrict__ coef,
             const double *__restrict__ W_all, const double *Y_all, double *Z_all)
{
    __shared__ double s_red[2][SW_WARPS];
    __shared__ double s_y[2];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        const int64_t n0 = n_off[b];
        const int64_t N = n_off[b + 1] - n0;
        const int64_t j0 = j_off[b];
        const int Jc = (int)(j_off[b + 1] - j0);
        const int J = 2 * Jc;
        const double *t = t_all + t_off[b];
        const double *W = W_all + w_off[b];
        const double *Y = Y_all + n0;
        double *Z = Z_all + n0;
        const bool act = tid < Jc;
        double ca = 0, cb = 0, cc = 0, cd = 0;
        if (act) {
            const double *cf = coef + 4 * (j0 + tid);
            ca = cf[0]; cb = cf[1]; cc = cf[2]; cd = cf[3];
        }
        double Fc = 0.0, Fs = 0.0;      // this term's two entries of F
        double prev = 0.0;              // z (solve) or y (matmul) of the neighbouring step
        double uc_nb = 0.0, us_nb = 0.0, wc_nb = 0.0, ws_nb = 0.0, t_nb = 0.0;  // neighbour row
        __syncthreads();

        for (int64_t base = 0; base < N; base += CH) {
            // rows of this chunk: independent of the chain
            double tt[CH], uc[CH], us[CH], wc[CH], ws[CH], yy[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int64_t m = base + k;
                const int64_t n = UPPER ? (N - 1 - m) : m;
                const bool ok = m < N;
                tt[k] = ok ? t[n] : 0.0;
                wc[k] = (ok && act) ? W[n * J + tid] : 0.0;
                ws[k] = (ok && act) ? W[n * J + Jc + tid] : 0.0;
                yy[k] = (ok && tid == 0) ? Y[n] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                double sn, cs;
                sincos(cd * tt[k], &sn, &cs);
                uc[k] = ca * cs + cb * sn;
                us[k] = ca * sn - cb * cs;
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int64_t m = base + k;
                if (m >= N) break;
                const int64_t n = UPPER ? (N - 1 - m) : m;
                const int par = (int)(m & 1);
                if (m > 0) {
                    // UPPER: t_n - t_{n+1}; lower: t_{n-1} - t_n  (both <= 0)
                    const double p = exp(cc * (UPPER ? (tt[k] - t_nb) : (t_nb - tt[k])));
                    if (UPPER) { Fc = p * (Fc + uc_nb * prev); Fs = p * (Fs + us_nb * prev); }
                    else       { Fc = p * (Fc + wc_nb * prev); Fs = p * (Fs + ws_nb * prev); }
                }
                double part = UPPER ? (wc[k] * Fc + ws[k] * Fs) : (uc[k] * Fc + us[k] * Fs);
                part = warp_sum(part);
                if (lane == 0) s_red[par][warp] = part;
                if (tid == 0) s_y[par] = yy[k];
                __syncthreads();
                double acc = 0.0;
#pragma unroll
                for (int w = 0; w < SW_WARPS; ++w) acc += s_red[par][w];
                const double yn = s_y[par];
                const double zn = SOLVE ? (yn - acc) : (yn + acc);
                if (tid == 0) Z[n] = zn;
                prev = SOLVE ? zn : yn;
                uc_nb = uc[k]; us_nb = us[k]; wc_nb = wc[k]; ws_nb = ws[k]; t_nb = tt[k];
            }
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_sweep(int op, int64_t B, const int64_t *n_off, const int64_t *t_off,
                         const int64_t *j_off, const int64_t *w_off, const double *t,
                         const double *coef, const double *W, const double *Y, double *Z,
                         cudaStream_t stream)
{
    const int grid = (int)(B < 65535 ? B : 65535);
    switch (op) {
    case 0: sweep_kernel<false, true><<<grid, SW_THREADS, 0, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    case 1: sweep_kernel<false, false><<<grid, SW_THREADS, 0, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    case 2: sweep_kernel<true, true><<<grid, SW_THREADS, 0, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    default: sweep_kernel<true, false><<<grid, SW_THREADS, 0, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    }
    return cudaGetLastError();
}

}  // namespace gf
