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
// overlap with it; per step the solve chain is one FMA pair, a warp-shuffle sum and one barrier.
// The matmul ops have no feedback from the outputs into the state: their CH dot products per
// chunk are reduced together (one halving butterfly, one barrier per chunk).
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
    __shared__ double s_chunk[2][SW_WARPS][CH];     // matmul: the CH sums of a chunk per warp
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
                yy[k] = (ok && (tid == 0 || !SOLVE)) ? Y[n] : 0.0;    // matmul: every thread needs y
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                double sn, cs;
                sincos_cw(cd * tt[k], &sn, &cs);
                uc[k] = ca * cs + cb * sn;
                us[k] = ca * sn - cb * cs;
            }
            // decay over each step of the chunk (UPPER: t_n - t_{n+1}; lower: t_{n-1} - t_n; both <= 0):
            // off the chain as well
            double pk[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const double tp = (k == 0) ? t_nb : tt[k - 1];
                pk[k] = (base + k > 0 && base + k < N) ? exp(cc * (UPPER ? (tt[k] - tp) : (tp - tt[k]))) : 0.0;
            }
            if (!SOLVE) {
                // matmul: the state does not depend on the outputs, so the CH dot products of the
                // chunk are independent -- one multi-value butterfly and one barrier per chunk
                double part[CH];
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    if (base + k > 0) {
                        const double pr = (k == 0) ? prev : yy[k - 1];
                        const double a0 = (k == 0) ? (UPPER ? uc_nb : wc_nb) : (UPPER ? uc[k - 1] : wc[k - 1]);
                        const double a1 = (k == 0) ? (UPPER ? us_nb : ws_nb) : (UPPER ? us[k - 1] : ws[k - 1]);
                        Fc = pk[k] * fma(a0, pr, Fc);
                        Fs = pk[k] * fma(a1, pr, Fs);
                    }
                    part[k] = UPPER ? (wc[k] * Fc + ws[k] * Fs) : (uc[k] * Fc + us[k] * Fs);
                }
                // halving butterfly: after the stages 16, 8, 4 every lane holds one of the CH = 8 sums
                // (index = lane bits 4..2), then two plain stages
                {
                    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
                    double q4[4], q2[2], q1;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double send = h16 ? part[i] : part[4 + i], keep = h16 ? part[4 + i] : part[i];
                        q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double send = h8 ? q4[i] : q4[2 + i], keep = h8 ? q4[2 + i] : q4[i];
                        q2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    {
                        const double send = h4 ? q2[0] : q2[1], keep = h4 ? q2[1] : q2[0];
                        q1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
                    q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
                    // value index held by this lane: bit 2 of k from lane bit 4, bit 1 from bit 3, bit 0 from bit 2
                    const int kidx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                    const int par = (int)((base / CH) & 1);
                    if ((lane & 3) == 0) s_chunk[par][warp][kidx] = q1;
                    __syncthreads();
                    if (tid < CH && base + tid < N) {
                        double acc = 0.0;
#pragma unroll
                        for (int w = 0; w < SW_WARPS; ++w) acc += s_chunk[par][w][tid];
                        const int64_t n = UPPER ? (N - 1 - (base + tid)) : (base + tid);
                        Z[n] = Y[n] + acc;
                    }
                }
                // carry the neighbour row into the next chunk
                const int last = (int)((N - base < CH ? N - base : CH) - 1);
#pragma unroll
                for (int k = 0; k < CH; ++k)
                    if (k == last) {
                        prev = yy[k];
                        uc_nb = uc[k]; us_nb = us[k]; wc_nb = wc[k]; ws_nb = ws[k]; t_nb = tt[k];
                    }
            } else {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int64_t m = base + k;
                if (m >= N) break;
                const int64_t n = UPPER ? (N - 1 - m) : m;
                const int par = (int)(m & 1);
                if (m > 0) {
                    const double p = pk[k];
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
                const double zn = yn - acc;
                if (tid == 0) Z[n] = zn;
                prev = zn;
                uc_nb = uc[k]; us_nb = us[k]; wc_nb = wc[k]; ws_nb = ws[k]; t_nb = tt[k];
            }
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
