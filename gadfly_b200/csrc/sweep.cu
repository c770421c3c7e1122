// K4: O(N J) sweeps over a stored factor W (celerite2 driver.solve_lower / matmul_lower /
// solve_upper / matmul_upper; reference call sites gadfly/gp.py:327,350,370).
//
//   lower, n = 0..N-1:   F <- p_n o (F + W_{n-1} prev),   z_n = y_n -/+ U_n . F
//   upper, n = N-1..0:   F <- p_n o (F + U_{n+1} prev),   z_n = y_n -/+ W_n . F
// with prev = z (solve) or y (matmul) of the neighbouring step and p_n the decay over the
// step.  U rows are regenerated from (t, coef); W is read once, coalesced.
//
// One CTA per sequence.  The dependency chain of a sweep is one FMA pair, one dot product and its
// reduction per step; regenerating the U rows (sincos, exp: ~85 FP64 instructions per term and
// step) in the same threads would triple the step.  So SW2_PROD (15) producer warps generate the rows
// of the coming steps into a shared-memory ring (each warp a step of its own, three terms per
// lane), and ONE chain warp holds the whole state F (three terms per lane): the dot product is a
// single warp butterfly -- no cross-warp stage, no __syncthreads on the chain.  Hand-over by
// mbarriers per ring half (8 steps): full[h] (one arrival per step's producer warp), empty[h]
// (one arrival of the chain).
#include "common.cuh"

namespace gf {

namespace {

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

#ifndef GF_SW2_PROD
#define GF_SW2_PROD 15
#endif
constexpr int SW2_PROD = GF_SW2_PROD;            // producer warps
constexpr int SW2_THREADS = 32 * (1 + SW2_PROD);
constexpr int SW2_RS = 16;                        // ring slots (two halves)
// A producer warp takes every SW2_PROD-th step.  Its mbarrier waits name a phase by parity only,
// which is unambiguous as long as the warp never asks for a phase two ahead of the barrier:
// consecutive steps of a warp must be at most two ring halves apart.
static_assert(SW2_PROD <= SW2_RS, "producer stride must not exceed the ring");
constexpr int SW2_HALF = 8;
// terms per lane: 3 (3 x 32 >= GF_MAX_J / 2) -- or 6 for the wide kernels (6 x 32 >= GF_MAX_J_WIDE / 2),
// a separate instantiation so that the ordinary widths keep their code
template <int SW2_TPL>
struct Sweep2Smem {
    static constexpr int SW2_JC = 32 * SW2_TPL;
    double2 dot[SW2_RS][SW2_JC];     // row the state is read through at this step (U_n lower, W_n upper)
    double2 upd[SW2_RS][SW2_JC];     // row that enters the state after this step (W_n lower, U_n upper)
    double dec[SW2_RS][SW2_JC];      // decay of the state over the step into this one
    double yv[SW2_RS];
    unsigned long long full[2], empty[2];
};

__device__ __forceinline__ void sw_mbar_init(unsigned long long *mb, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void sw_mbar_inval(unsigned long long *mb)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(mb)) : "memory");
}
__device__ __forceinline__ void sw_mbar_arrive(unsigned long long *mb)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(mb)) : "memory");
}
__device__ __forceinline__ void sw_mbar_wait(unsigned long long *mb, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SW_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni SW_MBAR_DONE;\n"
        "bra.uni SW_MBAR_WAIT;\n"
        "SW_MBAR_DONE:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(mb)), "r"(parity) : "memory");
}

template <bool UPPER, bool SOLVE, int SW2_TPL>
__global__ void __launch_bounds__(SW2_THREADS)
sweep2_kernel(int64_t B, const int64_t *__restrict__ n_off, const int64_t *__restrict__ t_off,
              const int64_t *__restrict__ j_off, const int64_t *__restrict__ w_off,
              const double *__restrict__ t_all, const double *__restrict__ coef,
              const double *__restrict__ W_all, const double *Y_all, double *Z_all)
{
    extern __shared__ __align__(16) unsigned char sw2_raw[];
    Sweep2Smem<SW2_TPL> &sm = *reinterpret_cast<Sweep2Smem<SW2_TPL> *>(sw2_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

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
        __syncthreads();                           // the ring of the previous sequence is dead
        if (tid == 0) {
            if (b != (int64_t)blockIdx.x) {        // not the first sequence of this CTA
                sw_mbar_inval(&sm.full[0]); sw_mbar_inval(&sm.full[1]);
                sw_mbar_inval(&sm.empty[0]); sw_mbar_inval(&sm.empty[1]);
            }
            sw_mbar_init(&sm.full[0], SW2_HALF); sw_mbar_init(&sm.full[1], SW2_HALF);
            sw_mbar_init(&sm.empty[0], 1);       sw_mbar_init(&sm.empty[1], 1);
        }
        __syncthreads();
        const int64_t n_half = (N + SW2_HALF - 1) / SW2_HALF;       // ring halves to go through

        if (warp == 0) {
            // ---------------- chain ----------------
            double Fc[SW2_TPL], Fs[SW2_TPL], gc[SW2_TPL], gs[SW2_TPL];
#pragma unroll
            for (int k = 0; k < SW2_TPL; ++k) { Fc[k] = Fs[k] = gc[k] = gs[k] = 0.0; }
            double prev = 0.0;
            for (int64_t hr = 0; hr < n_half; ++hr) {
                const int h = (int)(hr & 1);
                sw_mbar_wait(&sm.full[h], (uint32_t)((hr >> 1) & 1));
                const int cnt = (int)((N - hr * SW2_HALF < SW2_HALF) ? (N - hr * SW2_HALF) : SW2_HALF);
                auto step = [&](const int q) {
                    const int slot = h * SW2_HALF + q;
                    const int64_t m = hr * SW2_HALF + q;
                    double part0 = 0.0, part1 = 0.0;
#pragma unroll
                    for (int k = 0; k < SW2_TPL; ++k) {
                        const int j = lane + 32 * k;
                        const double p = sm.dec[slot][j];
                        const double2 d = sm.dot[slot][j];
                        // F <- p o (F + row_{m-1} prev_{m-1})   (p = 0 at the first step)
                        Fc[k] = p * fma(gc[k], prev, Fc[k]);
                        Fs[k] = p * fma(gs[k], prev, Fs[k]);
                        part0 = fma(d.x, Fc[k], part0);
                        part1 = fma(d.y, Fs[k], part1);
                        const double2 u = sm.upd[slot][j];
                        gc[k] = u.x; gs[k] = u.y;
                    }
                    const double acc = warp_sum(part0 + part1);
                    const double yn = sm.yv[slot];
                    const double zn = SOLVE ? (yn - acc) : (yn + acc);
                    const int64_t n = UPPER ? (N - 1 - m) : m;
                    if (lane == 0) Z[n] = zn;
                    prev = SOLVE ? zn : yn;
                };
                if (cnt == SW2_HALF) {
                    // unrolled: the ring reads of the next steps (and, for the matmul ops, their whole
                    // butterflies) are scheduled under the reduction of the current one
#pragma unroll
                    for (int q = 0; q < SW2_HALF; ++q) step(q);
                } else {
                    for (int q = 0; q < cnt; ++q) step(q);
                }
                __syncwarp();
                if (lane == 0) sw_mbar_arrive(&sm.empty[h]);
            }
        } else {
            // ---------------- producers: warp w takes the steps m = w - 1 (mod SW2_PROD) ----------------
            double ca[SW2_TPL], cb[SW2_TPL], cc[SW2_TPL], cd[SW2_TPL];
#pragma unroll
            for (int k = 0; k < SW2_TPL; ++k) {
                const int j = lane + 32 * k;
                ca[k] = cb[k] = cc[k] = cd[k] = 0.0;
                if (j < Jc) {
                    const double *cf = coef + 4 * (j0 + j);
                    ca[k] = cf[0]; cb[k] = cf[1]; cc[k] = cf[2]; cd[k] = cf[3];
                }
            }
            const int64_t m_end = n_half * SW2_HALF;     // dummy arrivals complete the last half
            // inputs of a step are fetched one step of this warp ahead (the W rows come from HBM:
            // their latency would otherwise sit in front of every step)
            double tn_n = 0.0, dt_n = 0.0, yv_n = 0.0, wc_n[SW2_TPL], ws_n[SW2_TPL];
            auto fetch = [&](const int64_t m) {
                if (m < N) {
                    const int64_t n = UPPER ? (N - 1 - m) : m;
                    tn_n = t[n];
                    dt_n = (m == 0) ? 0.0 : (UPPER ? (tn_n - t[n + 1]) : (t[n - 1] - tn_n));
                    yv_n = Y[n];
#pragma unroll
                    for (int k = 0; k < SW2_TPL; ++k) {
                        const int j = lane + 32 * k;
                        wc_n[k] = (j < Jc) ? W[n * J + j] : 0.0;
                        ws_n[k] = (j < Jc) ? W[n * J + Jc + j] : 0.0;
                    }
                }
            };
            fetch(warp - 1);
            for (int64_t m = warp - 1; m < m_end; m += SW2_PROD) {
                const int64_t hr = m / SW2_HALF;
                const int h = (int)(hr & 1);
                const int slot = (int)(m % SW2_RS);
                const double tn = tn_n, dt = dt_n, yv = yv_n;   // decay exponent dt <= 0 (previous step of the sweep)
                double wc[SW2_TPL], ws[SW2_TPL];
#pragma unroll
                for (int k = 0; k < SW2_TPL; ++k) { wc[k] = wc_n[k]; ws[k] = ws_n[k]; }
                fetch(m + SW2_PROD);
                if (hr >= 2) sw_mbar_wait(&sm.empty[h], (uint32_t)(((hr >> 1) - 1) & 1));
                if (m < N) {
#pragma unroll
                    for (int k = 0; k < SW2_TPL; ++k) {
                        const int j = lane + 32 * k;
                        const bool on = j < Jc;
                        double sn, cs;
                        sincos_cw(__dmul_rn(cd[k], tn), &sn, &cs);
                        const double uc = ca[k] * cs + cb[k] * sn;
                        const double us = ca[k] * sn - cb[k] * cs;
                        sm.dot[slot][j] = UPPER ? make_double2(wc[k], ws[k]) : make_double2(on ? uc : 0.0, on ? us : 0.0);
                        sm.upd[slot][j] = UPPER ? make_double2(on ? uc : 0.0, on ? us : 0.0) : make_double2(wc[k], ws[k]);
                        sm.dec[slot][j] = (m == 0) ? 0.0 : exp(cc[k] * dt);
                    }
                    if (lane == 0) sm.yv[slot] = yv;
                }
                __syncwarp();
                if (lane == 0) sw_mbar_arrive(&sm.full[h]);
            }
        }
    }
}

template <int TPL>
static cudaError_t launch_sweep_tpl(int op, int64_t B, const int64_t *n_off, const int64_t *t_off,
                                    const int64_t *j_off, const int64_t *w_off, const double *t,
                                    const double *coef, const double *W, const double *Y, double *Z,
                                    cudaStream_t stream)
{
    const int grid = (int)(B < 65535 ? B : 65535);
    static bool configured[64] = {};      // per device: a process may hold handles on several
    int dev = 0;
    cudaGetDevice(&dev);
    const int bytes = (int)sizeof(Sweep2Smem<TPL>);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sweep2_kernel<false, true, TPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sweep2_kernel<false, false, TPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sweep2_kernel<true, true, TPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(sweep2_kernel<true, false, TPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    switch (op) {
    case 0: sweep2_kernel<false, true, TPL><<<grid, SW2_THREADS, bytes, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    case 1: sweep2_kernel<false, false, TPL><<<grid, SW2_THREADS, bytes, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    case 2: sweep2_kernel<true, true, TPL><<<grid, SW2_THREADS, bytes, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    default: sweep2_kernel<true, false, TPL><<<grid, SW2_THREADS, bytes, stream>>>(B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z); break;
    }
    return cudaGetLastError();
}

}  // namespace

// jc_max: most complex terms of any sequence of the batch
cudaError_t launch_sweep(int op, int64_t B, const int64_t *n_off, const int64_t *t_off,
                         const int64_t *j_off, const int64_t *w_off, const double *t,
                         const double *coef, const double *W, const double *Y, double *Z,
                         int jc_max, cudaStream_t stream)
{
    if (jc_max <= 96) return launch_sweep_tpl<3>(op, B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z, stream);
    return launch_sweep_tpl<6>(op, B, n_off, t_off, j_off, w_off, t, coef, W, Y, Z, stream);
}

}  // namespace gf
