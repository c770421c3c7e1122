// K1 / K2 / K3 for narrow kernels (J <= 32: up to 16 SHO terms -- granulation-only kernels, single
// terms, the J = 10 option of BASELINE configs[3]): ONE WARP PER SEQUENCE.
//
// The big scan (scan_fast.cu) gives a whole SM to one sequence and its step time (~1900 cycles) is
// set by synchronisation latency, not by J: a J = 10 kernel would cost what J = 172 costs.  Here
// lane k owns column k of the symmetric state S (JT doubles in registers), so the matrix-vector
// product tmp_k = sum_i u_i S_ik needs no reduction at all, and the only cross-lane traffic per
// step is the broadcast of the row (u_i, p_i, d w_i) through a per-warp shared-memory line and two
// warp butterflies (pivot, forward substitution).  Many warps per SM hide each other's latency.
// Arithmetic in the order of the celerite2 recurrences, as scan_ref.cu:
//     S <- diag(p) (S + d_{n-1} w_{n-1}^T w_{n-1}) diag(p);  tmp = u_n S;
//     d_n = a_n - tmp.u_n;  w_n = (v_n - tmp) / d_n;  F <- p o (F + w_{n-1} z_{n-1});  z_n = y_n - u_n.F
// Column order: [cos block | sin block], k < Jc: cos column of term k, else sin column of term k - Jc.
#include "common.cuh"

namespace gf {

namespace {

constexpr int SS_JMAX = 32;
// CTA shape per tile width: 2 x 8 warps per SM at <= 128 registers for JT <= 16, 1 x 12 warps at
// <= 168 registers for JT = 32 (the 32-double column does not fit 128 registers without spills)
// (3 x 8 warps at <= 80 registers spills and is 27 % slower -- measured)
template <int JT> struct SmallShape { static constexpr int warps = (JT > 16) ? 12 : 8, ctas = (JT > 16) ? 1 : 2; };

__device__ __forceinline__ double ss_warp_sum(double x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

template <int MODE, int JT>
__global__ void __launch_bounds__(32 * SmallShape<JT>::warps, SmallShape<JT>::ctas) scan_small_kernel(ScanArgs A)
{
    constexpr int SS_WARPS = SmallShape<JT>::warps;
    __shared__ double2 s_up[SS_WARPS][SS_JMAX];     // (u_i, p_i) of the current row
    __shared__ double s_dw[SS_WARPS][SS_JMAX];      // d_{n-1} w_{n-1,i}
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(A.counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= A.B) break;
        const int b = A.order[item];
        const int64_t n0 = A.n_off[b];
        const int64_t N = A.n_off[b + 1] - n0;
        const int64_t j0 = A.j_off[b];
        const int Jc = (int)(A.j_off[b + 1] - j0);
        const int J = 2 * Jc;
        const double *t = A.t + A.t_off[b];
        const long long y0 = A.y_like_t ? A.t_off[b] : n0;
        const double *y = A.y ? A.y + y0 : nullptr;
        const double *dg = A.diag ? A.diag + y0 : nullptr;
        const double ddiag = A.ddiag[b];
        const bool on = lane < J;
        const bool is_sin = lane >= Jc;
        double ca = 0, cb = 0, cc = 0, cd = 0;
        if (on) {
            const double *cf = A.coef + 4 * (j0 + (is_sin ? lane - Jc : lane));
            ca = cf[0]; cb = cf[1]; cc = cf[2]; cd = cf[3];
        }
        double sum_a = 0.0;                       // sum of a' in term order (same on every lane)
        for (int j = 0; j < Jc; ++j) sum_a += A.coef[4 * (j0 + j)];

        double S[JT];
#pragma unroll
        for (int i = 0; i < JT; ++i) S[i] = 0.0;
        double wk = 0.0, Fk = 0.0, dprev = 0.0, zprev = 0.0, tprev = 0.0;
        double logdet = 0.0, prod = 1.0, quad = 0.0;
        int esum8 = 0;
        long long esum = 0;
        int32_t fail = 0;

        for (int64_t base = 0; base < N && fail == 0; base += 32) {
            // a chunk of 32 steps: every lane fetches the inputs of one of them (coalesced)
            const int64_t m = base + lane;
            const bool have = m < N;
            const double tl = have ? t[m] : 0.0;
            double yl = 0.0;
            if (have) {
                if (y) yl = y[m];
                else if (MODE == MODE_SAMPLE) yl = philox_normal(A.seed, A.seq0 + (uint64_t)b, (uint64_t)m);
            }
            const double dl = (have && dg) ? dg[m] : 0.0;
            double xl = 0.0;                       // this lane's output of the chunk (sample mode)
            const int cnt = (int)((N - base < 32) ? (N - base) : 32);
            for (int q = 0; q < cnt; ++q) {
                const int64_t n = base + q;
                const double tn = __shfl_sync(0xffffffffu, tl, q);
                const double yn = __shfl_sync(0xffffffffu, yl, q);
                const double dgn = __shfl_sync(0xffffffffu, dl, q);
                // ---- row n of U, V, the decay over the step, forward-substitution state -----
                double u = 0.0, v = 0.0, p = 1.0;
                if (on) {
                    double sn, cs;
                    sincos_cw(__dmul_rn(cd, tn), &sn, &cs);
                    if (is_sin) { u = ca * sn - cb * cs; v = sn; }
                    else        { u = ca * cs + cb * sn; v = cs; }
                    if (n > 0) {
                        p = exp(cc * (tprev - tn));
                        Fk = p * (Fk + wk * zprev);
                    }
                }
                s_up[warp][lane] = make_double2(u, p);
                s_dw[warp][lane] = dprev * wk;
                __syncwarp();
                // ---- S update and tmp = u S for this lane's column ---------------------------
                double tmp = 0.0;
                if (n > 0) {
                    double tmp1 = 0.0;
#pragma unroll
                    for (int i = 0; i < JT; i += 2) {
                        const double2 r0 = s_up[warp][i], r1 = s_up[warp][i + 1];
                        const double s0 = (r0.y * (S[i] + s_dw[warp][i] * wk)) * p;
                        const double s1 = (r1.y * (S[i + 1] + s_dw[warp][i + 1] * wk)) * p;
                        S[i] = s0; S[i + 1] = s1;
                        tmp += r0.x * s0;
                        tmp1 += r1.x * s1;
                    }
                    tmp += tmp1;
                }
                __syncwarp();                      // the row line is free for the next step
                // ---- pivot, new row of W, outputs ----------------------------------------------
                const double r1 = ss_warp_sum(tmp * u);
                const double r2 = ss_warp_sum(u * Fk);
                const double an = (dgn + ddiag) + sum_a;
                const double dn = an - r1;
                if (!(dn > 0.0)) { fail = (int32_t)(n + 1); break; }
                wk = on ? (v - tmp) / dn : 0.0;
                double zn;
                if (MODE == MODE_SAMPLE) {
                    zn = yn * sqrt(dn);
                    if (lane == q) xl = zn + r2;
                } else if (MODE == MODE_LOGLIKE) {
                    zn = yn - r2;
                    quad += zn * zn / dn;
                } else {
                    // factor: d_n and the row of W (celerite2's blocked column order is this kernel's)
                    zn = 0.0;
                    if (lane == q) xl = dn;
                    if (A.out_W && on) A.out_W[A.w_off[b] + n * (int64_t)J + lane] = wk;
                }
                logdet_push(dn, prod, esum8);
                if ((q & 7) == 7) { logdet += log(prod); prod = 1.0; esum += esum8; esum8 = 0; }
                dprev = dn; zprev = zn; tprev = tn;
            }
            if (MODE != MODE_LOGLIKE && have) A.out_x[n0 + m] = xl;    // (rows after a failure: unspecified)
        }
        if (lane == 0) {
            A.logdet[b] = logdet_total(logdet, prod, esum + esum8);
            if (MODE == MODE_LOGLIKE && A.quad) A.quad[b] = quad;
            A.status[b] = fail;
        }
    }
}

template <int MODE, int JT>
cudaError_t launch_small_shape(const ScanArgs &args, int sm_count, cudaStream_t stream)
{
    // persistent warps pull sequences from the queue: at most the resident CTAs, at least enough warps
    constexpr int warps = SmallShape<JT>::warps, ctas = SmallShape<JT>::ctas;
    const int64_t want = (args.B + warps - 1) / warps;
    const int grid = (int)(want < (int64_t)ctas * sm_count ? want : (int64_t)ctas * sm_count);
    scan_small_kernel<MODE, JT><<<grid, 32 * warps, 0, stream>>>(args);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_small_mode(const ScanArgs &args, int jmax, int sm_count, cudaStream_t stream)
{
    if (jmax <= 8)  return launch_small_shape<MODE, 8>(args, sm_count, stream);
    if (jmax <= 16) return launch_small_shape<MODE, 16>(args, sm_count, stream);
    return launch_small_shape<MODE, 32>(args, sm_count, stream);
}

}  // namespace

bool scan_small_supports(int mode, int jmax)
{
    (void)mode;
    return jmax <= SS_JMAX;
}

cudaError_t launch_scan_small(int mode, const ScanArgs &args, int jmax, int sm_count,
                              cudaStream_t stream, int *launches)
{
    *launches = 1;
    if (mode == MODE_LOGLIKE) return launch_small_mode<MODE_LOGLIKE>(args, jmax, sm_count, stream);
    if (mode == MODE_SAMPLE) return launch_small_mode<MODE_SAMPLE>(args, jmax, sm_count, stream);
    return launch_small_mode<MODE_FACTOR>(args, jmax, sm_count, stream);
}

}  // namespace gf
