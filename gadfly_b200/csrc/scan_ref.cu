// Reference-order semiseparable scan (validation kernel, GF_FLAG_REFERENCE_ORDER).
//
// One CTA per sequence, J x J state S held as 8x8 register tiles of its upper triangle
// (one tile per thread), plain __syncthreads() phases, arithmetic in the order of the
// celerite2 recurrences:
//     S <- diag(p) (S + d_{n-1} w_{n-1}^T w_{n-1}) diag(p);  tmp = u_n S;
//     d_n = a_n - tmp.u_n;  w_n = (v_n - tmp) / d_n
// The U/V rows are generated on the fly from t and (a', b', c, d); nothing of size N*J is
// read.  This kernel is the simple cross-check for the warp-specialised fast scan
// (scan_fast.cu); it shares the entry points and the batching/work-queue conventions.
#include "common.cuh"

namespace gf {

namespace {

constexpr int REF_THREADS = 256;

__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

// Sum (a, b) over the CTA; every thread returns the identical totals.
__device__ __forceinline__ void block_sum2(double &a, double &b, double (*red)[2], int tid)
{
    a = warp_sum(a);
    b = warp_sum(b);
    if ((tid & 31) == 0) { red[tid >> 5][0] = a; red[tid >> 5][1] = b; }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < REF_THREADS / 32; ++w) { sa += red[w][0]; sb += red[w][1]; }
    __syncthreads();
    a = sa; b = sb;
}

template <int MODE>
__global__ void __launch_bounds__(REF_THREADS, 1) scan_ref_kernel(ScanArgs A)
{
    __shared__ double s_u[JP_MAX], s_v[JP_MAX], s_p[JP_MAX], s_w[JP_MAX], s_dw[JP_MAX];
    __shared__ double s_part[NB_MAX][JP_MAX];
    __shared__ double s_red[REF_THREADS / 32][2];
    __shared__ int s_next;

    const int tid = threadIdx.x;

    for (;;) {
        if (tid == 0) s_next = atomicAdd(A.counter, 1);
        __syncthreads();
        const int item = s_next;
        __syncthreads();
        if (item >= A.B) break;
        const int b = A.order[item];

        const int64_t n0 = A.n_off[b];
        const int64_t N = A.n_off[b + 1] - n0;
        const int64_t j0 = A.j_off[b];
        const int Jc = (int)(A.j_off[b + 1] - j0);
        const int J = 2 * Jc;
        const int nb = (J + TILE - 1) / TILE;
        const int ntile = nb * (nb + 1) / 2;
        const double *t = A.t + A.t_off[b];
        const long long y0 = A.y_like_t ? A.t_off[b] : n0;
        const double *y = A.y ? A.y + y0 : nullptr;
        const double *dg = A.diag ? A.diag + y0 : nullptr;
        const double ddiag = A.ddiag[b];

        // column k = 2*term + s  (s = 0: cos column, s = 1: sin column)
        const bool colthread = tid < J;
        double ca = 0, cb = 0, cc = 0, cd = 0;
        if (colthread) {
            const double *cf = A.coef + 4 * (j0 + (tid >> 1));
            ca = cf[0]; cb = cf[1]; cc = cf[2]; cd = cf[3];
        }
        // sum of a' in term order (same on every thread)
        double sum_a = 0.0;
        for (int j = 0; j < Jc; ++j) sum_a += A.coef[4 * (j0 + j)];

        int bi = 0, bj = 0;
        const bool tilethread = tid < ntile;
        if (tilethread) tile_coords(tid, nb, bi, bj);

        double S[TILE][TILE];
#pragma unroll
        for (int i = 0; i < TILE; ++i)
#pragma unroll
            for (int j = 0; j < TILE; ++j) S[i][j] = 0.0;

        double wk = 0.0, Fk = 0.0;       // this column's w_{n-1}, F
        double dprev = 0.0, zprev = 0.0;
        double logdet = 0.0, quad = 0.0;
        int32_t fail = 0;
        double tprev = 0.0;

        if (N <= 0) {
            if (tid == 0) { A.logdet[b] = 0.0; if (A.quad) A.quad[b] = 0.0; A.status[b] = 0; }
            continue;
        }

        for (int64_t n = 0; n < N; ++n) {
            const double tn = t[n];
            // ---- phase 1: row generation and O(J) state ---------------------------------
            if (tid < JP_MAX) {
                double u = 0.0, v = 0.0, p = 1.0;
                if (colthread) {
                    double sn, cs;
                    sincos(__dmul_rn(cd, tn), &sn, &cs);
                    if (tid & 1) { u = ca * sn - cb * cs; v = sn; }
                    else         { u = ca * cs + cb * sn; v = cs; }
                    if (n > 0) {
                        p = exp(cc * (tprev - tn));
                        Fk = p * (Fk + wk * zprev);
                    }
                }
                s_u[tid] = u; s_v[tid] = v; s_p[tid] = p; s_w[tid] = wk; s_dw[tid] = dprev * wk;
            }
            __syncthreads();

            // ---- phase 2: S update + tmp = u S (tile threads) ---------------------------
            if (n > 0 && tilethread) {
                double pi_[TILE], dwi[TILE], ui[TILE], pj[TILE], wj[TILE], uj[TILE];
#pragma unroll
                for (int e = 0; e < TILE; ++e) {
                    pi_[e] = s_p[bi * TILE + e]; dwi[e] = s_dw[bi * TILE + e]; ui[e] = s_u[bi * TILE + e];
                    pj[e] = s_p[bj * TILE + e];  wj[e] = s_w[bj * TILE + e];   uj[e] = s_u[bj * TILE + e];
                }
                double rowp[TILE], colp[TILE];
#pragma unroll
                for (int e = 0; e < TILE; ++e) { rowp[e] = 0.0; colp[e] = 0.0; }
#pragma unroll
                for (int i = 0; i < TILE; ++i)
#pragma unroll
                    for (int j = 0; j < TILE; ++j) {
                        double s = (pi_[i] * (S[i][j] + dwi[i] * wj[j])) * pj[j];
                        S[i][j] = s;
                        colp[j] += ui[i] * s;     // tmp_j += u_i S_ij
                        rowp[i] += s * uj[j];     // tmp_i += S_ij u_j  (mirror element S_ji)
                    }
                // column block bj receives the column partials in slot bi; for an off-diagonal
                // tile, row block bi receives the row partials in slot bj
#pragma unroll
                for (int e = 0; e < TILE; ++e) s_part[bi][bj * TILE + e] = colp[e];
                if (bi != bj) {
#pragma unroll
                    for (int e = 0; e < TILE; ++e) s_part[bj][bi * TILE + e] = rowp[e];
                }
            }
            __syncthreads();

            // ---- phase 3: d_n, w_n, forward sweep ---------------------------------------
            double tmpk = 0.0, r1 = 0.0, r2 = 0.0, uk = 0.0, vk = 0.0;
            if (tid < JP_MAX) {
                uk = s_u[tid]; vk = s_v[tid];
                // columns beyond nb * TILE were never written by phase 2
                if (n > 0 && tid < nb * TILE) {
                    for (int s = 0; s < nb; ++s) tmpk += s_part[s][tid];
                }
                r1 = tmpk * uk;
                r2 = uk * Fk;
            }
            block_sum2(r1, r2, s_red, tid);
            const double an = ((dg ? dg[n] : 0.0) + ddiag) + sum_a;
            const double dn = an - r1;
            if (!(dn > 0.0)) { fail = (int32_t)(n + 1); break; }
            if (tid < JP_MAX) wk = (vk - tmpk) / dn;
            double zn;
            if (MODE == MODE_SAMPLE) {
                double nrm = y ? y[n] : philox_normal(A.seed, A.seq0 + (uint64_t)b, (uint64_t)n);
                double nu = nrm * sqrt(dn);
                zn = nu;
                if (tid == 0) A.out_x[n0 + n] = nu + r2;
            } else if (MODE == MODE_LOGLIKE) {
                zn = y[n] - r2;
            } else {
                zn = 0.0;
                if (tid == 0) A.out_x[n0 + n] = dn;
                if (A.out_W && colthread) {
                    // celerite2 blocked column order [cos-block | sin-block]
                    int col = (tid & 1) * Jc + (tid >> 1);
                    A.out_W[A.w_off[b] + n * (int64_t)J + col] = wk;
                }
            }
            if (tid == 0) {
                logdet += log(dn);
                if (MODE == MODE_LOGLIKE) quad += zn * zn / dn;
            }
            dprev = dn; zprev = zn; tprev = tn;
        }
        if (tid == 0) {
            A.logdet[b] = logdet;
            if (A.quad) A.quad[b] = quad;
            A.status[b] = fail;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_scan_ref(int mode, const ScanArgs &args, int grid, cudaStream_t stream)
{
    switch (mode) {
    case MODE_LOGLIKE: scan_ref_kernel<MODE_LOGLIKE><<<grid, REF_THREADS, 0, stream>>>(args); break;
    case MODE_SAMPLE:  scan_ref_kernel<MODE_SAMPLE><<<grid, REF_THREADS, 0, stream>>>(args); break;
    default:           scan_ref_kernel<MODE_FACTOR><<<grid, REF_THREADS, 0, stream>>>(args); break;
    }
    return cudaGetLastError();
}

}  // namespace gf
