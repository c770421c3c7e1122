// Reference-order semiseparable scan for kernels WIDER than the register-resident scans take
// (GF_MAX_J = 176 < J <= GF_MAX_J_WIDE = 352): the reference's ``kernel + more terms``
// (gadfly/core.py:405-427) accepts any number of extra SHO terms.
//
// The symmetric J x J state of such a kernel (up to 990 8x8 tiles = 507 KB) fits neither the register
// file (253 tiles are the limit of scan_fast / scan_ref) nor shared memory, so it lives in a global
// scratch buffer that stays L2-resident (126 MB L2; 148 CTAs x 507 KB = 75 MB): one CTA per sequence,
// every thread walks its tiles (element-major layout: the 64 loads / stores of a tile are coalesced
// across the threads of a warp), the arithmetic and its order are scan_ref's (celerite2's recurrences,
// SURVEY.md A.6).  Correct and honest rather than fast: ~1 MB of L2 traffic per time step; wide
// kernels are the exception (the solar kernel plus up to 90 extra terms).
#include "common.cuh"

namespace gf {

namespace {

constexpr int JW_MAX = 352;                 // GF_MAX_J_WIDE
constexpr int NBW = JW_MAX / TILE;          // 44
constexpr int NTW = NBW * (NBW + 1) / 2;    // 990 tiles
constexpr int WT = 384;                     // threads: >= JW_MAX column threads

struct WideSmem {
    double u[JW_MAX], v[JW_MAX], p[JW_MAX], w[JW_MAX], dw[JW_MAX];
    double part[NBW][JW_MAX];
    double red[WT / 32][2];
    int next;
};

__device__ __forceinline__ double warp_sum_w(double x)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

__device__ __forceinline__ void block_sum2_w(double &a, double &b, double (*red)[2], int tid)
{
    a = warp_sum_w(a);
    b = warp_sum_w(b);
    if ((tid & 31) == 0) { red[tid >> 5][0] = a; red[tid >> 5][1] = b; }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < WT / 32; ++w) { sa += red[w][0]; sb += red[w][1]; }
    __syncthreads();
    a = sa; b = sb;
}

template <int MODE>
__global__ void __launch_bounds__(WT, 1) scan_wide_kernel(ScanArgs A, double *state)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmem &sm = *reinterpret_cast<WideSmem *>(smem_raw);
    const int tid = threadIdx.x;
    // this CTA's state: element e of tile g at S[e * NTW + g]
    double *S = state + (size_t)blockIdx.x * (size_t)NTW * 64;

    for (;;) {
        if (tid == 0) sm.next = atomicAdd(A.counter, 1);
        __syncthreads();
        const int item = sm.next;
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
        double sum_a = 0.0;                       // sum of a' in term order (same on every thread)
        for (int j = 0; j < Jc; ++j) sum_a += A.coef[4 * (j0 + j)];

        for (int g = tid; g < ntile; g += WT)
#pragma unroll 8
            for (int e = 0; e < 64; ++e) S[e * NTW + g] = 0.0;

        double wk = 0.0, Fk = 0.0;
        double dprev = 0.0, zprev = 0.0;
        double logdet = 0.0, quad = 0.0;
        int32_t fail = 0;
        double tprev = 0.0;

        if (N <= 0) {
            if (tid == 0) { A.logdet[b] = 0.0; if (A.quad) A.quad[b] = 0.0; A.status[b] = 0; }
            continue;
        }
        __syncthreads();

        for (int64_t n = 0; n < N; ++n) {
            const double tn = t[n];
            // ---- phase 1: row generation and O(J) state ---------------------------------
            if (tid < JW_MAX) {
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
                sm.u[tid] = u; sm.v[tid] = v; sm.p[tid] = p; sm.w[tid] = wk; sm.dw[tid] = dprev * wk;
            }
            __syncthreads();

            // ---- phase 2: S update + tmp = u S, tile by tile ----------------------------
            if (n > 0) {
                for (int g = tid; g < ntile; g += WT) {
                    int bi, bj;
                    tile_coords(g, nb, bi, bj);
                    double rowp[TILE], colp[TILE];
#pragma unroll
                    for (int e = 0; e < TILE; ++e) { rowp[e] = 0.0; colp[e] = 0.0; }
#pragma unroll
                    for (int i = 0; i < TILE; ++i) {
                        const double pi = sm.p[bi * TILE + i], dwi = sm.dw[bi * TILE + i], ui = sm.u[bi * TILE + i];
#pragma unroll
                        for (int j = 0; j < TILE; ++j) {
                            double *sp = &S[(i * TILE + j) * NTW + g];
                            const double s = (pi * (*sp + dwi * sm.w[bj * TILE + j])) * sm.p[bj * TILE + j];
                            *sp = s;
                            colp[j] += ui * s;                       // tmp_j += u_i S_ij
                            rowp[i] += s * sm.u[bj * TILE + j];      // tmp_i += S_ij u_j  (mirror element)
                        }
                    }
#pragma unroll
                    for (int e = 0; e < TILE; ++e) sm.part[bi][bj * TILE + e] = colp[e];
                    if (bi != bj) {
#pragma unroll
                        for (int e = 0; e < TILE; ++e) sm.part[bj][bi * TILE + e] = rowp[e];
                    }
                }
            }
            __syncthreads();

            // ---- phase 3: d_n, w_n, forward sweep ---------------------------------------
            double tmpk = 0.0, r1 = 0.0, r2 = 0.0, uk = 0.0, vk = 0.0;
            if (tid < JW_MAX) {
                uk = sm.u[tid]; vk = sm.v[tid];
                if (n > 0 && tid < nb * TILE) {
                    for (int s = 0; s < nb; ++s) tmpk += sm.part[s][tid];
                }
                r1 = tmpk * uk;
                r2 = uk * Fk;
            }
            block_sum2_w(r1, r2, sm.red, tid);
            const double an = ((dg ? dg[n] : 0.0) + ddiag) + sum_a;
            const double dn = an - r1;
            if (!(dn > 0.0)) { fail = (int32_t)(n + 1); break; }
            if (tid < JW_MAX) wk = (vk - tmpk) / dn;
            double zn;
            if (MODE == MODE_SAMPLE) {
                const double nrm = y ? y[n] : philox_normal(A.seed, A.seq0 + (uint64_t)b, (uint64_t)n);
                const double nu = nrm * sqrt(dn);
                zn = nu;
                if (tid == 0) A.out_x[n0 + n] = nu + r2;
            } else if (MODE == MODE_LOGLIKE) {
                zn = y[n] - r2;
            } else {
                zn = 0.0;
                if (tid == 0) A.out_x[n0 + n] = dn;
                if (A.out_W && colthread) {
                    const int col = (tid & 1) * Jc + (tid >> 1);     // celerite2's blocked column order
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

size_t scan_wide_state_bytes(int grid) { return (size_t)grid * (size_t)NTW * 64 * sizeof(double); }
bool scan_wide_supports(int jmax) { return jmax <= JW_MAX; }

cudaError_t launch_scan_wide(int mode, const ScanArgs &args, int grid, double *state, cudaStream_t stream)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(scan_wide_kernel<MODE_LOGLIKE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WideSmem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_wide_kernel<MODE_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WideSmem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(scan_wide_kernel<MODE_FACTOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WideSmem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    switch (mode) {
    case MODE_LOGLIKE: scan_wide_kernel<MODE_LOGLIKE><<<grid, WT, sizeof(WideSmem), stream>>>(args, state); break;
    case MODE_SAMPLE:  scan_wide_kernel<MODE_SAMPLE><<<grid, WT, sizeof(WideSmem), stream>>>(args, state); break;
    default:           scan_wide_kernel<MODE_FACTOR><<<grid, WT, sizeof(WideSmem), stream>>>(args, state); break;
    }
    return cudaGetLastError();
}

}  // namespace gf
