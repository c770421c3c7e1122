// Blocked semiseparable scan (K1 log-likelihood, K2 sample): KB = 4 time steps per hand-over.
//
// The k-step blocked recurrence of DESIGN.md section 5b in its bulk-synchronous form.  Same CTA as
// scan_fast.cu (this file includes it for the producer warp, the row ring, the lazy decay frames and
// the tile map): 8 matrix warps with the J x J state in registers, 3 "chain" warps, 1 producer warp.
// Per block of kk <= 4 consecutive rows of one decay frame:
//
//   M  (matrix warps)  rank-k update with the t~, w~ of the PREVIOUS block, frame change if the block
//                      starts one, then kk matrix-vector products g_i = u~_i S~ against the same state
//   C1 (all 352)       g-sums of the partial products, T0_i = v~_i - g_i
//   C2 (all 352)       ONE batch of dot products: q_i = g_i.u~_i, f_i = u~_i.F~, M_im = u~_i.T0_m (m < i)
//   C3 (chain warps)   the scalar k x k recursion, every lane redundantly:
//                        c_im = M_im - sum_{l<m} (c_ml / d_l) c_il
//                        d_i  = a_i - q_i - sum_{m<i} c_im^2 / d_m
//                        z_i  = y_i - f_i - sum_{m<i} (c_im / d_m) z_m      (sampling: x_i = nu_i + f_i + ...)
//   C4 (chain warps)   t~_i = T0_i - sum_{m<i} (c_im / d_m) t~_m,  w~_i = t~_i / d_i,  F~ += sum w~_i z_i,
//                      operands of the next block, outputs
//
// No reduction is on a serial path: it is block LDL^T on the k x k Schur complement, the arithmetic of
// the step-by-step recurrence up to summation order (tools/blocked_recurrence_check.py).
#define GF_SCAN_FAST_AS_HEADER
#include "scan_fast.cu"

#include <cstddef>

namespace gf {

namespace {

constexpr int KB = 4;                                     // rows per block
constexpr int BLK_THREADS = MAT_THREADS + CH_THREADS;     // everybody but the producer warp
constexpr int BAR_D = 10, BAR_A = 11, BAR_B = 12, BAR_C = 13;

struct BlkSmem {
    double P23[2][NSLOT][JP_MAX];      // partial sums of rows 2, 3 of a block (rows 0, 1: FastSmem::P)
    double2 U2[KB][4][NB_PAD];         // rows u~_i of the block: element pairs (8 b + 2 q, + 1) at [q][b]
    double2 T2[KB][4][NB_PAD];         // t~ of the previous block
    double2 W2[KB][4][NB_PAD];         // w~ of the previous block
    double Rb[JP_MAX];                 // frame change factors applied to the state at the block start
    double2 T0[KB][JC_MAX];            // v~_i - g_i per term
    double2 G[KB][JC_MAX];             // g_i per term
    double2 F[JC_MAX];                 // forward-substitution state, frame of the current block
    double dot[16];
    int kk, kprev, flag, pad;          // block descriptor, written by chain thread 0
};
static_assert(sizeof(FastSmem) % 16 == 0, "BlkSmem follows FastSmem in dynamic shared memory");
static_assert(sizeof(FastSmem) + sizeof(BlkSmem) <= 227 * 1024, "shared memory");

// byte offset of the partial sums of block row I relative to FastSmem::P[0]
constexpr int P_OFF_23 = (int)(sizeof(FastSmem) - offsetof(FastSmem, P)) + (int)offsetof(BlkSmem, P23);
template <int I>
__host__ __device__ constexpr int p_vec_off() { return I < 2 ? I * P_PAR_BYTES : P_OFF_23 + (I - 2) * P_PAR_BYTES; }

__device__ __forceinline__ int dot_id_q(int i) { return i; }
__device__ __forceinline__ int dot_id_f(int i) { return 4 + i; }
__device__ __forceinline__ int dot_id_m(int i, int m) { return 8 + i * (i - 1) / 2 + m; }

// ------------------------------------------------------------------------------------------
// matrix side
// ------------------------------------------------------------------------------------------
// 16-byte store without the "memory" clobber of scan_fast.cu's sts_v2: the compiler may move the operand
// loads of the next row above the stores of this one (different arrays; the hand-over barriers are
// volatile asm statements and keep their order with respect to these)
template <int OFF>
__device__ __forceinline__ void sts_v2_free(const uint32_t addr, const double x, const double y)
{
    asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "d"(x), "d"(y));
}

template <int M>
__device__ __forceinline__ void blk_update(double (&S)[TILE][TILE], const double2 *tb, const double2 *wc)
{
    double t[TILE], w[TILE];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double2 a = tb[(M * 4 + q) * NB_PAD], c = wc[(M * 4 + q) * NB_PAD];
        t[2 * q] = a.x; t[2 * q + 1] = a.y; w[2 * q] = c.x; w[2 * q + 1] = c.y;
    }
#pragma unroll
    for (int i = 0; i < TILE; ++i)
#pragma unroll
        for (int j = 0; j < TILE; ++j) S[i][j] = fma(t[i], w[j], S[i][j]);
}

template <int I>
__device__ __forceinline__ void blk_load_row(const double2 *ub, const double2 *uc, double (&ur)[TILE], double (&uj)[TILE])
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double2 a = ub[(I * 4 + q) * NB_PAD], c = uc[(I * 4 + q) * NB_PAD];
        ur[2 * q] = a.x; ur[2 * q + 1] = a.y; uj[2 * q] = c.x; uj[2 * q + 1] = c.y;
    }
}

// One matrix-vector product of the block.  The operands of the NEXT row are fetched before this row's
// exchange and stores: ptxas never moves a shared-memory load above an earlier store it cannot
// disambiguate, so without this every row would start with an exposed load latency.
template <int I>
__device__ __forceinline__ void blk_matvec(const double (&S)[TILE][TILE], const MatConst &mc, const double2 *ub,
                                           const double2 *uc, double (&ur)[TILE], double (&uj)[TILE], const bool more)
{
    double rowp[TILE], colp[TILE];
#pragma unroll
    for (int j = 0; j < TILE; ++j) colp[j] = ur[0] * S[0][j];
#pragma unroll
    for (int i = 1; i < TILE; ++i)
#pragma unroll
        for (int j = 0; j < TILE; ++j) colp[j] = fma(ur[i], S[i][j], colp[j]);
#pragma unroll
    for (int i = 0; i < TILE; ++i) rowp[i] = S[i][0] * uj[0];
#pragma unroll
    for (int j = 1; j < TILE; ++j)
#pragma unroll
        for (int i = 0; i < TILE; ++i) rowp[i] = fma(S[i][j], uj[j], rowp[i]);
    if (I + 1 < KB && more) blk_load_row<(I + 1 < KB) ? I + 1 : I>(ub, uc, ur, uj);

    // 2x2 group exchange (unconditional: a warp mixes tile kinds) and stores: the layout of
    // scan_fast.cu's matrix_phase
    constexpr int OFF = p_vec_off<I>();
    const bool hr = geom_mr(mc.geom) != 0, hc = geom_mc(mc.geom) != 0;
    double rs[4], cs[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double send_r = hr ? rowp[q] : rowp[4 + q];
        const double keep_r = hr ? rowp[4 + q] : rowp[q];
        rs[q] = keep_r + shfl_xor_d(send_r, 1);
        const double send_c = hc ? colp[q] : colp[4 + q];
        const double keep_c = hc ? colp[4 + q] : colp[q];
        cs[q] = keep_c + shfl_xor_d(send_c, 2);
    }
    const int kind = geom_kind(mc.geom);
    if (kind == 0) {
        sts_v2_free<OFF>(mc.pr0, rs[0], rs[1]);
        sts_v2_free<OFF>(mc.pr0 ^ 16u, rs[2], rs[3]);
        sts_v2_free<OFF>(mc.pc0, cs[0], cs[1]);
        sts_v2_free<OFF>(mc.pc0 ^ 16u, cs[2], cs[3]);
    } else if (kind == 1) {
        // diagonal tile: its (symmetric) contribution is stored once
#pragma unroll
        for (int q = 0; q < 4; ++q) sts_v2_free<OFF>(mc.pc0 ^ (16u * q), colp[2 * q], colp[2 * q + 1]);
    } else if (kind == 2) {
        // off-diagonal tile of a diagonal super-block: both contributions, no partner lanes
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sts_v2_free<OFF>(mc.pr0 ^ (16u * q), rowp[2 * q], rowp[2 * q + 1]);
            sts_v2_free<OFF>(mc.pc0 ^ (16u * q), colp[2 * q], colp[2 * q + 1]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// C1 / C2: data-parallel over the 352 non-producer threads, x = 0 .. 351
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void blk_c1(FastSmem &sm, BlkSmem &ex, const int x, const int n0, const int kk, const int Jc)
{
    for (int item = x; item < KB * JC_MAX; item += BLK_THREADS) {
        const int i = item / JC_MAX, term = item - i * JC_MAX;
        if (i >= kk || term >= Jc) continue;
        const double2 *Pp = (i < 2) ? reinterpret_cast<const double2 *>(&sm.P[i][0][0])
                                    : reinterpret_cast<const double2 *>(&ex.P23[i - 2][0][0]);
        const int tixA = pchunk(term, 0), tixB = pchunk(term, 1);
        double gc = 0.0, gs = 0.0;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const double2 v = Pp[s * (JP_MAX / 2) + ((s & 1) ? tixB : tixA)];
            gc += v.x; gs += v.y;
        }
        const double2 vn = sm.RV[(n0 + i) & (RR - 1)][term];
        ex.G[i][term] = make_double2(gc, gs);
        ex.T0[i][term] = make_double2(vn.x - gc, vn.y - gs);
    }
}

__device__ __forceinline__ void blk_c2(FastSmem &sm, BlkSmem &ex, const int x, const int n0, const int kk, const int Jc)
{
    // one dot product per half-warp: 14 of the 22 half-warps, all in flight at once
    const int id = x >> 4, hl = x & 15;
    int i = 0, m = 0, kind = 0;
    if (id < 4) { kind = 0; i = id; }
    else if (id < 8) { kind = 1; i = id - 4; }
    else if (id < 14) { kind = 2; const int r = id - 8; i = (r < 1) ? 1 : (r < 3) ? 2 : 3; m = r - i * (i - 1) / 2; }
    const bool on = id < 14 && i < kk;
    double acc = 0.0;
    if (on) {
        const double2 *U = sm.RU[(n0 + i) & (RR - 1)];
        const double2 *Y = (kind == 0) ? ex.G[i] : (kind == 1) ? ex.F : ex.T0[m];
        for (int t = hl; t < Jc; t += 16) {
            const double2 u = U[t], y = Y[t];
            acc = fma(u.x, y.x, acc);
            acc = fma(u.y, y.y, acc);
        }
    }
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) acc += shfl_xor_d(acc, off);
    if (on && hl == 0) ex.dot[id] = acc;
}

__device__ __forceinline__ void blk_matrix_loop(FastSmem &sm, BlkSmem &ex, const int mt, const int nsb, const int Jc)
{
    const MatConst mc = make_mat_const(sm, make_tile_map(mt, nsb));
    const int bi = geom_bi(mc.geom), bj = geom_bj(mc.geom);
    const double2 *ub = &ex.U2[0][0][bi], *uc = &ex.U2[0][0][bj];
    const double2 *tb = &ex.T2[0][0][bi], *wc = &ex.W2[0][0][bj];
    double S[TILE][TILE];
#pragma unroll
    for (int i = 0; i < TILE; ++i)
#pragma unroll
        for (int j = 0; j < TILE; ++j) S[i][j] = 0.0;
    int n0 = 0;
#ifdef GF_BLK_TIMING
    long long tm[6] = {0, 0, 0, 0, 0, 0};
    long long tc = clock64();
#define GF_TICK(k) { const long long now_ = clock64(); tm[k] += now_ - tc; tc = now_; }
#else
#define GF_TICK(k)
#endif
    for (;;) {
        bar_sync(BAR_D, BLK_THREADS);
        GF_TICK(0)
        const int kk = ex.kk, kprev = ex.kprev, flag = ex.flag;
        if (kk == 0) break;
        if (kprev > 0) blk_update<0>(S, tb, wc);
        if (kprev > 1) blk_update<1>(S, tb, wc);
        if (kprev > 2) blk_update<2>(S, tb, wc);
        if (kprev > 3) blk_update<3>(S, tb, wc);
#ifdef GF_BLK_TIMING
        { double chk = S[0][0] + S[7][7] + S[3][4]; asm volatile("" :: "d"(chk)); }
        GF_TICK(5)
#endif
        if (flag) {
#pragma unroll
            for (int i = 0; i < TILE; ++i) {
                const double ri = ex.Rb[bi * TILE + i];
#pragma unroll
                for (int j = 0; j < TILE; ++j) S[i][j] = (ri * S[i][j]) * ex.Rb[bj * TILE + j];
            }
        }
        {
            double ur[TILE], uj[TILE];
            blk_load_row<0>(ub, uc, ur, uj);
            blk_matvec<0>(S, mc, ub, uc, ur, uj, kk > 1);
            if (kk > 1) blk_matvec<1>(S, mc, ub, uc, ur, uj, kk > 2);
            if (kk > 2) blk_matvec<2>(S, mc, ub, uc, ur, uj, kk > 3);
            if (kk > 3) blk_matvec<3>(S, mc, ub, uc, ur, uj, false);
        }
        GF_TICK(1)
        bar_sync(BAR_A, BLK_THREADS);
        GF_TICK(2)
        blk_c1(sm, ex, mt, n0, kk, Jc);
        bar_sync(BAR_B, BLK_THREADS);
        GF_TICK(3)
        blk_c2(sm, ex, mt, n0, kk, Jc);
        bar_sync(BAR_C, BLK_THREADS);
        GF_TICK(4)
        n0 += kk;
    }
#ifdef GF_BLK_TIMING
    if (blockIdx.x == 0 && (mt & 31) == 0)
        printf("matrix warp %d, cycles per row: wait D %.0f | updates %.0f | products %.0f | wait A %.0f | C1 + B %.0f | C2 + C %.0f\n", mt >> 5,
               (double)tm[0] / n0, (double)tm[5] / n0, (double)tm[1] / n0, (double)tm[2] / n0, (double)tm[3] / n0, (double)tm[4] / n0);
#endif
}

// rows of the block that starts at n1: up to KB rows of one ring half, none but the first changes the frame
__device__ __forceinline__ int blk_rows(const FastSmem &sm, const int n1, const int N)
{
    int k = 1;
    while (k < KB && n1 + k < N && ((n1 + k) & (HALF - 1)) != 0 && !sm.Rflag[(n1 + k) & (RR - 1)]) ++k;
    return k;
}

template <int MODE>
__device__ __forceinline__ void blk_chain_loop(FastSmem &sm, BlkSmem &ex, const ScanArgs &A, const int ht,
                                               const int b, const int N, const int Jc)
{
    const int x = MAT_THREADS + ht;
    const int term = ht;
    const bool act = term < Jc;
    const int tq = term & 3, tb = term >> 2;          // element pair / block of this term's two columns
    const long long n0g = A.n_off[b];
    const int nh = (int)ring_halves(N);

    bar_sync(BAR_FULL + 0, N_RING);                   // ring half 0: rows 0..7
    int n0 = 0;
    {
        const int kk0 = blk_rows(sm, 0, N);
        if (act) {
#pragma unroll
            for (int i = 0; i < KB; ++i)
                if (i < kk0) ex.U2[i][tq][tb] = sm.RU[i][term];
        }
        if (ht == 0) { ex.kk = kk0; ex.kprev = 0; ex.flag = 0; }
    }
    double logsum = 0.0, prod = 1.0, quad = 0.0;
    int esum = 0;
    int32_t fail = 0;

    for (;;) {
        bar_sync(BAR_D, BLK_THREADS);
        const int kk = ex.kk;
        if (kk == 0) break;
        bar_sync(BAR_A, BLK_THREADS);
        blk_c1(sm, ex, x, n0, kk, Jc);
        bar_sync(BAR_B, BLK_THREADS);
        blk_c2(sm, ex, x, n0, kk, Jc);
        bar_sync(BAR_C, BLK_THREADS);

        // ---- C3: scalar recursion (every lane) ---------------------------------------------
        double e[KB][KB], cc[KB][KB], rd[KB], dd[KB], zz[KB], xx[KB];
        int bad = -1;
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            rd[i] = 0.0; dd[i] = 1.0; zz[i] = 0.0; xx[i] = 0.0;
#pragma unroll
            for (int m = 0; m < KB; ++m) { e[i][m] = 0.0; cc[i][m] = 0.0; }
            if (i < kk) {
                const int slot = (n0 + i) & (RR - 1);
#pragma unroll
                for (int m = 0; m < i; ++m) {
                    double c = ex.dot[dot_id_m(i, m)];
#pragma unroll
                    for (int l = 0; l < m; ++l) c = fma(-e[m][l], cc[i][l], c);
                    cc[i][m] = c;
                }
                double acc = ex.dot[dot_id_q(i)];
#pragma unroll
                for (int m = 0; m < i; ++m) {
                    e[i][m] = cc[i][m] * rd[m];
                    acc = fma(e[i][m], cc[i][m], acc);
                }
                dd[i] = sm.Ra[slot] - acc;
                if (!(dd[i] > 0.0) && bad < 0) bad = i;
                rd[i] = fast_rcp(dd[i]);
                const double fi = ex.dot[dot_id_f(i)];
                double corr = fi;
#pragma unroll
                for (int m = 0; m < i; ++m) corr = fma(e[i][m], zz[m], corr);
                if (MODE == MODE_LOGLIKE) {
                    zz[i] = sm.Ry[slot] - corr;
                } else {
                    zz[i] = sm.Ry[slot] * sqrt(dd[i]);
                    xx[i] = zz[i] + corr;
                }
            }
        }
        const int good = (bad < 0) ? kk : bad;        // rows of this block with a positive pivot

        // ---- ring: rows n0 .. n0 + kk - 1 are touched, the next block's rows must be there ------
        for (int n = n0; n < n0 + kk; ++n)
            if (((n + 2) & (HALF - 1)) == 0 && (n + 2) / HALF < nh)
                bar_sync(BAR_FULL + (((n + 2) / HALF) & 1), N_RING);

        // ---- C4: vectors, next block ----------------------------------------------------------
        const int n1 = n0 + kk;
        const bool more = (bad < 0) && n1 < N;
        int kk1 = 0, flag1 = 0;
        if (more) { kk1 = blk_rows(sm, n1, N); flag1 = sm.Rflag[n1 & (RR - 1)]; }
        if (act) {
            double2 tt[KB], ww[KB];
            double2 F = ex.F[term];
#pragma unroll
            for (int i = 0; i < KB; ++i) {
                tt[i] = make_double2(0.0, 0.0); ww[i] = make_double2(0.0, 0.0);
                if (i < kk) {
                    double2 T = ex.T0[i][term];
#pragma unroll
                    for (int m = 0; m < i; ++m) { T.x = fma(-e[i][m], tt[m].x, T.x); T.y = fma(-e[i][m], tt[m].y, T.y); }
                    tt[i] = T;
                    ww[i] = make_double2(T.x * rd[i], T.y * rd[i]);
                    F.x = fma(ww[i].x, zz[i], F.x);
                    F.y = fma(ww[i].y, zz[i], F.y);
                }
            }
            const double r1 = (more && flag1) ? sm.Rr[n1 & (RR - 1)][term] : 1.0;
#pragma unroll
            for (int i = 0; i < KB; ++i) {
                ex.T2[i][tq][tb] = tt[i];
                ex.W2[i][tq][tb] = ww[i];
                if (i < kk1) ex.U2[i][tq][tb] = sm.RU[(n1 + i) & (RR - 1)][term];
            }
            ex.Rb[2 * term] = r1; ex.Rb[2 * term + 1] = r1;
            ex.F[term] = make_double2(F.x * r1, F.y * r1);
        }
        if (ht == 0) {
#pragma unroll
            for (int i = 0; i < KB; ++i)
                if (i < good) {
                    logdet_push(dd[i], prod, esum);
                    if (MODE == MODE_LOGLIKE) quad = fma(zz[i] * zz[i], rd[i], quad);
                    else A.out_x[n0g + n0 + i] = xx[i];
                }
            logsum += log(prod); prod = 1.0;
            ex.kk = more ? kk1 : 0; ex.kprev = kk; ex.flag = flag1;
            if (bad >= 0) sm.stop = n0 + bad;       // the producer keeps only the hand-shake going
        }
        if (bad >= 0) fail = n0 + bad + 1;
        // ring: the rows of this block are dead
        for (int n = n0; n < n0 + kk; ++n)
            if ((n & (HALF - 1)) == HALF - 1 && n / HALF + 2 < nh)
                bar_arrive(BAR_EMPTY + ((n / HALF) & 1), N_RING);
        n0 = n1;
    }
    // not positive definite: keep the ring hand-shake with the producer going until the natural end
    for (int n = n0; n < N; ++n) {
        if (((n + 2) & (HALF - 1)) == 0 && (n + 2) / HALF < nh)
            bar_sync(BAR_FULL + (((n + 2) / HALF) & 1), N_RING);
        if ((n & (HALF - 1)) == HALF - 1 && n / HALF + 2 < nh)
            bar_arrive(BAR_EMPTY + ((n / HALF) & 1), N_RING);
    }
    if (ht == 0) {
        A.logdet[b] = logdet_total(logsum, 1.0, esum);
        if (MODE == MODE_LOGLIKE && A.quad) A.quad[b] = quad;
        A.status[b] = fail;
    }
}

template <int MODE>
__global__ void __launch_bounds__(FT_THREADS, 1) scan_blk_kernel(ScanArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw);
    BlkSmem &ex = *reinterpret_cast<BlkSmem *>(smem_raw + sizeof(FastSmem));
    const int tid = threadIdx.x;
    // per-sequence reset of the blocked buffers (padding columns must read as zero, Rb as one)
    auto reset = [&](const int xx) {
        double *z = reinterpret_cast<double *>(&ex);
        const int nz = (int)(offsetof(BlkSmem, dot) / sizeof(double));
        for (int i = xx; i < nz; i += BLK_THREADS) z[i] = 0.0;
        for (int i = xx; i < JP_MAX; i += BLK_THREADS) ex.Rb[i] = 1.0;
    };
    if (tid < MAT_THREADS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_MAT));
        SeqInfo q;
        while (next_sequence(sm, A, tid, q)) {
            if (q.N > 0) {
                reset(tid);
                bar_sync(BAR_A, BLK_THREADS);
                blk_matrix_loop(sm, ex, tid, q.nsb, q.Jc);
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_HLP));
        const int ht = tid - MAT_THREADS;
        SeqInfo q;
        if (ht < CH_THREADS) {
            while (next_sequence(sm, A, tid, q)) {
                if (q.N > 0) {
                    reset(tid);
                    bar_sync(BAR_A, BLK_THREADS);
                    blk_chain_loop<MODE>(sm, ex, A, ht, q.b, q.N, q.Jc);
                } else if (ht == 0) {
                    A.logdet[q.b] = 0.0;
                    if (A.quad) A.quad[q.b] = 0.0;
                    A.status[q.b] = 0;
                }
            }
        } else {
            while (next_sequence(sm, A, tid, q)) {
                if (q.N > 0) producer_loop<MODE>(sm, A, ht - CH_THREADS, q.b, q.N, q.Jc);
            }
        }
    }
}

template <int MODE>
cudaError_t launch_blk_mode(const ScanArgs &args, int grid, cudaStream_t stream)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int bytes = (int)(sizeof(FastSmem) + sizeof(BlkSmem));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(scan_blk_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    scan_blk_kernel<MODE><<<grid, FT_THREADS, bytes, stream>>>(args);
    return cudaGetLastError();
}

}  // namespace

bool scan_blk_supports(int mode, int jmax) { return mode != MODE_FACTOR && jmax <= JP_MAX; }

cudaError_t launch_scan_blk(int mode, const ScanArgs &args, int sm_count, cudaStream_t stream)
{
    const int grid = (int)(args.B < sm_count ? args.B : sm_count);
    return mode == MODE_LOGLIKE ? launch_blk_mode<MODE_LOGLIKE>(args, grid, stream)
                                : launch_blk_mode<MODE_SAMPLE>(args, grid, stream);
}

}  // namespace gf
