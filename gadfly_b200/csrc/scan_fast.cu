// Fast semiseparable scan (K1 log-likelihood, K2 sample, K3 factor) for sm_100a, FP64.
//
// One CTA per sequence (persistent CTAs pull sequences from a queue), 1 CTA per SM, three
// warpgroups with re-balanced register budgets (setmaxnreg):
//
//   WG0, WG1  "matrix" threads.  The J x J symmetric state lives in REGISTERS: each thread owns
//             one 8x8 tile of the upper triangle (253 tiles at J = 176).  Per time step and per
//             stored element: one FMA for the rank-1 update and two FMAs for the two matrix-
//             vector partial products the symmetric tile contributes to (3 DFMA per element;
//             the FP64 pipe is the binding resource).  Tiles are grouped 2x2 per 4 lanes so
//             that half of the partial sums are combined with warp shuffles and the operand
//             vectors are read as conflict-free / broadcast LDS.128.
//   WG2       "vector" threads, one lane per complex term (cos and sin columns).  They generate
//             the rows u_n, v_n on the fly from t (sincos / exp, nothing of size N*J is read),
//             finish the matrix-vector product, form the pivot d_n, the new row w_n, the
//             forward-substitution state F and the outputs, and publish the operands of the
//             next matrix phase.
//
// Two algebraic rearrangements of the celerite recurrences make this fast;
// both are exact in exact arithmetic and differ from the reference order only in rounding:
//
//  (1) Lazy decay.  With q_n = exp(-c (t_n - t_ref)) and S = diag(q) S~ diag(q), the update
//      S <- P (S + d w w^T) P becomes a pure accumulation S~ += d w~ w~^T with w~ = w / q,
//      u~ = u q, v~ = v / q: the two multiplications per element per step disappear.  The
//      reference time is moved ("renormalisation": S~ <- r r^T o S~, r = exp(-c dt) <= 1)
//      every RENORM_STEPS steps or when c_max (t_n - t_ref) would exceed RENORM_LIMIT, so all
//      scaled quantities stay far inside the FP64 range; large gaps just drive r -> 0.
//  (2) One-step-stale matrix-vector product.  h_n = u~_n S~(n) is evaluated as
//      g_n + d_{n-1} (u~_n . w~_{n-1}) w~_{n-1} with g_n = u~_n S~(n-1), so matrix phase n needs
//      only w~_{n-2}: the vector work of step n-1 overlaps matrix phase n instead of
//      serialising with it.  Synchronisation is by named barriers (producer bar.arrive,
//      consumer bar.sync), double-buffered operands and partial sums.
#include "common.cuh"

namespace gf {

namespace {

constexpr int FT_THREADS = 384;
constexpr int MAT_THREADS = 256;
constexpr int HLP_THREADS = 128;
constexpr int HLP_WARPS = 4;
constexpr int TPW = 22;              // complex terms per vector warp (4 x 22 = 88 >= JP_MAX / 2)
constexpr int NSB_MAX = NB_MAX / 2;  // 16 x 16 super-blocks per side
constexpr int NSLOT = NSB_MAX + 1;   // partial-sum slots per column
constexpr int RING = 64;             // staged t / y / diag entries
constexpr int CHUNK = 32;
constexpr int REG_MAT = 208;
constexpr int REG_HLP = 88;
constexpr double RENORM_LIMIT = 64.0;
constexpr int RENORM_STEPS = 64;

constexpr int BAR_OPS = 1;    // ids 1, 2: operands of matrix phase (n & 1) are ready
constexpr int BAR_PART = 3;   // ids 3, 4: partial sums of matrix phase (n & 1) are ready
constexpr int BAR_HLP = 5;    // vector-warp internal

struct FastSmem {
    double2 A[2][TILE][NB_MAX];     // (u~_n[k], d w~[k]) for k = 8 b + e, indexed [e][b]
    double2 C[2][TILE][NB_MAX];     // (u~_n[k], w~[k])
    double R[2][JP_MAX];            // renormalisation factors r[k] of the phase
    double P[2][NSLOT][JP_MAX];     // partial sums of g_n, natural column order
    double2 red2[2][HLP_WARPS];     // (beta, gamma) per vector warp
    double red1[2][HLP_WARPS];      // alpha per vector warp
    double tbuf[RING], ybuf[RING], dbuf[RING];
    int renorm[2];
    long long stop;                 // first matrix phase that must not run
    int next;
};

__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ double shfl_xor_d(double x, int m)
{
    return __shfl_xor_sync(0xffffffffu, x, m);
}

struct Row {
    double uc, us, vc, vs;   // u~, v~ of the cos and sin column
    double r;                // frame change factor into this step's frame (1 if none)
    double q;                // exp(-c (t - t_ref)) of this step (1 on a renormalising step)
    int flag;                // this step renormalises
};

struct RowGen {
    double ca, cb, cc, cd;   // a', b', c, d of this lane's term
    double cmax;             // largest c of the sequence
    double t_ref;
    long long m_ref;

    __device__ __forceinline__ Row make(double tm, long long m, bool valid)
    {
        Row row;
        row.uc = row.us = row.vc = row.vs = 0.0;
        row.r = 1.0; row.q = 1.0; row.flag = 0;
        if (!valid) return row;
        const double dt = tm - t_ref;
        const bool rn = (cmax * dt > RENORM_LIMIT) || (m - m_ref >= RENORM_STEPS);
        const double E = cc * dt;
        double q = 1.0, qinv = 1.0;
        if (rn) {
            row.r = exp(-E);
            row.flag = 1;
            t_ref = tm;
            m_ref = m;
        } else {
            q = exp(-E);
            qinv = exp(E);
        }
        row.q = q;
        double sn, cs;
        sincos(cd * tm, &sn, &cs);
        row.uc = (ca * cs + cb * sn) * q;
        row.us = (ca * sn - cb * cs) * q;
        row.vc = cs * qinv;
        row.vs = sn * qinv;
        return row;
    }
};

// ------------------------------------------------------------------------------------------
// matrix warpgroups
// ------------------------------------------------------------------------------------------
struct TileMap {
    int bi, bj;        // block row / column of this thread's tile
    int kind;          // 0 off-diagonal 2x2 group lane, 1 diagonal tile, 2 off tile of a diagonal
                       // super-block, 3 idle
    int ri, cj;        // position inside the 2x2 group (kind 0)
    int slot_row, slot_col;
};

__device__ __forceinline__ TileMap make_tile_map(int mt, int nsb)
{
    TileMap m;
    m.bi = 0; m.bj = 0; m.kind = 3; m.ri = 0; m.cj = 0; m.slot_row = 0; m.slot_col = 0;
    const int n_off = 2 * nsb * (nsb - 1);
    if (mt < n_off) {
        int g = mt >> 2, l = mt & 3;
        int I = 0, rem = g;
        while (rem >= nsb - 1 - I) { rem -= nsb - 1 - I; ++I; }
        const int Jsb = I + 1 + rem;
        m.ri = l >> 1; m.cj = l & 1;
        m.bi = 2 * I + m.ri; m.bj = 2 * Jsb + m.cj;
        m.kind = 0;
        m.slot_row = Jsb;    // contribution to the columns of super-block I from (I, Jsb)
        m.slot_col = I;      // contribution to the columns of super-block Jsb from (I, Jsb)
    } else if (mt < n_off + 3 * nsb) {
        const int q = mt - n_off, I = q / 3, k = q - 3 * I;
        if (k == 0) { m.bi = m.bj = 2 * I; m.kind = 1; m.slot_col = I; }
        else if (k == 2) { m.bi = m.bj = 2 * I + 1; m.kind = 1; m.slot_col = I; }
        else { m.bi = 2 * I; m.bj = 2 * I + 1; m.kind = 2; m.slot_row = nsb; m.slot_col = nsb; }
    }
    return m;
}

__device__ __forceinline__ void matrix_loop(FastSmem &sm, const int mt, const int nsb,
                                            const long long N)
{
    const TileMap tm = make_tile_map(mt, nsb);
    const int bi = tm.bi, bj = tm.bj;
    double S[TILE][TILE];
#pragma unroll
    for (int i = 0; i < TILE; ++i)
#pragma unroll
        for (int j = 0; j < TILE; ++j) S[i][j] = 0.0;

    for (long long n = 0; n < N; ++n) {
        const int par = (int)(n & 1);
        bar_sync(BAR_OPS + par, FT_THREADS);
        if (n >= sm.stop) break;

        double uj[TILE], wj[TILE];
#pragma unroll
        for (int e = 0; e < TILE; ++e) {
            const double2 c = sm.C[par][e][bj];
            uj[e] = c.x; wj[e] = c.y;
        }
        if (sm.renorm[par]) {
            // frame change: S~ <- r r^T o (S~ + d w~ w~^T); the rank-1 term is consumed here
#pragma unroll
            for (int i = 0; i < TILE; ++i) {
                const double dwi = sm.A[par][i][bi].y;
                const double rgi = sm.R[par][bi * TILE + i];
#pragma unroll
                for (int j = 0; j < TILE; ++j) {
                    const double rgj = sm.R[par][bj * TILE + j];
                    S[i][j] = (rgi * fma(dwi, wj[j], S[i][j])) * rgj;
                }
            }
#pragma unroll
            for (int e = 0; e < TILE; ++e) wj[e] = 0.0;
        }

        double colp[TILE], rowp[TILE];
#pragma unroll
        for (int e = 0; e < TILE; ++e) { colp[e] = 0.0; rowp[e] = 0.0; }
#pragma unroll
        for (int i = 0; i < TILE; ++i) {
            const double2 a = sm.A[par][i][bi];
            const double ui = a.x, dwi = a.y;
            double rp = 0.0;
#pragma unroll
            for (int j = 0; j < TILE; ++j) {
                const double T = fma(dwi, wj[j], S[i][j]);
                S[i][j] = T;
                colp[j] = fma(ui, T, colp[j]);
                rp = fma(T, uj[j], rp);
            }
            rowp[i] = rp;
        }

        // 2x2 group: combine the two tiles of a block row (lane ^ 1) and of a block column
        // (lane ^ 2); each lane keeps four of the eight sums
        double rs[4], cs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double send_r = tm.cj ? rowp[q] : rowp[4 + q];
            const double keep_r = tm.cj ? rowp[4 + q] : rowp[q];
            rs[q] = keep_r + shfl_xor_d(send_r, 1);
            const double send_c = tm.ri ? colp[q] : colp[4 + q];
            const double keep_c = tm.ri ? colp[4 + q] : colp[q];
            cs[q] = keep_c + shfl_xor_d(send_c, 2);
        }
        if (tm.kind == 0) {
            double *pr = &sm.P[par][tm.slot_row][bi * TILE + tm.cj * 4];
            double *pc = &sm.P[par][tm.slot_col][bj * TILE + tm.ri * 4];
            *reinterpret_cast<double2 *>(pr) = make_double2(rs[0], rs[1]);
            *reinterpret_cast<double2 *>(pr + 2) = make_double2(rs[2], rs[3]);
            *reinterpret_cast<double2 *>(pc) = make_double2(cs[0], cs[1]);
            *reinterpret_cast<double2 *>(pc + 2) = make_double2(cs[2], cs[3]);
        } else if (tm.kind == 1) {
            double *pc = &sm.P[par][tm.slot_col][bj * TILE];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<double2 *>(pc + 2 * q) = make_double2(colp[2 * q], colp[2 * q + 1]);
        } else if (tm.kind == 2) {
            double *pr = &sm.P[par][tm.slot_row][bi * TILE];
            double *pc = &sm.P[par][tm.slot_col][bj * TILE];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                *reinterpret_cast<double2 *>(pr + 2 * q) = make_double2(rowp[2 * q], rowp[2 * q + 1]);
                *reinterpret_cast<double2 *>(pc + 2 * q) = make_double2(colp[2 * q], colp[2 * q + 1]);
            }
        }
        bar_arrive(BAR_PART + par, FT_THREADS);
    }
}

// ------------------------------------------------------------------------------------------
// vector warpgroup
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void hreduce2(FastSmem &sm, int par, int hw, int lane, double &a, double &b)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { a += shfl_xor_d(a, off); b += shfl_xor_d(b, off); }
    if (lane == 0) sm.red2[par][hw] = make_double2(a, b);
    bar_sync(BAR_HLP, HLP_THREADS);
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int w = 0; w < HLP_WARPS; ++w) { const double2 v = sm.red2[par][w]; sa += v.x; sb += v.y; }
    a = sa; b = sb;
}

__device__ __forceinline__ void hreduce1(FastSmem &sm, int par, int hw, int lane, double &a)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += shfl_xor_d(a, off);
    if (lane == 0) sm.red1[par][hw] = a;
    bar_sync(BAR_HLP, HLP_THREADS);
    double sa = 0.0;
#pragma unroll
    for (int w = 0; w < HLP_WARPS; ++w) sa += sm.red1[par][w];
    a = sa;
}

template <int MODE>
__device__ __forceinline__ void vector_loop(FastSmem &sm, const ScanArgs &A, const int ht,
                                            const int b, const int nsb, const long long N,
                                            const int Jc)
{
    const int hw = ht >> 5, lane = ht & 31;
    const int term = hw * TPW + lane;
    const bool act = (lane < TPW) && (term < Jc);
    const int k0 = 2 * term;                 // cos column; sin column is k0 + 1
    const int kb = k0 >> 3, ke = k0 & 7;     // both columns sit in block kb (ke is even)
    const long long n0 = A.n_off[b];
    const long long j0 = A.j_off[b];
    const double *t = A.t + A.t_off[b];
    const double *y = A.y ? A.y + n0 : nullptr;
    const double *dg = A.diag ? A.diag + n0 : nullptr;
    const double ddiag = A.ddiag[b];
    const int J = 2 * Jc;
    // sampling without external normals: the draws enter through the staging ring like y
    const bool philox = (MODE == MODE_SAMPLE) && (A.y == nullptr);
    const uint64_t seq = A.seq0 + (uint64_t)b;

    RowGen gen;
    gen.ca = gen.cb = gen.cc = gen.cd = 0.0;
    if (act) {
        const double *cf = A.coef + 4 * (j0 + term);
        gen.ca = cf[0]; gen.cb = cf[1]; gen.cc = cf[2]; gen.cd = cf[3];
    }
    // sum of a' in term order; largest decay rate
    double sum_a = 0.0, cmax = 0.0;
    for (int j = 0; j < Jc; ++j) {
        sum_a += A.coef[4 * (j0 + j)];
        cmax = fmax(cmax, A.coef[4 * (j0 + j) + 2]);
    }
    gen.cmax = cmax;

    // stage the first RING entries of t / y / diag
    {
        const int role = ht >> 5, l = ht & 31;
        for (int c = 0; c < RING; c += CHUNK) {
            const long long m = c + l;
            if (m < N) {
                if (role == 0) sm.tbuf[m] = t[m];
                if (role == 1) sm.ybuf[m] = y ? y[m] : (philox ? philox_normal(A.seed, seq, (uint64_t)m) : 0.0);
                if (role == 2) sm.dbuf[m] = dg ? dg[m] : 0.0;
            }
        }
    }
    bar_sync(BAR_HLP, HLP_THREADS);

    gen.t_ref = sm.tbuf[0];
    gen.m_ref = 0;
    Row r0 = gen.make(sm.tbuf[0], 0, act);
    Row r1 = gen.make(N > 1 ? sm.tbuf[1] : 0.0, 1, act && N > 1);
    Row r2 = gen.make(N > 2 ? sm.tbuf[2] : 0.0, 2, act && N > 2);
    // the renormalisation decision is uniform, but idle lanes skipped make(): recompute it
    // for them from lane 0 of the warp (all active lanes agree)
    r0.flag = __shfl_sync(0xffffffffu, r0.flag, 0);
    r1.flag = __shfl_sync(0xffffffffu, r1.flag, 0);
    r2.flag = __shfl_sync(0xffffffffu, r2.flag, 0);

    // operands of matrix phases 0 and 1: rows 0 / 1, no rank-1 term yet
    if (act) {
        sm.A[0][ke][kb] = make_double2(r0.uc, 0.0);     sm.C[0][ke][kb] = make_double2(r0.uc, 0.0);
        sm.A[0][ke + 1][kb] = make_double2(r0.us, 0.0); sm.C[0][ke + 1][kb] = make_double2(r0.us, 0.0);
        sm.A[1][ke][kb] = make_double2(r1.uc, 0.0);     sm.C[1][ke][kb] = make_double2(r1.uc, 0.0);
        sm.A[1][ke + 1][kb] = make_double2(r1.us, 0.0); sm.C[1][ke + 1][kb] = make_double2(r1.us, 0.0);
        sm.R[0][k0] = 1.0; sm.R[0][k0 + 1] = 1.0;
        sm.R[1][k0] = r1.r; sm.R[1][k0 + 1] = r1.r;
    }
    if (ht == 0) { sm.renorm[0] = 0; sm.renorm[1] = r1.flag; }
    bar_arrive(BAR_OPS + 0, FT_THREADS);
    if (N > 1) bar_arrive(BAR_OPS + 1, FT_THREADS);

    double wc = 0.0, ws = 0.0;       // w~_{n-1}, in the frame of step n
    double Fc = 0.0, Fs = 0.0;       // F~ of this term
    double kappa = 0.0;              // d_{n-1} (u~_n . w~_{n-1})
    double logdet = 0.0, prod = 1.0, quad = 0.0;
    int32_t fail = 0;
    double pend = 0.0;               // staged global load in flight

    for (long long n = 0; n < N; ++n) {
        const int par = (int)(n & 1);
        // staging of the input ring, one chunk ahead: issue at n % CHUNK == 0, store one step later
        {
            const int role = ht >> 5, l = ht & 31;
            const long long base = (n & ~(long long)(CHUNK - 1)) + CHUNK;
            const long long m = base + l;
            const int ph = (int)(n & (CHUNK - 1));
            {
                if (ph == 0 && base >= RING && m < N) {
                    if (role == 0) pend = t[m];
                    if (role == 1) pend = y ? y[m] : (philox ? philox_normal(A.seed, seq, (uint64_t)m) : 0.0);
                    if (role == 2) pend = dg ? dg[m] : 0.0;
                }
                if (ph == 1 && base >= RING && m < N) {
                    if (role == 0) sm.tbuf[m & (RING - 1)] = pend;
                    if (role == 1) sm.ybuf[m & (RING - 1)] = pend;
                    if (role == 2) sm.dbuf[m & (RING - 1)] = pend;
                }
            }
        }

        bar_sync(BAR_PART + par, FT_THREADS);
        double gc = 0.0, gs = 0.0;
        if (act) {
            for (int s = 0; s <= nsb; ++s) {
                const double2 v = *reinterpret_cast<const double2 *>(&sm.P[par][s][k0]);
                gc += v.x; gs += v.y;
            }
        }
        const double hc = fma(kappa, wc, gc), hs = fma(kappa, ws, gs);
        double beta = hc * r0.uc + hs * r0.us;
        double gamma = r0.uc * Fc + r0.us * Fs;
        hreduce2(sm, par, hw, lane, beta, gamma);

        const int slot = (int)(n & (RING - 1));
        const double an = (sm.dbuf[slot] + ddiag) + sum_a;
        const double dn = an - beta;
        if (!(dn > 0.0)) {
            // not positive definite: stop the matrix warps at phase n + 2 and absorb the
            // arrival of phase n + 1 (already released) so that the barriers end up balanced
            fail = (int32_t)(n + 1);
            if (ht == 0) sm.stop = n + 2;
            if (n + 2 < N) bar_arrive(BAR_OPS + par, FT_THREADS);
            if (n + 1 < N) bar_sync(BAR_PART + (par ^ 1), FT_THREADS);
            break;
        }
        const double rd = 1.0 / dn;
        double wcn = (r0.vc - hc) * rd, wsn = (r0.vs - hs) * rd;   // w~_n, frame of step n
        double zp;
        if (MODE == MODE_LOGLIKE) {
            const double zn = sm.ybuf[slot] - gamma;
            quad = fma(zn * zn, rd, quad);
            zp = zn;
        } else if (MODE == MODE_SAMPLE) {
            zp = sm.ybuf[slot] * sqrt(dn);
            if (ht == 0) A.out_x[n0 + n] = zp + gamma;
        } else {
            zp = 0.0;
            if (ht == 0) A.out_x[n0 + n] = dn;
            if (A.out_W && act) {
                double *Wn = A.out_W + A.w_off[b] + n * (long long)J;
                Wn[term] = wcn * r0.q;
                Wn[Jc + term] = wsn * r0.q;
            }
        }
        prod *= dn;
        if ((n & 7) == 7) { if (hw == 0) logdet += log(prod); prod = 1.0; }
        Fc = fma(wcn, zp, Fc); Fs = fma(wsn, zp, Fs);
        // into the frame of step n + 1
        wcn *= r1.r; wsn *= r1.r; Fc *= r1.r; Fs *= r1.r;

        // operands of matrix phase n + 2: row n + 2 and the rank-1 term of step n
        if (n + 2 < N) {
            if (act) {
                sm.A[par][ke][kb] = make_double2(r2.uc, dn * wcn);
                sm.A[par][ke + 1][kb] = make_double2(r2.us, dn * wsn);
                sm.C[par][ke][kb] = make_double2(r2.uc, wcn);
                sm.C[par][ke + 1][kb] = make_double2(r2.us, wsn);
                sm.R[par][k0] = r2.r; sm.R[par][k0 + 1] = r2.r;
            }
            if (ht == 0) sm.renorm[par] = r2.flag;
            bar_arrive(BAR_OPS + par, FT_THREADS);
        }
        wc = wcn; ws = wsn;
        if (n + 1 < N) {
            double alpha = r1.uc * wc + r1.us * ws;
            hreduce1(sm, par, hw, lane, alpha);
            kappa = dn * alpha;
        }
        r0 = r1; r1 = r2;
        const long long m3 = n + 3;
        r2 = gen.make(m3 < N ? sm.tbuf[m3 & (RING - 1)] : 0.0, m3, act && m3 < N);
        r2.flag = __shfl_sync(0xffffffffu, r2.flag, 0);
    }
    if (ht == 0) {
        if (prod != 1.0) logdet += log(prod);
        A.logdet[b] = logdet;
        if (MODE == MODE_LOGLIKE && A.quad) A.quad[b] = quad;
        A.status[b] = fail;
    }
}

// Per-sequence prologue shared by both roles: claim the next sequence, clear the buffers.
// Returns false when the queue is empty.
struct SeqInfo {
    int b, Jc, nsb;
    long long N;
};

__device__ __forceinline__ bool next_sequence(FastSmem &sm, const ScanArgs &A, const int tid, SeqInfo &q)
{
    __syncthreads();   // everybody is done with the previous sequence
    if (tid == 0) sm.next = atomicAdd(A.counter, 1);
    // operand buffers (padding columns must read as zero) and partial sums
    {
        double *z = reinterpret_cast<double *>(&sm.A[0][0][0]);
        const int nz = (int)((sizeof(sm.A) + sizeof(sm.C) + sizeof(sm.R) + sizeof(sm.P)) / sizeof(double));
        for (int i = tid; i < nz; i += FT_THREADS) z[i] = 0.0;
    }
    __syncthreads();
    const int item = sm.next;
    if (item >= A.B) return false;
    q.b = A.order[item];
    q.N = A.n_off[q.b + 1] - A.n_off[q.b];
    q.Jc = (int)(A.j_off[q.b + 1] - A.j_off[q.b]);
    const int nb = (2 * q.Jc + TILE - 1) / TILE;
    q.nsb = (nb + 1) / 2;
    if (tid == 0) sm.stop = q.N;
    __syncthreads();
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(FT_THREADS, 1) scan_fast_kernel(ScanArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw);
    const int tid = threadIdx.x;
    // the two roles never rejoin: each has its own persistent loop, so that the register
    // budgets set by setmaxnreg apply to the whole role
    if (tid < MAT_THREADS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_MAT));
        SeqInfo q;
        while (next_sequence(sm, A, tid, q)) {
            if (q.N > 0) matrix_loop(sm, tid, q.nsb, q.N);
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_HLP));
        SeqInfo q;
        while (next_sequence(sm, A, tid, q)) {
            if (q.N > 0) {
                vector_loop<MODE>(sm, A, tid - MAT_THREADS, q.b, q.nsb, q.N, q.Jc);
            } else if (tid == MAT_THREADS) {
                A.logdet[q.b] = 0.0;
                if (A.quad) A.quad[q.b] = 0.0;
                A.status[q.b] = 0;
            }
        }
    }
}

template <int MODE>
cudaError_t launch_mode(const ScanArgs &args, int grid, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(scan_fast_kernel<MODE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(FastSmem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    scan_fast_kernel<MODE><<<grid, FT_THREADS, sizeof(FastSmem), stream>>>(args);
    return cudaGetLastError();
}

}  // namespace

bool scan_fast_supports(int mode, int jmax)
{
    (void)mode;
    return jmax <= JP_MAX;
}

cudaError_t launch_scan_fast(int mode, const ScanArgs &args, int jmax, int sm_count,
                             cudaStream_t stream, int *launches)
{
    (void)jmax;
    const int grid = (int)(args.B < sm_count ? args.B : sm_count);
    *launches = 1;
    switch (mode) {
    case MODE_LOGLIKE: return launch_mode<MODE_LOGLIKE>(args, grid, stream);
    case MODE_SAMPLE:  return launch_mode<MODE_SAMPLE>(args, grid, stream);
    default:           return launch_mode<MODE_FACTOR>(args, grid, stream);
    }
}

}  // namespace gf
