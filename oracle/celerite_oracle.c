/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
 *
 * CPU restatement of the semiseparable ("celerite") recurrences that gadfly's
 * GP hot path runs inside celerite2:
 *   gadfly/gp.py:59,167-204   GaussianProcess.compute        -> factor
 *   gadfly/gp.py:329-350      log_likelihood                 -> solve_lower + norm
 *   gadfly/gp.py:308-327      dot_tril                       -> matmul_lower
 *   gadfly/gp.py:372-395      sample                         -> dot_tril(randn)
 *   gadfly/gp.py:352-370      apply_inverse                  -> solve_lower, /d, solve_upper
 * The arithmetic itself lives in celerite2 (PyPI "celerite2", un-pinned in the
 * reference's pyproject.toml:20, C++ header c++/include/celerite2/core.hpp), which
 * is NOT under /root/reference and cannot be installed here.  This file restates
 * its published algorithm (Foreman-Mackey et al. 2017, AJ 154, 220; celerite2
 * docs) -- SURVEY.md Appendix A.2/A.6 -- and is validated against dense
 * numpy.linalg.cholesky on the kernel definition (tests/test_oracle.py).  The
 * reference ships no golden vectors for this path, so every "matches celerite2"
 * claim means "matches this validated restatement": PARITY UNPINNED.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (gadfly_b200/) never does.
 *
 * Column order of the J-wide state follows celerite2: [real | cos-block | sin-block],
 * J = Jr + 2*Jc.  All arrays are row-major float64.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* c[J] and one row of (U, V): Term.get_celerite_matrices, SURVEY A.2 */
static void fill_c(int Jr, int Jc, const double *cr, const double *cc, double *c)
{
    for (int j = 0; j < Jr; ++j) c[j] = cr[j];
    for (int j = 0; j < Jc; ++j) { c[Jr + j] = cc[j]; c[Jr + Jc + j] = cc[j]; }
}

static void fill_row(int Jr, int Jc, double tn, const double *ar, const double *ac,
                     const double *bc, const double *dc, double *u, double *v)
{
    for (int j = 0; j < Jr; ++j) { u[j] = ar[j]; v[j] = 1.0; }
    for (int j = 0; j < Jc; ++j) {
        double arg = dc[j] * tn;
        double cs = cos(arg), sn = sin(arg);
        u[Jr + j] = ac[j] * cs + bc[j] * sn;
        u[Jr + Jc + j] = ac[j] * sn - bc[j] * cs;
        v[Jr + j] = cs;
        v[Jr + Jc + j] = sn;
    }
}

static double sum_a(int Jr, int Jc, const double *ar, const double *ac)
{
    double s = 0.0;
    for (int j = 0; j < Jr; ++j) s += ar[j];
    for (int j = 0; j < Jc; ++j) s += ac[j];
    return s;
}

/* Materialise c[J], a[N], U[N,J], V[N,J].  diag may be NULL (zeros). */
void orc_matrices(long N, int Jr, int Jc, const double *t, const double *diag, double ddiag,
                  const double *ar, const double *cr, const double *ac, const double *bc,
                  const double *cc, const double *dc, double *c, double *a, double *U, double *V)
{
    int J = Jr + 2 * Jc;
    fill_c(Jr, Jc, cr, cc, c);
    double sa = sum_a(Jr, Jc, ar, ac);
    for (long n = 0; n < N; ++n) {
        a[n] = ((diag ? diag[n] : 0.0) + ddiag) + sa;
        fill_row(Jr, Jc, t[n], ar, ac, bc, dc, U + n * J, V + n * J);
    }
}

/* factor: K = L D L^T.  Returns 0, or 1 + index of the first d[n] <= 0.  SURVEY A.6 */
long orc_factor(long N, int J, const double *t, const double *c, const double *a,
                const double *U, const double *V, double *d, double *W)
{
    double *S = (double *)calloc((size_t)J * J, sizeof(double));
    double *p = (double *)malloc(sizeof(double) * J);
    double *tmp = (double *)malloc(sizeof(double) * J);
    long fail = 0;
    d[0] = a[0];
    if (!(d[0] > 0.0)) { fail = 1; goto done; }
    for (int j = 0; j < J; ++j) W[j] = V[j] / d[0];
    for (long n = 1; n < N; ++n) {
        const double *wp = W + (n - 1) * J, *un = U + n * J, *vn = V + n * J;
        double *wn = W + n * J;
        double dt = t[n - 1] - t[n];
        for (int j = 0; j < J; ++j) p[j] = exp(c[j] * dt);
        /* S += d[n-1] w^T w ; S = diag(p) S diag(p) */
        for (int j = 0; j < J; ++j) {
            double dw = d[n - 1] * wp[j];
            double *Sj = S + (size_t)j * J;
            for (int k = 0; k < J; ++k) Sj[k] += dw * wp[k];
        }
        for (int j = 0; j < J; ++j) {
            double *Sj = S + (size_t)j * J;
            for (int k = 0; k < J; ++k) Sj[k] = (p[j] * Sj[k]) * p[k];
        }
        /* tmp = u S */
        for (int k = 0; k < J; ++k) tmp[k] = 0.0;
        for (int j = 0; j < J; ++j) {
            const double *Sj = S + (size_t)j * J;
            double uj = un[j];
            for (int k = 0; k < J; ++k) tmp[k] += uj * Sj[k];
        }
        double dn = a[n];
        double acc = 0.0;
        for (int k = 0; k < J; ++k) acc += tmp[k] * un[k];
        dn -= acc;
        d[n] = dn;
        if (!(dn > 0.0)) { fail = n + 1; goto done; }
        for (int k = 0; k < J; ++k) wn[k] = (vn[k] - tmp[k]) / dn;
    }
done:
    free(S); free(p); free(tmp);
    return fail;
}

/* Z = L^{-1} Y (sign = -1) or Z = L Y (sign = +1); Y, Z are [N, nrhs]; may alias. */
static void lower_sweep(int sign, long N, int J, int nrhs, const double *t, const double *c,
                        const double *U, const double *W, const double *Y, double *Z)
{
    double *F = (double *)calloc((size_t)J * nrhs, sizeof(double));
    double *p = (double *)malloc(sizeof(double) * J);
    double *prev = (double *)malloc(sizeof(double) * nrhs);
    for (int r = 0; r < nrhs; ++r) { prev[r] = (sign < 0) ? Y[r] : Y[r]; Z[r] = Y[r]; }
    for (long n = 1; n < N; ++n) {
        double dt = t[n - 1] - t[n];
        const double *wp = W + (n - 1) * J, *un = U + n * J;
        for (int j = 0; j < J; ++j) p[j] = exp(c[j] * dt);
        /* solve: F += w^T z[n-1] (the already-solved value); matmul: F += w^T y[n-1] (the input) */
        for (int j = 0; j < J; ++j)
            for (int r = 0; r < nrhs; ++r)
                F[(size_t)j * nrhs + r] = p[j] * (F[(size_t)j * nrhs + r] + wp[j] * prev[r]);
        for (int r = 0; r < nrhs; ++r) {
            double acc = 0.0;
            for (int j = 0; j < J; ++j) acc += un[j] * F[(size_t)j * nrhs + r];
            double yn = Y[n * nrhs + r];
            double zn = (sign < 0) ? yn - acc : yn + acc;
            prev[r] = (sign < 0) ? zn : yn;
            Z[n * nrhs + r] = zn;
        }
    }
    free(F); free(p); free(prev);
}

void orc_solve_lower(long N, int J, int nrhs, const double *t, const double *c, const double *U,
                     const double *W, const double *Y, double *Z)
{ lower_sweep(-1, N, J, nrhs, t, c, U, W, Y, Z); }

void orc_matmul_lower(long N, int J, int nrhs, const double *t, const double *c, const double *U,
                      const double *W, const double *Y, double *Z)
{ lower_sweep(+1, N, J, nrhs, t, c, U, W, Y, Z); }

/* Z = L^{-T} Y (sign=-1) or Z = L^T Y (sign=+1) */
static void upper_sweep(int sign, long N, int J, int nrhs, const double *t, const double *c,
                        const double *U, const double *W, const double *Y, double *Z)
{
    double *F = (double *)calloc((size_t)J * nrhs, sizeof(double));
    double *p = (double *)malloc(sizeof(double) * J);
    double *prev = (double *)malloc(sizeof(double) * nrhs);
    for (int r = 0; r < nrhs; ++r) { Z[(N - 1) * nrhs + r] = Y[(N - 1) * nrhs + r]; prev[r] = Y[(N - 1) * nrhs + r]; }
    for (long n = N - 2; n >= 0; --n) {
        double dt = t[n] - t[n + 1];
        const double *un1 = U + (n + 1) * J, *wn = W + n * J;
        for (int j = 0; j < J; ++j) p[j] = exp(c[j] * dt);
        for (int j = 0; j < J; ++j)
            for (int r = 0; r < nrhs; ++r)
                F[(size_t)j * nrhs + r] = p[j] * (F[(size_t)j * nrhs + r] + un1[j] * prev[r]);
        for (int r = 0; r < nrhs; ++r) {
            double acc = 0.0;
            for (int j = 0; j < J; ++j) acc += wn[j] * F[(size_t)j * nrhs + r];
            double yn = Y[n * nrhs + r];
            double zn = (sign < 0) ? yn - acc : yn + acc;
            prev[r] = (sign < 0) ? zn : yn;
            Z[n * nrhs + r] = zn;
        }
    }
    free(F); free(p); free(prev);
}

void orc_solve_upper(long N, int J, int nrhs, const double *t, const double *c, const double *U,
                     const double *W, const double *Y, double *Z)
{ upper_sweep(-1, N, J, nrhs, t, c, U, W, Y, Z); }

void orc_matmul_upper(long N, int J, int nrhs, const double *t, const double *c, const double *U,
                      const double *W, const double *Y, double *Z)
{ upper_sweep(+1, N, J, nrhs, t, c, U, W, Y, Z); }

/*
 * Streaming (nothing materialised) fused pass: factor + forward solve (mode 0) or
 * factor + L*(sqrt(d) o y) (mode 1).  Same operation order as the functions above,
 * so results are bit-identical to orc_matrices -> orc_factor -> orc_*_lower.
 * This is the "celerite2-equivalent CPU path" the benchmark times: dense J x J state,
 * 6 J^2 flop per step.
 *   mode 0: out[0] = sum log d, out[1] = sum z^2/d        (x may be NULL)
 *   mode 1: x[N] = L (sqrt(d) o y)  (+ out[0] = sum log d)
 * Returns 0 or 1 + index of first non-positive d.
 */
long orc_stream(int mode, long N, int Jr, int Jc, const double *t, const double *y,
                const double *diag, double ddiag, const double *ar, const double *cr,
                const double *ac, const double *bc, const double *cc, const double *dc,
                double *out, double *x)
{
    int J = Jr + 2 * Jc;
    size_t JJ = (size_t)J * J;
    double *S = (double *)calloc(JJ + 7 * (size_t)J, sizeof(double));
    double *c = S + JJ, *p = c + J, *u = p + J, *v = u + J, *w = v + J, *tmp = w + J, *F = tmp + J;
    fill_c(Jr, Jc, cr, cc, c);
    double sa = sum_a(Jr, Jc, ar, ac);
    double logdet = 0.0, quad = 0.0, dprev, zprev;
    long fail = 0;
    fill_row(Jr, Jc, t[0], ar, ac, bc, dc, u, v);
    dprev = ((diag ? diag[0] : 0.0) + ddiag) + sa;
    if (!(dprev > 0.0)) { fail = 1; goto done; }
    for (int j = 0; j < J; ++j) w[j] = v[j] / dprev;
    logdet = log(dprev);
    if (mode == 0) { zprev = y[0]; quad = zprev * zprev / dprev; }
    else { zprev = y[0] * sqrt(dprev); x[0] = zprev; }
    for (long n = 1; n < N; ++n) {
        double dt = t[n - 1] - t[n];
        for (int j = 0; j < J; ++j) p[j] = exp(c[j] * dt);
        for (int j = 0; j < J; ++j) F[j] = p[j] * (F[j] + w[j] * zprev);
        fill_row(Jr, Jc, t[n], ar, ac, bc, dc, u, v);
        for (int k = 0; k < J; ++k) tmp[k] = 0.0;
        /* one pass over S: rank-1 update, decay, and tmp = u S -- per element the same
         * operations in the same order as orc_factor's three passes (bit-identical results
         * in the non-contracting build), but S streams through the cache once per step */
        for (int j = 0; j < J; ++j) {
            const double dw = dprev * w[j], pj = p[j], uj = u[j];
            double *restrict Sj = S + (size_t)j * J;
            for (int k = 0; k < J; ++k) {
                double sjk = Sj[k] + dw * w[k];
                sjk = (pj * sjk) * p[k];
                Sj[k] = sjk;
                tmp[k] += uj * sjk;
            }
        }
        double dn = ((diag ? diag[n] : 0.0) + ddiag) + sa;
        double acc = 0.0;
        for (int k = 0; k < J; ++k) acc += tmp[k] * u[k];
        dn -= acc;
        if (!(dn > 0.0)) { fail = n + 1; goto done; }
        for (int k = 0; k < J; ++k) w[k] = (v[k] - tmp[k]) / dn;
        double uf = 0.0;
        for (int j = 0; j < J; ++j) uf += u[j] * F[j];
        logdet += log(dn);
        if (mode == 0) {
            double zn = y[n] - uf;
            quad += zn * zn / dn;
            zprev = zn;
        } else {
            double yn = y[n] * sqrt(dn);
            x[n] = yn + uf;
            zprev = yn;
        }
        dprev = dn;
    }
done:
    out[0] = logdet; out[1] = quad;
    free(S);
    return fail;
}

/* Batch driver for the CPU baseline: sequence b uses t/y[n_off[b]..n_off[b+1]) and complex
 * terms [j_off[b], j_off[b+1]) of (ac,bc,cc,dc) (complex terms only -- all gadfly kernels are).
 * One sequence per OpenMP thread.  out is [B][2]; x (mode 1) is laid out like y. */
void orc_stream_batch(int mode, long B, const long *n_off, const long *t_off, const long *j_off,
                      const double *t, const double *y, const double *ddiag, const double *ac,
                      const double *bc, const double *cc, const double *dc, double *out,
                      double *x, long *status, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < B; ++b) {
        long N = n_off[b + 1] - n_off[b];
        int Jc = (int)(j_off[b + 1] - j_off[b]);
        long j0 = j_off[b];
        status[b] = orc_stream(mode, N, 0, Jc, t + t_off[b], y + n_off[b], NULL, ddiag[b],
                               NULL, NULL, ac + j0, bc + j0, cc + j0, dc + j0, out + 2 * b,
                               x ? x + n_off[b] : NULL);
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Term.get_psd, SURVEY A.5: un-convolved coefficients, times sinc^2(delta*omega/2). */
void orc_psd(long F, const double *omega, int Jr, int Jc, const double *ar, const double *cr,
             const double *ac, const double *bc, const double *cc, const double *dc, double delta,
             double *out)
{
    const double pre = sqrt(2.0 / M_PI);
    for (long i = 0; i < F; ++i) {
        double w = omega[i], w2 = w * w, acc = 0.0;
        for (int j = 0; j < Jr; ++j) acc += ar[j] * cr[j] / (cr[j] * cr[j] + w2);
        for (int j = 0; j < Jc; ++j) {
            double c2 = cc[j] * cc[j], d2 = dc[j] * dc[j];
            double ac_ = ac[j] * cc[j], bd = bc[j] * dc[j];
            double w02 = c2 + d2;
            acc += ((ac_ + bd) * w02 + (ac_ - bd) * w2) / (w2 * w2 + 2.0 * (c2 - d2) * w2 + w02 * w02);
        }
        double psd = pre * acc;
        if (delta > 0.0) {
            double arg = 0.5 * delta * w;
            double sinc = (arg == 0.0) ? 1.0 : sin(arg) / arg;
            psd *= sinc * sinc;
        }
        out[i] = psd;
    }
}
