// Hyper-parameter feeder on the device: stellar parameters -> SHO hyper-parameters
// (reference gadfly/core.py:107-333, Hyperparameters.for_star; scaling relations gadfly/scale.py) ->
// celerite coefficients incl. the exposure-time transform and its diagonal correction (celerite2
// SHOTerm.get_coefficients / TermConvolution.get_coefficients; called from
// gadfly/core.py:345-394), for B stars at once, plus the bandpass amplitude ratio of Morris+ (2020)
// Eqn 11 (gadfly/scale.py:635-729).
//
//   hyper_kernel     one CTA per star, one thread per solar term: scaled (S0, w0, Q) and the keep mask
//                    (the reference drops terms whose scaled frequency or power is not positive), count
//   coef_kernel      one CTA per star: compaction to the CSR layout, (a, b, c, d), (a', b'), Delta-diag
//   coef_csr_kernel  the same arithmetic for (S0, w0, Q) already in CSR layout, one warp per kernel
//   bandpass_kernel  one CTA per star: four Planck-weighted quadratures over the wavelength grid
//
// The arithmetic follows gadfly_b200/feeder.py (the vectorised host restatement) operation by operation.
// The Voigt envelope of Kiefer et al. (2018) needs Re w(z): the reference takes it from astropy /
// scipy's Faddeeva function; here it is the trapezoid rule on (y/pi) int exp(-t^2) / ((x-t)^2 + y^2) dt,
// which converges like exp(-2 pi y / h) -- Im z is the same constant 0.587 for every star (the ratio of
// the Lorentzian and Gaussian widths), h = 0.1 gives 1e-14 relative against scipy.special.wofz.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace gf {

namespace {

constexpr int FEED_THREADS = 128;      // >= terms per star (5 granulation + 81 p-modes in the solar fit)
constexpr int VOIGT_K = 64;
constexpr double VOIGT_H = 0.1;
constexpr double PI = 3.141592653589793;
constexpr double T_SUN = 5777.0, NUMAX_SUN = 3090.0, DNU_SUN = 135.1;

__constant__ double c_voigt_w[2 * VOIGT_K + 1];     // exp(-(k h)^2)

__device__ double voigt_re_w(double x, double y)
{
    double s = 0.0;
    const double y2 = y * y;
#pragma unroll 4
    for (int k = -VOIGT_K; k <= VOIGT_K; ++k) {
        const double u = x - VOIGT_H * (double)k;
        s += c_voigt_w[k + VOIGT_K] / (u * u + y2);
    }
    return y * VOIGT_H / PI * s;
}

// gadfly_b200/scale.py::_v_osc_kiefer_scaled (reference gadfly/scale.py:515-539), m^2/s^2
__device__ double v_osc_kiefer(double freq, double nu_max, double dnu)
{
    const double w = DNU_SUN / dnu;
    const double sigma = 181.8 / w, gamma = 150.9 / w, Sigma = 611.8 / w;
    const double S = -0.1, a = 3299 * 1e4, b = -581.0;
    const double A = 1 / PI * (atan(S * (freq - nu_max) / Sigma) + 0.5);
    const double fwhm_L = 2 * gamma, fwhm_G = 2.355 * sigma;
    const double sqrt_ln2 = sqrt(log(2.0));
    const double zr = 2.0 * (freq - nu_max) * sqrt_ln2 / fwhm_G;
    const double zi = fwhm_L * sqrt_ln2 / fwhm_G;
    const double voigt = voigt_re_w(zr, zi) * sqrt(log(2.0) * PI) / fwhm_G * fwhm_L * a;
    return A * __dadd_rn(b, voigt) * 1e-6;
}

// Kjeldsen & Bedding (1995) Eqn 5 (reference gadfly/scale.py:579-588)
__device__ double velocity_to_intensity(double v, double T, double wl_nm)
{
    const double r = T / 5777.0;
    return 20.1 * (v / (wl_nm / 550.0) / (r * r));
}

struct StarScalars {
    double amp, nu_max, gran_amp, gran_tau, scale_dnu, dnu, ratio_huber, Gamma, i_numax;
};

__global__ void __launch_bounds__(FEED_THREADS)
hyper_kernel(FeedArgs A, double *sho_all, unsigned char *keep_all, int32_t *count)
{
    __shared__ StarScalars sc;
    const int nt = A.n_gran + A.n_modes;
    for (int64_t b = blockIdx.x; b < A.B; b += gridDim.x) {
        const double M = A.mass[b], R = A.radius[b], T = A.temperature[b], L = A.luminosity[b];
        if (threadIdx.x == 0) {
            sc.amp = A.alpha ? A.alpha[b] : 1.0;
            sc.nu_max = NUMAX_SUN * (M * pow(R, -2.0) * pow(T / T_SUN, -0.5));
            sc.gran_amp = (L * L / (pow(M, 3.0) * pow(T, 5.5))) / A.gran_power_sun;
            sc.gran_tau = (L / (M * pow(T, 3.5))) / A.tau_sun;
            sc.scale_dnu = sqrt(M) * pow(R, -1.5);
            sc.dnu = DNU_SUN * sc.scale_dnu;
            const double c_K = pow(T / 5934.0, 0.8);
            const double amp_huber = pow(L, 0.886) / (pow(M, 1.89) * T * c_K);
            sc.ratio_huber = amp_huber / A.amp_huber_sun;
            sc.Gamma = 1.02 * exp((T - T_SUN) / 436.0);
            sc.i_numax = velocity_to_intensity(v_osc_kiefer(sc.nu_max, sc.nu_max, sc.dnu), T, A.wl_nm);
        }
        __syncthreads();
        const int i = threadIdx.x;
        double S0 = 0.0, w0 = 0.0, Q = 0.0;
        bool keep = false;
        if (i < A.n_gran) {
            S0 = A.gran[3 * i] * sc.gran_amp * sc.amp;
            w0 = A.gran[3 * i + 1] / sc.gran_tau;
            Q = A.gran[3 * i + 2];
            keep = w0 > 0.0;
        } else if (i < nt) {
            const double *m = A.modes + (size_t)(i - A.n_gran) * (4 + A.n_gran);
            const double solar_nu = m[0], Q_fit = m[1], solar_Gamma = m[2], unscaled_height = m[3];
            double bg_sum = 0.0;
            for (int g = 0; g < A.n_gran; ++g) bg_sum = __dadd_rn(bg_sum, __dmul_rn(m[4 + g], sc.amp));
            const double nu = __dadd_rn(sc.nu_max, __dmul_rn(solar_nu - NUMAX_SUN, sc.scale_dnu));
            w0 = 2 * PI * nu;
            const bool positive = w0 > 0.0;
            const double nu_safe = positive ? nu : 1.0;
            const double i_freq = velocity_to_intensity(v_osc_kiefer(nu_safe, sc.nu_max, sc.dnu), T, A.wl_nm);
            const double factor = (i_freq / sc.i_numax) * sc.ratio_huber;
            Q = Q_fit * sc.Gamma / solar_Gamma;
            const double height = unscaled_height * factor;
            const double amp_A = sqrt(PI * sc.Gamma * height / 2);
            const double half = amp_A / 2;
            const double peak = half * half / (4 * PI * nu_safe);
            S0 = (0.5 * sqrt(PI / 2) * peak / (Q * Q)) * bg_sum;
            keep = positive && S0 > 0.0;
        }
        if (i < nt) {
            double *o = sho_all + ((size_t)b * nt + i) * 3;
            o[0] = S0; o[1] = w0; o[2] = Q;
            keep_all[(size_t)b * nt + i] = keep ? 1 : 0;
        }
        const int kept = __syncthreads_count(keep ? 1 : 0);
        const int over = __syncthreads_count((keep && Q < 0.5) ? 1 : 0);
        // a kept overdamped term (Q < 0.5) has no (a, b, c, d) in this layout: reported as a negative count
        if (threadIdx.x == 0) count[b] = over ? -kept - 1 : kept;
    }
}

// One SHO term -> (a, b, c, d), (a', b') and its share of the diagonal correction.
// SHOTerm.get_coefficients, underdamped branch (eps = 1e-5 as in celerite2), then the
// exposure-time transform in the expression order of terms.TermConvolution: it is
// cancellation-sensitive, so nothing here is contracted into FMAs.
struct TermCoef { double a, b, c, d, a_new, b_new, dterm; };
__device__ TermCoef term_coefficients(double S0, double w0, double Q, double dt)
{
    const double f = sqrt(fmax(__dsub_rn(__dmul_rn(4.0, __dmul_rn(Q, Q)), 1.0), 1e-5));
    const double a = S0 * w0 * Q;
    const double bb = a / f;
    const double c = 0.5 * w0 / Q;
    const double d = c * f;
    const double cd = __dmul_rn(c, dt), dd = __dmul_rn(d, dt);
    const double c2 = __dmul_rn(c, c), d2 = __dmul_rn(d, d);
    const double c2pd2 = __dadd_rn(c2, d2), c2md2 = __dsub_rn(c2, d2);
    const double q = __dmul_rn(dt, c2pd2);
    const double factor = 2.0 / __dmul_rn(q, q);
    const double ch = cosh(cd), sh = sinh(cd);
    double sn, cs;
    sincos(dd, &sn, &cs);
    const double cos_term = __dsub_rn(__dmul_rn(ch, cs), 1.0);
    const double sin_term = __dmul_rn(sh, sn);
    const double C1 = __dadd_rn(__dmul_rn(a, c2md2), __dmul_rn(__dmul_rn(__dmul_rn(2.0, bb), c), d));
    const double C2 = __dsub_rn(__dmul_rn(bb, c2md2), __dmul_rn(__dmul_rn(__dmul_rn(2.0, a), c), d));
    TermCoef r;
    r.a = a; r.b = bb; r.c = c; r.d = d;
    r.a_new = __dmul_rn(factor, __dsub_rn(__dmul_rn(C1, cos_term), __dmul_rn(C2, sin_term)));
    r.b_new = __dmul_rn(factor, __dadd_rn(__dmul_rn(C2, cos_term), __dmul_rn(C1, sin_term)));
    const double norm = __dmul_rn(q, q);
    const double acbd = __dadd_rn(__dmul_rn(a, c), __dmul_rn(bb, d));
    const double num = __dadd_rn(__dsub_rn(__dmul_rn(__dmul_rn(C2, ch), sn), __dmul_rn(__dmul_rn(C1, sh), cs)),
                                 __dmul_rn(__dmul_rn(acbd, dt), c2pd2));
    r.dterm = num / norm;
    return r;
}

__global__ void __launch_bounds__(FEED_THREADS)
coef_kernel(FeedArgs A, const double *sho_all, const unsigned char *keep_all, const int64_t *j_off,
            double *sho, double *coef, double *base, double *ddiag)
{
    __shared__ int warp_kept[FEED_THREADS / 32];
    __shared__ double dterm_s[FEED_THREADS];
    const int nt = A.n_gran + A.n_modes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b = blockIdx.x; b < A.B; b += gridDim.x) {
        const int i = threadIdx.x;
        const bool keep = i < nt && keep_all[(size_t)b * nt + i] != 0;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_kept[warp] = __popc(mask);
        __syncthreads();
        int rank = __popc(mask & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) rank += warp_kept[w];
        int total = 0;
        for (int w = 0; w < FEED_THREADS / 32; ++w) total += warp_kept[w];
        if (keep) {
            const double *s = sho_all + ((size_t)b * nt + i) * 3;
            const double S0 = s[0], w0 = s[1], Q = s[2];
            const TermCoef r = term_coefficients(S0, w0, Q, A.delta[b]);
            dterm_s[rank] = r.dterm;
            const size_t row = (size_t)j_off[b] + rank;
            if (sho) { sho[3 * row] = S0; sho[3 * row + 1] = w0; sho[3 * row + 2] = Q; }
            if (base) { base[4 * row] = r.a; base[4 * row + 1] = r.b; base[4 * row + 2] = r.c; base[4 * row + 3] = r.d; }
            coef[4 * row] = r.a_new; coef[4 * row + 1] = r.b_new; coef[4 * row + 2] = r.c; coef[4 * row + 3] = r.d;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int k = 0; k < total; ++k) s += dterm_s[k];
            ddiag[b] = 2 * s;
        }
        __syncthreads();
    }
}

// (S0, w0, Q) already in CSR layout (a hyper-parameter lattice, the perturbed kernels of a gradient):
// one warp per kernel, lanes stride over its terms; Delta-diag summed in term order by lane 0 of a
// per-term scratch so that the result does not depend on the launch shape.
__global__ void __launch_bounds__(FEED_THREADS)
coef_csr_kernel(int64_t B, const int64_t *j_off, const double *sho, const double *delta, double *coef,
                double *base, double *dterm, double *ddiag, int32_t *overdamped)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (FEED_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t nwarp = (int64_t)gridDim.x * (FEED_THREADS / 32);
    for (int64_t b = warp0; b < B; b += nwarp) {
        const int64_t j0 = j_off[b], j1 = j_off[b + 1];
        const double dt = delta[b];
        for (int64_t j = j0 + lane; j < j1; j += 32) {
            const double Q = sho[3 * j + 2];
            if (Q < 0.5) atomicAdd(overdamped, 1);
            const TermCoef r = term_coefficients(sho[3 * j], sho[3 * j + 1], Q, dt);
            if (base) { base[4 * j] = r.a; base[4 * j + 1] = r.b; base[4 * j + 2] = r.c; base[4 * j + 3] = r.d; }
            coef[4 * j] = r.a_new; coef[4 * j + 1] = r.b_new; coef[4 * j + 2] = r.c; coef[4 * j + 3] = r.d;
            dterm[j] = r.dterm;
        }
        __syncwarp();
        if (lane == 0) {
            double s = 0.0;
            for (int64_t j = j0; j < j1; ++j) s += dterm[j];
            ddiag[b] = 2 * s;
        }
    }
}

constexpr int BP_THREADS = 256;

__device__ double planck_nu(double wl_um, double T)
{
    const double h = 6.62607015e-34, c = 299792458.0, kB = 1.380649e-23;
    const double nu = c / (wl_um * 1e-6);
    return 2.0 * h * (nu * nu * nu) / (c * c) / expm1(h * nu / (kB * T));
}

// reference gadfly/scale.py:635-729: ratio_0 = int dI/dT wl F / int dI/dT wl,  ratio_1 = int I wl / int I wl F
// (dI/dT by the reference's +-10 K difference), trapezoid rule on the caller's wavelength grid
__global__ void __launch_bounds__(BP_THREADS)
bandpass_kernel(int64_t B, const double *T, int64_t n_wl, const double *wl, const double *filt, double *out)
{
    __shared__ double red[4][BP_THREADS / 32];
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        const double Tb = T[b];
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int64_t k = threadIdx.x; k + 1 < n_wl; k += BP_THREADS) {
            double y[2][4];
            for (int e = 0; e < 2; ++e) {
                const double w = wl[k + e], F = filt[k + e];
                const double I = planck_nu(w, Tb);
                const double dI = (planck_nu(w, Tb + 10.0) - planck_nu(w, Tb - 10.0)) / 20.0;
                y[e][0] = dI * w * F; y[e][1] = dI * w; y[e][2] = I * w; y[e][3] = I * w * F;
            }
            const double dx = wl[k + 1] - wl[k];
            for (int q = 0; q < 4; ++q) acc[q] += 0.5 * (y[1][q] + y[0][q]) * dx;
        }
        for (int q = 0; q < 4; ++q) {
            double v = acc[q];
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s[4];
            for (int q = 0; q < 4; ++q) {
                s[q] = 0.0;
                for (int w = 0; w < BP_THREADS / 32; ++w) s[q] += red[q][w];
            }
            out[b] = (s[0] / s[1]) * (s[2] / s[3]);
        }
        __syncthreads();
    }
}

bool voigt_ready[64] = {};

}  // namespace

int feed_max_terms() { return FEED_THREADS; }

cudaError_t launch_feed_hyper(const FeedArgs &A, double *sho_all, unsigned char *keep_all, int32_t *count,
                              cudaStream_t stream)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && !voigt_ready[dev]) {
        double w[2 * VOIGT_K + 1];
        for (int k = -VOIGT_K; k <= VOIGT_K; ++k) { const double t = VOIGT_H * k; w[k + VOIGT_K] = std::exp(-t * t); }
        e = cudaMemcpyToSymbolAsync(c_voigt_w, w, sizeof(w), 0, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(stream);      // w is on this stack frame
        if (e != cudaSuccess) return e;
        voigt_ready[dev] = true;
    }
    const unsigned grid = (unsigned)std::min<int64_t>(A.B, 1 << 20);
    hyper_kernel<<<grid, FEED_THREADS, 0, stream>>>(A, sho_all, keep_all, count);
    return cudaGetLastError();
}

cudaError_t launch_feed_coef(const FeedArgs &A, const double *sho_all, const unsigned char *keep_all,
                             const int64_t *j_off, double *sho, double *coef, double *base, double *ddiag,
                             cudaStream_t stream)
{
    const unsigned grid = (unsigned)std::min<int64_t>(A.B, 1 << 20);
    coef_kernel<<<grid, FEED_THREADS, 0, stream>>>(A, sho_all, keep_all, j_off, sho, coef, base, ddiag);
    return cudaGetLastError();
}

cudaError_t launch_feed_sho(int64_t B, const int64_t *j_off, const double *sho, const double *delta,
                            double *coef, double *base, double *dterm, double *ddiag, int32_t *overdamped,
                            cudaStream_t stream)
{
    const int64_t blocks = (B + FEED_THREADS / 32 - 1) / (FEED_THREADS / 32);
    const unsigned grid = (unsigned)std::min<int64_t>(blocks, 1 << 20);
    coef_csr_kernel<<<grid, FEED_THREADS, 0, stream>>>(B, j_off, sho, delta, coef, base, dterm, ddiag, overdamped);
    return cudaGetLastError();
}

cudaError_t launch_bandpass(int64_t B, const double *T, int64_t n_wl, const double *wl, const double *filt,
                            double *out, cudaStream_t stream)
{
    const unsigned grid = (unsigned)std::min<int64_t>(B, 1 << 20);
    bandpass_kernel<<<grid, BP_THREADS, 0, stream>>>(B, T, n_wl, wl, filt, out);
    return cudaGetLastError();
}

}  // namespace gf
