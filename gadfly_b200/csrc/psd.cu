// K5: kernel power spectral density on dense frequency grids (data-parallel).
//
//   psd_b(w) = sqrt(2/pi) * sum_j ((a c + b d)(c^2+d^2) + (a c - b d) w^2)
//                                 / (w^4 + 2 (c^2 - d^2) w^2 + (c^2+d^2)^2)   * sinc^2(delta_b w / 2)
//
// (celerite2 Term.get_psd / TermConvolution.get_psd; reference call sites gadfly/psd.py:151,
// gadfly/tests/test_core.py:34).  For the high-Q p-mode terms (Q ~ 1e3..1e4) the denominator
// cancels by a factor ~Q^2 near resonance, so the 1e-12 parity bar is only reachable by
// evaluating every product and sum in the oracle's order with no FMA contraction: all
// cancellation-sensitive arithmetic uses the __d*_rn intrinsics, which nvcc never fuses.
//
// Mapping: grid = (frequency chunks, stars).  The per-term constants of one star live in
// shared memory; each thread owns PSD_BINS_PER_THREAD bins (independent division chains for
// ILP) and loops over the terms; omega loads and psd stores are coalesced.
#include "common.cuh"

namespace gf {

namespace {

// 1 / x for a positive normal x to ~1 ulp: hardware seed (>= 20 bits) and one third-order step.
// (The IEEE division of __ddiv_rn costs ~25 FP64 instructions per (term, bin) and dominated the
// kernel; the quotient is not cancellation-sensitive, the 1e-12 bar needs 1e-15, not 0.5 ulp.)
__device__ __forceinline__ double psd_rcp(const double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double p = fma(e, e, e);
    return fma(r, p, r);
}

constexpr int PSD_THREADS = 256;
constexpr int PSD_BPT = 4;           // bins per thread
constexpr int PSD_MAX_TERMS = 512;   // terms staged per pass

__global__ void __launch_bounds__(PSD_THREADS)
psd_kernel(const int64_t *__restrict__ j_off, const double *__restrict__ coef,
           const double *__restrict__ delta, const double *__restrict__ omega, int64_t F,
           double *__restrict__ out)
{
    __shared__ double4 s_k[PSD_MAX_TERMS];   // (k1, k2, k3, k4) per term
    const int b = blockIdx.y;
    const int64_t j0 = j_off[b];
    const int Jc = (int)(j_off[b + 1] - j0);
    const double dl = delta ? delta[b] : 0.0;

    const int64_t base = ((int64_t)blockIdx.x * PSD_THREADS) * PSD_BPT + threadIdx.x;
    double w2[PSD_BPT], w4[PSD_BPT], acc[PSD_BPT], w[PSD_BPT];
#pragma unroll
    for (int q = 0; q < PSD_BPT; ++q) {
        int64_t f = base + (int64_t)q * PSD_THREADS;
        w[q] = (f < F) ? omega[f] : 1.0;
        w2[q] = __dmul_rn(w[q], w[q]);
        w4[q] = __dmul_rn(w2[q], w2[q]);
        acc[q] = 0.0;
    }

    for (int jbase = 0; jbase < Jc; jbase += PSD_MAX_TERMS) {
        const int cnt = min(PSD_MAX_TERMS, Jc - jbase);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += PSD_THREADS) {
            const double *cf = coef + 4 * (j0 + jbase + j);
            const double a = cf[0], bb = cf[1], c = cf[2], d = cf[3];
            const double c2 = __dmul_rn(c, c), d2 = __dmul_rn(d, d);
            const double w02 = __dadd_rn(c2, d2);
            const double ac = __dmul_rn(a, c), bd = __dmul_rn(bb, d);
            double4 k;
            k.x = __dmul_rn(2.0, __dsub_rn(c2, d2));        // k1 = 2 (c^2 - d^2)
            k.y = __dmul_rn(w02, w02);                      // k2 = (c^2 + d^2)^2
            k.z = __dmul_rn(__dadd_rn(ac, bd), w02);        // k3 = (ac + bd) w0^2
            k.w = __dsub_rn(ac, bd);                        // k4 = ac - bd
            s_k[j] = k;
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const double4 k = s_k[j];
#pragma unroll
            for (int q = 0; q < PSD_BPT; ++q) {
                // Only the DENOMINATOR is cancellation-sensitive (by ~Q^2 next to a resonance) and keeps
                // the oracle's un-fused order; the numerator (k4 = a c - b d is a rounding residue
                // for an SHO term) and the accumulation of the positive quotients are fused: 8
                // instead of 10 FP64 instructions per (term, bin), differences <= 1 ulp of the sum.
                const double num = fma(k.w, w2[q], k.z);
                const double den = __dadd_rn(__dadd_rn(w4[q], __dmul_rn(k.x, w2[q])), k.y);
                acc[q] = fma(num, psd_rcp(den), acc[q]);
            }
        }
    }

    const double pre = 0.79788456080286535588;  // sqrt(2/pi)
#pragma unroll
    for (int q = 0; q < PSD_BPT; ++q) {
        int64_t f = base + (int64_t)q * PSD_THREADS;
        if (f < F) {
            double psd = __dmul_rn(pre, acc[q]);
            if (dl > 0.0) {
                const double arg = __dmul_rn(__dmul_rn(0.5, dl), w[q]);
                const double sinc = (fabs(arg) > 0.0) ? __ddiv_rn(sin(arg), arg) : 1.0;
                psd = __dmul_rn(psd, __dmul_rn(sinc, sinc));
            }
            out[(int64_t)b * F + f] = psd;
        }
    }
}

}  // namespace

cudaError_t launch_psd(int64_t B, const int64_t *j_off, const double *coef, const double *delta,
                       const double *omega, int64_t F, double *out, cudaStream_t stream)
{
    const int64_t per_block = (int64_t)PSD_THREADS * PSD_BPT;
    // gridDim.y is limited to 65535 stars per launch
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
        int64_t nbatch = (B - b0 < 65535) ? (B - b0) : 65535;
        dim3 grid((unsigned)((F + per_block - 1) / per_block), (unsigned)nbatch);
        psd_kernel<<<grid, PSD_THREADS, 0, stream>>>(j_off + b0, coef, delta ? delta + b0 : nullptr,
                                                     omega, F, out + b0 * F);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace gf
