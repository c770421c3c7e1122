// Observed power spectrum of evenly sampled light curves and its (log-)binning, on the device
// (closes sample -> PSD -> compare without leaving HBM).
//
//   power_kernel  |rfft(flux)|^2 * norm, norm = d / sqrt(2 pi) / N  -- the reference's
//                 normalisation (gadfly/psd.py:566-587); the transform itself is cuFFT D2Z (library
//                 FFT: not the north-star path), batched over the light curves
//   bin_kernel    per (bin, light curve): trapezoidal mean of the power over the bin's points on
//                 the (log10) frequency axis and the reference's error estimate
//                 std(y) / sqrt(n) * mean(x) / (x_hi - x_lo) / constant  (gadfly/psd.py:186-297,
//                 spectral_binning / spectral_binning_err); a bin with one point returns that point
#include "common.cuh"
#include <cufft.h>
#include <dlfcn.h>

namespace gf {

namespace {

constexpr int OT = 256;

__global__ void __launch_bounds__(OT) power_kernel(int64_t B, int64_t NC, int64_t skip, double norm,
                                                   const double2 *spec, double *power)
{
    const int64_t nout = NC - skip;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y)
        for (int64_t i = (int64_t)blockIdx.x * OT + threadIdx.x; i < nout; i += (int64_t)gridDim.x * OT) {
            const double2 z = spec[b * NC + skip + i];
            power[b * nout + i] = (z.x * z.x + z.y * z.y) * norm;
        }
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < OT / 32; ++w) s += red[w];
    return s;
}

// bins are index ranges [lo[k], lo[k] + cnt[k]) of the shared, monotone axis x[F]
__global__ void __launch_bounds__(OT) bin_kernel(int64_t B, int64_t F, int64_t nb, const int64_t *lo,
                                                 const int64_t *cnt, const double *x, const double *power,
                                                 double constant, double *stat, double *err)
{
    __shared__ double red[OT / 32];
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const double *y = power + b * F;
        for (int64_t k = blockIdx.x; k < nb; k += gridDim.x) {
            const int64_t i0 = lo[k], n = cnt[k];
            double s = nan(""), e = nan("");
            if (n > 0) {
                const double span = x[i0 + n - 1] - x[i0];
                if (n > 1 && span > 0.0) {
                    double tz = 0.0, sy = 0.0, sx = 0.0;
                    for (int64_t i = threadIdx.x; i < n; i += OT) {
                        sy += y[i0 + i];
                        sx += x[i0 + i];
                        if (i + 1 < n) tz = fma(0.5 * (y[i0 + i + 1] + y[i0 + i]), x[i0 + i + 1] - x[i0 + i], tz);
                    }
                    tz = block_sum(tz, red);
                    sy = block_sum(sy, red);
                    sx = block_sum(sx, red);
                    const double mean_y = sy / (double)n;
                    double dev = 0.0;
                    for (int64_t i = threadIdx.x; i < n; i += OT) {
                        const double dlt = y[i0 + i] - mean_y;
                        dev = fma(dlt, dlt, dev);
                    }
                    dev = block_sum(dev, red);
                    s = tz / span;
                    e = sqrt(dev / (double)n) / sqrt((double)n) * ((sx / (double)n) / span) / constant;
                } else {
                    s = y[i0];
                    e = y[i0];
                }
            }
            if (threadIdx.x == 0) { stat[b * nb + k] = s; err[b * nb + k] = e; }
            __syncthreads();
        }
    }
}

}  // namespace

// cuFFT is bound at first use (dlopen), so that the library loads -- and the GP hot path runs --
// on a machine without libcufft; only this entry point then reports an error.
struct CufftApi {
    cufftResult (*PlanMany)(cufftHandle *, int, int *, int *, int, int, int *, int, int, cufftType, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal *, cufftDoubleComplex *) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool ok = false;
};
static CufftApi &cufft_api()
{
    static CufftApi api = [] {
        CufftApi a;
        void *lib = nullptr;
        for (const char *name : {"libcufft.so.11", "libcufft.so.12", "libcufft.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return a;
        a.PlanMany = (decltype(a.PlanMany))dlsym(lib, "cufftPlanMany");
        a.SetStream = (decltype(a.SetStream))dlsym(lib, "cufftSetStream");
        a.ExecD2Z = (decltype(a.ExecD2Z))dlsym(lib, "cufftExecD2Z");
        a.Destroy = (decltype(a.Destroy))dlsym(lib, "cufftDestroy");
        a.ok = a.PlanMany && a.SetStream && a.ExecD2Z && a.Destroy;
        return a;
    }();
    return api;
}

// cuFFT plan cache: one plan per handle, rebuilt when (N, B) changes
struct FftPlan {
    cufftHandle plan = 0;
    int64_t n = 0, batch = 0;
    bool valid = false;
};

cudaError_t launch_obs_power(FftPlan *fp, int64_t B, int64_t N, const double *flux, double d, int include_zero,
                             double2 *spec, double *power, cudaStream_t stream)
{
    if (B == 0 || N == 0) return cudaSuccess;
    CufftApi &cf = cufft_api();
    if (!cf.ok) return cudaErrorSharedObjectInitFailed;
    if (!fp->valid || fp->n != N || fp->batch != B) {
        if (fp->valid) cf.Destroy(fp->plan);
        fp->valid = false;
        int n[1] = {(int)N};
        if (cf.PlanMany(&fp->plan, 1, n, nullptr, 1, (int)N, nullptr, 1, (int)(N / 2 + 1), CUFFT_D2Z, (int)B) != CUFFT_SUCCESS)
            return cudaErrorUnknown;
        fp->n = N; fp->batch = B; fp->valid = true;
    }
    if (cf.SetStream(fp->plan, stream) != CUFFT_SUCCESS) return cudaErrorUnknown;
    if (cf.ExecD2Z(fp->plan, const_cast<double *>(flux), reinterpret_cast<cufftDoubleComplex *>(spec)) != CUFFT_SUCCESS)
        return cudaErrorUnknown;
    const int64_t NC = N / 2 + 1, skip = include_zero ? 0 : 1;
    const double norm = d / sqrt(2.0 * 3.14159265358979323846) / (double)N;
    int gx = (int)((NC + OT - 1) / OT);
    if (gx > 2048) gx = 2048;
    dim3 grid(gx, (unsigned)(B < 65535 ? B : 65535));
    power_kernel<<<grid, OT, 0, stream>>>(B, NC, skip, norm, spec, power);
    return cudaGetLastError();
}

void destroy_fft_plan(FftPlan *fp)
{
    if (fp && fp->valid) { cufft_api().Destroy(fp->plan); fp->valid = false; }
}

FftPlan *new_fft_plan() { return new FftPlan; }
void delete_fft_plan(FftPlan *fp) { destroy_fft_plan(fp); delete fp; }

cudaError_t launch_bin_power(int64_t B, int64_t F, int64_t nb, const int64_t *lo, const int64_t *cnt,
                             const double *x, const double *power, double constant, double *stat,
                             double *err, cudaStream_t stream)
{
    if (B == 0 || nb == 0) return cudaSuccess;
    dim3 grid((unsigned)(nb < 4096 ? nb : 4096), (unsigned)(B < 65535 ? B : 65535));
    bin_kernel<<<grid, OT, 0, stream>>>(B, F, nb, lo, cnt, x, power, constant, stat, err);
    return cudaGetLastError();
}

}  // namespace gf
