// k right-hand sides per sequence on ONE factor: the data-parallel helpers around the O(N J) sweeps.
//
//   prep_kernel   Y[b][r][n] = sqrt(d_b[n]) * n_{b,r}[n]   (the input of matmul_lower for sampling,
//                 celerite2 dot_tril: reference gadfly/gp.py:308-327,391), normal draws either read or
//                 drawn here from Philox4x32-10 with the global realisation index seq0 + b k + r
//   quad_kernel   quad[b][r] = sum_n z_{b,r}[n]^2 / d_b[n]   (log-likelihood after solve_lower, gp.py:350)
#include "common.cuh"

namespace gf {

namespace {

constexpr int MT = 256;

// virtual sequence v = b k + r; its samples start at n_off_v[v]; d of its sequence at d_off[v]
__global__ void __launch_bounds__(MT) prep_kernel(int64_t V, const int64_t *n_off_v, const int64_t *d_off,
                                                  const double *d, const double *normals, uint64_t seed,
                                                  uint64_t seq0, double *out)
{
    for (int64_t v = blockIdx.y; v < V; v += gridDim.y) {
        const int64_t n0 = n_off_v[v], N = n_off_v[v + 1] - n0;
        const double *dv = d + d_off[v];
        for (int64_t n = (int64_t)blockIdx.x * MT + threadIdx.x; n < N; n += (int64_t)gridDim.x * MT) {
            const double z = normals ? normals[n0 + n] : philox_normal(seed, seq0 + (uint64_t)v, (uint64_t)n);
            out[n0 + n] = sqrt(dv[n]) * z;
        }
    }
}

__global__ void __launch_bounds__(MT) quad_kernel(int64_t V, const int64_t *n_off_v, const int64_t *d_off,
                                                  const double *d, const double *z, double *quad)
{
    __shared__ double red[MT / 32];
    for (int64_t v = blockIdx.x; v < V; v += gridDim.x) {
        const int64_t n0 = n_off_v[v], N = n_off_v[v + 1] - n0;
        const double *dv = d + d_off[v];
        // sequential chunks per thread, then a fixed tree: the result does not depend on the launch
        double acc = 0.0;
        for (int64_t n = threadIdx.x; n < N; n += MT) {
            const double zn = z[n0 + n];
            acc = fma(zn * zn, 1.0 / dv[n], acc);
        }
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < MT / 32; ++w) s += red[w];
            quad[v] = s;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_multi_prep(int64_t V, int64_t max_n, const int64_t *n_off_v, const int64_t *d_off,
                              const double *d, const double *normals, uint64_t seed, uint64_t seq0,
                              double *out, cudaStream_t stream)
{
    if (V == 0 || max_n == 0) return cudaSuccess;
    int gx = (int)((max_n + MT - 1) / MT);
    if (gx > 1024) gx = 1024;
    dim3 grid(gx, (unsigned)(V < 65535 ? V : 65535));
    prep_kernel<<<grid, MT, 0, stream>>>(V, n_off_v, d_off, d, normals, seed, seq0, out);
    return cudaGetLastError();
}

cudaError_t launch_multi_quad(int64_t V, const int64_t *n_off_v, const int64_t *d_off, const double *d,
                              const double *z, double *quad, cudaStream_t stream)
{
    if (V == 0) return cudaSuccess;
    const int grid = (int)(V < 65535 ? V : 65535);
    quad_kernel<<<grid, MT, 0, stream>>>(V, n_off_v, d_off, d, z, quad);
    return cudaGetLastError();
}

}  // namespace gf
