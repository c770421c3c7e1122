// Copy / kernel overlap of the LIBRARY's scan kernels through its C ABI, without torch in the process.
// Input: variants/solar_coef.bin (Jc, coef[Jc][4], ddiag) written by tools/e2e_probe.py --dump or the
// snippet in profiles/r2_overlap.txt.
// nvcc -O2 -o variants/overlap_probe2 tools/overlap_probe2.cu -Iinclude -Lgadfly_b200 -lgadfly_b200 -Xlinker -rpath -Xlinker $PWD/gadfly_b200
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "gadfly_b200.h"

static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { int rc_ = (int)(x); if (rc_ != 0) { printf("FAILED %s -> %d (%s)\n", #x, rc_, h ? gf_last_error(h) : ""); exit(1); } } while (0)

int main(int argc, char **argv)
{
    const int64_t N = argc > 1 ? atoll(argv[1]) : (1 << 18);
    const int64_t B = 148;
    gf_handle h = nullptr;
    FILE *f = fopen("variants/solar_coef.bin", "rb");
    if (!f) { printf("no variants/solar_coef.bin\n"); return 1; }
    double jc_d; fread(&jc_d, 8, 1, f);
    const int64_t Jc = (int64_t)jc_d;
    std::vector<double> c1((size_t)Jc * 4); fread(c1.data(), 8, c1.size(), f);
    double dd1; fread(&dd1, 8, 1, f); fclose(f);
    std::vector<double> coef((size_t)B * Jc * 4), ddiag((size_t)B, dd1);
    for (int64_t b = 0; b < B; ++b) std::copy(c1.begin(), c1.end(), coef.begin() + b * Jc * 4);
    std::vector<int64_t> n_off(B + 1), t_off(B, 0), j_off(B + 1);
    for (int64_t b = 0; b <= B; ++b) { n_off[b] = b * N; j_off[b] = b * Jc; }
    CK(gf_create(0, &h));
    double *t_d, *y_d, *x_d, *y_h, *x_h, *big_h, *big_d, *ld_d, *q_d; int32_t *st_d;
    const size_t big = 1u << 30;
    cudaMalloc(&t_d, N * 8); cudaMalloc(&y_d, B * N * 8); cudaMalloc(&x_d, B * N * 8);
    cudaMalloc(&ld_d, B * 8); cudaMalloc(&q_d, B * 8); cudaMalloc(&st_d, B * 4);
    cudaMallocHost(&y_h, B * N * 8); cudaMallocHost(&x_h, B * N * 8);
    cudaMallocHost(&big_h, big); cudaMalloc(&big_d, big);
    {
        std::vector<double> t((size_t)N);
        for (int64_t i = 0; i < N; ++i) t[i] = 6e-5 * (double)i;
        cudaMemcpy(t_d, t.data(), N * 8, cudaMemcpyHostToDevice);
    }
    cudaStream_t sc; cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
    // y: a draw of the process itself (K2), so that K1 sees sensible data
    CK(gf_sample_batched(h, B, n_off.data(), t_off.data(), j_off.data(), t_d, N, nullptr, coef.data(), ddiag.data(),
                         nullptr, 1, 0, y_d, ld_d, st_d, 0));
    cudaMemcpy(y_h, y_d, B * N * 8, cudaMemcpyDeviceToHost);
    auto K1 = [&](const double *y, uint32_t fl) {
        CK(gf_loglike_batched(h, B, n_off.data(), t_off.data(), j_off.data(), t_d, N, y, nullptr, coef.data(),
                              ddiag.data(), ld_d, q_d, st_d, fl)); };
    auto K2 = [&](double *x, uint32_t fl) {
        CK(gf_sample_batched(h, B, n_off.data(), t_off.data(), j_off.data(), t_d, N, nullptr, coef.data(),
                             ddiag.data(), nullptr, 2, 0, x, ld_d, st_d, fl)); };
    auto sync = [&]() { CK(gf_synchronize(h)); cudaDeviceSynchronize(); };
    auto timed = [&](auto fn) { fn(); sync(); double best = 1e30; for (int r = 0; r < 3; ++r) { double t0 = now(); fn(); sync(); best = std::min(best, now() - t0); } return best; };
    const double k1 = timed([&] { K1(y_d, 0); }), k2 = timed([&] { K2(x_d, 0); });
    printf("device-resident: K1 (log-likelihood) %.1f ms, K2 (sample) %.1f ms\n", k1, k2);
    for (int kern = 0; kern < 2; ++kern)
        for (int dir = 0; dir < 2; ++dir) {
            sync();
            const double t0 = now();
            if (kern == 0) K1(y_d, 1); else K2(x_d, 1);
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
            if (dir == 0) cudaMemcpyAsync(big_d, big_h, big, cudaMemcpyHostToDevice, sc);
            else cudaMemcpyAsync(big_h, big_d, big, cudaMemcpyDeviceToHost, sc);
            cudaStreamSynchronize(sc);
            const double t1 = now();
            sync();
            printf("%s running, %s of 1 GiB issued at 50 ms on another stream: copy done at %.0f ms, kernel done at %.0f ms\n",
                   kern == 0 ? "K1" : "K2", dir == 0 ? "H2D" : "D2H", t1 - t0, now() - t0);
        }
    const double a = timed([&] { K2(x_d, 1); K1(y_h, 0); });
    printf("K2 async (device out) then K1 with y in pinned host memory: %.1f ms (+%.1f over K1 + K2)\n", a, a - k1 - k2);
    const double b = timed([&] { K1(y_d, 1); K1(y_h, 0); });
    printf("K1 async (device)     then K1 with y in pinned host memory: %.1f ms (+%.1f over 2 K1)\n", b, b - 2 * k1);
    const double c = timed([&] { K2(x_h, 1); K1(y_d, 0); });
    printf("K2 async to pinned host then K1 device-resident:            %.1f ms (+%.1f over K1 + K2)\n", c, c - k1 - k2);
    const double d = timed([&] { K2(x_h, 1); K2(x_d, 0); });
    printf("K2 async to pinned host then K2 device-resident:            %.1f ms (+%.1f over 2 K2)\n", d, d - 2 * k2);
    // the Python binding's default: small outputs (log det, status) in ordinary pageable memory
    std::vector<double> ld_p((size_t)B); std::vector<int32_t> st_p((size_t)B);
    const double e = timed([&] {
        CK(gf_sample_batched(h, B, n_off.data(), t_off.data(), j_off.data(), t_d, N, nullptr, coef.data(),
                             ddiag.data(), nullptr, 2, 0, x_d, ld_p.data(), st_p.data(), 1));
        K1(y_h, 0); });
    printf("K2 async with PAGEABLE log det / status outputs, then K1 with host y: %.1f ms (+%.1f over K1 + K2)\n", e, e - k1 - k2);
    gf_destroy(h);
    return 0;
}
