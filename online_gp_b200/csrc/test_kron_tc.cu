// Standalone check + timing of the tcgen05 / TMEM / TMA pair kernels (kron_tc.cu) against a CPU fp64 reference and
// against the SIMT pair kernels (kron_fused.cu).  Build: csrc/build.sh; run on the B200 box:  ./test_kron_tc [time]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace wiski {
int tc_pair_apply(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X, float* Y, int64_t c,
                  cudaStream_t st, const int64_t* h_lay, long long* prof);
int tc_pair_grad_dir(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int pair, const float* Z,
                     const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st, const int64_t* h_lay);
int fused_pair_apply(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X, float* Y,
                     int64_t c, cudaStream_t st, const int64_t* h_lay);
int fused_pair_grad_jvp(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int pair,
                        const float* Z, const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st,
                        const int64_t* h_lay);
}
extern "C" int wiski_kron_tc_enable(int on);

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

struct Case {
    int d;
    int64_t g[4];
    int pair;
    int64_t c;
    int chunks_in, chunks_out;     // 1 = plain row-major
};

// element (row, col) of a (possibly column-chunked) panel
static inline size_t at(int64_t row, int64_t col, int64_t rows, int64_t c, int chunks) {
    const int64_t cw = c / chunks;
    return (size_t)((col / cw) * rows * cw + row * cw + col % cw);
}

static void toeplitz_apply(const std::vector<double>& col, const double* x, double* y) {
    for (int a = 0; a < 32; ++a) {
        double s = 0;
        for (int b = 0; b < 32; ++b) s += col[std::abs(a - b)] * x[b];
        y[a] = s;
    }
}

static int run_case(const Case& cs) {
    const int d = cs.d;
    int64_t m = 1;
    for (int i = 0; i < d; ++i) m *= cs.g[i];
    const int64_t c = cs.c, gmax = 32;
    const int u_ax = 2 * cs.pair, v_ax = u_ax + 1;
    int64_t sv = 1, nb = 1;
    for (int j = v_ax + 1; j < d; ++j) sv *= cs.g[j];
    for (int j = 0; j < u_ax; ++j) nb *= cs.g[j];
    std::vector<float> cols(d * gmax), dirs(d * gmax), X((size_t)m * c), Z((size_t)m * c);
    for (int i = 0; i < d; ++i)
        for (int k = 0; k < 32; ++k) {
            cols[i * gmax + k] = expf(-0.004f * (1 + i) * k * k) * (1.f + 0.1f * i);
            dirs[i * gmax + k] = cols[i * gmax + k] * 0.01f * k * k;
        }
    for (auto& v : X) v = frand();
    for (auto& v : Z) v = frand();
    float *dcols, *ddirs, *dX, *dZ, *dY, *dY2;
    double* dout;
    cudaMalloc(&dcols, cols.size() * 4); cudaMalloc(&ddirs, dirs.size() * 4);
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dY, X.size() * 4); cudaMalloc(&dY2, X.size() * 4);
    cudaMalloc(&dout, 6 * 8);
    cudaMemcpy(dcols, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ddirs, dirs.data(), dirs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dY, 0xff, X.size() * 4);
    cudaMemset(dout, 0, 6 * 8);
    int64_t lay[9];
    auto set_lay = [&](int i, int chunks) {
        const int64_t cw = c / chunks;
        lay[3 * i] = cw; lay[3 * i + 1] = cw; lay[3 * i + 2] = chunks > 1 ? m * cw : 0;
    };
    set_lay(0, cs.chunks_in); set_lay(1, cs.chunks_out); set_lay(2, 1);
    const bool plain = cs.chunks_in == 1 && cs.chunks_out == 1;
    int rc = wiski::tc_pair_apply(dcols, d, cs.g, gmax, cs.pair, dX, dY, c, 0, plain ? nullptr : lay, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case d=%d g0=%lld pair=%d c=%lld chunks=(%d,%d) m=%lld: apply rc=%d (%s) cuda=%s\n", d, (long long)cs.g[0], cs.pair,
           (long long)c, cs.chunks_in, cs.chunks_out, (long long)m, rc, rc ? wiski_last_error() : "", cudaGetErrorString(e));
    if (rc != 0 || e != cudaSuccess) return 1;
    std::vector<float> Y(X.size());
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<double> cu(cols.begin() + u_ax * gmax, cols.begin() + u_ax * gmax + 32);
    std::vector<double> cv(cols.begin() + v_ax * gmax, cols.begin() + v_ax * gmax + 32);
    std::vector<double> du(dirs.begin() + u_ax * gmax, dirs.begin() + u_ax * gmax + 32);
    std::vector<double> dv(dirs.begin() + v_ax * gmax, dirs.begin() + v_ax * gmax + 32);
    // check a sample of (ob, oa, column) lines completely
    double maxerr = 0, maxref = 0;
    int nchk = 0;
    for (int64_t ob = 0; ob < nb; ob += (nb > 7 ? nb / 7 : 1))
        for (int64_t oa = 0; oa < sv; oa += (sv > 5 ? sv / 5 : 1))
            for (int64_t col = 0; col < c; col += (c > 48 ? 37 : 5)) {
                double t[32][32], y1[32][32];
                for (int u = 0; u < 32; ++u) {
                    double x[32];
                    for (int v = 0; v < 32; ++v)
                        x[v] = X[at(ob * 1024 * sv + (u * 32 + v) * sv + oa, col, m, c, cs.chunks_in)];
                    toeplitz_apply(cv, x, t[u]);
                }
                for (int v = 0; v < 32; ++v) {
                    double x[32], y[32];
                    for (int u = 0; u < 32; ++u) x[u] = t[u][v];
                    toeplitz_apply(cu, x, y);
                    for (int u = 0; u < 32; ++u) y1[u][v] = y[u];
                }
                for (int u = 0; u < 32; ++u)
                    for (int v = 0; v < 32; ++v) {
                        const double got = Y[at(ob * 1024 * sv + (u * 32 + v) * sv + oa, col, m, c, cs.chunks_out)];
                        maxerr = fmax(maxerr, fabs(got - y1[u][v]));
                        maxref = fmax(maxref, fabs(y1[u][v]));
                    }
                ++nchk;
            }
    printf("  apply: %d lines checked, max abs err %.3e, max |ref| %.3e, rel %.3e %s\n", nchk, maxerr, maxref, maxerr / maxref,
           maxerr / maxref < 2e-5 ? "OK" : "FAIL");
    int bad = !(maxerr / maxref < 2e-5);
    // directional gradient pass (plain layouts or chunked Z), STORE and no-store variants
    if (cs.chunks_out == 1) {
        for (int store = 1; store >= 0; --store) {
            cudaMemset(dout, 0, 6 * 8);
            cudaMemset(dY2, 0xff, X.size() * 4);
            set_lay(0, cs.chunks_in); set_lay(1, 1); set_lay(2, 1);
            // Z operand: reuse dX's layout convention (chunks_in) by passing dX-like data: Z is stored plain, so re-lay it
            std::vector<float> Zl(Z.size());
            for (int64_t r = 0; r < m; ++r)
                for (int64_t col = 0; col < c; ++col) Zl[at(r, col, m, c, cs.chunks_in)] = Z[(size_t)r * c + col];
            cudaMemcpy(dZ, Zl.data(), Zl.size() * 4, cudaMemcpyHostToDevice);
            // P operand = X stored plain
            std::vector<float> Xp(X.size());
            for (int64_t r = 0; r < m; ++r)
                for (int64_t col = 0; col < c; ++col) Xp[(size_t)r * c + col] = X[at(r, col, m, c, cs.chunks_in)];
            cudaMemcpy(dY, Xp.data(), Xp.size() * 4, cudaMemcpyHostToDevice);
            rc = wiski::tc_pair_grad_dir(dcols, ddirs, d, cs.g, gmax, cs.pair, dZ, dY, store ? dY2 : nullptr, c, dout, 0,
                                         cs.chunks_in == 1 ? nullptr : lay);
            e = cudaDeviceSynchronize();
            printf("  grad_dir store=%d rc=%d (%s) cuda=%s\n", store, rc, rc ? wiski_last_error() : "", cudaGetErrorString(e));
            if (rc != 0 || e != cudaSuccess) return 1;
            double out3[3];
            cudaMemcpy(out3, dout, 24, cudaMemcpyDeviceToHost);
            std::vector<float> Zo(X.size());
            cudaMemcpy(Zo.data(), dY2, Zo.size() * 4, cudaMemcpyDeviceToHost);
            // CPU reference over ALL tiles (double)
            double r0 = 0, r1 = 0, r2 = 0, a0 = 0, a1 = 0, a2 = 0, zerr = 0, zref = 0;
            for (int64_t ob = 0; ob < nb; ++ob)
                for (int64_t oa = 0; oa < sv; ++oa)
                    for (int64_t col = 0; col < c; ++col) {
                        static double S[32][32], zu[32][32], zd[32][32];
                        for (int u = 0; u < 32; ++u) {
                            double x[32];
                            for (int v = 0; v < 32; ++v) x[v] = Xp[(size_t)(ob * 1024 * sv + (u * 32 + v) * sv + oa) * c + col];
                            toeplitz_apply(cv, x, S[u]);
                        }
                        for (int v = 0; v < 32; ++v) {
                            double x[32], y[32], y2[32];
                            for (int u = 0; u < 32; ++u) x[u] = Z[(size_t)(ob * 1024 * sv + (u * 32 + v) * sv + oa) * c + col];
                            toeplitz_apply(cu, x, y);
                            toeplitz_apply(du, x, y2);
                            for (int u = 0; u < 32; ++u) { zu[u][v] = y[u]; zd[u][v] = y2[u]; }
                        }
                        for (int u = 0; u < 32; ++u) {
                            double y[32], y2[32];
                            toeplitz_apply(dv, zu[u], y2);
                            toeplitz_apply(cv, zu[u], y);
                            for (int v = 0; v < 32; ++v) {
                                const double pval = Xp[(size_t)(ob * 1024 * sv + (u * 32 + v) * sv + oa) * c + col];
                                r0 += zd[u][v] * S[u][v]; a0 += fabs(zd[u][v] * S[u][v]);
                                r1 += y2[v] * pval;       a1 += fabs(y2[v] * pval);
                                r2 += zu[u][v] * S[u][v]; a2 += fabs(zu[u][v] * S[u][v]);
                                if (store) {
                                    const double got = Zo[(size_t)(ob * 1024 * sv + (u * 32 + v) * sv + oa) * c + col];
                                    zerr = fmax(zerr, fabs(got - y[v]));
                                    zref = fmax(zref, fabs(y[v]));
                                }
                            }
                        }
                    }
            const double e0 = fabs(out3[0] - r0) / a0, e1 = fabs(out3[1] - r1) / a1, e2 = fabs(out3[2] - r2) / a2;
            printf("    out3 = %.9e %.9e %.9e\n    ref  = %.9e %.9e %.9e\n    err / sum|terms| = %.2e %.2e %.2e %s", out3[0], out3[1],
                   out3[2], r0, r1, r2, e0, e1, e2, (e0 < 1e-6 && e1 < 1e-6 && e2 < 1e-6) ? "OK" : "FAIL");
            bad |= !(e0 < 1e-6 && e1 < 1e-6 && e2 < 1e-6);
            if (store) {
                printf("   Zout rel err %.3e %s", zerr / zref, zerr / zref < 2e-5 ? "OK" : "FAIL");
                bad |= !(zerr / zref < 2e-5);
            }
            printf("\n");
        }
    }
    cudaFree(dcols); cudaFree(ddirs); cudaFree(dX); cudaFree(dZ); cudaFree(dY); cudaFree(dY2); cudaFree(dout);
    return bad;
}

static void timeit(int64_t c) {
    wiski_kron_tc_enable(0);      // the fused_* entry points below are the SIMT reference timings
    const int d = 4;
    const int64_t g[4] = {32, 32, 32, 32}, gmax = 32, m = 1 << 20;
    std::vector<float> cols(d * gmax), dirs(d * gmax);
    for (int i = 0; i < d; ++i)
        for (int k = 0; k < 32; ++k) {
            cols[i * gmax + k] = expf(-0.004f * (1 + i) * k * k);
            dirs[i * gmax + k] = cols[i * gmax + k] * 0.01f * k * k;
        }
    float *dcols, *ddirs, *dX, *dZ, *dY;
    double* dout;
    cudaMalloc(&dcols, cols.size() * 4); cudaMalloc(&ddirs, dirs.size() * 4);
    cudaMalloc(&dX, (size_t)m * c * 4); cudaMalloc(&dZ, (size_t)m * c * 4); cudaMalloc(&dY, (size_t)m * c * 4);
    cudaMalloc(&dout, 24);
    cudaMemcpy(dcols, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ddirs, dirs.data(), dirs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dX, 0, (size_t)m * c * 4); cudaMemset(dZ, 0, (size_t)m * c * 4); cudaMemset(dout, 0, 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)m * c * 4;
    auto report = [&](const char* name, double nb, auto fn) {
        fn();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: cuda=%s\n", name, cudaGetErrorString(e)); return; }
        const int reps = 10;
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) fn();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        printf("%-34s c=%lld  %.3f ms  %.0f GB/s (algorithmic %.2f GB)\n", name, (long long)c, ms, nb / ms / 1e6, nb / 1e9);
    };
    for (int pair = 0; pair < 2; ++pair) {
        char nm[64];
        snprintf(nm, 64, "tc pair_apply pair %d", pair);
        report(nm, 2 * bytes, [&] { wiski::tc_pair_apply(dcols, d, g, gmax, pair, dX, dY, c, 0, nullptr, nullptr); });
        snprintf(nm, 64, "simt pair_apply pair %d", pair);
        report(nm, 2 * bytes, [&] { wiski::fused_pair_apply(dcols, d, g, gmax, pair, dX, dY, c, 0, nullptr); });
        snprintf(nm, 64, "tc grad_dir store pair %d", pair);
        report(nm, 3 * bytes, [&] { wiski::tc_pair_grad_dir(dcols, ddirs, d, g, gmax, pair, dZ, dX, dY, c, dout, 0, nullptr); });
        snprintf(nm, 64, "tc grad_dir nostore pair %d", pair);
        report(nm, 2 * bytes, [&] { wiski::tc_pair_grad_dir(dcols, ddirs, d, g, gmax, pair, dZ, dX, nullptr, c, dout, 0, nullptr); });
        snprintf(nm, 64, "simt grad_dir store pair %d", pair);
        report(nm, 3 * bytes, [&] { wiski::fused_pair_grad_jvp(dcols, ddirs, d, g, gmax, pair, dZ, dX, dY, c, dout, 0, nullptr); });
    }
    // access pattern of a pairing (1,2) of the same 32^4 grid (rows 55 KB / 1.7 MB apart, 32 pages per tile), emulated as
    // pair 1 of the 5-D grid [32, 1, 32, 32, 32] (n_before = 32, sv = 32)
    {
        const int64_t g5[5] = {32, 1, 32, 32, 32};
        std::vector<float> c5(5 * gmax), d5(5 * gmax);
        for (int i = 0; i < 5; ++i)
            for (int k = 0; k < 32; ++k) { c5[i * gmax + k] = expf(-0.004f * (1 + i) * k * k); d5[i * gmax + k] = c5[i * gmax + k] * 0.01f * k * k; }
        float *dc5, *dd5;
        cudaMalloc(&dc5, c5.size() * 4); cudaMalloc(&dd5, d5.size() * 4);
        cudaMemcpy(dc5, c5.data(), c5.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dd5, d5.data(), d5.size() * 4, cudaMemcpyHostToDevice);
        report("tc pair_apply axes (1,2) emulated", 2 * bytes, [&] { wiski::tc_pair_apply(dc5, 5, g5, gmax, 1, dX, dY, c, 0, nullptr, nullptr); });
        report("tc grad_dir store axes (1,2) emul", 3 * bytes, [&] { wiski::tc_pair_grad_dir(dc5, dd5, 5, g5, gmax, 1, dZ, dX, dY, c, dout, 0, nullptr); });
        report("tc grad_dir nostore axes (1,2) emul", 2 * bytes, [&] { wiski::tc_pair_grad_dir(dc5, dd5, 5, g5, gmax, 1, dZ, dX, nullptr, c, dout, 0, nullptr); });
        cudaFree(dc5); cudaFree(dd5);
    }
    // page / DRAM locality of the strided pair (0,1): the same pass with the input and / or the output panel stored as
    // column blocks of 16 columns ([27][m][16]: a tile's 1024 rows are then 64 KB apart instead of 1.7 MB)
    {
        int64_t lay[6];
        const int64_t cwv[2] = {c, 16};
        for (int li = 0; li < 2; ++li)
            for (int lo = 0; lo < 2; ++lo) {
                lay[0] = cwv[li]; lay[1] = cwv[li]; lay[2] = li ? m * 16 : 0;
                lay[3] = cwv[lo]; lay[4] = cwv[lo]; lay[5] = lo ? m * 16 : 0;
                char nm[64];
                snprintf(nm, 64, "tc pair_apply pair 0 in cw=%lld out cw=%lld", (long long)cwv[li], (long long)cwv[lo]);
                report(nm, 2 * bytes, [&] { wiski::tc_pair_apply(dcols, d, g, gmax, 0, dX, dY, c, 0, lay, nullptr); });
            }
    }
    // role / phase cycle breakdown of the apply kernel (PROF instantiation), averaged per CTA and tile
    long long* dprof;
    cudaMalloc(&dprof, 16 * 8);
    for (int pair = 0; pair < 2; ++pair) {
        cudaMemset(dprof, 0, 16 * 8);
        wiski::tc_pair_apply(dcols, d, g, gmax, pair, dX, dY, c, 0, nullptr, dprof);
        cudaDeviceSynchronize();
        long long pr[16];
        cudaMemcpy(pr, dprof, 16 * 8, cudaMemcpyDeviceToHost);
        const double tiles = (double)m / 1024 * (c / 16);
        const char* names[13] = {"issuer0 wait a_ready", "issuer0 issue+commit", "w: wait full", "w: p1 load+split+st", "w: p1 wait d",
                                 "w: p1 ld+sts", "w: bar1", "w: p2 load+split+st", "w: bar2", "w: p2 wait d", "w: p2 ld+stg",
                                 "(unused)", "producer wait empty"};
        printf("apply pair %d cycles per tile (sum over CTAs / tiles):\n", pair);
        for (int i = 0; i < 13; ++i) printf("  %-24s %9.0f\n", names[i], pr[i] / tiles);
    }
    cudaFree(dprof);
    cudaFree(dcols); cudaFree(ddirs); cudaFree(dX); cudaFree(dZ); cudaFree(dY); cudaFree(dout);
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc > 1 && !strcmp(argv[1], "once")) {
        // one launch of each production kernel at the bench shape (for ncu --set full)
        const int d = 4;
        const int64_t g[4] = {32, 32, 32, 32}, gmax = 32, m = 1 << 20, c = argc > 2 ? atoll(argv[2]) : 432;
        std::vector<float> cols(d * gmax), dirs(d * gmax);
        for (int i = 0; i < d; ++i)
            for (int k = 0; k < 32; ++k) {
                cols[i * gmax + k] = expf(-0.004f * (1 + i) * k * k);
                dirs[i * gmax + k] = cols[i * gmax + k] * 0.01f * k * k;
            }
        float *dcols, *ddirs, *dX, *dZ, *dY;
        double* dout;
        cudaMalloc(&dcols, cols.size() * 4); cudaMalloc(&ddirs, dirs.size() * 4);
        cudaMalloc(&dX, (size_t)m * c * 4); cudaMalloc(&dZ, (size_t)m * c * 4); cudaMalloc(&dY, (size_t)m * c * 4);
        cudaMalloc(&dout, 24);
        cudaMemcpy(dcols, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(ddirs, dirs.data(), dirs.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dX, 0, (size_t)m * c * 4); cudaMemset(dZ, 0, (size_t)m * c * 4); cudaMemset(dout, 0, 24);
        wiski::tc_pair_apply(dcols, d, g, gmax, 1, dX, dY, c, 0, nullptr, nullptr);
        wiski::tc_pair_apply(dcols, d, g, gmax, 0, dX, dY, c, 0, nullptr, nullptr);
        wiski::tc_pair_grad_dir(dcols, ddirs, d, g, gmax, 0, dZ, dX, dY, c, dout, 0, nullptr);
        wiski::tc_pair_grad_dir(dcols, ddirs, d, g, gmax, 1, dZ, dX, nullptr, c, dout, 0, nullptr);
        printf("once: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "time")) {
        timeit(argc > 2 ? atoll(argv[2]) : 432);
        return 0;
    }
    const Case cases[] = {
        {2, {32, 32, 1, 1}, 0, 16, 1, 1},           // a single tile
        {2, {32, 32, 1, 1}, 0, 432, 1, 1},          // 27 tiles
        {4, {4, 4, 32, 32}, 1, 48, 1, 1},           // pair 1: sv = 1, n_before = 16
        {4, {32, 32, 4, 4}, 0, 32, 1, 1},           // pair 0: sv = 16
        {4, {32, 32, 4, 4}, 0, 64, 2, 1},           // column-chunked input (sharded path: received blocks)
        {4, {4, 4, 32, 32}, 1, 64, 1, 2},           // column-chunked output (send buffer)
        {4, {32, 32, 32, 32}, 0, 16, 1, 1},         // full 32^4, more tiles than SMs (1024 tiles)
        {4, {32, 32, 32, 32}, 1, 16, 1, 1},
    };
    for (const Case& cs : cases) bad |= run_case(cs);
    printf(bad ? "SOME CHECKS FAILED\n" : "ALL CHECKS PASSED\n");
    return bad;
}
