// Standalone check + timing of the tcgen05 / TMEM / TMA pair kernels (kron_tc.cu) against a CPU fp64 reference and
// against the SIMT pair kernels (kron_fused.cu).  Build: csrc/build.sh; run on the B200 box:
//   ./test_kron_tc                 correctness (any axis pair, chunked layouts, both store paths)
//   ./test_kron_tc time [c]        timings at the bench shape (32^4 grid, c = 432 columns) + role / phase cycle breakdown
//   ./test_kron_tc once [c]        one launch of each production kernel (for ncu --set full)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace wiski {
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
    int au, av;
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

struct Strides {
    int64_t before, mid, after, su, sv, smid, sbefore;
    int64_t row(int64_t bi, int u, int64_t mi, int v, int64_t ai) const { return bi * sbefore + u * su + mi * smid + v * sv + ai; }
};
static Strides strides_of(const Case& cs) {
    Strides s;
    s.before = s.mid = s.after = 1;
    for (int j = cs.av + 1; j < cs.d; ++j) s.after *= cs.g[j];
    for (int j = cs.au + 1; j < cs.av; ++j) s.mid *= cs.g[j];
    for (int j = 0; j < cs.au; ++j) s.before *= cs.g[j];
    s.sv = s.after;
    s.smid = 32 * s.after;
    s.su = s.mid * 32 * s.after;
    s.sbefore = 32 * s.mid * 32 * s.after;
    return s;
}

static int run_case(const Case& cs) {
    const int d = cs.d;
    int64_t m = 1;
    for (int i = 0; i < d; ++i) m *= cs.g[i];
    const int64_t c = cs.c, gmax = 32;
    const Strides S = strides_of(cs);
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
    cudaMemset(dY, 0xff, X.size() * 4);
    cudaMemset(dout, 0, 6 * 8);
    int64_t lay[9];
    auto set_lay = [&](int i, int chunks) {
        const int64_t cw = c / chunks;
        lay[3 * i] = cw; lay[3 * i + 1] = cw; lay[3 * i + 2] = chunks > 1 ? m * cw : 0;
    };
    set_lay(0, cs.chunks_in); set_lay(1, cs.chunks_out); set_lay(2, 1);
    const bool plain = cs.chunks_in == 1 && cs.chunks_out == 1;
    int rc = wiski::tc_pair_apply_axes(dcols, d, cs.g, gmax, cs.au, cs.av, dX, dY, c, 0, plain ? nullptr : lay, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    printf("case d=%d g=[%lld,%lld,%lld,%lld] axes=(%d,%d) c=%lld chunks=(%d,%d): apply rc=%d (%s) cuda=%s\n", d, (long long)cs.g[0],
           (long long)cs.g[1], (long long)cs.g[2], (long long)cs.g[3], cs.au, cs.av, (long long)c, cs.chunks_in, cs.chunks_out, rc,
           rc ? wiski_last_error() : "", cudaGetErrorString(e));
    if (rc != 0 || e != cudaSuccess) return 1;
    std::vector<float> Y(X.size());
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<double> cu(cols.begin() + cs.au * gmax, cols.begin() + cs.au * gmax + 32);
    std::vector<double> cv(cols.begin() + cs.av * gmax, cols.begin() + cs.av * gmax + 32);
    std::vector<double> du(dirs.begin() + cs.au * gmax, dirs.begin() + cs.au * gmax + 32);
    std::vector<double> dv(dirs.begin() + cs.av * gmax, dirs.begin() + cs.av * gmax + 32);
    auto step_of = [](int64_t n, int64_t k) { return n > k ? n / k : (int64_t)1; };
    // check a sample of (before, mid, after, column) line sets completely
    double maxerr = 0, maxref = 0;
    int nchk = 0;
    for (int64_t bi = 0; bi < S.before; bi += step_of(S.before, 3))
        for (int64_t mi = 0; mi < S.mid; mi += step_of(S.mid, 3))
            for (int64_t ai = 0; ai < S.after; ai += step_of(S.after, 3))
                for (int64_t col = 0; col < c; col += (c > 48 ? 37 : 5)) {
                    double t[32][32], y1[32][32];
                    for (int u = 0; u < 32; ++u) {
                        double x[32];
                        for (int v = 0; v < 32; ++v) x[v] = X[at(S.row(bi, u, mi, v, ai), col, m, c, cs.chunks_in)];
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
                            const double got = Y[at(S.row(bi, u, mi, v, ai), col, m, c, cs.chunks_out)];
                            maxerr = fmax(maxerr, fabs(got - y1[u][v]));
                            maxref = fmax(maxref, fabs(y1[u][v]));
                        }
                    ++nchk;
                }
    printf("  apply: %d line sets checked, max abs err %.3e, max |ref| %.3e, rel %.3e %s\n", nchk, maxerr, maxref, maxerr / maxref,
           maxerr / maxref < 2e-5 ? "OK" : "FAIL");
    int bad = !(maxerr / maxref < 2e-5);
    // directional gradient pass (chunked Z allowed), STORE and no-store variants
    if (cs.chunks_out == 1) {
        std::vector<float> Zl(Z.size()), Xp(X.size());
        for (int64_t r = 0; r < m; ++r)
            for (int64_t col = 0; col < c; ++col) {
                Zl[at(r, col, m, c, cs.chunks_in)] = Z[(size_t)r * c + col];
                Xp[(size_t)r * c + col] = X[at(r, col, m, c, cs.chunks_in)];
            }
        cudaMemcpy(dZ, Zl.data(), Zl.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dY, Xp.data(), Xp.size() * 4, cudaMemcpyHostToDevice);      // P operand, plain
        // CPU reference over ALL tiles (double)
        double r0 = 0, r1 = 0, r2 = 0, a0 = 0, a1 = 0, a2 = 0;
        std::vector<float> Zref(X.size());
        for (int64_t bi = 0; bi < S.before; ++bi)
            for (int64_t mi = 0; mi < S.mid; ++mi)
                for (int64_t ai = 0; ai < S.after; ++ai)
                    for (int64_t col = 0; col < c; ++col) {
                        static double Sm[32][32], zu[32][32], zd[32][32];
                        for (int u = 0; u < 32; ++u) {
                            double x[32];
                            for (int v = 0; v < 32; ++v) x[v] = Xp[(size_t)S.row(bi, u, mi, v, ai) * c + col];
                            toeplitz_apply(cv, x, Sm[u]);
                        }
                        for (int v = 0; v < 32; ++v) {
                            double x[32], y[32], y2[32];
                            for (int u = 0; u < 32; ++u) x[u] = Z[(size_t)S.row(bi, u, mi, v, ai) * c + col];
                            toeplitz_apply(cu, x, y);
                            toeplitz_apply(du, x, y2);
                            for (int u = 0; u < 32; ++u) { zu[u][v] = y[u]; zd[u][v] = y2[u]; }
                        }
                        for (int u = 0; u < 32; ++u) {
                            double y[32], y2[32];
                            toeplitz_apply(dv, zu[u], y2);
                            toeplitz_apply(cv, zu[u], y);
                            for (int v = 0; v < 32; ++v) {
                                const double pval = Xp[(size_t)S.row(bi, u, mi, v, ai) * c + col];
                                r0 += zd[u][v] * Sm[u][v]; a0 += fabs(zd[u][v] * Sm[u][v]);
                                r1 += y2[v] * pval;        a1 += fabs(y2[v] * pval);
                                r2 += zu[u][v] * Sm[u][v]; a2 += fabs(zu[u][v] * Sm[u][v]);
                                Zref[(size_t)S.row(bi, u, mi, v, ai) * c + col] = (float)y[v];
                            }
                        }
                    }
        for (int store = 1; store >= 0; --store) {
            cudaMemset(dout, 0, 6 * 8);
            cudaMemset(dY2, 0xff, X.size() * 4);
            set_lay(0, cs.chunks_in); set_lay(1, 1); set_lay(2, 1);
            rc = wiski::tc_pair_grad_dir_axes(dcols, ddirs, d, cs.g, gmax, cs.au, cs.av, dZ, dY, store ? dY2 : nullptr, c, dout, 0,
                                              cs.chunks_in == 1 ? nullptr : lay);
            e = cudaDeviceSynchronize();
            printf("  grad_dir store=%d rc=%d (%s) cuda=%s\n", store, rc, rc ? wiski_last_error() : "", cudaGetErrorString(e));
            if (rc != 0 || e != cudaSuccess) return 1;
            double out3[3];
            cudaMemcpy(out3, dout, 24, cudaMemcpyDeviceToHost);
            const double e0 = fabs(out3[0] - r0) / a0, e1 = fabs(out3[1] - r1) / a1, e2 = fabs(out3[2] - r2) / a2;
            printf("    out3 = %.9e %.9e %.9e\n    ref  = %.9e %.9e %.9e\n    err / sum|terms| = %.2e %.2e %.2e %s", out3[0], out3[1],
                   out3[2], r0, r1, r2, e0, e1, e2, (e0 < 1e-6 && e1 < 1e-6 && e2 < 1e-6) ? "OK" : "FAIL");
            bad |= !(e0 < 1e-6 && e1 < 1e-6 && e2 < 1e-6);
            if (store) {
                std::vector<float> Zo(X.size());
                cudaMemcpy(Zo.data(), dY2, Zo.size() * 4, cudaMemcpyDeviceToHost);
                double zerr = 0, zref = 0;
                for (size_t i = 0; i < Zo.size(); ++i) { zerr = fmax(zerr, fabs((double)Zo[i] - Zref[i])); zref = fmax(zref, fabs((double)Zref[i])); }
                printf("   Zout rel err %.3e %s", zerr / zref, zerr / zref < 2e-5 ? "OK" : "FAIL");
                bad |= !(zerr / zref < 2e-5);
            }
            printf("\n");
        }
    }
    cudaFree(dcols); cudaFree(ddirs); cudaFree(dX); cudaFree(dZ); cudaFree(dY); cudaFree(dY2); cudaFree(dout);
    return bad;
}

struct Bench {
    int d = 4;
    int64_t g[4] = {32, 32, 32, 32}, gmax = 32, m = 1 << 20, c;
    float *dcols, *ddirs, *dX, *dZ, *dY;
    double* dout;
    explicit Bench(int64_t c_) : c(c_) {
        std::vector<float> cols(d * gmax), dirs(d * gmax);
        for (int i = 0; i < d; ++i)
            for (int k = 0; k < 32; ++k) {
                cols[i * gmax + k] = expf(-0.004f * (1 + i) * k * k);
                dirs[i * gmax + k] = cols[i * gmax + k] * 0.01f * k * k;
            }
        cudaMalloc(&dcols, cols.size() * 4); cudaMalloc(&ddirs, dirs.size() * 4);
        cudaMalloc(&dX, (size_t)m * c * 4); cudaMalloc(&dZ, (size_t)m * c * 4); cudaMalloc(&dY, (size_t)m * c * 4);
        cudaMalloc(&dout, 24);
        cudaMemcpy(dcols, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(ddirs, dirs.data(), dirs.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dX, 0, (size_t)m * c * 4); cudaMemset(dZ, 0, (size_t)m * c * 4); cudaMemset(dout, 0, 24);
    }
    ~Bench() { cudaFree(dcols); cudaFree(ddirs); cudaFree(dX); cudaFree(dZ); cudaFree(dY); cudaFree(dout); }
};

static void timeit(int64_t c) {
    wiski_kron_tc_enable(0);      // the fused_* entry points below are the SIMT reference timings
    Bench b(c);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)b.m * c * 4;
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
        printf("%-40s c=%lld  %.3f ms  %.0f GB/s (algorithmic %.2f GB)\n", name, (long long)c, ms, nb / ms / 1e6, nb / 1e9);
    };
    const int axes[4][2] = {{0, 3}, {1, 2}, {0, 1}, {2, 3}};
    for (auto& ax : axes) {
        char nm[64];
        snprintf(nm, 64, "tc pair_apply axes (%d,%d)", ax[0], ax[1]);
        report(nm, 2 * bytes, [&] { wiski::tc_pair_apply_axes(b.dcols, b.d, b.g, b.gmax, ax[0], ax[1], b.dX, b.dY, c, 0, nullptr, nullptr); });
        snprintf(nm, 64, "tc grad_dir store axes (%d,%d)", ax[0], ax[1]);
        report(nm, 3 * bytes, [&] { wiski::tc_pair_grad_dir_axes(b.dcols, b.ddirs, b.d, b.g, b.gmax, ax[0], ax[1], b.dZ, b.dX, b.dY, c, b.dout, 0, nullptr); });
        snprintf(nm, 64, "tc grad_dir nostore axes (%d,%d)", ax[0], ax[1]);
        report(nm, 2 * bytes, [&] { wiski::tc_pair_grad_dir_axes(b.dcols, b.ddirs, b.d, b.g, b.gmax, ax[0], ax[1], b.dZ, b.dX, nullptr, c, b.dout, 0, nullptr); });
    }
    for (int pair = 0; pair < 2; ++pair) {
        char nm[64];
        snprintf(nm, 64, "simt pair_apply pair %d", pair);
        report(nm, 2 * bytes, [&] { wiski::fused_pair_apply(b.dcols, b.d, b.g, b.gmax, pair, b.dX, b.dY, c, 0, nullptr); });
        snprintf(nm, 64, "simt grad_dir store pair %d", pair);
        report(nm, 3 * bytes, [&] { wiski::fused_pair_grad_jvp(b.dcols, b.ddirs, b.d, b.g, b.gmax, pair, b.dZ, b.dX, b.dY, c, b.dout, 0, nullptr); });
    }
    // role / phase cycle breakdown of the apply kernel (PROF instantiation), averaged per CTA and tile
    long long* dprof;
    cudaMalloc(&dprof, 16 * 8);
    for (int k = 0; k < 2; ++k) {
        cudaMemset(dprof, 0, 16 * 8);
        wiski::tc_pair_apply_axes(b.dcols, b.d, b.g, b.gmax, axes[k][0], axes[k][1], b.dX, b.dY, c, 0, nullptr, dprof);
        cudaDeviceSynchronize();
        long long pr[16];
        cudaMemcpy(pr, dprof, 16 * 8, cudaMemcpyDeviceToHost);
        const double tiles = (double)b.m / 1024 * (c / 16);
        const char* names[13] = {"issuer0 wait a_ready", "issuer0 issue+commit", "w: wait full", "w: p1 load+split+st", "w: p1 wait d",
                                 "w: p1 ld+sts", "w: bar1", "w: p2 load+split+st", "w: bar2", "w: p2 wait d", "w: p2 ld+store",
                                 "(unused)", "producer wait empty"};
        printf("apply axes (%d,%d) cycles per tile (sum over CTAs / tiles):\n", axes[k][0], axes[k][1]);
        for (int i = 0; i < 13; ++i) printf("  %-24s %9.0f\n", names[i], pr[i] / tiles);
    }
    cudaFree(dprof);
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc > 1 && !strcmp(argv[1], "once")) {
        Bench b(argc > 2 ? atoll(argv[2]) : 432);     // forward (1,2) then (0,3); backward (0,3) with store, (1,2) without
        wiski::tc_pair_apply_axes(b.dcols, b.d, b.g, b.gmax, 1, 2, b.dX, b.dY, b.c, 0, nullptr, nullptr);
        wiski::tc_pair_apply_axes(b.dcols, b.d, b.g, b.gmax, 0, 3, b.dX, b.dY, b.c, 0, nullptr, nullptr);
        wiski::tc_pair_grad_dir_axes(b.dcols, b.ddirs, b.d, b.g, b.gmax, 0, 3, b.dZ, b.dX, b.dY, b.c, b.dout, 0, nullptr);
        wiski::tc_pair_grad_dir_axes(b.dcols, b.ddirs, b.d, b.g, b.gmax, 1, 2, b.dZ, b.dX, nullptr, b.c, b.dout, 0, nullptr);
        printf("once: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "time")) {
        timeit(argc > 2 ? atoll(argv[2]) : 432);
        return 0;
    }
    const Case cases[] = {
        {2, {32, 32, 1, 1}, 0, 1, 16, 1, 1},           // a single tile
        {2, {32, 32, 1, 1}, 0, 1, 432, 1, 1},          // 27 tiles
        {4, {4, 4, 32, 32}, 2, 3, 48, 1, 1},           // adjacent inner pair: after = 1, before = 16
        {4, {32, 32, 4, 4}, 0, 1, 32, 1, 1},           // adjacent outer pair: after = 16
        {4, {32, 4, 4, 32}, 0, 3, 32, 1, 1},           // outer / inner pairing: mid = 16
        {4, {4, 32, 32, 4}, 1, 2, 32, 1, 1},           // before = 4, after = 4
        {4, {32, 4, 4, 32}, 0, 3, 64, 2, 1},           // column-chunked input (sharded path: received blocks)
        {4, {4, 32, 32, 4}, 1, 2, 64, 1, 2},           // column-chunked output (send buffer)
        {4, {4, 32, 32, 4}, 1, 2, 64, 2, 1},           // chunked Z in the backward of the slab-local pair
        {3, {32, 4, 32, 1}, 0, 2, 32, 1, 1},           // d = 3: mid only
        {4, {32, 32, 32, 32}, 0, 3, 16, 1, 1},         // full 32^4, more tiles than SMs (1024 tiles)
        {4, {32, 32, 32, 32}, 1, 2, 16, 1, 1},
        {4, {32, 32, 32, 32}, 0, 1, 16, 1, 1},         // direct-store path (tile rows 1.7 MB apart at c = 432; 64 KB here)
    };
    for (const Case& cs : cases) bad |= run_case(cs);
    printf(bad ? "SOME CHECKS FAILED\n" : "ALL CHECKS PASSED\n");
    return bad;
}
