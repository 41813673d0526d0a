// Standalone check + timing of the tcgen05 3xTF32 Gram / panel-rmul kernels against a CPU fp64 reference.
// Build: nvcc ... test_gemm_tc.cu gemm_tc.o interp.o -o test_gemm_tc ; run on the B200 box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

namespace wiski {
int tc_gram_f32(const float*, const float*, int64_t, int64_t, int64_t, float*, float*, cudaStream_t, int64_t nblk = 1, bool symmetric = false);
int tc_panel_rmul_f32(const float*, int64_t, int64_t, const float*, int64_t, float*, cudaStream_t, int64_t nblk = 1, int terms = 3, float* work = nullptr, const PushDst* push = nullptr);
int64_t tc_gram_work_elems(int64_t, int64_t, int64_t);
int tc_panel_rmul_nt_f32(const float*, int64_t, int64_t, const float*, int64_t, float*, cudaStream_t);
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

static int check(int64_t m, int64_t r, int64_t r2) {
    std::vector<float> A(m * r), B(m * r2), Mm(r * r2);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand();
    for (auto& v : Mm) v = frand();
    float *dA, *dB, *dM, *dG, *dW, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dM, Mm.size() * 4);
    cudaMalloc(&dG, r * r2 * 4); cudaMalloc(&dO, m * r2 * 4);
    int64_t we = wiski::tc_gram_work_elems(m, r, r2);
    cudaMalloc(&dW, (we > 0 ? we : 1) * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dM, Mm.data(), Mm.size() * 4, cudaMemcpyHostToDevice);
    int rc = wiski::tc_gram_f32(dA, dB, m, r, r2, dG, dW, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("gram m=%lld r=%lld r2=%lld rc=%d (%s) cuda=%s\n", (long long)m, (long long)r, (long long)r2, rc,
           wiski_last_error(), cudaGetErrorString(e));
    if (rc == 0 && e == cudaSuccess) {
        std::vector<float> G(r * r2);
        cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int64_t i = 0; i < r; i += 7)
            for (int64_t j = 0; j < r2; j += 5) {
                double s = 0;
                for (int64_t k = 0; k < m; ++k) s += (double)A[k * r + i] * (double)B[k * r2 + j];
                maxerr = fmax(maxerr, fabs(s - G[i * r2 + j]));
                maxref = fmax(maxref, fabs(s));
            }
        printf("  gram max abs err %.3e (max |ref| %.3e, sqrt(m)=%.1f) rel-to-scale %.3e\n", maxerr, maxref, sqrt((double)m),
               maxerr / sqrt((double)m));
    }
    rc = wiski::tc_panel_rmul_f32(dA, m, r, dM, r2, dO, 0);
    e = cudaDeviceSynchronize();
    printf("rmul rc=%d (%s) cuda=%s\n", rc, wiski_last_error(), cudaGetErrorString(e));
    if (rc == 0 && e == cudaSuccess) {
        std::vector<float> O(m * r2);
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int64_t i = 0; i < m; i += 97)
            for (int64_t j = 0; j < r2; j += 3) {
                double s = 0;
                for (int64_t k = 0; k < r; ++k) s += (double)A[i * r + k] * (double)Mm[k * r2 + j];
                maxerr = fmax(maxerr, fabs(s - O[i * r2 + j]));
                maxref = fmax(maxref, fabs(s));
            }
        // last rows / columns explicitly
        for (int64_t i = m - 3; i < m; ++i)
            for (int64_t j = r2 - 3; j < r2; ++j) {
                double s = 0;
                for (int64_t k = 0; k < r; ++k) s += (double)A[i * r + k] * (double)Mm[k * r2 + j];
                maxerr = fmax(maxerr, fabs(s - O[i * r2 + j]));
            }
        printf("  rmul max abs err %.3e (max |ref| %.3e) rel-to-scale %.3e\n", maxerr, maxref, maxerr / sqrt((double)r));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dM); cudaFree(dG); cudaFree(dW); cudaFree(dO);
    return 0;
}

static void timeit(int64_t m, int64_t r) {
    float *dA, *dB, *dM, *dG, *dW, *dO;
    cudaMalloc(&dA, m * r * 4); cudaMalloc(&dB, m * r * 4); cudaMalloc(&dM, r * r * 4);
    cudaMalloc(&dG, r * r * 4); cudaMalloc(&dO, m * r * 4);
    cudaMalloc(&dW, wiski::tc_gram_work_elems(m, r, r) * 4);
    cudaMemset(dA, 0, m * r * 4); cudaMemset(dB, 0, m * r * 4); cudaMemset(dM, 0, r * r * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) { wiski::tc_gram_f32(dA, dB, m, r, r, dG, dW, 0); wiski::tc_panel_rmul_f32(dA, m, r, dM, r, dO, 0); }
    cudaDeviceSynchronize();
    float ms;
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) wiski::tc_gram_f32(dA, dB, m, r, r, dG, dW, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("time gram  m=%lld r=%lld: %.3f ms  (%.1f TFLOP/s fp32-equivalent)\n", (long long)m, (long long)r, ms / 5,
           2.0 * m * r * r / (ms / 5 * 1e-3) / 1e12);
    cudaEventRecord(e0);
    for (int it = 0; it < 5; ++it) wiski::tc_panel_rmul_f32(dA, m, r, dM, r, dO, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("time rmul  m=%lld r=%lld: %.3f ms  (%.1f TFLOP/s fp32-equivalent) cuda=%s\n", (long long)m, (long long)r, ms / 5,
           2.0 * m * r * r / (ms / 5 * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(dA); cudaFree(dB); cudaFree(dM); cudaFree(dG); cudaFree(dW); cudaFree(dO);
}

static void ones_probe(int64_t m, int64_t r) {
    std::vector<float> A(m * r, 1.0f), Mm(r * r, 1.0f);
    for (int64_t k = 0; k < m; ++k) for (int64_t i = 0; i < r; ++i) A[k * r + i] = 1.0f + (float)(i % 4);   // col pattern
    float *dA, *dM, *dG, *dW, *dO;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dM, Mm.size() * 4); cudaMalloc(&dG, r * r * 4); cudaMalloc(&dO, m * r * 4);
    cudaMalloc(&dW, wiski::tc_gram_work_elems(m, r, r) * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dM, Mm.data(), Mm.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dW, 0xff, wiski::tc_gram_work_elems(m, r, r) * 4);
    wiski::tc_gram_f32(dA, dA, m, r, r, dG, dW, 0);
    cudaDeviceSynchronize();
    std::vector<float> G(r * r), O(m * r);
    cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost);
    printf("ones gram m=%lld: expect G[i][j]=m*(1+i%%4)*(1+j%%4): ", (long long)m);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) printf("%g ", G[i * r + j]);
    printf("| G[40][70]=%g G[127][127]=%g\n", G[40 * r + 70], G[127 * r + 127]);
    wiski::tc_panel_rmul_nt_f32(dA, m, r, dM, r, dO, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    printf("ones rmul NT (both K-major): expect %g: ", 2.5 * r);
    for (int j = 0; j < 6; ++j) printf("%g ", O[j]);
    printf("| O[5][100]=%g O[m-1][r-1]=%g\n", O[5 * r + 100], O[(m - 1) * r + r - 1]);
    wiski::tc_panel_rmul_f32(dA, m, r, dM, r, dO, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    printf("ones rmul: expect %g: ", 2.5 * r);
    for (int j = 0; j < 6; ++j) printf("%g ", O[j]);
    printf("| O[5][100]=%g O[m-1][r-1]=%g\n", O[5 * r + 100], O[(m - 1) * r + r - 1]);
    cudaFree(dA); cudaFree(dM); cudaFree(dG); cudaFree(dW); cudaFree(dO);
}

int main() {
    ones_probe(4096, 128);
    check(8192, 128, 128);
    check(8192 + 40, 432, 432);
    check(20000, 512, 256);
    check(4096, 100, 64);
    timeit(1 << 20, 432);
    timeit(1 << 20, 512);
    return 0;
}
