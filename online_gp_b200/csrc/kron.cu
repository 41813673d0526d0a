// k9 / k15: Kronecker-Toeplitz MVM on m x c panels and its gradient w.r.t. the Toeplitz columns (sm_100a).
//
// Replaces GPyTorch KroneckerProductLazyTensor._matmul + ToeplitzLazyTensor._matmul (FFT circulant embedding, one
// pass + two transposes per grid dimension) reached from
// online_gp/models/batched_fixed_noise_online_gp.py:348 (Kuu @ L), :366 (Kuu @ interpolation_cache) and
// online_gp/mlls/streaming_partial_mll.py:29, and the autograd of those products.
//
// Layout: X is [g_0, ..., g_{d-1}, c] row-major (grid axis 0 slowest, panel column fastest).  Applying factor i is
// a batch of independent "lines": line l = (o, w) with o in [0, prod_{j<i} g_j), w in [0, inner), inner =
// (prod_{j>i} g_j) * c; element b of the line sits at X[o*g_i*inner + b*inner + w].  Consecutive threads take
// consecutive w, so every load/store of a warp is one contiguous segment (no transposes, unlike the reference).
//   * g_i <= 32: register kernel — a thread owns a line, keeps the g_i outputs and the Toeplitz column in
//     registers, fully unrolled (direct apply: 2*g_i flop per element at 2*b bytes; see DESIGN.md roofline).
//   * otherwise: shared-memory tile kernel (any g_i).
#include "common.cuh"

namespace wiski {

// gemm_tc.cu: batched tensor-core GEMM forms of one Kronecker axis (fp32, g >= 64)
int64_t tc_axis_work_elems(int64_t g, int64_t outer, int64_t inner, int contract);
int tc_axis_apply_f32(const float* X, float* Y, const float* col, int64_t g, int64_t outer, int64_t inner, float* work,
                      cudaStream_t st);
int tc_axis_contract_f32(const float* Z, const float* P, int64_t g, int64_t outer, int64_t inner, double* acc64,
                         float* work, cudaStream_t st);


// kron_fused.cu: two axes per pass for 32-point axes (fp32)
bool fused_supported(int d, const int64_t* h_g, int64_t c);
int fused_kron_mm(const float* cols, int d, const int64_t* h_g, int64_t gmax, const float* X, int64_t c, float* Y,
                  float* work, cudaStream_t st);

// ------------------------------------------------------------------ register kernels (g <= G, G in {8,16,32})
template <typename T, int G>
__global__ void __launch_bounds__(128) axis_apply_reg_kernel(const T* __restrict__ X, T* __restrict__ Y,
                                                             const T* __restrict__ col, int g, int64_t inner,
                                                             int64_t nlines) {
    T t[G];
#pragma unroll
    for (int k = 0; k < G; ++k) t[k] = (k < g) ? col[k] : T(0);
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = l / inner, w = l - o * inner;
        const T* xp = X + o * g * inner + w;
        T* yp = Y + o * g * inner + w;
        T acc[G];
#pragma unroll
        for (int a = 0; a < G; ++a) acc[a] = T(0);
#pragma unroll
        for (int b0 = 0; b0 < G; b0 += 8) {
            T xb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) xb[j] = (b0 + j < g) ? xp[(int64_t)(b0 + j) * inner] : T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int a = 0; a < G; ++a) {
                    int k = a - (b0 + j);
                    k = k < 0 ? -k : k;
                    acc[a] += t[k] * xb[j];
                }
            }
        }
#pragma unroll
        for (int a = 0; a < G; ++a)
            if (a < g) yp[(int64_t)a * inner] = acc[a];
    }
}

// grad[k] += sum over lines of sum_{|a-b|=k} z[a] p[b]; per-block partial -> double atomics into acc64[g].
template <typename T, int G>
__global__ void __launch_bounds__(128) axis_contract_reg_kernel(const T* __restrict__ Z, const T* __restrict__ P,
                                                                int g, int64_t inner, int64_t nlines,
                                                                double* __restrict__ acc64) {
    T acc[G];
#pragma unroll
    for (int k = 0; k < G; ++k) acc[k] = T(0);
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = l / inner, w = l - o * inner;
        const T* zp = Z + o * g * inner + w;
        const T* pp = P + o * g * inner + w;
        T pv[G];
#pragma unroll
        for (int b = 0; b < G; ++b) pv[b] = (b < g) ? pp[(int64_t)b * inner] : T(0);
#pragma unroll
        for (int a0 = 0; a0 < G; a0 += 8) {
            T za[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) za[j] = (a0 + j < g) ? zp[(int64_t)(a0 + j) * inner] : T(0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int b = 0; b < G; ++b) {
                    int k = (a0 + j) - b;
                    k = k < 0 ? -k : k;
                    acc[k] += za[j] * pv[b];
                }
            }
        }
    }
    __shared__ T red[4][G];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        T v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < G && threadIdx.x < g) {
        double s = (double)red[0][threadIdx.x] + (double)red[1][threadIdx.x] + (double)red[2][threadIdx.x] +
                   (double)red[3][threadIdx.x];
        atomicAdd(&acc64[threadIdx.x], s);
    }
}

// ------------------------------------------------------------------ shared-memory tile kernels (any g)
// Block = (W lanes over lines) x (TY thread rows); tile[g][W+1] in smem; each thread produces AT consecutive
// outputs per step with a sliding window of the Toeplitz column.
template <typename T, int W, int TY>
__global__ void axis_apply_smem_kernel(const T* __restrict__ X, T* __restrict__ Y, const T* __restrict__ col, int g,
                                       int64_t inner, int64_t nlines) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);   // [g][W+1]
    T* cs = tile + (size_t)g * (W + 1);         // [g]
    constexpr int AT = 4;
    int tx = threadIdx.x % W, ty = threadIdx.x / W;
    for (int k = threadIdx.x; k < g; k += W * TY) cs[k] = col[k];
    for (int64_t l0 = (int64_t)blockIdx.x * W; l0 < nlines; l0 += (int64_t)gridDim.x * W) {
        int64_t l = l0 + tx;
        bool ok = l < nlines;
        int64_t o = ok ? l / inner : 0, w = ok ? l - o * inner : 0;
        const T* xp = X + o * g * inner + w;
        T* yp = Y + o * g * inner + w;
        __syncthreads();
        for (int b = ty; b < g; b += TY) tile[b * (W + 1) + tx] = ok ? xp[(int64_t)b * inner] : T(0);
        __syncthreads();
        for (int a0 = ty * AT; a0 < g; a0 += TY * AT) {
            T acc[AT];
#pragma unroll
            for (int j = 0; j < AT; ++j) acc[j] = T(0);
            for (int b = 0; b < g; ++b) {
                T xv = tile[b * (W + 1) + tx];
#pragma unroll
                for (int j = 0; j < AT; ++j) {
                    int k = a0 + j - b;
                    k = k < 0 ? -k : k;
                    acc[j] += (k < g ? cs[k] : T(0)) * xv;
                }
            }
            if (ok) {
#pragma unroll
                for (int j = 0; j < AT; ++j)
                    if (a0 + j < g) yp[(int64_t)(a0 + j) * inner] = acc[j];
            }
        }
    }
}

// Contraction for any g: thread k (strided) accumulates offset k over the tile's lines.
template <typename T, int W>
__global__ void axis_contract_smem_kernel(const T* __restrict__ Z, const T* __restrict__ P, int g, int64_t inner,
                                          int64_t nlines, double* __restrict__ acc64) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* zt = reinterpret_cast<T*>(smem_raw);     // [g][W+1]
    T* pt = zt + (size_t)g * (W + 1);           // [g][W+1]
    int nthr = blockDim.x;
    for (int64_t l0 = (int64_t)blockIdx.x * W; l0 < nlines; l0 += (int64_t)gridDim.x * W) {
        __syncthreads();
        for (int e = threadIdx.x; e < g * W; e += nthr) {
            int b = e / W, tx = e % W;
            int64_t l = l0 + tx;
            T zv = T(0), pv = T(0);
            if (l < nlines) {
                int64_t o = l / inner, w = l - o * inner;
                int64_t off = o * g * inner + (int64_t)b * inner + w;
                zv = Z[off];
                pv = P[off];
            }
            zt[b * (W + 1) + tx] = zv;
            pt[b * (W + 1) + tx] = pv;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < g; k += nthr) {
            double acc = 0.0;
            for (int tx = 0; tx < W; ++tx) {
                T part = T(0);
                for (int a = 0; a + k < g; ++a) {
                    T za = zt[a * (W + 1) + tx], pa = pt[a * (W + 1) + tx];
                    T zb = zt[(a + k) * (W + 1) + tx], pb = pt[(a + k) * (W + 1) + tx];
                    part += (k == 0) ? za * pa : (za * pb + zb * pa);
                }
                acc += (double)part;
            }
            atomicAdd(&acc64[k], acc);
        }
    }
}

template <typename T>
__global__ void zero_f64_kernel(double* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}
template <typename T>
__global__ void cvt_f64_kernel(const double* __restrict__ src, T* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (T)src[i];
}

template <typename T>
static int launch_axis_apply(const T* X, T* Y, const T* col, int64_t g, int64_t outer, int64_t inner,
                             cudaStream_t st) {
    int64_t nlines = outer * inner;
    if (nlines == 0) return 0;
    if (g <= 32) {
        int64_t blocks = ceil_div(nlines, 128);
        int64_t cap = (int64_t)kNumSMs * 32;
        if (blocks > cap) blocks = cap;
        if (g <= 8) axis_apply_reg_kernel<T, 8><<<(unsigned)blocks, 128, 0, st>>>(X, Y, col, (int)g, inner, nlines);
        else if (g <= 16) axis_apply_reg_kernel<T, 16><<<(unsigned)blocks, 128, 0, st>>>(X, Y, col, (int)g, inner, nlines);
        else axis_apply_reg_kernel<T, 32><<<(unsigned)blocks, 128, 0, st>>>(X, Y, col, (int)g, inner, nlines);
    } else {
        // tile width so that g*(W+1)*sizeof(T) stays within ~96 KB
        size_t budget = 96 * 1024;
        int W = 32;
        while (W > 4 && ((size_t)g * (W + 1) + g) * sizeof(T) > budget) W >>= 1;
        size_t smem = ((size_t)g * (W + 1) + g) * sizeof(T);
        WISKI_CHECK_ARG(smem <= 200 * 1024, "kron_toeplitz_mm: grid size %lld too large for the direct kernel",
                        (long long)g);
        int64_t blocks = ceil_div(nlines, W);
        int64_t cap = (int64_t)kNumSMs * 8;
        if (blocks > cap) blocks = cap;
#define LAUNCH_SMEM(WW, TYY)                                                                              \
    do {                                                                                                  \
        auto kfn = axis_apply_smem_kernel<T, WW, TYY>;                                                    \
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                         "kron_toeplitz_mm(attr)");                                                       \
        kfn<<<(unsigned)blocks, WW * TYY, smem, st>>>(X, Y, col, (int)g, inner, nlines);                  \
    } while (0)
        if (W == 32) LAUNCH_SMEM(32, 8);
        else if (W == 16) LAUNCH_SMEM(16, 16);
        else if (W == 8) LAUNCH_SMEM(8, 32);
        else LAUNCH_SMEM(4, 64);
#undef LAUNCH_SMEM
    }
    WISKI_CHECK_LAUNCH("kron_toeplitz_mm");
    count_launches(1);
    return 0;
}

template <typename T>
static int launch_axis_contract(const T* Z, const T* P, int64_t g, int64_t outer, int64_t inner, double* acc64,
                                cudaStream_t st) {
    int64_t nlines = outer * inner;
    if (nlines == 0) return 0;
    if (g <= 32) {
        int64_t blocks = ceil_div(nlines, 128);
        int64_t cap = (int64_t)kNumSMs * 16;
        if (blocks > cap) blocks = cap;
        if (g <= 8) axis_contract_reg_kernel<T, 8><<<(unsigned)blocks, 128, 0, st>>>(Z, P, (int)g, inner, nlines, acc64);
        else if (g <= 16) axis_contract_reg_kernel<T, 16><<<(unsigned)blocks, 128, 0, st>>>(Z, P, (int)g, inner, nlines, acc64);
        else axis_contract_reg_kernel<T, 32><<<(unsigned)blocks, 128, 0, st>>>(Z, P, (int)g, inner, nlines, acc64);
    } else {
        size_t budget = 96 * 1024;
        int W = 16;
        while (W > 2 && (size_t)2 * g * (W + 1) * sizeof(T) > budget) W >>= 1;
        size_t smem = (size_t)2 * g * (W + 1) * sizeof(T);
        WISKI_CHECK_ARG(smem <= 200 * 1024, "kron_toeplitz_bwd_cols: grid size %lld too large", (long long)g);
        int64_t blocks = ceil_div(nlines, W);
        int64_t cap = (int64_t)kNumSMs * 8;
        if (blocks > cap) blocks = cap;
        int threads = g >= 256 ? 256 : (int)(ceil_div(g, 32) * 32);
#define LAUNCH_C(WW)                                                                                      \
    do {                                                                                                  \
        auto kfn = axis_contract_smem_kernel<T, WW>;                                                      \
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                         "kron_toeplitz_bwd_cols(attr)");                                                 \
        kfn<<<(unsigned)blocks, threads, smem, st>>>(Z, P, (int)g, inner, nlines, acc64);                 \
    } while (0)
        if (W == 16) LAUNCH_C(16);
        else if (W == 8) LAUNCH_C(8);
        else if (W == 4) LAUNCH_C(4);
        else LAUNCH_C(2);
#undef LAUNCH_C
    }
    WISKI_CHECK_LAUNCH("kron_toeplitz_bwd_cols");
    count_launches(1);
    return 0;
}

template <typename T>
static int check_dims(int d, const int64_t* h_g, int64_t gmax, int64_t* m_out) {
    WISKI_CHECK_ARG(d >= 1 && d <= WISKI_MAX_DIMS, "kron_toeplitz: d=%d outside [1,%d]", d, WISKI_MAX_DIMS);
    int64_t m = 1;
    for (int i = 0; i < d; ++i) {
        WISKI_CHECK_ARG(h_g[i] >= 1 && h_g[i] <= gmax, "kron_toeplitz: g[%d]=%lld not in [1,gmax=%lld]", i,
                        (long long)h_g[i], (long long)gmax);
        m *= h_g[i];
    }
    *m_out = m;
    return 0;
}

// Y = K X with d axis passes, ping-ponging so that the last pass lands in Y.
template <typename T>
static int kron_mm(const T* cols, int d, const int64_t* h_g, int64_t gmax, const T* X, int64_t c, T* Y, T* work,
                   void* stream) {
    int64_t m;
    if (int rc = check_dims<T>(d, h_g, gmax, &m)) return rc;
    WISKI_CHECK_ARG(c >= 1, "kron_toeplitz_mm: c=%lld", (long long)c);
    WISKI_CHECK_ARG(X != Y && (d == 1 || (work != nullptr && work != X && work != Y)),
                    "kron_toeplitz_mm: X, Y, work must be distinct");
    cudaStream_t st = as_stream(stream);
    const T* src = X;
    int64_t outer = 1;
    for (int i = 0; i < d; ++i) {
        int64_t inner = (m / (outer * h_g[i])) * c;
        // passes remaining after this one: d-1-i; the last must write Y
        T* dst = ((d - 1 - i) % 2 == 0) ? Y : work;
        if (int rc = launch_axis_apply<T>(src, dst, cols + (int64_t)i * gmax, h_g[i], outer, inner, st)) return rc;
        src = dst;
        outer *= h_g[i];
    }
    return 0;
}

// grad_cols[i][k] = d/dcol_i[k] sum(Z * (K X)).
// S_i = T_{i+1}..T_{d-1} X (suffix chain, stored), Pz_i = T_0..T_{i-1} Z (prefix chain, ping-pong);
// grad_i = fold_{|a-b|} sum_lines Pz_i[a] S_i[b].
template <typename T>
static int kron_bwd_cols(const T* cols, int d, const int64_t* h_g, int64_t gmax, const T* Z, const T* X, int64_t c,
                         T* grad_cols, T* work, void* stream) {
    int64_t m;
    if (int rc = check_dims<T>(d, h_g, gmax, &m)) return rc;
    WISKI_CHECK_ARG(c >= 1 && work != nullptr, "kron_toeplitz_bwd_cols: bad arguments");
    cudaStream_t st = as_stream(stream);
    int64_t panel = m * c;
    // work layout: [0, d-1) suffix panels S_0..S_{d-2}; then two prefix ping-pong panels; the double accumulator
    // lives at the head of the second ping-pong panel's tail? -> keep it simple: it aliases nothing: we carve it
    // from the end of the last panel region (requires panel*sizeof(T) >= d*gmax*8, else a dedicated tail is used).
    T* suf = work;                                  // S_i at suf + i*panel, i in [0, d-1)
    T* pz0 = work + (int64_t)(d - 1) * panel;
    T* pz1 = pz0 + panel;
    // caller provides (d+1) panels + d*gmax doubles (+ alignment slack): wiski_kron_toeplitz_bwd_work_elems()
    uintptr_t acc_addr = (reinterpret_cast<uintptr_t>(pz1 + panel) + 7) & ~uintptr_t(7);
    double* acc64 = reinterpret_cast<double*>(acc_addr);
    int64_t nacc = (int64_t)d * gmax;
    zero_f64_kernel<T><<<(unsigned)ceil_div(nacc, 256), 256, 0, st>>>(acc64, nacc);
    // suffix chain
    {
        const T* src = X;
        for (int i = d - 2; i >= 0; --i) {
            int64_t outer = 1;
            for (int j = 0; j <= i; ++j) outer *= h_g[j];        // axes before i+1
            int64_t inner = (m / (outer * h_g[i + 1])) * c;
            T* dst = suf + (int64_t)i * panel;
            if (int rc = launch_axis_apply<T>(src, dst, cols + (int64_t)(i + 1) * gmax, h_g[i + 1], outer, inner, st))
                return rc;
            src = dst;
        }
    }
    // prefix chain + contractions
    {
        const T* pz = Z;
        int64_t outer = 1;
        for (int i = 0; i < d; ++i) {
            int64_t inner = (m / (outer * h_g[i])) * c;
            const T* S = (i == d - 1) ? X : suf + (int64_t)i * panel;
            if (int rc = launch_axis_contract<T>(pz, S, h_g[i], outer, inner, acc64 + (int64_t)i * gmax, st)) return rc;
            if (i < d - 1) {
                T* dst = (pz == pz0) ? pz1 : pz0;
                if (int rc = launch_axis_apply<T>(pz, dst, cols + (int64_t)i * gmax, h_g[i], outer, inner, st)) return rc;
                pz = dst;
            }
            outer *= h_g[i];
        }
    }
    cvt_f64_kernel<T><<<(unsigned)ceil_div(nacc, 256), 256, 0, st>>>(acc64, grad_cols, nacc);
    WISKI_CHECK_LAUNCH("kron_toeplitz_bwd_cols");
    count_launches(2);
    return 0;
}

}  // namespace wiski

extern "C" {
// single-axis building blocks (used by the row-sharded multi-GPU path, online_gp_b200/parallel.py)
int wiski_kron_axis_apply_f32(const float* X, float* Y, const float* col, int64_t g, int64_t outer, int64_t inner,
                              void* stream) {
    WISKI_CHECK_ARG(X != Y && g >= 1 && outer >= 0 && inner >= 1, "kron_axis_apply: bad arguments");
    return wiski::launch_axis_apply<float>(X, Y, col, g, outer, inner, wiski::as_stream(stream));
}
int wiski_kron_axis_apply_f64(const double* X, double* Y, const double* col, int64_t g, int64_t outer, int64_t inner,
                              void* stream) {
    WISKI_CHECK_ARG(X != Y && g >= 1 && outer >= 0 && inner >= 1, "kron_axis_apply: bad arguments");
    return wiski::launch_axis_apply<double>(X, Y, col, g, outer, inner, wiski::as_stream(stream));
}
int wiski_kron_axis_contract_f32(const float* Z, const float* P, int64_t g, int64_t outer, int64_t inner, double* acc64,
                                 void* stream) {
    WISKI_CHECK_ARG(g >= 1 && outer >= 0 && inner >= 1 && acc64 != nullptr, "kron_axis_contract: bad arguments");
    return wiski::launch_axis_contract<float>(Z, P, g, outer, inner, acc64, wiski::as_stream(stream));
}
int wiski_kron_axis_contract_f64(const double* Z, const double* P, int64_t g, int64_t outer, int64_t inner,
                                 double* acc64, void* stream) {
    WISKI_CHECK_ARG(g >= 1 && outer >= 0 && inner >= 1 && acc64 != nullptr, "kron_axis_contract: bad arguments");
    return wiski::launch_axis_contract<double>(Z, P, g, outer, inner, acc64, wiski::as_stream(stream));
}
// tensor-core (tcgen05) form of the two building blocks for axes with g >= 64 points (gemm_tc.cu)
int64_t wiski_kron_axis_tc_work_elems(int64_t g, int64_t outer, int64_t inner, int contract) {
    return wiski::tc_axis_work_elems(g, outer, inner, contract);
}
int wiski_kron_axis_apply_tc_f32(const float* X, float* Y, const float* col, int64_t g, int64_t outer, int64_t inner,
                                 float* work, void* stream) {
    WISKI_CHECK_ARG(X != Y && work != nullptr, "kron_axis_apply_tc: bad arguments");
    int rc = wiski::tc_axis_apply_f32(X, Y, col, g, outer, inner, work, wiski::as_stream(stream));
    if (rc == 3) wiski::set_error("kron_axis_apply_tc: shape g=%lld outer=%lld inner=%lld not supported", (long long)g,
                                  (long long)outer, (long long)inner);
    return rc;
}
int wiski_kron_axis_contract_tc_f32(const float* Z, const float* P, int64_t g, int64_t outer, int64_t inner,
                                    double* acc64, float* work, void* stream) {
    WISKI_CHECK_ARG(acc64 != nullptr && work != nullptr, "kron_axis_contract_tc: bad arguments");
    int rc = wiski::tc_axis_contract_f32(Z, P, g, outer, inner, acc64, work, wiski::as_stream(stream));
    if (rc == 3) wiski::set_error("kron_axis_contract_tc: shape g=%lld outer=%lld inner=%lld not supported", (long long)g,
                                  (long long)outer, (long long)inner);
    return rc;
}
int wiski_kron_toeplitz_mm_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, const float* X, int64_t c,
                               float* Y, float* work, void* stream) {
    if (d >= 1 && d <= WISKI_MAX_DIMS && wiski::fused_supported(d, h_g, c) && X != Y && work != nullptr && work != X &&
        work != Y)
        return wiski::fused_kron_mm(cols, d, h_g, gmax, X, c, Y, work, wiski::as_stream(stream));
    return wiski::kron_mm<float>(cols, d, h_g, gmax, X, c, Y, work, stream);
}
int wiski_kron_toeplitz_mm_f64(const double* cols, int d, const int64_t* h_g, int64_t gmax, const double* X,
                               int64_t c, double* Y, double* work, void* stream) {
    return wiski::kron_mm<double>(cols, d, h_g, gmax, X, c, Y, work, stream);
}
int64_t wiski_kron_toeplitz_bwd_work_elems(int d, int64_t m, int64_t c, int64_t gmax, int elem_size) {
    // (d+1) panels + d*gmax doubles (rounded up in elements), + 2 elements of alignment slack
    int64_t acc = ((int64_t)d * gmax * 8 + elem_size - 1) / elem_size;
    return (int64_t)(d + 1) * m * c + acc + 2;
}
int wiski_kron_toeplitz_bwd_cols_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, const float* Z,
                                     const float* X, int64_t c, float* grad_cols, float* work, void* stream) {
    return wiski::kron_bwd_cols<float>(cols, d, h_g, gmax, Z, X, c, grad_cols, work, stream);
}
int wiski_kron_toeplitz_bwd_cols_f64(const double* cols, int d, const int64_t* h_g, int64_t gmax, const double* Z,
                                     const double* X, int64_t c, double* grad_cols, double* work, void* stream) {
    return wiski::kron_bwd_cols<double>(cols, d, h_g, gmax, Z, X, c, grad_cols, work, stream);
}
}
