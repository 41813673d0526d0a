// k5/k7/k10/k11: dense m x r panel kernels — right-multiply, row-local low-rank root update, Gram, fused Q-MVM
// and the conjugate-gradient driver (sm_100a).  SIMT (FFMA/DFMA) versions for every dtype and shape; the fp32
// tensor-core (tcgen05) versions of the two real contractions live in gemm_tc.cu and are dispatched from here.
//
// Reference operations replaced (online_gp/lazy/updated_root_lazy_tensor.py:79,97-100,115-117;
// online_gp/models/batched_fixed_noise_online_gp.py:346-361,375-376; GPyTorch linear_cg).
#include "common.cuh"

namespace wiski {

// gemm_tc.cu (fp32, tcgen05): return 0 if handled, 3 if the shape is not supported by the tensor-core path.
int tc_gram_f32(const float* A, const float* Bm, int64_t m, int64_t r, int64_t r2, float* G, float* work,
                cudaStream_t st, int64_t nblk = 1, bool symmetric = false);
int tc_panel_rmul_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, float* Out, cudaStream_t st,
                      int64_t nblk = 1, int terms = 3, float* work = nullptr, const PushDst* push = nullptr);
int64_t tc_rmul_work_elems(int64_t r, int64_t r2);
int64_t tc_gram_work_elems(int64_t m, int64_t r, int64_t r2);

// ------------------------------------------------------------------ Out[M x N] = P[M x K] @ Mm[K x N]
template <typename T>
__global__ void __launch_bounds__(256) rmul_tile_kernel(const T* __restrict__ P, const T* __restrict__ Mm,
                                                        T* __restrict__ Out, int64_t M, int64_t K, int64_t N) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN];
    int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    int64_t row0 = (int64_t)blockIdx.x * BM, col0 = (int64_t)blockIdx.y * BN;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int64_t k0 = 0; k0 < K; k0 += BK) {
        // A tile: 64 rows x 16 k  (4 elements per thread), stored k-major
        for (int e = threadIdx.x; e < BM * BK; e += 256) {
            int rr = e / BK, kk = e % BK;
            int64_t gr = row0 + rr, gk = k0 + kk;
            As[kk][rr] = (gr < M && gk < K) ? P[gr * K + gk] : T(0);
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            int kk = e / BN, cc = e % BN;
            int64_t gk = k0 + kk, gc = col0 + cc;
            Bs[kk][cc] = (gk < K && gc < N) ? Mm[gk * N + gc] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t gr = row0 + ty * 4 + i;
        if (gr >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t gc = col0 + tx + 16 * j;
            if (gc < N) Out[gr * N + gc] = acc[i][j];
        }
    }
}

// N small (<= 8): one warp per row, lanes strided over K.
template <typename T, int NP>
__global__ void rmul_skinny_kernel(const T* __restrict__ P, const T* __restrict__ Mm, T* __restrict__ Out, int64_t M,
                                   int64_t K, int N) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < M; i += 2 * nwarps) {
        const int64_t i2 = i + nwarps;
        const bool has2 = i2 < M;
        T acc[NP], acc2[NP];
#pragma unroll
        for (int t = 0; t < NP; ++t) { acc[t] = T(0); acc2[t] = T(0); }
#pragma unroll 4
        for (int64_t j = lane; j < K; j += 32) {
            T x = P[i * K + j];
            T x2 = has2 ? P[i2 * K + j] : T(0);
#pragma unroll
            for (int t = 0; t < NP; ++t)
                if (t < N) {
                    T mv = Mm[j * N + t];
                    acc[t] += x * mv;
                    acc2[t] += x2 * mv;
                }
        }
#pragma unroll
        for (int t = 0; t < NP; ++t) {
            T v = warp_sum(acc[t]);
            T v2 = warp_sum(acc2[t]);
            if (lane == 0 && t < N) {
                Out[i * N + t] = v;
                if (has2) Out[i2 * N + t] = v2;
            }
        }
    }
}

// ------------------------------------------------------------------ P[i,:] += (P[i,:] @ U) @ Vt   (row-local)
template <typename T, int QP>
__global__ void lowrank_update_kernel(T* __restrict__ P, int64_t m, int64_t r, const T* __restrict__ U,
                                      const T* __restrict__ Vt, int q) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < m; i += nwarps) {
        T* row = P + i * r;
        T dot[QP];
#pragma unroll
        for (int t = 0; t < QP; ++t) dot[t] = T(0);
        for (int64_t j = lane; j < r; j += 32) {
            T x = row[j];
#pragma unroll
            for (int t = 0; t < QP; ++t)
                if (t < q) dot[t] += x * U[j * q + t];
        }
#pragma unroll
        for (int t = 0; t < QP; ++t) dot[t] = warp_sum(dot[t]);
        for (int64_t j = lane; j < r; j += 32) {
            T x = row[j];
#pragma unroll
            for (int t = 0; t < QP; ++t)
                if (t < q) x += dot[t] * Vt[(int64_t)t * r + j];
            row[j] = x;
        }
    }
}

// ------------------------------------------------------------------ two panels, one launch, coefficients in smem
// P_k[i,:] += (P_k[i,:] @ U) @ Vt_k  for k = 0, 1 (root L and inverse root B of the rank-q update share U = p and differ
// in Vt: C p^T / C' p^T; updated_root_lazy_tensor.py:97-117).  U [r][q] and both Vt [q][r] are staged in shared memory
// ([QP][r] each, t-major: lane j reads word j -> conflict free); a warp owns ROWS rows at a time so that every
// coefficient read from shared memory feeds ROWS FMAs (q = 8: 2 x 8 FMA per element against 8 B of HBM traffic — the
// per-row kernel above re-reads U and Vt from L1 for every row and is L1-bound for q > 2).
// The ROWS x QP partial dots are reduced across the warp by recursive halving (N - 1 shuffles for N values).
// VEC elements per lane and load (16-byte vectors when r % VEC == 0: 4 x fewer load / LDS instructions and 4 x the bytes in
// flight per warp — with q = 8 the 48 KB of coefficients leave room for 16 warps per SM only, and scalar loads then kept
// just 8 KB per SM in flight: 33 % of HBM on the 128^3 / q = 8 workload before, see profiles/r02_*road3d*).
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) VecT {
    T v[VEC];
};
template <typename T, int QP, int ROWS, int VEC>
__global__ void __launch_bounds__(256) lowrank_update2_kernel(T* __restrict__ P0, T* __restrict__ P1, int64_t m, int64_t r,
                                                              const T* __restrict__ U, const T* __restrict__ Vt0,
                                                              const T* __restrict__ Vt1, int q, T* __restrict__ t_out) {
    extern __shared__ __align__(16) unsigned char smem_lr[];
    T* Us = reinterpret_cast<T*>(smem_lr);          // [QP][r]
    T* Vs = Us + (int64_t)QP * r;                    // [2][QP][r]
    T* red = Vs + (int64_t)2 * QP * r;               // [8 warps][ROWS * QP]
    constexpr int N = ROWS * QP;
    using V = VecT<T, VEC>;
    const int npanels = P1 != nullptr ? 2 : 1;
    for (int64_t e = threadIdx.x; e < (int64_t)QP * r; e += blockDim.x) {
        const int t = (int)(e / r);
        const int64_t j = e - (int64_t)t * r;
        Us[e] = t < q ? U[j * q + t] : T(0);
        Vs[e] = t < q ? Vt0[(int64_t)t * r + j] : T(0);
        if (npanels == 2) Vs[(int64_t)QP * r + e] = t < q ? Vt1[(int64_t)t * r + j] : T(0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    T* myred = red + wib * N;
    const int64_t groups_per_panel = (m + ROWS - 1) / ROWS;
    const int64_t ngroups = groups_per_panel * npanels;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t gidx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; gidx < ngroups; gidx += nwarps) {
        const int k = gidx >= groups_per_panel ? 1 : 0;
        const int64_t row0 = (gidx - (int64_t)k * groups_per_panel) * ROWS;
        T* base = (k ? P1 : P0) + row0 * r;
        const T* Vk = Vs + (int64_t)k * QP * r;
        const int nvalid = (int)((m - row0) < ROWS ? (m - row0) : ROWS);
        T dot[N];
#pragma unroll
        for (int i = 0; i < N; ++i) dot[i] = T(0);
        for (int64_t j = (int64_t)lane * VEC; j < r; j += 32 * VEC) {
            V x[ROWS];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                if (rr < nvalid) {
                    x[rr] = *reinterpret_cast<const V*>(base + (int64_t)rr * r + j);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) x[rr].v[e] = T(0);
                }
            }
#pragma unroll
            for (int t = 0; t < QP; ++t) {
                const V u = *reinterpret_cast<const V*>(Us + (int64_t)t * r + j);
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) dot[rr * QP + t] += x[rr].v[e] * u.v[e];
            }
        }
        // recursive halving: afterwards every lane holds the warp total of value `idx`
        int idx = 0;
        {
            int n = N;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                if (n > 1) {
                    const int half = n / 2;
                    const bool upper = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < N / 2; ++i) {
                        if (i < half) {
                            const T keep = upper ? dot[i + half] : dot[i];
                            const T send = upper ? dot[i] : dot[i + half];
                            dot[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    if (upper) idx += half;
                    n = half;
                } else {
                    dot[0] += __shfl_xor_sync(0xffffffffu, dot[0], o);
                }
            }
        }
        __syncwarp();
        myred[idx] = dot[0];             // lanes holding the same idx hold the same total
        __syncwarp();
        T dsum[N];
#pragma unroll
        for (int i = 0; i < N; ++i) dsum[i] = myred[i];
        if (t_out != nullptr && k == 0 && lane < N) {          // by-product: (P0 U)[row, t] of the rows BEFORE the update
            const int rr = lane / QP, t = lane - rr * QP;
            if (rr < nvalid && t < q) t_out[(row0 + rr) * q + t] = myred[lane];
        }
        for (int64_t j = (int64_t)lane * VEC; j < r; j += 32 * VEC) {
            V x[ROWS];
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                if (rr < nvalid) {
                    x[rr] = *reinterpret_cast<const V*>(base + (int64_t)rr * r + j);
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) x[rr].v[e] = T(0);
                }
            }
#pragma unroll
            for (int t = 0; t < QP; ++t) {
                const V v = *reinterpret_cast<const V*>(Vk + (int64_t)t * r + j);
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) x[rr].v[e] += dsum[rr * QP + t] * v.v[e];
            }
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr)
                if (rr < nvalid) *reinterpret_cast<V*>(base + (int64_t)rr * r + j) = x[rr];
        }
    }
}

// ------------------------------------------------------------------ Gram: G[r x r2] = A^T Bm, split over rows
template <typename T>
__global__ void __launch_bounds__(256) gram_tile_kernel(const T* __restrict__ A, const T* __restrict__ Bm, int64_t m,
                                                        int64_t r, int64_t r2, int64_t rows_per_split,
                                                        T* __restrict__ part) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ T As[BK][BM];
    __shared__ T Bs[BK][BN];
    int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    int64_t i0 = (int64_t)blockIdx.x * BM, j0 = (int64_t)blockIdx.y * BN;
    int64_t k_begin = (int64_t)blockIdx.z * rows_per_split;
    int64_t k_end = k_begin + rows_per_split;
    if (k_end > m) k_end = m;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
        for (int e = threadIdx.x; e < BK * BM; e += 256) {
            int kk = e / BM, cc = e % BM;
            int64_t gk = k0 + kk, gc = i0 + cc;
            As[kk][cc] = (gk < k_end && gc < r) ? A[gk * r + gc] : T(0);
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            int kk = e / BN, cc = e % BN;
            int64_t gk = k0 + kk, gc = j0 + cc;
            Bs[kk][cc] = (gk < k_end && gc < r2) ? Bm[gk * r2 + gc] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
    T* out = part + (int64_t)blockIdx.z * r * r2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t gi = i0 + ty + 16 * i;
        if (gi >= r) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t gj = j0 + tx + 16 * j;
            if (gj < r2) out[gi * r2 + gj] = acc[i][j];
        }
    }
}

// r2 small (<= 4): block accumulates A[i, j] * Bm[i, t] over its row slab; thread j strided over r.
template <typename T, int NP>
__global__ void gram_skinny_kernel(const T* __restrict__ A, const T* __restrict__ Bm, int64_t m, int64_t r, int r2,
                                   int64_t rows_per_block, T* __restrict__ part) {
    int64_t k_begin = (int64_t)blockIdx.x * rows_per_block;
    int64_t k_end = k_begin + rows_per_block;
    if (k_end > m) k_end = m;
    for (int64_t j = threadIdx.x; j < r; j += blockDim.x) {
        T acc[NP];
#pragma unroll
        for (int t = 0; t < NP; ++t) acc[t] = T(0);
        for (int64_t i = k_begin; i < k_end; ++i) {
            T a = A[i * r + j];
#pragma unroll
            for (int t = 0; t < NP; ++t)
                if (t < r2) acc[t] += a * Bm[i * r2 + t];
        }
#pragma unroll
        for (int t = 0; t < NP; ++t)
            if (t < r2) part[((int64_t)blockIdx.x * r + j) * r2 + t] = acc[t];
    }
}

// r2 small, r <= 32*RJ: warp per row, lanes strided over the r columns, per-lane register accumulators; rows are
// streamed fully coalesced (this is the L^T u half of the fused Q-MVM).  part: [gridDim.x][r][r2].
template <typename T, int RJ, int NP>
__global__ void __launch_bounds__(256) gram_skinny_warp_kernel(const T* __restrict__ A, const T* __restrict__ Bm,
                                                               int64_t m, int64_t r, int r2, T* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* red = reinterpret_cast<T*>(smem_raw);   // [8][r*r2]
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T acc[RJ][NP];
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj)
#pragma unroll
        for (int t = 0; t < NP; ++t) acc[jj][t] = T(0);
    int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
    for (int64_t i = gw; i < m; i += 2 * nw) {
        const int64_t i2 = i + nw;
        const bool has2 = i2 < m;
        T a0[RJ], a1[RJ];
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj) {
            int64_t j = lane + 32 * jj;
            a0[jj] = (j < r) ? A[i * r + j] : T(0);
            a1[jj] = (has2 && j < r) ? A[i2 * r + j] : T(0);
        }
        T b0[NP], b1[NP];
#pragma unroll
        for (int t = 0; t < NP; ++t) {
            b0[t] = (t < r2) ? Bm[i * r2 + t] : T(0);
            b1[t] = (has2 && t < r2) ? Bm[i2 * r2 + t] : T(0);
        }
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj)
#pragma unroll
            for (int t = 0; t < NP; ++t) acc[jj][t] += a0[jj] * b0[t] + a1[jj] * b1[t];
    }
    int64_t rc = r * r2;
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) {
        int64_t j = lane + 32 * jj;
#pragma unroll
        for (int t = 0; t < NP; ++t)
            if (j < r && t < r2) red[warp * rc + j * r2 + t] = acc[jj][t];
    }
    __syncthreads();
    for (int64_t e = threadIdx.x; e < rc; e += blockDim.x) {
        T sum = T(0);
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w * rc + e];
        part[(int64_t)blockIdx.x * rc + e] = sum;
    }
}

// out[e] = (init ? init[e] : 0) + sum_s part[s][e]   (summed in double)
template <typename T>
__global__ void reduce_parts_kernel(const T* __restrict__ part, int64_t nparts, int64_t n, const T* __restrict__ init,
                                    T* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double s = init ? (double)init[e] : 0.0;
    for (int64_t p = 0; p < nparts; ++p) s += (double)part[p * n + e];
    out[e] = (T)s;
}

// ------------------------------------------------------------------ fused Q-MVM partials
// Each warp walks rows i: t[cc] = KL[i,:] . v[:,cc];  acc[j][cc] += L[i,j] * t[cc].  One pass over both panels.
template <typename T, int RJ, int C>
__global__ void __launch_bounds__(256) qmv_kernel(const T* __restrict__ L, const T* __restrict__ KL, int64_t m,
                                                  int64_t r, const T* __restrict__ v, int c, T* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* red = reinterpret_cast<T*>(smem_raw);  // [8 warps][r*C]
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T vr[RJ][C], acc[RJ][C];
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) {
        int64_t j = lane + 32 * jj;
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            vr[jj][cc] = (j < r && cc < c) ? v[j * c + cc] : T(0);
            acc[jj][cc] = T(0);
        }
    }
    int64_t gw = (int64_t)blockIdx.x * 8 + warp, nw = (int64_t)gridDim.x * 8;
    for (int64_t i = gw; i < m; i += nw) {
        const T* kr = KL + i * r;
        const T* lr = L + i * r;
        T kv[RJ], lv[RJ];
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj) {
            int64_t j = lane + 32 * jj;
            kv[jj] = (j < r) ? kr[j] : T(0);
            lv[jj] = (j < r) ? lr[j] : T(0);
        }
        T t[C];
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            T s = T(0);
#pragma unroll
            for (int jj = 0; jj < RJ; ++jj) s += kv[jj] * vr[jj][cc];
            t[cc] = warp_sum(s);
        }
#pragma unroll
        for (int jj = 0; jj < RJ; ++jj)
#pragma unroll
            for (int cc = 0; cc < C; ++cc) acc[jj][cc] += lv[jj] * t[cc];
    }
    int64_t rc = r * c;
#pragma unroll
    for (int jj = 0; jj < RJ; ++jj) {
        int64_t j = lane + 32 * jj;
#pragma unroll
        for (int cc = 0; cc < C; ++cc)
            if (j < r && cc < c) red[warp * rc + j * c + cc] = acc[jj][cc];
    }
    __syncthreads();
    for (int64_t e = threadIdx.x; e < rc; e += blockDim.x) {
        T s = T(0);
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * rc + e];
        part[(int64_t)blockIdx.x * rc + e] = s;
    }
}

// ------------------------------------------------------------------ CG vector kernels (single block; r*c is tiny)
template <typename T>
struct CgState {   // all device pointers into the work buffer
    T* x; T* rs; T* p; T* Ap; T* rz; T* rhs_norm; T* resid;   // rz, rhs_norm [c]; resid [1] mean residual norm
};

template <typename T>
__device__ T block_sum(T v, T* sh) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    T s = T(0);
    int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) s += sh[w];
    return s;
}

template <typename T>
__global__ void cg_init_kernel(const T* __restrict__ rhs, int64_t r, int c, CgState<T> s) {
    __shared__ T sh[32];
    for (int cc = 0; cc < c; ++cc) {
        T a = T(0);
        for (int64_t j = threadIdx.x; j < r; j += blockDim.x) a += rhs[j * c + cc] * rhs[j * c + cc];
        T nrm = sqrt(block_sum(a, sh));
        if (nrm < T(1e-10)) nrm = T(1);
        T rz = T(0);
        for (int64_t j = threadIdx.x; j < r; j += blockDim.x) {
            T v = rhs[j * c + cc] / nrm;
            s.x[j * c + cc] = T(0);
            s.rs[j * c + cc] = v;
            s.p[j * c + cc] = v;
            rz += v * v;
        }
        rz = block_sum(rz, sh);
        if (threadIdx.x == 0) { s.rz[cc] = rz; s.rhs_norm[cc] = nrm; }
    }
    if (threadIdx.x == 0) s.resid[0] = T(1);
}

template <typename T>
__global__ void cg_update_kernel(int64_t r, int c, T tol, CgState<T> s) {
    __shared__ T sh[32];
    T mean = T(0);
    for (int cc = 0; cc < c; ++cc) {
        T a = T(0);
        for (int64_t j = threadIdx.x; j < r; j += blockDim.x) a += s.p[j * c + cc] * s.Ap[j * c + cc];
        T pAp = block_sum(a, sh);
        T rz = s.rz[cc];
        T alpha = (fabs(pAp) > T(1e-30)) ? rz / pAp : T(0);
        T rz_new = T(0);
        for (int64_t j = threadIdx.x; j < r; j += blockDim.x) {
            s.x[j * c + cc] += alpha * s.p[j * c + cc];
            T rv = s.rs[j * c + cc] - alpha * s.Ap[j * c + cc];
            s.rs[j * c + cc] = rv;
            rz_new += rv * rv;
        }
        rz_new = block_sum(rz_new, sh);
        T beta = (rz > T(1e-30)) ? rz_new / rz : T(0);
        for (int64_t j = threadIdx.x; j < r; j += blockDim.x)
            s.p[j * c + cc] = s.rs[j * c + cc] + beta * s.p[j * c + cc];
        __syncthreads();
        if (threadIdx.x == 0) s.rz[cc] = rz_new;
        mean += sqrt(rz_new);
    }
    if (threadIdx.x == 0) s.resid[0] = mean / T(c);
}

template <typename T>
__global__ void cg_finish_kernel(int64_t r, int c, CgState<T> s, T* __restrict__ xout) {
    for (int64_t e = threadIdx.x; e < r * c; e += blockDim.x) xout[e] = s.x[e] * s.rhs_norm[e % c];
}

// ------------------------------------------------------------------ host dispatch
template <typename T>
static int panel_rmul(const T* P, int64_t m, int64_t r, const T* M, int64_t r2, T* Out, void* stream) {
    WISKI_CHECK_ARG(m >= 0 && r >= 1 && r2 >= 1, "panel_rmul: bad sizes");
    WISKI_CHECK_ARG((const void*)P != (const void*)Out, "panel_rmul: Out must not alias P");
    if (m == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (r2 <= 8) {
        int64_t blocks = ceil_div(m, 8);
        const int64_t cap = (int64_t)kNumSMs * (background_mode() ? 2 : 16);
        if (blocks > cap) blocks = cap;
        if (r2 == 1) {
            WISKI_CHECK_CUDA(apply_background_carveout(rmul_skinny_kernel<T, 1>), "panel_rmul(carveout)");
            rmul_skinny_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(P, M, Out, m, r, (int)r2);
        }
        else if (r2 <= 2) rmul_skinny_kernel<T, 2><<<(unsigned)blocks, 256, 0, st>>>(P, M, Out, m, r, (int)r2);
        else if (r2 <= 4) rmul_skinny_kernel<T, 4><<<(unsigned)blocks, 256, 0, st>>>(P, M, Out, m, r, (int)r2);
        else rmul_skinny_kernel<T, 8><<<(unsigned)blocks, 256, 0, st>>>(P, M, Out, m, r, (int)r2);
    } else {
        dim3 grid((unsigned)ceil_div(m, 64), (unsigned)ceil_div(r2, 64));
        rmul_tile_kernel<T><<<grid, 256, 0, st>>>(P, M, Out, m, r, r2);
    }
    WISKI_CHECK_LAUNCH("panel_rmul");
    count_launches(1);
    return 0;
}

template <typename T>
static int lowrank_update(T* P, int64_t m, int64_t r, const T* U, const T* Vt, int64_t q, void* stream) {
    WISKI_CHECK_ARG(m >= 0 && r >= 1 && q >= 1 && q <= 32, "panel_lowrank_update: need 1 <= q <= 32 (q=%lld)",
                    (long long)q);
    if (m == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = ceil_div(m, 8);
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
#define LR(QP) lowrank_update_kernel<T, QP><<<(unsigned)blocks, 256, 0, st>>>(P, m, r, U, Vt, (int)q)
    if (q == 1) LR(1);
    else if (q <= 2) LR(2);
    else if (q <= 4) LR(4);
    else if (q <= 8) LR(8);
    else if (q <= 16) LR(16);
    else LR(32);
#undef LR
    WISKI_CHECK_LAUNCH("panel_lowrank_update");
    count_launches(1);
    return 0;
}

// P1 / Vt1 may be NULL (single panel).  Falls back to the per-row kernel when U and Vt do not fit in shared memory.
template <typename T>
static int lowrank_update2(T* P0, T* P1, int64_t m, int64_t r, const T* U, const T* Vt0, const T* Vt1, int64_t q, void* stream,
                           T* t_out = nullptr) {
    WISKI_CHECK_ARG(m >= 0 && r >= 1 && q >= 1 && q <= 32, "panel_lowrank_update2: need 1 <= q <= 32 (q=%lld)", (long long)q);
    WISKI_CHECK_ARG((P1 == nullptr) == (Vt1 == nullptr), "panel_lowrank_update2: P1 and Vt1 go together");
    if (m == 0) return 0;
    const int qp = q == 1 ? 1 : q <= 2 ? 2 : q <= 4 ? 4 : q <= 8 ? 8 : q <= 16 ? 16 : 32;
    const size_t smem = ((size_t)3 * qp * r + 8 * 32) * sizeof(T);
    if (smem > 200 * 1024) {
        if (t_out != nullptr) {
            if (int rc = panel_rmul<T>(P0, m, r, U, q, t_out, stream)) return rc;
        }
        if (int rc = lowrank_update<T>(P0, m, r, U, Vt0, q, stream)) return rc;
        return P1 != nullptr ? lowrank_update<T>(P1, m, r, U, Vt1, q, stream) : 0;
    }
    cudaStream_t st = as_stream(stream);
    const bool vec = r % (16 / (int64_t)sizeof(T)) == 0 && (reinterpret_cast<uintptr_t>(P0) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(P1) & 15) == 0;
#define LR2(QP, ROWS)                                                                                                        \
    do {                                                                                                                     \
        auto kfn = vec ? lowrank_update2_kernel<T, QP, ROWS, 16 / (int)sizeof(T)> : lowrank_update2_kernel<T, QP, ROWS, 1>;   \
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "lowrank2(attr)"); \
        int64_t groups = ceil_div(m, ROWS) * (P1 != nullptr ? 2 : 1);                                                        \
        int per_sm = smem > 100 * 1024 ? 1 : smem > 48 * 1024 ? 2 : 4;                                                       \
        if (background_mode()) per_sm = 1;                                                                                   \
        WISKI_CHECK_CUDA(apply_background_carveout(kfn), "lowrank2(carveout)");                                              \
        int64_t blocks = ceil_div(groups, 8);                                                                                \
        if (blocks > (int64_t)kNumSMs * per_sm) blocks = (int64_t)kNumSMs * per_sm;                                          \
        kfn<<<(unsigned)blocks, 256, smem, st>>>(P0, P1, m, r, U, Vt0, Vt1, (int)q, t_out);                                  \
    } while (0)
    if (qp == 1) LR2(1, 8);
    else if (qp == 2) LR2(2, 8);
    else if (qp == 4) LR2(4, 8);
    else if (qp == 8) LR2(8, 4);
    else if (qp == 16) LR2(16, 2);
    else LR2(32, 1);
#undef LR2
    WISKI_CHECK_LAUNCH("panel_lowrank_update2");
    count_launches(1);
    return 0;
}


// ------------------------------------------------------------------ P[m x c] += T[m x q] W[q x c]   (q <= 32)
// The column-sharded copy of the root panel follows the rank-q update through the m x q vector T = L p gathered from all
// ranks: one streaming pass (16-byte vectors, W in shared memory), HBM-bound.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) outer_add_kernel(T* __restrict__ P, int64_t m, int64_t c, const T* __restrict__ Tv,
                                                        int q, const T* __restrict__ W) {
    extern __shared__ __align__(16) unsigned char smem_oa[];
    T* Ws = reinterpret_cast<T*>(smem_oa);          // [q][c]
    for (int64_t e = threadIdx.x; e < (int64_t)q * c; e += blockDim.x) Ws[e] = W[e];
    __syncthreads();
    using V = VecT<T, VEC>;
    const int64_t vec_per_row = c / VEC;
    const int64_t total = m * vec_per_row;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / vec_per_row, j = (e - row * vec_per_row) * VEC;
        V x = *reinterpret_cast<const V*>(P + row * c + j);
        for (int k = 0; k < q; ++k) {
            const T t = Tv[row * q + k];
            const V w = *reinterpret_cast<const V*>(Ws + (int64_t)k * c + j);
#pragma unroll
            for (int v = 0; v < VEC; ++v) x.v[v] += t * w.v[v];
        }
        *reinterpret_cast<V*>(P + row * c + j) = x;
    }
}
template <typename T>
static int outer_add(T* P, int64_t m, int64_t c, const T* Tv, int64_t q, const T* W, void* stream) {
    WISKI_CHECK_ARG(m >= 0 && c >= 1 && q >= 1 && q <= 32, "panel_outer_add: need 1 <= q <= 32 (q=%lld)", (long long)q);
    if (m == 0) return 0;
    const size_t smem = (size_t)q * c * sizeof(T);
    WISKI_CHECK_ARG(smem <= 200 * 1024, "panel_outer_add: q x c coefficients do not fit in shared memory");
    constexpr int VEC = 16 / (int)sizeof(T);
    const bool vec = c % VEC == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0;
    auto kfn = vec ? outer_add_kernel<T, VEC> : outer_add_kernel<T, 1>;
    if (smem > 48 * 1024)
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "outer_add(attr)");
    const int64_t work = m * (c / (vec ? VEC : 1));
    int64_t blocks = ceil_div(work, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    kfn<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(P, m, c, Tv, (int)q, W);
    WISKI_CHECK_LAUNCH("panel_outer_add");
    count_launches(1);
    return 0;
}

static inline int64_t gram_splits(int64_t m, int64_t r, int64_t r2) {
    int64_t tiles = ceil_div(r, 64) * ceil_div(r2, 64);
    int64_t ks = ceil_div((int64_t)kNumSMs * 4, tiles);
    int64_t max_ks = ceil_div(m, 256);
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    return ks;
}
static inline int64_t gram_skinny_blocks(int64_t m) {
    int64_t nb = ceil_div(m, 64);
    const int64_t cap = (int64_t)kNumSMs * (background_mode() ? 1 : 4);
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return nb;
}

template <typename T>
static int gram(const T* A, const T* Bm, int64_t m, int64_t r, int64_t r2, T* G, T* work, void* stream) {
    WISKI_CHECK_ARG(m >= 1 && r >= 1 && r2 >= 1 && work != nullptr, "gram: bad arguments");
    cudaStream_t st = as_stream(stream);
    int64_t n = r * r2;
    if (r2 <= 2 && r <= 1024 && (size_t)8 * r * r2 * sizeof(T) <= 96 * 1024) {
        int64_t nb = gram_skinny_blocks(m);
        size_t smem = (size_t)8 * r * r2 * sizeof(T);
        int rj = (int)ceil_div(r, 32);
#define GSW(RJ, NP)                                                                                            \
    do {                                                                                                       \
        auto kfn = gram_skinny_warp_kernel<T, RJ, NP>;                                                         \
        WISKI_CHECK_CUDA(apply_background_carveout(kfn), "gram(carveout)");                                    \
        if (smem > 48 * 1024)                                                                                  \
            WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                             "gram(attr)");                                                                    \
        kfn<<<(unsigned)nb, 256, smem, st>>>(A, Bm, m, r, (int)r2, work);                                      \
    } while (0)
#define GSW_NP(RJ)                  \
    do {                            \
        if (r2 == 1) GSW(RJ, 1);    \
        else GSW(RJ, 2);            \
    } while (0)
        if (rj <= 4) GSW_NP(4);
        else if (rj <= 8) GSW_NP(8);
        else if (rj <= 16) GSW_NP(16);
        else GSW_NP(32);
#undef GSW_NP
#undef GSW
        reduce_parts_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(work, nb, n, (const T*)nullptr, G);
    } else if (r2 <= 4) {
        int64_t nb = gram_skinny_blocks(m);
        int64_t rpb = ceil_div(m, nb);
        nb = ceil_div(m, rpb);
        if (r2 == 1) gram_skinny_kernel<T, 1><<<(unsigned)nb, 256, 0, st>>>(A, Bm, m, r, (int)r2, rpb, work);
        else if (r2 == 2) gram_skinny_kernel<T, 2><<<(unsigned)nb, 256, 0, st>>>(A, Bm, m, r, (int)r2, rpb, work);
        else gram_skinny_kernel<T, 4><<<(unsigned)nb, 256, 0, st>>>(A, Bm, m, r, (int)r2, rpb, work);
        reduce_parts_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(work, nb, n, (const T*)nullptr, G);
    } else {
        int64_t ks = gram_splits(m, r, r2);
        int64_t rps = ceil_div(ceil_div(m, ks), 16) * 16;
        ks = ceil_div(m, rps);
        dim3 grid((unsigned)ceil_div(r, 64), (unsigned)ceil_div(r2, 64), (unsigned)ks);
        gram_tile_kernel<T><<<grid, 256, 0, st>>>(A, Bm, m, r, r2, rps, work);
        reduce_parts_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(work, ks, n, (const T*)nullptr, G);
    }
    WISKI_CHECK_LAUNCH("gram");
    count_launches(2);
    return 0;
}

static inline int64_t qmv_blocks(int64_t m) {
    int64_t nb = ceil_div(m, 8);
    if (nb > (int64_t)kNumSMs * 2) nb = (int64_t)kNumSMs * 2;
    if (nb < 1) nb = 1;
    return nb;
}

template <typename T>
static int q_matvec(const T* L, const T* KL, int64_t m, int64_t r, const T* v, int64_t c, T* w, T* work,
                    void* stream) {
    WISKI_CHECK_ARG(m >= 1 && r >= 1 && c >= 1 && work != nullptr, "q_matvec: bad arguments");
    WISKI_CHECK_ARG(r <= 1024, "q_matvec: fused kernel supports r <= 1024 (r=%lld)", (long long)r);
    WISKI_CHECK_ARG(c <= 4, "q_matvec: c <= 4 per call (c=%lld); split the right-hand sides", (long long)c);
    cudaStream_t st = as_stream(stream);
    int64_t nb = qmv_blocks(m);
    size_t smem = (size_t)8 * r * c * sizeof(T);
    WISKI_CHECK_ARG(smem <= 200 * 1024, "q_matvec: r*c too large for the fused kernel (r=%lld, c=%lld)", (long long)r,
                    (long long)c);
    int rj = (int)ceil_div(r, 32);
#define QMV(RJ, C)                                                                                          \
    do {                                                                                                    \
        auto kfn = qmv_kernel<T, RJ, C>;                                                                    \
        if (smem > 48 * 1024)                                                                               \
            WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                             "q_matvec(attr)");                                                             \
        kfn<<<(unsigned)nb, 256, smem, st>>>(L, KL, m, r, v, (int)c, work);                                 \
    } while (0)
#define QMV_C(RJ)                      \
    do {                               \
        if (c == 1) QMV(RJ, 1);        \
        else if (c == 2) QMV(RJ, 2);   \
        else QMV(RJ, 4);               \
    } while (0)
    if (rj <= 4) QMV_C(4);
    else if (rj <= 8) QMV_C(8);
    else if (rj <= 16) QMV_C(16);
    else QMV_C(32);
#undef QMV_C
#undef QMV
    int64_t n = r * c;
    reduce_parts_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(work, nb, n, v, w);
    WISKI_CHECK_LAUNCH("q_matvec");
    count_launches(2);
    return 0;
}

template <typename T>
static int cg_solve(const T* L, const T* KL, int64_t m, int64_t r, const T* rhs, int64_t c, T tol, int max_iter,
                    int check_every, T* x, int* h_iters, T* h_resid, T* work, void* stream) {
    WISKI_CHECK_ARG(m >= 1 && r >= 1 && c >= 1 && c <= 4 && work != nullptr && max_iter >= 1, "cg_solve: bad arguments");
    if (check_every < 1) check_every = 1;
    cudaStream_t st = as_stream(stream);
    int64_t n = r * c;
    CgState<T> s;
    s.x = work; s.rs = s.x + n; s.p = s.rs + n; s.Ap = s.p + n; s.rz = s.Ap + n; s.rhs_norm = s.rz + c;
    s.resid = s.rhs_norm + c;
    T* qwork = s.resid + 8;
    cg_init_kernel<T><<<1, 256, 0, st>>>(rhs, r, (int)c, s);
    int min_iter = max_iter - 1 < 10 ? max_iter - 1 : 10;
    int it = 0;
    T resid = T(1);
    while (it < max_iter) {
        if (int rc = q_matvec<T>(L, KL, m, r, s.p, c, s.Ap, qwork, stream)) return rc;
        cg_update_kernel<T><<<1, 256, 0, st>>>(r, (int)c, tol, s);
        count_launches(1);
        ++it;
        if (it % check_every == 0 || it == max_iter) {
            WISKI_CHECK_CUDA(cudaMemcpyAsync(&resid, s.resid, sizeof(T), cudaMemcpyDeviceToHost, st), "cg_solve");
            WISKI_CHECK_CUDA(cudaStreamSynchronize(st), "cg_solve");
            if (it >= min_iter && resid < tol) break;
        }
    }
    cg_finish_kernel<T><<<1, 256, 0, st>>>(r, (int)c, s, x);
    WISKI_CHECK_LAUNCH("cg_solve");
    count_launches(2);
    if (h_iters) *h_iters = it;
    if (h_resid) *h_resid = resid;
    return 0;
}

}  // namespace wiski

extern "C" {
int wiski_panel_rmul_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, float* Out, void* stream) {
    int rc = wiski::tc_panel_rmul_f32(P, m, r, M, r2, Out, wiski::as_stream(stream));
    if (rc != 3) return rc;
    return wiski::panel_rmul<float>(P, m, r, M, r2, Out, stream);
}
int wiski_panel_rmul_f64(const double* P, int64_t m, int64_t r, const double* M, int64_t r2, double* Out,
                         void* stream) {
    return wiski::panel_rmul<double>(P, m, r, M, r2, Out, stream);
}
int wiski_panel_lowrank_update_f32(float* P, int64_t m, int64_t r, const float* U, const float* Vt, int64_t q,
                                   void* stream) {
    return wiski::lowrank_update<float>(P, m, r, U, Vt, q, stream);
}
int wiski_panel_lowrank_update2_f32(float* P0, float* P1, int64_t m, int64_t r, const float* U, const float* Vt0,
                                    const float* Vt1, int64_t q, void* stream) {
    return wiski::lowrank_update2<float>(P0, P1, m, r, U, Vt0, Vt1, q, stream);
}
int wiski_panel_lowrank_update2_t_f32(float* P0, float* P1, int64_t m, int64_t r, const float* U, const float* Vt0,
                                      const float* Vt1, int64_t q, float* Tout, void* stream) {
    return wiski::lowrank_update2<float>(P0, P1, m, r, U, Vt0, Vt1, q, stream, Tout);
}
int wiski_panel_lowrank_update2_t_f64(double* P0, double* P1, int64_t m, int64_t r, const double* U, const double* Vt0,
                                      const double* Vt1, int64_t q, double* Tout, void* stream) {
    return wiski::lowrank_update2<double>(P0, P1, m, r, U, Vt0, Vt1, q, stream, Tout);
}
int wiski_panel_outer_add_f32(float* P, int64_t m, int64_t c, const float* T, int64_t q, const float* W, void* stream) {
    return wiski::outer_add<float>(P, m, c, T, q, W, stream);
}
int wiski_panel_outer_add_f64(double* P, int64_t m, int64_t c, const double* T, int64_t q, const double* W, void* stream) {
    return wiski::outer_add<double>(P, m, c, T, q, W, stream);
}
int wiski_panel_lowrank_update2_f64(double* P0, double* P1, int64_t m, int64_t r, const double* U, const double* Vt0,
                                    const double* Vt1, int64_t q, void* stream) {
    return wiski::lowrank_update2<double>(P0, P1, m, r, U, Vt0, Vt1, q, stream);
}
int wiski_panel_lowrank_update_f64(double* P, int64_t m, int64_t r, const double* U, const double* Vt, int64_t q,
                                   void* stream) {
    return wiski::lowrank_update<double>(P, m, r, U, Vt, q, stream);
}
int64_t wiski_gram_work_elems(int64_t m, int64_t r, int64_t r2) {
    int64_t simt = (r2 <= 4) ? wiski::gram_skinny_blocks(m) * r * r2 : wiski::gram_splits(m, r, r2) * r * r2;
    int64_t tc = wiski::tc_gram_work_elems(m, r, r2);
    return simt > tc ? simt : tc;
}
int wiski_gram_f32(const float* A, const float* Bm, int64_t m, int64_t r, int64_t r2, float* G, float* work,
                   void* stream) {
    int rc = wiski::tc_gram_f32(A, Bm, m, r, r2, G, work, wiski::as_stream(stream));
    if (rc != 3) return rc;
    return wiski::gram<float>(A, Bm, m, r, r2, G, work, stream);
}
int wiski_gram_f64(const double* A, const double* Bm, int64_t m, int64_t r, int64_t r2, double* G, double* work,
                   void* stream) {
    return wiski::gram<double>(A, Bm, m, r, r2, G, work, stream);
}
int wiski_gram_sym_f32(const float* A, const float* Bm, int64_t m, int64_t r, float* G, float* work, void* stream) {
    int rc = wiski::tc_gram_f32(A, Bm, m, r, r, G, work, wiski::as_stream(stream), 1, true);
    if (rc != 3) return rc;
    return wiski::gram<float>(A, Bm, m, r, r, G, work, stream);
}
int64_t wiski_panel_rmul_ex_work_elems(int64_t r, int64_t r2) { return wiski::tc_rmul_work_elems(r, r2); }
int wiski_panel_rmul_ex_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, int64_t nblk, int terms,
                            float* Out, float* work, void* stream) {
    WISKI_CHECK_ARG(nblk >= 1 && r2 % nblk == 0 && terms >= 1 && terms <= 3, "panel_rmul_ex: bad nblk / terms");
    WISKI_CHECK_ARG(terms != 2 || work != nullptr, "panel_rmul_ex: terms = 2 needs the scratch buffer");
    int rc = wiski::tc_panel_rmul_f32(P, m, r, M, r2, Out, wiski::as_stream(stream), nblk, terms, work);
    if (rc != 3) return rc;
    if (nblk == 1) return wiski::panel_rmul<float>(P, m, r, M, r2, Out, stream);
    wiski::set_error("panel_rmul_ex: shape m=%lld r=%lld r2=%lld nblk=%lld not supported by the tensor-core path",
                     (long long)m, (long long)r, (long long)r2, (long long)nblk);
    return 3;
}
/* Column-chunked variants for the row-sharded path (K L and its gradient live as the all-to-all's receive / send
 * buffer: nblk blocks [m, r2 / nblk], block j = columns [j r2/nblk, (j+1) r2/nblk)).  One tensor-core launch when the
 * shape allows it (fp32, (r2 / nblk) % 32 == 0), otherwise one plain call per block. */
int wiski_gram_chunked_f32(const float* A, const float* Bb, int64_t m, int64_t r, int64_t r2, int64_t nblk, float* G,
                           float* work, void* stream) {
    WISKI_CHECK_ARG(nblk >= 1 && r2 % nblk == 0, "gram_chunked: r2=%lld not divisible into %lld blocks", (long long)r2,
                    (long long)nblk);
    int rc = wiski::tc_gram_f32(A, Bb, m, r, r2, G, work, wiski::as_stream(stream), nblk);
    if (rc != 3) return rc;
    wiski::set_error("gram_chunked: shape m=%lld r=%lld r2=%lld nblk=%lld not supported by the tensor-core path",
                     (long long)m, (long long)r, (long long)r2, (long long)nblk);
    return 3;
}
int wiski_gram_chunked_sym_f32(const float* A, const float* Bb, int64_t m, int64_t r, int64_t nblk, float* G, float* work,
                               void* stream) {
    WISKI_CHECK_ARG(nblk >= 1 && r % nblk == 0, "gram_chunked_sym: r=%lld not divisible into %lld blocks", (long long)r,
                    (long long)nblk);
    int rc = wiski::tc_gram_f32(A, Bb, m, r, r, G, work, wiski::as_stream(stream), nblk, true);
    if (rc != 3) return rc;
    wiski::set_error("gram_chunked_sym: shape m=%lld r=%lld nblk=%lld not supported by the tensor-core path", (long long)m,
                     (long long)r, (long long)nblk);
    return 3;
}
int wiski_panel_rmul_chunked_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, int64_t nblk,
                                 float* Outb, void* stream) {
    WISKI_CHECK_ARG(nblk >= 1 && r2 % nblk == 0, "panel_rmul_chunked: r2=%lld not divisible into %lld blocks",
                    (long long)r2, (long long)nblk);
    int rc = wiski::tc_panel_rmul_f32(P, m, r, M, r2, Outb, wiski::as_stream(stream), nblk);
    if (rc != 3) return rc;
    wiski::set_error("panel_rmul_chunked: shape m=%lld r=%lld r2=%lld nblk=%lld not supported by the tensor-core path",
                     (long long)m, (long long)r, (long long)r2, (long long)nblk);
    return 3;
}
/* The same product with column block j written at dst[j] (row-major [m, r2 / n_dst]): dst[j] points into rank j's
 * NVLink-mapped receive buffer, so the row -> column layout change of the sharded backward happens in the GEMM's own
 * epilogue stores (no separate all-to-all pass). */
int wiski_panel_rmul_push_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, int terms, float* const* dst,
                              int n_dst, float* work, void* stream) {
    WISKI_CHECK_ARG(dst != nullptr && n_dst >= 2 && n_dst <= 8 && r2 % n_dst == 0 && terms >= 1 && terms <= 3,
                    "panel_rmul_push: bad destinations / terms");
    for (int j = 0; j < n_dst; ++j) WISKI_CHECK_ARG(dst[j] != nullptr, "panel_rmul_push: NULL destination %d", j);
    const wiski::PushDst pd{dst, n_dst, 1};
    int rc = wiski::tc_panel_rmul_f32(P, m, r, M, r2, nullptr, wiski::as_stream(stream), n_dst, terms, work, &pd);
    if (rc != 3) return rc;
    wiski::set_error("panel_rmul_push: shape m=%lld r=%lld r2=%lld n_dst=%d not supported by the tensor-core path",
                     (long long)m, (long long)r, (long long)r2, n_dst);
    return 3;
}
int64_t wiski_qmv_work_elems(int64_t m, int64_t r, int64_t c) { return wiski::qmv_blocks(m) * r * c; }
int wiski_q_matvec_f32(const float* L, const float* KL, int64_t m, int64_t r, const float* v, int64_t c, float* w,
                       float* work, void* stream) {
    return wiski::q_matvec<float>(L, KL, m, r, v, c, w, work, stream);
}
int wiski_q_matvec_f64(const double* L, const double* KL, int64_t m, int64_t r, const double* v, int64_t c,
                       double* w, double* work, void* stream) {
    return wiski::q_matvec<double>(L, KL, m, r, v, c, w, work, stream);
}
int64_t wiski_cg_work_elems(int64_t m, int64_t r, int64_t c) {
    return 4 * r * c + 2 * c + 8 + wiski::qmv_blocks(m) * r * c;
}
int wiski_cg_solve_f32(const float* L, const float* KL, int64_t m, int64_t r, const float* rhs, int64_t c,
                       float tol, int max_iter, int check_every, float* x, int* h_iters, float* h_resid,
                       float* work, void* stream) {
    return wiski::cg_solve<float>(L, KL, m, r, rhs, c, tol, max_iter, check_every, x, h_iters, h_resid, work, stream);
}
int wiski_cg_solve_f64(const double* L, const double* KL, int64_t m, int64_t r, const double* rhs, int64_t c,
                       double tol, int max_iter, int check_every, double* x, int* h_iters, double* h_resid,
                       double* work, void* stream) {
    return wiski::cg_solve<double>(L, KL, m, r, rhs, c, tol, max_iter, check_every, x, h_iters, h_resid, work, stream);
}
}
