// k9 / k15 fast path: two Kronecker axes per pass over an m x c fp32 panel, for grids whose axes have 32 points
// (BASELINE config 2: 32^4).  A CTA stages a [32 x 32 x 16-column] grid tile in shared memory (cp.async, padded so
// that both axis orientations are bank-conflict free), applies both symmetric-Toeplitz factors there and streams
// the result back, so K X costs d/2 panel passes instead of d and the column gradient 2.5 instead of 2d.
//
// Symmetric Toeplitz trick: T is centro-symmetric, so with s = x + Jx, a = x - Jx (J = reversal) the product splits
// into two half-size products  (T x)[i] = (P s)[i] + (M a)[i],  (T x)[31-i] = (P s)[i] - (M a)[i],  i < 16, with
// P[i][j] = (t|i-j| + t(31-i-j)) / 2, M[i][j] = (t|i-j| - t(31-i-j)) / 2: 512 FMA per line instead of 1024.
// P and M live in __constant__ memory (filled on the stream from the device-resident Toeplitz columns), so every
// FFMA takes its coefficient as a uniform constant-bank operand: no register or shared-memory traffic for them.
//
// Reference operations replaced: KroneckerProductLazyTensor._matmul / ToeplitzLazyTensor._matmul and their autograd
//  reached from online_gp/models/batched_fixed_noise_online_gp.py:348.
//
// NOTE: the constant coefficient slots are process-global; calls must be issued on one stream at a time (the host
// layer uses torch's current stream, like the reference's single default stream).
#include "common.cuh"

#include <stdlib.h>

namespace wiski {

// kron_tc.cu: the same two passes on the tensor pipe (tcgen05 / TMEM / TMA); return 3 for shapes they do not take
int tc_pair_apply(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X, float* Y,
                  int64_t c, cudaStream_t st, const int64_t* h_lay, long long* prof);
int tc_pair_grad_dir(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int pair, const float* Z,
                     const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st, const int64_t* h_lay);

// tensor-core pair kernels on (default) / off (SIMT kernels of this file): env WISKI_KRON_TC=0 or wiski_kron_tc_enable(0)
static int g_use_tc = -1;
static bool use_tc() {
    if (g_use_tc < 0) {
        const char* e = getenv("WISKI_KRON_TC");
        g_use_tc = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return g_use_tc != 0;
}

constexpr int G = 32;          // grid points per fused axis
constexpr int H = 16;          // half
constexpr int CB = 16;         // panel columns per tile
constexpr int UP = G * CB + 16;   // pitch (floats) of one u-group in shared memory: 32 rows of 16 + 16 pad
constexpr int TILE_FLOATS = G * UP;   // 16896 floats = 67584 B

__constant__ float c_coef[4][2][H][H];          // [slot][P|M][i][j]; slots 0,1 = T_u, T_v; 2,3 = direction matrices T'_u, T'_v
__device__ float g_coef_stage[4 * 2 * H * H];

__global__ void prep_coef_kernel(const float* __restrict__ col0, const float* __restrict__ col1,
                                 const float* __restrict__ col2, const float* __restrict__ col3) {
    int t = threadIdx.x;            // 256 threads: (i, j)
    int i = t / H, j = t % H;
    int d = i - j;
    d = d < 0 ? -d : d;
    const float* cols[4] = {col0, col1, col2, col3};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (cols[s] == nullptr) continue;
        float a = cols[s][d], b = cols[s][G - 1 - i - j];
        g_coef_stage[((s * 2 + 0) * H + i) * H + j] = 0.5f * (a + b);
        g_coef_stage[((s * 2 + 1) * H + i) * H + j] = 0.5f * (a - b);
    }
}

static int set_coefficients(const float* col0, const float* col1, cudaStream_t st, const float* dir0 = nullptr,
                            const float* dir1 = nullptr) {
    prep_coef_kernel<<<1, 256, 0, st>>>(col0, col1, dir0, dir1);
    void* stage = nullptr;
    WISKI_CHECK_CUDA(cudaGetSymbolAddress(&stage, g_coef_stage), "kron_fused(symbol)");
    const int nslots = (dir0 != nullptr) ? 4 : 2;
    WISKI_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_coef, stage, sizeof(float) * nslots * 2 * H * H, 0,
                                             cudaMemcpyDeviceToDevice, st), "kron_fused(coef)");
    count_launches(1);
    return 0;
}

// x <- T x for the symmetric Toeplitz factor in constant slot SLOT (in registers, fully unrolled)
template <int SLOT>
__device__ __forceinline__ void sym_apply32(float (&x)[G]) {
    float s[H], a[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        s[j] = x[j] + x[G - 1 - j];
        a[j] = x[j] - x[G - 1 - j];
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        float ys = 0.f, ya = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            ys = fmaf(c_coef[SLOT][0][i][j], s[j], ys);
            ya = fmaf(c_coef[SLOT][1][i][j], a[j], ya);
        }
        x[i] = ys + ya;
        x[G - 1 - i] = ys - ya;
    }
}

// two lines at once: every coefficient fetched from the constant bank feeds two FFMAs
template <int SLOT>
__device__ __forceinline__ void sym_apply32x2(float (&x0)[G], float (&x1)[G]) {
    float s0[H], a0[H], s1[H], a1[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        s0[j] = x0[j] + x0[G - 1 - j];
        a0[j] = x0[j] - x0[G - 1 - j];
        s1[j] = x1[j] + x1[G - 1 - j];
        a1[j] = x1[j] - x1[G - 1 - j];
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        float ys0 = 0.f, ya0 = 0.f, ys1 = 0.f, ya1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float cp = c_coef[SLOT][0][i][j], cm = c_coef[SLOT][1][i][j];
            ys0 = fmaf(cp, s0[j], ys0);
            ys1 = fmaf(cp, s1[j], ys1);
            ya0 = fmaf(cm, a0[j], ya0);
            ya1 = fmaf(cm, a1[j], ya1);
        }
        x0[i] = ys0 + ya0;
        x0[G - 1 - i] = ys0 - ya0;
        x1[i] = ys1 + ya1;
        x1[G - 1 - i] = ys1 - ya1;
    }
}

// ya = T_SA x and yb = T_SB x sharing the symmetric / antisymmetric parts of x
template <int SA, int SB>
__device__ __forceinline__ void sym_dual32(const float (&x)[G], float (&ya_)[G], float (&yb_)[G]) {
    float s[H], a[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
        s[j] = x[j] + x[G - 1 - j];
        a[j] = x[j] - x[G - 1 - j];
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
        float s0 = 0.f, a0 = 0.f, s1 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            s0 = fmaf(c_coef[SA][0][i][j], s[j], s0);
            a0 = fmaf(c_coef[SA][1][i][j], a[j], a0);
            s1 = fmaf(c_coef[SB][0][i][j], s[j], s1);
            a1 = fmaf(c_coef[SB][1][i][j], a[j], a1);
        }
        ya_[i] = s0 + a0;
        ya_[G - 1 - i] = s0 - a0;
        yb_[i] = s1 + a1;
        yb_[G - 1 - i] = s1 - a1;
    }
}

// acc[k] += sum_{|a-b|=k} z[a] p[b]
__device__ __forceinline__ void contract32(const float (&z)[G], const float (&p)[G], float (&acc)[G]) {
#pragma unroll
    for (int a = 0; a < G; ++a)
#pragma unroll
        for (int b = 0; b < G; ++b) {
            int k = a - b;
            k = k < 0 ? -k : k;
            acc[k] = fmaf(z[a], p[b], acc[k]);
        }
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Geometry of one axis pair (u slower, v faster; both of size 32):  row(u, v) = base + u*32*sv + v*sv,
// base = ob * (1024 * sv) + oa,  ob in [0, n_before), oa in [0, sv).  Tile id -> (ob, oa, column chunk), chunk fastest.
struct PairGeom {
    int64_t sv;          // row stride of v (product of the grid sizes after the pair)
    int64_t n_before;    // product of the grid sizes before the pair
    int64_t c;           // panel columns
    int64_t n_chunks;    // c / CB
    int64_t n_tiles;     // n_before * sv * n_chunks
};

// Where element (row, col) of an operand panel lives: plain row-major (ld = cw = c, cstride = 0) or column-chunked —
// columns cut into blocks of cw, block j a row-major [rows, cw] matrix at ptr + j * cstride (the send / receive
// layout of the row <-> column all-to-all of the sharded path, online_gp_b200/parallel.py).  cw % 16 == 0, so a
// 16-column tile never straddles two blocks.
struct PanelLay {
    int64_t ld, cw, cstride;
};
__device__ __forceinline__ int64_t tile_offset(const PanelLay& l, int64_t col0) {
    const int64_t j = col0 / l.cw;
    return j * l.cstride + (col0 - j * l.cw);
}

__device__ __forceinline__ void tile_coords(const PairGeom& g, int64_t tile, int64_t& rowbase, int64_t& col0) {
    int64_t cc = tile % g.n_chunks;
    int64_t o = tile / g.n_chunks;
    int64_t oa = o % g.sv, ob = o / g.sv;
    rowbase = ob * (1024 * g.sv) + oa;
    col0 = cc * CB;
}

// stage one [32][32][16] tile: 4096 16-byte pieces.  Thread t copies pieces t, t + NT, ...: piece q = (u, v, part)
// with part = q & 3, v = (q >> 2) & 31, u = q >> 7, so consecutive pieces of a thread differ by NT/128 in u only and
// both the global and the shared address advance by a constant stride (no per-piece index arithmetic).
// LAY = false: plain row-major [m, c] operands (ld = c, offset = col0), resolved at compile time — the single-GPU
// path pays nothing for the layout generality.
template <bool LAY>
__device__ __forceinline__ int64_t lay_ld(const PanelLay& l, const PairGeom& g) { return LAY ? l.ld : g.c; }
template <bool LAY>
__device__ __forceinline__ int64_t lay_off(const PanelLay& l, int64_t col0) { return LAY ? tile_offset(l, col0) : col0; }

template <int NT>
__device__ __forceinline__ void load_tile_async(float* buf, const float* __restrict__ X, const PairGeom& g,
                                                int64_t rowbase, int64_t off, int64_t ld) {
    static_assert(NT % 128 == 0, "NT must be a multiple of 128");
    constexpr int USTEP = NT / 128;
    const int q = threadIdx.x;
    const int part = q & 3, v = (q >> 2) & 31, u0 = q >> 7;
    const float* src = X + off + (rowbase + ((int64_t)u0 * G + v) * g.sv) * ld + part * 4;
    float* dst = buf + u0 * UP + v * CB + part * 4;
    const int64_t sstride = (int64_t)USTEP * G * g.sv * ld;
#pragma unroll
    for (int k = 0; k < G / USTEP; ++k) {
        cp_async16(dst, src);
        src += sstride;
        dst += USTEP * UP;
    }
}

// ------------------------------------------------------------------ Y = (T_u x T_v) X   (forward pair apply)
// slot 0 = factor of axis u, slot 1 = factor of axis v.
template <int NT, bool LAY>   // NT must be 256 (two lines per thread and phase)
__global__ void __launch_bounds__(NT, 1) pair_apply_kernel(const float* __restrict__ X, float* __restrict__ Y, PairGeom g,
                                                           PanelLay lx, PanelLay ly) {
    const int64_t ldx = lay_ld<LAY>(lx, g), ldy = lay_ld<LAY>(ly, g);
    extern __shared__ __align__(16) float smem[];
    int64_t tile = blockIdx.x;
    int it = 0;
    if (tile < g.n_tiles) {
        int64_t rb, c0;
        tile_coords(g, tile, rb, c0);
        load_tile_async<NT>(smem, X, g, rb, lay_off<LAY>(lx, c0), ldx);
    }
    cp_async_commit();
    const int64_t ustride = (int64_t)G * g.sv * ldy;      // elements between consecutive u rows of the output
    for (; tile < g.n_tiles; tile += gridDim.x, ++it) {
        float* buf = smem + (it & 1) * TILE_FLOATS;       // derived from the __shared__ base so that LDS/STS are emitted
        int64_t next = tile + gridDim.x;
        if (next < g.n_tiles) {
            int64_t rb, c0;
            tile_coords(g, next, rb, c0);
            load_tile_async<NT>(smem + ((it + 1) & 1) * TILE_FLOATS, X, g, rb, lay_off<LAY>(lx, c0), ldx);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        // phase 1: along v; lines (u, w) and (u + 16, w) of this thread   [NT == 256: two lines per thread]
        {
            const int l = threadIdx.x;
            const int w = l & (CB - 1), u = l / CB;
            float* p0 = buf + u * UP + w;
            float* p1 = p0 + (NT / CB) * UP;
            float x0[G], x1[G];
#pragma unroll
            for (int v = 0; v < G; ++v) { x0[v] = p0[v * CB]; x1[v] = p1[v * CB]; }
            sym_apply32x2<1>(x0, x1);
#pragma unroll
            for (int v = 0; v < G; ++v) { p0[v * CB] = x0[v]; p1[v * CB] = x1[v]; }
        }
        __syncthreads();
        // phase 2: along u; lines (v, w) and (v + 16, w); results go straight to global memory
        {
            int64_t rb, c0;
            tile_coords(g, tile, rb, c0);
            const int l = threadIdx.x;
            const int w = l & (CB - 1), v = l / CB;
            const float* p0 = buf + v * CB + w;
            const float* p1 = p0 + (NT / CB) * CB;
            float x0[G], x1[G];
#pragma unroll
            for (int u = 0; u < G; ++u) { x0[u] = p0[u * UP]; x1[u] = p1[u * UP]; }
            sym_apply32x2<0>(x0, x1);
            float* yp0 = Y + lay_off<LAY>(ly, c0) + (rb + (int64_t)v * g.sv) * ldy + w;
            float* yp1 = yp0 + (int64_t)(NT / CB) * g.sv * ldy;
#pragma unroll
            for (int u = 0; u < G; ++u) {
                *yp0 = x0[u];
                *yp1 = x1[u];
                yp0 += ustride;
                yp1 += ustride;
            }
        }
        __syncthreads();   // buffer may be refilled by the prefetch of the next iteration
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------ backward pair kernel
// Inputs: Z (incoming gradient side) and P (operand side), both m x c.  Per tile:
//   acc_u += contract_u(Z, T_v P);   acc_v += contract_v(T_u Z, P);   if STORE: Zout = T_v T_u Z.
template <bool STORE, bool LAY>
__global__ void __launch_bounds__(256, 1)
pair_grad_kernel(const float* __restrict__ Z, const float* __restrict__ P, float* __restrict__ Zout, PairGeom g,
                 double* __restrict__ acc_u64, double* __restrict__ acc_v64, PanelLay lz, PanelLay lp, PanelLay lo) {
    constexpr int NT = 256;
    const int64_t ldz = lay_ld<LAY>(lz, g), ldp = lay_ld<LAY>(lp, g), ldo = lay_ld<LAY>(lo, g);
    extern __shared__ __align__(16) float smem[];
    float* zt = smem;
    float* pt = smem + TILE_FLOATS;
    float* st = smem + 2 * TILE_FLOATS;      // scratch: T_v P
    float acc_u[G], acc_v[G];
#pragma unroll
    for (int k = 0; k < G; ++k) { acc_u[k] = 0.f; acc_v[k] = 0.f; }
    for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        int64_t rb, c0;
        tile_coords(g, tile, rb, c0);
        load_tile_async<NT>(zt, Z, g, rb, lay_off<LAY>(lz, c0), ldz);
        load_tile_async<NT>(pt, P, g, rb, lay_off<LAY>(lp, c0), ldp);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // A: st = T_v P  (lines along v)
        for (int l = threadIdx.x; l < G * CB; l += NT) {
            int w = l & (CB - 1), u = l / CB;
            float x[G];
#pragma unroll
            for (int v = 0; v < G; ++v) x[v] = pt[u * UP + v * CB + w];
            sym_apply32<1>(x);
#pragma unroll
            for (int v = 0; v < G; ++v) st[u * UP + v * CB + w] = x[v];
        }
        __syncthreads();
        // B: acc_u += contract_u(Z, st);  then Z <- T_u Z in place  (lines along u)
        for (int l = threadIdx.x; l < G * CB; l += NT) {
            int w = l & (CB - 1), v = l / CB;
            float z[G], p[G];
#pragma unroll
            for (int u = 0; u < G; ++u) { z[u] = zt[u * UP + v * CB + w]; p[u] = st[u * UP + v * CB + w]; }
            contract32(z, p, acc_u);
            sym_apply32<0>(z);
#pragma unroll
            for (int u = 0; u < G; ++u) zt[u * UP + v * CB + w] = z[u];
        }
        __syncthreads();
        // C: acc_v += contract_v(T_u Z, P);  STORE: Zout = T_v (T_u Z)   (lines along v)
        for (int l = threadIdx.x; l < G * CB; l += NT) {
            int w = l & (CB - 1), u = l / CB;
            float z[G], p[G];
#pragma unroll
            for (int v = 0; v < G; ++v) { z[v] = zt[u * UP + v * CB + w]; p[v] = pt[u * UP + v * CB + w]; }
            contract32(z, p, acc_v);
            if (STORE) {
                sym_apply32<1>(z);
                float* yp = Zout + lay_off<LAY>(lo, c0) + (rb + (int64_t)u * G * g.sv) * ldo + w;
#pragma unroll
                for (int v = 0; v < G; ++v) yp[(int64_t)v * g.sv * ldo] = z[v];
            }
        }
        __syncthreads();
    }
    // block reduction of the two column-gradient accumulators -> double atomics
    __shared__ float red[2][8][G];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        float a = warp_sum(acc_u[k]), b = warp_sum(acc_v[k]);
        if (lane == 0) { red[0][warp][k] = a; red[1][warp][k] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * G) {
        int which = threadIdx.x / G, k = threadIdx.x % G;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += (double)red[which][w][k];
        atomicAdd(which == 0 ? &acc_u64[k] : &acc_v64[k], s);
    }
}

// ------------------------------------------------------------------ backward pair kernel, directional form
// The loss depends on column i only through <grad_i, dcol_i> (dcol_i = d col_i / d lengthscale_i) and <grad_i, col_i>
// (every scale-type parameter), and  <grad_i, dcol_i> = sum_lines z^T T'_i p,  <grad_i, col_i> = <Z, K X>  for all i.
// So instead of the full 32-entry column gradient (1024 FMA per grid line) the pass applies the *direction* matrix
// T'_i (512 FMA) and takes a dot product.  Per tile:
//   1: S = T_v P (lines along v)            2: zu = T_u z, zd = T'_u z;  s_u += zd . S;  z <- zu   (lines along u)
//   3: s_v += (T'_v zu) . p;  s_scale += zu . S;  STORE: Zout = T_v zu                        (lines along v)
// out3: [s_u, s_v, s_scale] doubles (accumulated).
template <bool STORE>
__global__ void __launch_bounds__(256, 1)
pair_grad_jvp_kernel(const float* __restrict__ Z, const float* __restrict__ P, float* __restrict__ Zout, PairGeom g,
                     double* __restrict__ out3) {
    constexpr int NT = 256;
    extern __shared__ __align__(16) float smem[];
    float* zt = smem;
    float* pt = smem + TILE_FLOATS;
    float* st = smem + 2 * TILE_FLOATS;
    float su = 0.f, sv = 0.f, ss = 0.f;
    const int64_t vstride = g.sv * g.c;
    for (int64_t tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        int64_t rb, c0;
        tile_coords(g, tile, rb, c0);
        load_tile_async<NT>(zt, Z, g, rb, c0, g.c);
        load_tile_async<NT>(pt, P, g, rb, c0, g.c);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // 1: S = T_v P
#pragma unroll 1
        for (int l = threadIdx.x; l < G * CB; l += NT) {
            const int w = l & (CB - 1), u = l / CB;
            float x[G];
#pragma unroll
            for (int v = 0; v < G; ++v) x[v] = pt[u * UP + v * CB + w];
            sym_apply32<1>(x);
#pragma unroll
            for (int v = 0; v < G; ++v) st[u * UP + v * CB + w] = x[v];
        }
        __syncthreads();
        // 2: lines along u
#pragma unroll 1
        for (int l = threadIdx.x; l < G * CB; l += NT) {
            const int w = l & (CB - 1), v = l / CB;
            float z[G], zu[G], zd[G];
#pragma unroll
            for (int u = 0; u < G; ++u) z[u] = zt[u * UP + v * CB + w];
            sym_dual32<0, 2>(z, zu, zd);
#pragma unroll
            for (int u = 0; u < G; ++u) {
                su = fmaf(zd[u], st[u * UP + v * CB + w], su);
                zt[u * UP + v * CB + w] = zu[u];
            }
        }
        __syncthreads();
        // 3: lines along v
#pragma unroll 1
        for (int l = threadIdx.x; l < G * CB; l += NT) {
            const int w = l & (CB - 1), u = l / CB;
            float zu[G], zv[G], zd[G];
#pragma unroll
            for (int v = 0; v < G; ++v) {
                zu[v] = zt[u * UP + v * CB + w];
                ss = fmaf(zu[v], st[u * UP + v * CB + w], ss);
            }
            if (STORE) {
                sym_dual32<1, 3>(zu, zv, zd);
            } else {
#pragma unroll
                for (int v = 0; v < G; ++v) zd[v] = zu[v];
                sym_apply32<3>(zd);
            }
#pragma unroll
            for (int v = 0; v < G; ++v) sv = fmaf(zd[v], pt[u * UP + v * CB + w], sv);
            if (STORE) {
                float* yp = Zout + (rb + (int64_t)u * G * g.sv) * g.c + c0 + w;
#pragma unroll
                for (int v = 0; v < G; ++v) {
                    *yp = zv[v];
                    yp += vstride;
                }
            }
        }
        __syncthreads();
    }
    __shared__ float red[3][8];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    su = warp_sum(su); sv = warp_sum(sv); ss = warp_sum(ss);
    if (lane == 0) { red[0][warp] = su; red[1][warp] = sv; red[2][warp] = ss; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += (double)red[threadIdx.x][w];
        atomicAdd(&out3[threadIdx.x], s);
    }
}

// ------------------------------------------------------------------ host side
static bool make_geom(PairGeom& g, int d, const int64_t* h_g, int pair, int64_t c) {
    int u = 2 * pair, v = u + 1;
    if (v >= d || h_g[u] != G || h_g[v] != G || c % CB != 0) return false;
    g.sv = 1;
    for (int j = v + 1; j < d; ++j) g.sv *= h_g[j];
    g.n_before = 1;
    for (int j = 0; j < u; ++j) g.n_before *= h_g[j];
    g.c = c;
    g.n_chunks = c / CB;
    g.n_tiles = g.n_before * g.sv * g.n_chunks;
    return true;
}

// h_lay: 3 int64 per operand (ld, cw, cstride), nullptr = every operand plain row-major [m, c]
static bool make_lays(PanelLay* out, int n, const int64_t* h_lay, int64_t c) {
    for (int i = 0; i < n; ++i) {
        if (h_lay == nullptr) {
            out[i] = PanelLay{c, c, 0};
        } else {
            out[i] = PanelLay{h_lay[3 * i], h_lay[3 * i + 1], h_lay[3 * i + 2]};
            if (out[i].cw < CB || out[i].cw % CB != 0 || out[i].ld % 4 != 0 || out[i].cstride % 4 != 0 || out[i].ld < 1)
                return false;
        }
    }
    return true;
}

bool fused_supported(int d, const int64_t* h_g, int64_t c) {
    if (d < 2 || (d % 2) != 0 || c % CB != 0 || c < CB) return false;
    for (int i = 0; i < d; ++i)
        if (h_g[i] != G) return false;
    return true;
}

// Y = K X via d/2 fused pair passes (pairs applied last to first); `mid` receives the panel after every pair but the
// last one when d == 4 it is exactly X23 = (T_2 x T_3) X, which the backward pass reuses.
int fused_kron_mm(const float* cols, int d, const int64_t* h_g, int64_t gmax, const float* X, int64_t c, float* Y,
                  float* work, cudaStream_t st) {
    const int npairs = d / 2;
    const float* src = X;
    if (d == 4 && use_tc()) {
        // pairing (1,2) + (0,3): every tile then touches rows that are close in memory (32 consecutive rows per outer
        // index, or rows 32 apart), instead of the 1024 rows 1024 apart of the pair (0,1) — see DESIGN.md
        int rc = tc_pair_apply_axes(cols, d, h_g, gmax, 1, 2, X, work, c, st, nullptr, nullptr);
        if (rc == 0) rc = tc_pair_apply_axes(cols, d, h_g, gmax, 0, 3, work, Y, c, st, nullptr, nullptr);
        if (rc != 3) return rc;
    }
    size_t smem = 2 * TILE_FLOATS * sizeof(float);
    auto kfn = pair_apply_kernel<256, false>;
    WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_fused(attr)");
    for (int p = npairs - 1; p >= 0; --p) {
        PairGeom g;
        if (!make_geom(g, d, h_g, p, c)) { set_error("kron_fused: unsupported shape"); return 3; }
        // passes remaining after this one: p ; the last pass (p == 0) must write Y
        float* dst = (p % 2 == 0) ? Y : work;
        if (use_tc()) {
            const int rc = tc_pair_apply(cols, d, h_g, gmax, p, src, dst, c, st, nullptr, nullptr);
            if (rc == 0) { src = dst; continue; }
            if (rc != 3) return rc;
        }
        if (int rc = set_coefficients(cols + (int64_t)(2 * p) * gmax, cols + (int64_t)(2 * p + 1) * gmax, st)) return rc;
        int64_t grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
        const PanelLay plain = {c, c, 0};
        kfn<<<(unsigned)grid, 256, smem, st>>>(src, dst, g, plain, plain);   // <.., false>: layouts ignored
        WISKI_CHECK_LAUNCH("kron_fused(pair_apply)");
        count_launches(1);
        src = dst;
    }
    return 0;
}

// One forward pair pass: Y = (T_{2 pair} x T_{2 pair + 1}) X.
int fused_pair_apply(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X, float* Y,
                     int64_t c, cudaStream_t st, const int64_t* h_lay = nullptr) {
    if (use_tc()) {
        const int rc = tc_pair_apply(cols, d, h_g, gmax, pair, X, Y, c, st, h_lay, nullptr);
        if (rc != 3) return rc;
    }
    PairGeom g;
    if (!make_geom(g, d, h_g, pair, c)) { set_error("kron_fused: unsupported shape"); return 3; }
    PanelLay lay[2];
    if (!make_lays(lay, 2, h_lay, c)) { set_error("kron_fused: bad operand layout"); return 1; }
    if (int rc = set_coefficients(cols + (int64_t)(2 * pair) * gmax, cols + (int64_t)(2 * pair + 1) * gmax, st)) return rc;
    size_t smem = 2 * TILE_FLOATS * sizeof(float);
    auto kfn = (h_lay != nullptr) ? pair_apply_kernel<256, true> : pair_apply_kernel<256, false>;
    WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_fused(attr)");
    int64_t grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
    kfn<<<(unsigned)grid, 256, smem, st>>>(X, Y, g, lay[0], lay[1]);
    WISKI_CHECK_LAUNCH("kron_fused(pair_apply)");
    count_launches(1);
    return 0;
}

// One backward pair pass (see pair_grad_kernel).  acc_u64 / acc_v64: g doubles each, accumulated.
int fused_pair_grad(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* Z, const float* P,
                    float* Zout, int64_t c, double* acc_u64, double* acc_v64, cudaStream_t st,
                    const int64_t* h_lay = nullptr) {
    PairGeom g;
    if (!make_geom(g, d, h_g, pair, c)) { set_error("kron_fused: unsupported shape"); return 3; }
    PanelLay lay[3];
    if (!make_lays(lay, 3, h_lay, c)) { set_error("kron_fused: bad operand layout"); return 1; }
    if (int rc = set_coefficients(cols + (int64_t)(2 * pair) * gmax, cols + (int64_t)(2 * pair + 1) * gmax, st)) return rc;
    size_t smem = 3 * TILE_FLOATS * sizeof(float);
    int64_t grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
    if (Zout != nullptr) {
        auto kfn = (h_lay != nullptr) ? pair_grad_kernel<true, true> : pair_grad_kernel<true, false>;
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_fused(attr)");
        kfn<<<(unsigned)grid, 256, smem, st>>>(Z, P, Zout, g, acc_u64, acc_v64, lay[0], lay[1], lay[2]);
    } else {
        auto kfn = (h_lay != nullptr) ? pair_grad_kernel<false, true> : pair_grad_kernel<false, false>;
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_fused(attr)");
        kfn<<<(unsigned)grid, 256, smem, st>>>(Z, P, nullptr, g, acc_u64, acc_v64, lay[0], lay[1], lay[2]);
    }
    WISKI_CHECK_LAUNCH("kron_fused(pair_grad)");
    count_launches(1);
    return 0;
}

// Directional backward pair pass (see pair_grad_jvp_kernel).  out3: 3 doubles, accumulated.
int fused_pair_grad_jvp(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int pair,
                        const float* Z, const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st,
                        const int64_t* h_lay = nullptr) {
    if (use_tc()) {
        const int rc = tc_pair_grad_dir(cols, dirs, d, h_g, gmax, pair, Z, P, Zout, c, out3, st, h_lay);
        if (rc != 3) return rc;
    }
    if (h_lay != nullptr) { set_error("kron_fused: the SIMT directional pass takes plain row-major operands only"); return 3; }
    PairGeom g;
    if (!make_geom(g, d, h_g, pair, c)) { set_error("kron_fused: unsupported shape"); return 3; }
    if (int rc = set_coefficients(cols + (int64_t)(2 * pair) * gmax, cols + (int64_t)(2 * pair + 1) * gmax, st,
                                  dirs + (int64_t)(2 * pair) * gmax, dirs + (int64_t)(2 * pair + 1) * gmax))
        return rc;
    size_t smem = 3 * TILE_FLOATS * sizeof(float);
    int64_t grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
    if (Zout != nullptr) {
        auto kfn = pair_grad_jvp_kernel<true>;
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_fused(attr)");
        kfn<<<(unsigned)grid, 256, smem, st>>>(Z, P, Zout, g, out3);
    } else {
        auto kfn = pair_grad_jvp_kernel<false>;
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_fused(attr)");
        kfn<<<(unsigned)grid, 256, smem, st>>>(Z, P, nullptr, g, out3);
    }
    WISKI_CHECK_LAUNCH("kron_fused(pair_grad_jvp)");
    count_launches(1);
    return 0;
}

}  // namespace wiski

extern "C" {
int wiski_kron_tc_enable(int on) {
    const int prev = wiski::use_tc() ? 1 : 0;
    if (on >= 0) wiski::g_use_tc = on ? 1 : 0;          // on < 0: query only
    return prev;
}

int wiski_kron_pair_apply_axes_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int axis_u, int axis_v,
                                   const float* X, float* Y, int64_t c, const int64_t* h_lay, void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || axis_u < 0 || axis_v <= axis_u || axis_v >= d || X == Y) {
        wiski::set_error("kron_pair_apply_axes: bad axes / operands");
        return 1;
    }
    if (wiski::use_tc()) {
        const int rc = wiski::tc_pair_apply_axes(cols, d, h_g, gmax, axis_u, axis_v, X, Y, c, wiski::as_stream(stream), h_lay,
                                                 nullptr);
        if (rc != 3) return rc;
    }
    if (axis_v == axis_u + 1 && axis_u % 2 == 0)          // the SIMT kernels know the pairs (2p, 2p + 1) only
        return wiski::fused_pair_apply(cols, d, h_g, gmax, axis_u / 2, X, Y, c, wiski::as_stream(stream), h_lay);
    wiski::set_error("kron_pair_apply_axes: axes (%d, %d) need the tensor-core path", axis_u, axis_v);
    return 3;
}

int wiski_kron_pair_grad_dir_axes_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                      int axis_u, int axis_v, const float* Z, const float* P, float* Zout, int64_t c,
                                      double* out3, const int64_t* h_lay, void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || axis_u < 0 || axis_v <= axis_u || axis_v >= d || dirs == nullptr) {
        wiski::set_error("kron_pair_grad_dir_axes: bad axes / operands");
        return 1;
    }
    if (wiski::use_tc()) {
        const int rc = wiski::tc_pair_grad_dir_axes(cols, dirs, d, h_g, gmax, axis_u, axis_v, Z, P, Zout, c, out3,
                                                    wiski::as_stream(stream), h_lay);
        if (rc != 3) return rc;
    }
    if (axis_v == axis_u + 1 && axis_u % 2 == 0)
        return wiski::fused_pair_grad_jvp(cols, dirs, d, h_g, gmax, axis_u / 2, Z, P, Zout, c, out3, wiski::as_stream(stream),
                                          h_lay);
    wiski::set_error("kron_pair_grad_dir_axes: axes (%d, %d) need the tensor-core path", axis_u, axis_v);
    return 3;
}

/* Pushing variants for the row-sharded multi-GPU path (tensor-core kernels only; 3 = shape not supported): the result
 * leaves through TMA stores into the peers' NVLink-mapped buffers, dst[j] = where this rank's part starts on rank j.
 *   mode 1: X is a row slab [m_loc, c]; column block j (c / n_dst columns) goes to dst[j], a [m_loc, c / n_dst] panel
 *   mode 2: X is a column block [m, c] (all rows), axis_u = 0; the rows of axis-0 range j go to dst[j], a [m / n_dst, c] panel
 * h_lay_x: (ld, cw, cstride) of X or NULL (plain). */
int wiski_kron_pair_apply_push_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int axis_u, int axis_v,
                                   const float* X, int64_t c, const int64_t* h_lay_x, float* const* dst, int n_dst, int mode,
                                   void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || axis_u < 0 || axis_v <= axis_u || axis_v >= d || dst == nullptr || n_dst < 1 || n_dst > 8) {
        wiski::set_error("kron_pair_apply_push: bad axes / destinations");
        return 1;
    }
    int64_t lay[6] = {c, c, 0, c, c, 0};
    if (h_lay_x != nullptr) { lay[0] = h_lay_x[0]; lay[1] = h_lay_x[1]; lay[2] = h_lay_x[2]; }
    const wiski::PushDst pd{dst, n_dst, mode};
    const int rc = wiski::tc_pair_apply_axes(cols, d, h_g, gmax, axis_u, axis_v, X, nullptr, c, wiski::as_stream(stream), lay,
                                             nullptr, &pd);
    if (rc == 3) wiski::set_error("kron_pair_apply_push: axes (%d, %d), c=%lld, n_dst=%d, mode %d not supported", axis_u, axis_v,
                                  (long long)c, n_dst, mode);
    return rc;
}

/* Directional backward pair pass whose Zout = T_v T_u Z is pushed (mode 2: axis_u = 0, rows of axis-0 range j -> dst[j]).
 * h_lay_zp: (ld, cw, cstride) of Z then of P, or NULL (both plain [m, c]). */
int wiski_kron_pair_grad_dir_push_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                      int axis_u, int axis_v, const float* Z, const float* P, int64_t c, double* out3,
                                      const int64_t* h_lay_zp, float* const* dst, int n_dst, void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || axis_u < 0 || axis_v <= axis_u || axis_v >= d || dirs == nullptr || dst == nullptr ||
        n_dst < 1 || n_dst > 8) {
        wiski::set_error("kron_pair_grad_dir_push: bad axes / operands");
        return 1;
    }
    int64_t lay[9] = {c, c, 0, c, c, 0, c, c, 0};
    if (h_lay_zp != nullptr)
        for (int i = 0; i < 6; ++i) lay[i] = h_lay_zp[i];
    const wiski::PushDst pd{dst, n_dst, 2};
    const int rc = wiski::tc_pair_grad_dir_axes(cols, dirs, d, h_g, gmax, axis_u, axis_v, Z, P, nullptr, c, out3,
                                                wiski::as_stream(stream), lay, &pd);
    if (rc == 3) wiski::set_error("kron_pair_grad_dir_push: axes (%d, %d), c=%lld, n_dst=%d not supported", axis_u, axis_v,
                                  (long long)c, n_dst);
    return rc;
}

int wiski_kron_fused_pair_grad_dir_lay_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                           int pair, const float* Z, const float* P, float* Zout, int64_t c, double* out3,
                                           const int64_t* h_lay, void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || pair < 0 || 2 * pair + 1 >= d || dirs == nullptr || h_lay == nullptr) {
        wiski::set_error("kron_fused_pair_grad_dir_lay: unsupported shape");
        return 3;
    }
    return wiski::fused_pair_grad_jvp(cols, dirs, d, h_g, gmax, pair, Z, P, Zout, c, out3, wiski::as_stream(stream), h_lay);
}

int wiski_kron_fused_pair_grad_dir_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                       int pair, const float* Z, const float* P, float* Zout, int64_t c, double* out3,
                                       void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || pair < 0 || 2 * pair + 1 >= d || dirs == nullptr) {
        wiski::set_error("kron_fused_pair_grad_dir: unsupported shape");
        return 3;
    }
    return wiski::fused_pair_grad_jvp(cols, dirs, d, h_g, gmax, pair, Z, P, Zout, c, out3, wiski::as_stream(stream));
}

/* Fused two-axes-per-pass Kronecker-Toeplitz MVM (fp32, every grid axis 32 points, d even, c % 16 == 0).
 * Returns 3 when the shape is not supported (callers fall back to wiski_kron_toeplitz_mm_f32). */
int wiski_kron_fused_supported(int d, const int64_t* h_g, int64_t c) { return wiski::fused_supported(d, h_g, c) ? 1 : 0; }

int wiski_kron_fused_pair_apply_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X,
                                    float* Y, int64_t c, void* stream) {
    // only the two axes of the requested pair must have 32 points (a row-sharded slab has a shorter axis 0)
    if (d < 2 || d > WISKI_MAX_DIMS || pair < 0 || 2 * pair + 1 >= d || X == Y) {
        wiski::set_error("kron_fused_pair_apply: unsupported shape");
        return 3;
    }
    return wiski::fused_pair_apply(cols, d, h_g, gmax, pair, X, Y, c, wiski::as_stream(stream));
}

int wiski_kron_fused_pair_apply_lay_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair,
                                        const float* X, float* Y, int64_t c, const int64_t* h_lay, void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || pair < 0 || 2 * pair + 1 >= d || X == Y || h_lay == nullptr) {
        wiski::set_error("kron_fused_pair_apply_lay: unsupported shape");
        return 3;
    }
    return wiski::fused_pair_apply(cols, d, h_g, gmax, pair, X, Y, c, wiski::as_stream(stream), h_lay);
}

int wiski_kron_fused_pair_grad_lay_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair,
                                       const float* Z, const float* P, float* Zout, int64_t c, double* acc_u64,
                                       double* acc_v64, const int64_t* h_lay, void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || pair < 0 || 2 * pair + 1 >= d || h_lay == nullptr) {
        wiski::set_error("kron_fused_pair_grad_lay: unsupported shape");
        return 3;
    }
    return wiski::fused_pair_grad(cols, d, h_g, gmax, pair, Z, P, Zout, c, acc_u64, acc_v64, wiski::as_stream(stream),
                                  h_lay);
}

int wiski_kron_fused_pair_grad_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* Z,
                                   const float* P, float* Zout, int64_t c, double* acc_u64, double* acc_v64,
                                   void* stream) {
    if (d < 2 || d > WISKI_MAX_DIMS || pair < 0 || 2 * pair + 1 >= d) {
        wiski::set_error("kron_fused_pair_grad: unsupported shape");
        return 3;
    }
    return wiski::fused_pair_grad(cols, d, h_g, gmax, pair, Z, P, Zout, c, acc_u64, acc_v64, wiski::as_stream(stream));
}
}
