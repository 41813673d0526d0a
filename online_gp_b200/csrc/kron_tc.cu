// k9 / k15 on the tensor pipe: two Kronecker axes per pass over an m x c fp32 panel for 32-point axes (BASELINE
// config 2: 32^4), with tcgen05.mma (kind::tf32, 3xTF32 split operands), accumulators AND the data operand in tensor
// memory, grid tiles staged by TMA (5-D tensor maps: column, inner offset, axis v, axis u, outer offset).
//
// Why this shape.  Applying T_v (32 x 32) to a [32 u][32 v][16 w] grid tile is a GEMM whose *natural* M is only 32.
// Written transposed,  D[(u,w), v] = sum_v' X[(u,w), v'] T_v[v', v],  the panel data is the M x K operand (M = 512 rows
// (u,w) per tile = four 128-row MMA tiles, K = 32) and the Toeplitz factor the K x N operand (N = 32): no padding, no
// wasted MACs.  The catch is the operand layout: the tile sits in shared memory as [u][v][w] (w contiguous, what the
// panel's row-major global layout gives TMA), which is no canonical UMMA layout for 32-bit MN-major data (only
// SWIZZLE_128B_BASE32B with 32-float atoms exists).  So worker threads (thread = one (u,w) row = one TMEM lane) read
// their 32 K-values from shared memory into registers, split them there into tf32-big / fp32-remainder (no shared
// memory read-modify-write), and tcgen05.st them into TMEM, from where tcgen05.mma takes A directly ("TS" form).
// The second axis needs rows (v,w) and K = u: the first result goes TMEM -> registers -> an XOR-swizzled shared
// buffer [u][v][w] and is re-read transposed the same way.  Per tile and axis: 4 MMA tiles x 4 k-steps x 3 split
// products of 128 x 32 x 8 = 768 tensor cycles, against ~5 900 cycles of HBM time for the tile (64 KB in + 64 KB out
// at 6.4 TB/s / 148 SMs): the pass is bound by HBM, not by FMA issue as the SIMT form in kron_fused.cu is.
//
// Reference operations replaced: KroneckerProductLazyTensor._matmul / ToeplitzLazyTensor._matmul and their autograd
// (SURVEY.md App. A.4, 2b k9 / k15), reached from online_gp/models/batched_fixed_noise_online_gp.py:348 (Kuu @ L) and
// online_gp/models/online_ski_regression.py:141 (loss.backward()).
//
// CTA = 384 threads, 1 CTA / SM, persistent over tiles:
//   warp 0      TMA producer (one lane): 16 KB stages = one 128-row MMA tile of input, mbarrier ring
//   warps 1, 3  MMA issuers (one lane each, one per worker group): 12 tcgen05.mma per MMA tile, tcgen05.commit -> mbarrier
//   warp 2      TMEM allocator
//   warps 4-11  workers: two groups of 4 warps (one TMEM lane quadrant each); group g owns MMA tiles g and g + 2
// TMEM: 4 slots (group x tile) of 128 columns: A raw [0,32) | A remainder [32,64) | D [64,128).
#include <stdlib.h>
#include <string.h>
#include "tc_ptx.cuh"

namespace wiski {
namespace ktc {
using namespace tc;

constexpr int G = 32;                     // grid points per fused axis
constexpr int CB = 16;                    // panel columns per tile
constexpr int STAGE_F = 8 * G * CB;       // floats per stage: 8 lines of one axis x 32 x 16 columns = 16 KB
constexpr int STAGE_B = STAGE_F * 4;
constexpr int TILE_B = G * G * CB * 4;    // 64 KB
constexpr int NWORK = 256;
constexpr int NTHREADS = 128 + NWORK;
constexpr int SLOT_COLS = 128;
constexpr int TIMG_B = 4096;              // one 32 x 32 fp32 Toeplitz image (K-major, SWIZZLE_128B)

// Geometry of one axis pair (a, b), a < b, both of 32 points, on a grid [.., g_a, .., g_b, ..]:
//   row(u, v, f0, f1) = u * su + v * sv + f0 * sf0 + f1 * sf1      (row strides; sv < su)
// f0, f1 = the (at most two) non-trivial free index groups among {axes after b, axes between a and b, axes before a},
// f0 the faster one in memory.  The 5-D tensor map of an operand lists (column, then the row groups in memory order);
// pos_* are the positions of u, v, f0, f1 in it; the column-block index of a chunked operand folds into dimension
// pos_blk (the slowest real one, extent size_blk).  Tile id -> (column chunk, f0, f1), chunk fastest.
struct Geom {
    int nf0, nf1;        // extents of the free groups
    int n_chunks;        // c / CB
    long long n_tiles;   // nf0 * nf1 * n_chunks
    int pos_u, pos_v, pos_f0, pos_f1, pos_blk, size_blk;
    long long su, sv, sf0, sf1;
};
// where element (row, col) of a panel lives: ptr + (col / cw) * cstride + row * ld + col % cw
struct Lay {
    long long ld, cw, cstride;
};

// Pushing results into the peers' memory (row-sharded multi-GPU path): instead of ONE output tensor map the kernel gets
// one per destination rank, each over that rank's peer-mapped receive buffer (NVLink / NVSwitch), so the layout change
// between the row-sharded and the column-sharded panel happens in the producing kernel's own TMA stores:
//   mode 1  column block j of the tile's columns belongs to rank j      (row slab -> all rows of my columns)
//   mode 2  the tile's axis-u range splits into n_dst equal parts, part i belongs to rank i (u = grid axis 0:
//           column-sharded panel -> my rows of every column block); box_u = axis-u extent of one store
//   mode 3  (apply pass only) mode 2 with 32-column store boxes: a CTA works on the two 16-column tiles of a 32-column
//           group back to back, keeps the first one's result in the spare TMEM columns [96,128) of its slots and stores
//           both as ONE [u_loc][8 v][32 columns] box per rank — 128-byte row pieces, which NVLink carries at 665 GB/s
//           against 417 GB/s for 64-byte ones (profiles/r02_p2p_store_bandwidth_n2.txt)
constexpr int MAXP = 8;
struct PushMaps {
    CUtensorMap m[MAXP];
};
struct Push {
    int mode;            // 0 = off
    int box_u;           // mode 2: axis-u lines per store (<= 8 for the gradient pass, = u_loc for the apply pass)
    int u_loc;           // mode 2: axis-u lines per destination rank (32 / n_dst)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major SWIZZLE_128B image of the symmetric Toeplitz matrix T[n][k] = col[|n - k|] (32 x 32) and of its fp32
// remainder after tf32 truncation; row n lives at n * 128 B, its 16-byte chunk j at ((j ^ (n & 7)) * 16).
__device__ __forceinline__ void build_toeplitz_image(uint8_t* img_big, uint8_t* img_small, const float* __restrict__ col,
                                                     int row0, int tid, int nthreads) {
    for (int e = tid; e < G * G; e += nthreads) {
        const int n = e >> 5, k = e & 31;
        const int dist = n > k ? n - k : k - n;
        const uint32_t val = __float_as_uint(col[dist]);
        const int rn = row0 + n;
        const uint32_t off = (uint32_t)rn * 128u + (uint32_t)((((k >> 2) ^ (rn & 7)) << 4) + ((k & 3) << 2));
        *reinterpret_cast<uint32_t*>(img_big + off) = val;
        *reinterpret_cast<uint32_t*>(img_small + off) = tf32_small(val);
    }
}

// A (raw | remainder) of one MMA tile: registers -> TMEM columns [0,32) and [32,64) of the slot
__device__ __forceinline__ void split_to_tmem(uint32_t taddr, const uint32_t (&x)[32]) {
    tmem_st32(taddr, x);
    uint32_t s[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) s[j] = tf32_small(x[j]);
    tmem_st32(taddr + 32, s);
    tmem_st_wait();
    tc_fence_before();
}

// B operand (32 x N Toeplitz image pair, K-major SWIZZLE_128B): the 4 k-step descriptors of the big and of the
// remainder image, built once per kernel (the single issuing thread then only moves registers per MMA)
struct BDesc {
    uint64_t big[4], small[4];
};
__device__ __forceinline__ BDesc make_bdesc(uint32_t b_big, uint32_t b_small) {
    BDesc b;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        b.big[ks] = make_smem_desc(b_big + ks * 32, 0, 1024, 2);
        b.small[ks] = make_smem_desc(b_small + ks * 32, 0, 1024, 2);
    }
    return b;
}
// D[slot] = A[slot] (128 x 32, in TMEM) * B (32 x N), 3xTF32: 12 tcgen05.mma of 128 x N x 8.  Warp-collective (all lanes).
__device__ __forceinline__ void issue_mma_tile(uint32_t tslot, const BDesc& b, uint32_t idesc, uint32_t d_off = 64) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        umma_tf32_ts_elect(tslot + d_off, tslot + 32 + ks * 8, b.big[ks], idesc, ks > 0 ? 1u : 0u);   // remainder(A) * big(B)
        umma_tf32_ts_elect(tslot + d_off, tslot + ks * 8, b.small[ks], idesc, 1u);                    // big(A) * remainder(B)
        umma_tf32_ts_elect(tslot + d_off, tslot + ks * 8, b.big[ks], idesc, 1u);                      // big(A) * big(B)
    }
}
// it-th tile of this CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ..; in pair mode (Push mode 3) the two 16-column
// tiles (2 s, 2 s + 1) of 32-column group s = blockIdx.x + (it / 2) gridDim.x, one after the other (chunk index is the
// fastest part of the tile id, so the two differ in the column chunk only)
__device__ __forceinline__ long long tile_of(long long it, bool pair) {
    return pair ? 2 * ((long long)blockIdx.x + (it >> 1) * gridDim.x) + (it & 1) : (long long)blockIdx.x + it * gridDim.x;
}

__device__ __forceinline__ void decode_tile(const Geom& g, long long tile, int& cc, int& f0, int& f1) {
    cc = (int)(tile % g.n_chunks);
    const long long o = tile / g.n_chunks;
    f0 = (int)(o % g.nf0);
    f1 = (int)(o / g.nf0);
}

// TMA coordinates of the box of tile (f0, f1) that starts at (u0, v0) and column c0 of column block blk
struct Coords {
    int c[5];
};
__device__ __forceinline__ Coords tile_coords(const Geom& g, int c0, int blk, int f0, int f1, int u0, int v0) {
    Coords r;
    r.c[0] = c0;
#pragma unroll
    for (int i = 1; i < 5; ++i) {
        int x = 0;
        if (i == g.pos_u) x = u0;
        if (i == g.pos_v) x = v0;
        if (i == g.pos_f0) x = f0;
        if (i == g.pos_f1) x = f1;
        if (i == g.pos_blk) x += blk * g.size_blk;
        r.c[i] = x;
    }
    return r;
}
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* tm, uint64_t* bar, const Coords& k) {
    tma_load_5d(dst, tm, bar, k.c[0], k.c[1], k.c[2], k.c[3], k.c[4]);
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, const void* src, const Coords& k) {
    tma_store_5d(tm, src, k.c[0], k.c[1], k.c[2], k.c[3], k.c[4]);
}

// XOR-swizzled [u][v][w] buffer: two consecutive u (two lanes groups of a warp in the (u,w)-row steps) hit different
// bank halves; for a fixed u the map is a bijection of (v, w), so the (v,w)-row steps stay conflict free as well
__device__ __forceinline__ int ybuf_index(int u, int v, int w) { return u * (G * CB) + ((v * CB + w) ^ ((u & 1) << 4)); }

struct Bars {
    uint64_t* full;
    uint64_t* empty;
    uint64_t* a_ready;
    uint64_t* d_ready;
};

// ------------------------------------------------------------------------------------------ Y = (T_u x T_v) X
struct ApplyParams {
    int cwx, cwy;            // block widths of X and Y (TMA coordinates)
    float* Ydirect;          // non-NULL: results leave through per-thread global stores (layout ly) instead of TMA stores
    Lay ly;
    Geom g;
    const float* col_u;      // 32 floats each
    const float* col_v;
    long long* prof;         // PROF instantiation only: 13 counters
    Push push;
};

// PROF: accumulate clock64() deltas of the roles' phases into p.prof (test_kron_tc prof); compiled out otherwise
#define KTC_TICK(slot) do { if (PROF) { const long long now_ = clock64(); pr[slot] += now_ - tl; tl = now_; } } while (0)
template <int NST, bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
pair_apply_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ PushMaps pm, const ApplyParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer stays derived from the __shared__ array: LDS / STS
    uint8_t* timg = smem;                                                   // v big | v small | u big | u small
    float* ring = reinterpret_cast<float*>(smem + 4 * TIMG_B);
    float* ybuf = reinterpret_cast<float*>(smem + 4 * TIMG_B + NST * STAGE_B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * TIMG_B + NST * STAGE_B + TILE_B);
    Bars B{bars, bars + NST, bars + 2 * NST, bars + 2 * NST + 4};
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NST + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&B.full[s], 1);
            mbar_init(&B.empty[s], 128);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&B.a_ready[s], 128);
            mbar_init(&B.d_ready[s], 1);
        }
        mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    build_toeplitz_image(timg, timg + TIMG_B, p.col_v, 0, threadIdx.x, NTHREADS);
    build_toeplitz_image(timg + 2 * TIMG_B, timg + 3 * TIMG_B, p.col_u, 0, threadIdx.x, NTHREADS);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const Geom g = p.g;

    if (warp == 0) {
        if (lane == 0) {
            long long seq = 0;
            long long pr[2] = {0, 0}, tl = PROF ? clock64() : 0;
            for (long long itp = 0;; ++itp) {
                const long long tile = tile_of(itp, p.push.mode == 3);
                if (tile >= g.n_tiles) break;
                int cc, f0, f1;
                decode_tile(g, tile, cc, f0, f1);
                const int col0 = cc * CB, blk = col0 / p.cwx;
                const int c0 = col0 - blk * p.cwx;
                for (int j = 0; j < 4; ++j, ++seq) {
                    const int s = (int)(seq % NST);
                    KTC_TICK(1);
                    mbar_wait(&B.empty[s], (uint32_t)(((seq / NST) & 1) ^ 1));
                    KTC_TICK(0);
                    mbar_arrive_expect_tx(&B.full[s], STAGE_B);
                    tma_load_box(ring + s * STAGE_F, &tmX, &B.full[s], tile_coords(g, c0, blk, f0, f1, 8 * j, 0));
                }
            }
            if (PROF) atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 12), (unsigned long long)pr[0]);
        }
    } else if (warp == 1 || warp == 3) {
        // MMA issuers: warp 1 serves worker group 0 (slots 0, 1), warp 3 group 1 (slots 2, 3).  The whole warp runs the
        // loop (warp-uniform operands; one elected lane issues, see umma_tf32_ts_elect).
        {
            const int ig = __shfl_sync(0xffffffffu, warp >> 1, 0);
            const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
            const uint32_t idesc = make_idesc_tf32(false, false, 32);
            const uint32_t timg_a = smem_u32(timg);
            const BDesc bv = make_bdesc(timg_a, timg_a + TIMG_B), bu = make_bdesc(timg_a + 2 * TIMG_B, timg_a + 3 * TIMG_B);
            uint32_t use = 0;
            long long pr[2] = {0, 0}, tl = PROF ? clock64() : 0;
            const bool pair = p.push.mode == 3;
            for (long long itm = 0;; ++itm) {
                if (tile_of(itm, pair) >= g.n_tiles) break;
                // pair mode: the final result of the first half-tile goes to the spare columns [96,128) and waits there
                const uint32_t d_keep = (pair && (itm & 1) == 0) ? 96u : 64u;
#pragma unroll
                for (int phase = 0; phase < 2; ++phase, ++use) {
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int slot = ig * 2 + jj;
                        mbar_wait(&B.a_ready[slot], use & 1);
                        KTC_TICK(0);
                        tc_fence_after();
                        issue_mma_tile(tmem_base + slot * SLOT_COLS, phase == 0 ? bv : bu, idesc, phase == 0 ? 64u : d_keep);
                        umma_commit_elect(&B.d_ready[slot]);
                        KTC_TICK(1);
                    }
                }
            }
            if (PROF && ig == 0 && lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 0), (unsigned long long)pr[0]);
                atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 1), (unsigned long long)pr[1]);
            }
        }
    } else if (warp >= 4) {
        const int wq = warp - 4, grp = wq >> 2, quad = wq & 3;
        const int rho = quad * 32 + lane, hi = rho >> 4, w = rho & 15;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool direct = p.Ydirect != nullptr;
        const bool store_thread = !direct && quad == 0 && lane == 0;      // one per group: issues the group's TMA stores
        const long long ustride = g.su * p.ly.ld;
        uint32_t use = 0;
        long long it = 0;
        long long pr[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tl = PROF ? clock64() : 0;
        const bool pair = p.push.mode == 3;
        for (;; ++it) {
            const long long tile = tile_of(it, pair);
            if (tile >= g.n_tiles) break;
            int cc, f0, f1;
            decode_tile(g, tile, cc, f0, f1);
            // ---- axis v: rows (u, w), K = v'
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const long long seq = 4 * it + j;
                const int s = (int)(seq % NST);
                mbar_wait(&B.full[s], (uint32_t)((seq / NST) & 1));
                KTC_TICK(0);
                const float* st = ring + s * STAGE_F + hi * (G * CB) + w;
                uint32_t x[32];
#pragma unroll
                for (int v = 0; v < 32; ++v) x[v] = __float_as_uint(st[v * CB]);
                split_to_tmem(tlane + slot * SLOT_COLS, x);
                mbar_arrive(&B.a_ready[slot]);
                mbar_arrive(&B.empty[s]);
                KTC_TICK(1);
            }
            // ybuf doubles as the staging buffer of the previous tile's output stores: they must have been read out
            if (!direct) {
                if (store_thread) tma_store_wait_read<0>();
                named_bar_sync(3, NWORK);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                mbar_wait(&B.d_ready[slot], use & 1);
                KTC_TICK(2);
                tc_fence_after();
                uint32_t d[32];
                tmem_ld32(tlane + slot * SLOT_COLS + 64, d);
                const int u = 8 * j + hi;
#pragma unroll
                for (int v = 0; v < 32; ++v) ybuf[ybuf_index(u, v, w)] = __uint_as_float(d[v]);
                KTC_TICK(3);
            }
            ++use;
            tc_fence_before();
            named_bar_sync(1, NWORK);
            KTC_TICK(4);
            // ---- axis u: rows (v, w), K = u'
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const int v = 8 * j + hi;
                uint32_t x[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) x[u] = __float_as_uint(ybuf[ybuf_index(u, v, w)]);
                split_to_tmem(tlane + slot * SLOT_COLS, x);
                mbar_arrive(&B.a_ready[slot]);
            }
            KTC_TICK(5);
            named_bar_sync(2, NWORK);             // every read of ybuf is done: the next tile may overwrite it
            KTC_TICK(6);
            // results -> ybuf as four [32 u][8 v][16 w] boxes (conflict free) -> one TMA store per MMA tile
            const int col0 = cc * CB, blk = col0 / p.cwy;
            const int c0y = col0 - blk * p.cwy;
            if (pair) {
                // first half-tile: its result stays in TMEM columns [96,128); second: both halves leave as 32-column boxes,
                // one MMA tile (= [32 u][8 v][32 columns], 32 KB) per group at a time through ybuf
                const bool second = (it & 1) != 0;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = grp + 2 * jj, slot = grp * 2 + jj;
                    mbar_wait(&B.d_ready[slot], use & 1);
                    if (!second) continue;
                    tc_fence_after();
                    float* og = ybuf + grp * (2 * STAGE_F) + hi * (2 * CB) + w;
                    if (jj == 1) {                                   // the group's box of jj = 0 must have been read out
                        if (store_thread) tma_store_wait_read<0>();
                        named_bar_sync(4 + grp, 128);
                    }
                    uint32_t d[32];
                    tmem_ld32(tlane + slot * SLOT_COLS + 96, d);     // columns [c0, c0 + 16): the first half-tile
#pragma unroll
                    for (int u = 0; u < 32; ++u) og[u * (8 * 2 * CB)] = __uint_as_float(d[u]);
                    tmem_ld32(tlane + slot * SLOT_COLS + 64, d);     // columns [c0 + 16, c0 + 32)
#pragma unroll
                    for (int u = 0; u < 32; ++u) og[u * (8 * 2 * CB) + CB] = __uint_as_float(d[u]);
                    fence_proxy_async_smem();
                    named_bar_sync(4 + grp, 128);
                    if (store_thread) {
                        const Coords kc = tile_coords(g, c0y - CB, 0, f0, f1, 0, 8 * j);
                        for (int i = 0; i * p.push.u_loc < G; ++i)
                            tma_store_box(&pm.m[i], ybuf + grp * (2 * STAGE_F) + i * p.push.u_loc * (8 * 2 * CB), kc);
                        tma_store_commit();
                    }
                }
                ++use;
                tc_fence_before();
                continue;
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                mbar_wait(&B.d_ready[slot], use & 1);
                KTC_TICK(7);
                tc_fence_after();
                uint32_t d[32];
                tmem_ld32(tlane + slot * SLOT_COLS + 64, d);
                if (direct) {
                    const int blkd = col0 / (int)p.ly.cw;
                    float* yp = p.Ydirect + blkd * p.ly.cstride + (col0 - blkd * (int)p.ly.cw) + w +
                                (f0 * g.sf0 + f1 * g.sf1 + (long long)(8 * j + hi) * g.sv) * p.ly.ld;
#pragma unroll
                    for (int u = 0; u < 32; ++u) {
                        *yp = __uint_as_float(d[u]);
                        yp += ustride;
                    }
                } else {
                    float* og = ybuf + j * STAGE_F + hi * CB + w;
#pragma unroll
                    for (int u = 0; u < 32; ++u) og[u * (8 * CB)] = __uint_as_float(d[u]);
                    fence_proxy_async_smem();
                    named_bar_sync(4 + grp, 128);
                    if (store_thread) {
                        if (p.push.mode == 0) {
                            tma_store_box(&tmY, ybuf + j * STAGE_F, tile_coords(g, c0y, blk, f0, f1, 0, 8 * j));
                        } else if (p.push.mode == 1) {
                            tma_store_box(&pm.m[blk], ybuf + j * STAGE_F, tile_coords(g, c0y, 0, f0, f1, 0, 8 * j));
                        } else {                                           // [32 u][8 v][16 w]: u_loc lines per rank
                            const Coords kc = tile_coords(g, c0y, 0, f0, f1, 0, 8 * j);
                            for (int i = 0; i * p.push.u_loc < G; ++i)
                                tma_store_box(&pm.m[i], ybuf + j * STAGE_F + i * p.push.u_loc * (8 * CB), kc);
                        }
                        tma_store_commit();
                    }
                }
                KTC_TICK(8);
            }
            ++use;
            tc_fence_before();
        }
        if (store_thread) tma_store_wait<0>();
        if (PROF && wq == 0 && lane == 0) {
            for (int i = 0; i < 9; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(p.prof + 2 + i), (unsigned long long)pr[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ directional gradient pass
// Inputs Z (incoming-gradient side) and P (operand side), both m x c.  With the direction matrices T'_i = Toeplitz of
// dirs_i (d col_i / d lengthscale_i) the pass accumulates, per tile,
//   out3[0] += <T'_u Z, T_v P>     (= <grad col_u, dirs_u>)
//   out3[1] += <T'_v T_u Z, P>     (= <grad col_v, dirs_v>)
//   out3[2] += <T_u Z, T_v P>      (= <grad col_i, col_i> for every i)
// and, with STORE, writes Zout = T_v T_u Z for the next pair.  Steps per tile:
//   A  rows (u,w): S = T_v P                         (P streamed by u-blocks; S -> sbuf)
//   B  rows (v,w): [zu | zd] = [T_u | T'_u] Z         (Z streamed by v-blocks); out3[0] += zd . S, out3[2] += zu . S; zu -> zbuf
//   C  rows (u,w): [zuv | zd2] = [T_v | T'_v] zu      (zbuf); out3[1] += zd2 . P (P streamed again: L2 hits); Zout = zuv
struct GradParams {
    int cwz, cwp, cwo;       // block widths of Z, P and Zout (TMA coordinates)
    float* Odirect;          // non-NULL: Zout leaves through per-thread global stores (layout lo) instead of TMA stores
    Lay lo;
    Geom g;
    const float* col_u;
    const float* col_v;
    const float* dir_u;
    const float* dir_v;
    double* out3;
    Push push;
};

template <bool STORE, int NST>
__global__ void __launch_bounds__(NTHREADS, 1)
pair_grad_dir_tc_kernel(const __grid_constant__ CUtensorMap tmPu, const __grid_constant__ CUtensorMap tmZv,
                        const __grid_constant__ CUtensorMap tmO, const __grid_constant__ PushMaps pm, const GradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer stays derived from the __shared__ array: LDS / STS
    // images: [T_v ; T'_v] big (64 rows, 8 KB) | the same, remainder | [T_u ; T'_u] big | remainder
    uint8_t* timg = smem;
    float* ring = reinterpret_cast<float*>(smem + 8 * TIMG_B);
    float* sbuf = reinterpret_cast<float*>(smem + 8 * TIMG_B + NST * STAGE_B);
    float* zbuf = reinterpret_cast<float*>(smem + 8 * TIMG_B + NST * STAGE_B + TILE_B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 8 * TIMG_B + NST * STAGE_B + 2 * TILE_B);
    Bars B{bars, bars + NST, bars + 2 * NST, bars + 2 * NST + 4};
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NST + 8);
    float* red = reinterpret_cast<float*>(bars + 2 * NST + 9);               // [3][8] block reduction

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&B.full[s], 1);
            mbar_init(&B.empty[s], 128);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&B.a_ready[s], 128);
            mbar_init(&B.d_ready[s], 1);
        }
        mbar_fence_init();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmPu);
        tma_prefetch_desc(&tmZv);
        if (STORE) tma_prefetch_desc(&tmO);
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    build_toeplitz_image(timg, timg + 2 * TIMG_B, p.col_v, 0, threadIdx.x, NTHREADS);
    build_toeplitz_image(timg, timg + 2 * TIMG_B, p.dir_v, 32, threadIdx.x, NTHREADS);
    build_toeplitz_image(timg + 4 * TIMG_B, timg + 6 * TIMG_B, p.col_u, 0, threadIdx.x, NTHREADS);
    build_toeplitz_image(timg + 4 * TIMG_B, timg + 6 * TIMG_B, p.dir_u, 32, threadIdx.x, NTHREADS);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const Geom g = p.g;
    float acc_u = 0.f, acc_v = 0.f, acc_s = 0.f;

    if (warp == 0) {
        if (lane == 0) {
            long long seq = 0;
            for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                int cc, f0, f1;
                decode_tile(g, tile, cc, f0, f1);
                const int col0 = cc * CB;
                const int blkp = col0 / p.cwp, blkz = col0 / p.cwz;
                const int p0 = col0 - blkp * p.cwp, z0 = col0 - blkz * p.cwz;
                for (int k = 0; k < 12; ++k, ++seq) {
                    const int s = (int)(seq % NST);
                    mbar_wait(&B.empty[s], (uint32_t)(((seq / NST) & 1) ^ 1));
                    mbar_arrive_expect_tx(&B.full[s], STAGE_B);
                    const int j = k & 3;
                    if (k >= 4 && k < 8)
                        tma_load_box(ring + s * STAGE_F, &tmZv, &B.full[s], tile_coords(g, z0, blkz, f0, f1, 0, 8 * j));     // [32 u][8 v][16 w]
                    else
                        tma_load_box(ring + s * STAGE_F, &tmPu, &B.full[s], tile_coords(g, p0, blkp, f0, f1, 8 * j, 0));     // [8 u][32 v][16 w]
                }
            }
        }
    } else if (warp == 1 || warp == 3) {
        {
            const int ig = __shfl_sync(0xffffffffu, warp >> 1, 0);
            const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
            const uint32_t idesc32 = make_idesc_tf32(false, false, 32), idesc64 = make_idesc_tf32(false, false, 64);
            const uint32_t timg_a = smem_u32(timg);
            const BDesc bv = make_bdesc(timg_a, timg_a + 2 * TIMG_B);                                // [T_v ; T'_v]
            const BDesc bu = make_bdesc(timg_a + 4 * TIMG_B, timg_a + 6 * TIMG_B);                    // [T_u ; T'_u]
            const BDesc bdv = make_bdesc(timg_a + TIMG_B, timg_a + 3 * TIMG_B);                       // T'_v alone
            uint32_t use = 0;
            for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
#pragma unroll
                for (int step = 0; step < 3; ++step, ++use) {
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int slot = ig * 2 + jj;
                        mbar_wait(&B.a_ready[slot], use & 1);
                        tc_fence_after();
                        const uint32_t ts = tmem_base + slot * SLOT_COLS;
                        if (step == 0) issue_mma_tile(ts, bv, idesc32);                 // S = T_v P
                        else if (step == 1) issue_mma_tile(ts, bu, idesc64);            // [zu | zd] = [T_u | T'_u] Z
                        else if (STORE) issue_mma_tile(ts, bv, idesc64);                // [zuv | zd2] = [T_v | T'_v] zu
                        else issue_mma_tile(ts, bdv, idesc32);                          // zd2 = T'_v zu
                        umma_commit_elect(&B.d_ready[slot]);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        const int wq = warp - 4, grp = wq >> 2, quad = wq & 3;
        const int rho = quad * 32 + lane, hi = rho >> 4, w = rho & 15;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool direct = p.Odirect != nullptr;
        const bool store_thread = STORE && !direct && quad == 0 && lane == 0;
        const long long vstride = g.sv * p.lo.ld;
        uint32_t use = 0;
        long long it = 0;
        for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
            int cc, f0, f1;
            decode_tile(g, tile, cc, f0, f1);
            // ---- step A: S = T_v P, rows (u, w)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const long long seq = 12 * it + j;
                const int s = (int)(seq % NST);
                mbar_wait(&B.full[s], (uint32_t)((seq / NST) & 1));
                const float* st = ring + s * STAGE_F + hi * (G * CB) + w;
                uint32_t x[32];
#pragma unroll
                for (int v = 0; v < 32; ++v) x[v] = __float_as_uint(st[v * CB]);
                split_to_tmem(tlane + slot * SLOT_COLS, x);
                mbar_arrive(&B.a_ready[slot]);
                mbar_arrive(&B.empty[s]);
            }
            if (STORE && !direct) {        // sbuf doubles as the staging buffer of the previous tile's Zout stores
                if (store_thread) tma_store_wait_read<0>();
                named_bar_sync(3, NWORK);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                mbar_wait(&B.d_ready[slot], use & 1);
                tc_fence_after();
                uint32_t d[32];
                tmem_ld32(tlane + slot * SLOT_COLS + 64, d);
                const int u = 8 * j + hi;
#pragma unroll
                for (int v = 0; v < 32; ++v) sbuf[ybuf_index(u, v, w)] = __uint_as_float(d[v]);
            }
            ++use;
            tc_fence_before();
            named_bar_sync(1, NWORK);
            // ---- step B: [zu | zd] = [T_u | T'_u] Z, rows (v, w)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const long long seq = 12 * it + 4 + j;
                const int s = (int)(seq % NST);
                mbar_wait(&B.full[s], (uint32_t)((seq / NST) & 1));
                const float* st = ring + s * STAGE_F + hi * CB + w;             // [u'][8 v][16 w]
                uint32_t x[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) x[u] = __float_as_uint(st[u * (8 * CB)]);
                split_to_tmem(tlane + slot * SLOT_COLS, x);
                mbar_arrive(&B.a_ready[slot]);
                mbar_arrive(&B.empty[s]);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const int v = 8 * j + hi;
                float sv_[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) sv_[u] = sbuf[ybuf_index(u, v, w)];
                mbar_wait(&B.d_ready[slot], use & 1);
                tc_fence_after();
                uint32_t d[32];
                tmem_ld32(tlane + slot * SLOT_COLS + 64 + 32, d);               // zd = T'_u z
#pragma unroll
                for (int u = 0; u < 32; ++u) acc_u = fmaf(__uint_as_float(d[u]), sv_[u], acc_u);
                tmem_ld32(tlane + slot * SLOT_COLS + 64, d);                    // zu = T_u z
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    acc_s = fmaf(__uint_as_float(d[u]), sv_[u], acc_s);
                    zbuf[ybuf_index(u, v, w)] = __uint_as_float(d[u]);
                }
            }
            ++use;
            tc_fence_before();
            named_bar_sync(2, NWORK);
            // ---- step C: [zuv | zd2] = [T_v | T'_v] zu, rows (u, w)
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const int u = 8 * j + hi;
                uint32_t x[32];
#pragma unroll
                for (int v = 0; v < 32; ++v) x[v] = __float_as_uint(zbuf[ybuf_index(u, v, w)]);
                split_to_tmem(tlane + slot * SLOT_COLS, x);
                mbar_arrive(&B.a_ready[slot]);
            }
            const int col0 = cc * CB, blko = col0 / p.cwo;
            const int c0o = col0 - blko * p.cwo;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = grp + 2 * jj, slot = grp * 2 + jj;
                const long long seq = 12 * it + 8 + j;
                const int s = (int)(seq % NST);
                mbar_wait(&B.full[s], (uint32_t)((seq / NST) & 1));
                const float* st = ring + s * STAGE_F + hi * (G * CB) + w;
                float pv[32];
#pragma unroll
                for (int v = 0; v < 32; ++v) pv[v] = st[v * CB];
                mbar_arrive(&B.empty[s]);
                mbar_wait(&B.d_ready[slot], use & 1);
                tc_fence_after();
                uint32_t d[32];
                tmem_ld32(tlane + slot * SLOT_COLS + 64 + (STORE ? 32 : 0), d);  // zd2 = T'_v zu
#pragma unroll
                for (int v = 0; v < 32; ++v) acc_v = fmaf(__uint_as_float(d[v]), pv[v], acc_v);
                if (STORE) {
                    tmem_ld32(tlane + slot * SLOT_COLS + 64, d);                // zuv = T_v zu
                    if (direct) {
                        const int blkd = col0 / (int)p.lo.cw;
                        float* yp = p.Odirect + blkd * p.lo.cstride + (col0 - blkd * (int)p.lo.cw) + w +
                                    (f0 * g.sf0 + f1 * g.sf1 + (long long)(8 * j + hi) * g.su) * p.lo.ld;
#pragma unroll
                        for (int v = 0; v < 32; ++v) {
                            *yp = __uint_as_float(d[v]);
                            yp += vstride;
                        }
                    } else {                                                    // -> sbuf box [8 u][32 v][16 w] -> TMA store
                        float* og = sbuf + j * STAGE_F + hi * (G * CB) + w;
#pragma unroll
                        for (int v = 0; v < 32; ++v) og[v * CB] = __uint_as_float(d[v]);
                        fence_proxy_async_smem();
                        named_bar_sync(4 + grp, 128);
                        if (store_thread) {
                            if (p.push.mode == 0) {
                                tma_store_box(&tmO, sbuf + j * STAGE_F, tile_coords(g, c0o, blko, f0, f1, 8 * j, 0));
                            } else {                                       // [8 u][32 v][16 w] in pieces of box_u lines
                                for (int q = 0; q * p.push.box_u < 8; ++q) {
                                    const int u = 8 * j + q * p.push.box_u, rk = u / p.push.u_loc;
                                    tma_store_box(&pm.m[rk], sbuf + j * STAGE_F + q * p.push.box_u * (G * CB),
                                                  tile_coords(g, c0o, 0, f0, f1, u - rk * p.push.u_loc, 0));
                                }
                            }
                            tma_store_commit();
                        }
                    }
                }
            }
            ++use;
            tc_fence_before();
        }
        if (store_thread) tma_store_wait<0>();
        // ---- reduction of the three partial sums: warp shuffle, then one double atomic per CTA and sum
        acc_u = warp_sum(acc_u);
        acc_v = warp_sum(acc_v);
        acc_s = warp_sum(acc_s);
        if (lane == 0) {
            red[0 * 8 + wq] = acc_u;
            red[1 * 8 + wq] = acc_v;
            red[2 * 8 + wq] = acc_s;
        }
        named_bar_sync(3, NWORK);
        if (threadIdx.x - 128 < 3) {
            const int k = threadIdx.x - 128;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += (double)red[k * 8 + i];
            atomicAdd(&p.out3[k], s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ host side
// Geometry of the pair (au, av), au < av.  dims[] receives (extent, row stride) of the four row groups of the tensor
// map in memory order (positions 1..4 of the 5-D map; padded with extent-1 groups).
struct MapDims {
    long long ext[4], stride[4];
    int n;               // real groups
};
static bool make_geom(Geom& g, MapDims& md, int d, const int64_t* h_g, int au, int av, int64_t c) {
    if (au < 0 || av <= au || av >= d || h_g[au] != G || h_g[av] != G || c % CB != 0 || c < CB) return false;
    long long after = 1, mid = 1, before = 1;
    for (int j = av + 1; j < d; ++j) after *= h_g[j];
    for (int j = au + 1; j < av; ++j) mid *= h_g[j];
    for (int j = 0; j < au; ++j) before *= h_g[j];
    if ((after > 1) + (mid > 1) + (before > 1) > 2) return false;          // needs a 6-D map: not supported
    if (after >= (1 << 30) || mid >= (1 << 30) || before >= (1 << 30) || c / CB >= (1 << 30)) return false;
    const long long s_after = 1, s_v = after, s_mid = G * after, s_u = mid * G * after, s_before = G * mid * G * after;
    g.su = s_u;
    g.sv = s_v;
    g.nf0 = g.nf1 = 1;
    g.sf0 = g.sf1 = 0;
    g.pos_f0 = g.pos_f1 = 0;                                               // 0 = unused (never matches a row position)
    md.n = 0;
    int nfree = 0;
    auto add = [&](long long ext, long long stride) { md.ext[md.n] = ext; md.stride[md.n] = stride; return ++md.n; };
    auto add_free = [&](long long ext, long long stride) {
        const int pos = add(ext, stride);
        if (nfree == 0) { g.nf0 = (int)ext; g.sf0 = stride; g.pos_f0 = pos; }
        else { g.nf1 = (int)ext; g.sf1 = stride; g.pos_f1 = pos; }
        ++nfree;
    };
    if (after > 1) add_free(after, s_after);
    g.pos_v = add(G, s_v);
    if (mid > 1) add_free(mid, s_mid);
    g.pos_u = add(G, s_u);
    if (before > 1) add_free(before, s_before);
    g.pos_blk = md.n;                                                      // slowest real group
    g.size_blk = (int)md.ext[md.n - 1];
    const long long rows = before * G * mid * G * after;
    while (md.n < 4) { md.ext[md.n] = 1; md.stride[md.n] = rows; ++md.n; }
    g.n_chunks = (int)(c / CB);
    g.n_tiles = (long long)g.nf0 * g.nf1 * g.n_chunks;
    return true;
}

// 5-D map over a panel operand: (column within block, then the row groups of MapDims).
// Column-chunked operands (cw < c) must be stacked blocks [nblk][rows][cw] (ld = cw, cstride = rows * cw): the block
// index then folds into the slowest real row group.
static int make_map5(CUtensorMap* map, const float* base, const Geom& g, const MapDims& md, const Lay& l, int64_t c,
                     int box_v, int box_u, int box_cols = CB) {
    TcEncodeTiledFn enc = tc_encode_fn();
    if (enc == nullptr) {
        set_error("kron_tc: cuTensorMapEncodeTiled unavailable");
        return 2;
    }
    const int64_t rows = g.su * G * (g.pos_blk == g.pos_u ? 1 : (int64_t)g.size_blk);   // su * 32 * (extent of the groups before u)
    int64_t nblk = 1;
    if (l.cw < c) {
        if (l.ld != l.cw || l.cstride != rows * l.cw || c % l.cw != 0) return 3;
        nblk = c / l.cw;
    }
    if (l.cw % CB != 0 || l.ld % 4 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return 3;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t box[5], estr[5] = {1u, 1u, 1u, 1u, 1u};
    gdim[0] = (cuuint64_t)l.cw;
    box[0] = (cuuint32_t)box_cols;
    for (int i = 0; i < 4; ++i) {
        const int pos = i + 1;
        gdim[pos] = (cuuint64_t)md.ext[i] * (pos == g.pos_blk ? (cuuint64_t)nblk : 1u);
        gstr[i] = (cuuint64_t)md.stride[i] * l.ld * 4;
        box[pos] = pos == g.pos_u ? (cuuint32_t)box_u : pos == g.pos_v ? (cuuint32_t)box_v : 1u;
    }
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("kron_tc: cuTensorMapEncodeTiled failed (%d) su=%lld sv=%lld ld=%lld cw=%lld", (int)rc, g.su, g.sv,
                  (long long)l.ld, (long long)l.cw);
        return 2;
    }
    return 0;
}

static Lay lay_of(const int64_t* h_lay, int i, int64_t c) {
    if (h_lay == nullptr) return Lay{c, c, 0};
    return Lay{h_lay[3 * i], h_lay[3 * i + 1], h_lay[3 * i + 2]};
}

constexpr int kApplyStages = 8;
constexpr int kGradStages = 4;

// How results leave the SM: TMA stores (one bulk tensor store per 16 KB box) when the rows of a tile are close together
// in memory, per-thread stores when every 64-byte row of the tile lands in a different 2 MB page (pair (0,1) of a
// 32^4 grid in a plain row-major panel: 1024 rows 1.7 MB apart) — measured on B200 the TMA store engine then falls
// behind the 256 storing threads.  WISKI_KTC_STORE=tma|stg overrides (A/B timings).
static bool use_direct_store(const Geom& g, const Lay& l) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("WISKI_KTC_STORE");
        mode = (e != nullptr && e[0] == 't') ? 1 : (e != nullptr && e[0] == 's') ? 2 : 0;
    }
    if (mode == 1) return false;
    if (mode == 2) return true;
    return g.sv * l.ld * 4 >= (1ll << 19);
}

// Destination maps of a pushing launch (see PushMaps).  dst[j] = where this rank's part starts in rank j's buffer:
//   mode 1: a [rows, c / n_dst] panel (ld = c / n_dst) holding MY rows of rank j's column block
//   mode 2: a [rows / n_dst, c] panel (ld = c) holding rank j's axis-0 lines of MY columns
static int fill_push(PushMaps& pm, Push& push, const Geom& g, const MapDims& md, const PushDst& pd, int64_t c, int box_v,
                     bool grad) {
    memset(&pm, 0, sizeof(pm));
    push.mode = pd.mode;
    push.box_u = push.u_loc = G;
    if (pd.n_dst < 1 || pd.n_dst > MAXP || pd.dst == nullptr) return 3;
    if (pd.mode == 1) {
        if (c % ((int64_t)pd.n_dst * CB) != 0) return 3;
        const int64_t cw = c / pd.n_dst;
        for (int j = 0; j < pd.n_dst; ++j)
            if (int rc = make_map5(&pm.m[j], pd.dst[j], g, md, Lay{cw, cw, 0}, cw, box_v, grad ? 8 : G)) return rc;
        return 0;
    }
    if ((pd.mode != 2 && pd.mode != 3) || g.pos_u != g.pos_blk || G % pd.n_dst != 0) return 3;     // u = the slowest grid axis
    if (pd.mode == 3 && (grad || c % (2 * CB) != 0)) return 3;
    push.u_loc = G / pd.n_dst;
    push.box_u = grad ? (push.u_loc < 8 ? push.u_loc : 8) : push.u_loc;
    MapDims md2 = md;
    md2.ext[g.pos_u - 1] = push.u_loc;
    for (int j = 0; j < pd.n_dst; ++j)
        if (int rc = make_map5(&pm.m[j], pd.dst[j], g, md2, Lay{c, c, 0}, c, box_v, push.box_u, pd.mode == 3 ? 2 * CB : CB))
            return rc;
    return 0;
}

}  // namespace ktc

// Y = (T_au x T_av) X on the tensor pipe for any two 32-point grid axes au < av (the other axes act as batch indices).
// Returns 3 when the shape / layout is not supported.
// push != NULL: Y is ignored, the result goes to the peers' buffers (ktc::PushMaps).
int tc_pair_apply_axes(const float* cols, int d, const int64_t* h_g, int64_t gmax, int au, int av, const float* X, float* Y,
                       int64_t c, cudaStream_t st, const int64_t* h_lay, long long* prof, const PushDst* push) {
    using namespace ktc;
    Geom g;
    MapDims md;
    if (!make_geom(g, md, d, h_g, au, av, c)) return 3;
    const Lay lx = lay_of(h_lay, 0, c);
    Lay ly = lay_of(h_lay, 1, c);
    CUtensorMap tmX, tmY;
    PushMaps pm;
    ApplyParams p;
    p.push = Push{0, G, G};
    if (int rc = make_map5(&tmX, X, g, md, lx, c, G, 8)) return rc;
    if (push != nullptr) {
        if (int rc = fill_push(pm, p.push, g, md, *push, c, 8, false)) return rc;
        tmY = tmX;
        ly = push->mode == 1 ? Lay{c / push->n_dst, c / push->n_dst, 0} : Lay{c, c, 0};
        if (push->mode == 3 && getenv("WISKI_PUSH_64B") != nullptr) {
            // (A/B switch for the timings in profiles/: WISKI_PUSH_64B=1 keeps the 16-column boxes of mode 2)
            p.push.mode = 2;
            if (int rc = fill_push(pm, p.push, g, md, PushDst{push->dst, push->n_dst, 2}, c, 8, false)) return rc;
        }
    } else {
        memset(&pm, 0, sizeof(pm));
        if (int rc = make_map5(&tmY, Y, g, md, ly, c, 8, G)) return rc;
    }
    p.cwx = (int)lx.cw;
    p.cwy = (int)ly.cw;
    p.ly = ly;
    p.Ydirect = (push == nullptr && use_direct_store(g, ly)) ? Y : nullptr;
    p.g = g;
    p.col_u = cols + (int64_t)au * gmax;
    p.col_v = cols + (int64_t)av * gmax;
    p.prof = prof;
    const size_t smem = 1024 + 4 * TIMG_B + (size_t)kApplyStages * STAGE_B + TILE_B + (2 * kApplyStages + 8) * 8 + 16;
    auto kfn = prof != nullptr ? pair_apply_tc_kernel<kApplyStages, true> : pair_apply_tc_kernel<kApplyStages, false>;
    WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_tc(attr)");
    const long long units = p.push.mode == 3 ? g.n_tiles / 2 : g.n_tiles;      // pair mode: a CTA takes 32-column groups
    const long long grid = units < kNumSMs ? units : kNumSMs;
    kfn<<<(unsigned)grid, NTHREADS, smem, st>>>(tmX, tmY, pm, p);
    WISKI_CHECK_LAUNCH("kron_tc(pair_apply)");
    count_launches(1);
    return 0;
}

int tc_pair_apply(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X, float* Y,
                  int64_t c, cudaStream_t st, const int64_t* h_lay, long long* prof) {
    return tc_pair_apply_axes(cols, d, h_g, gmax, 2 * pair, 2 * pair + 1, X, Y, c, st, h_lay, prof, nullptr);
}

// Directional backward pair pass on the tensor pipe for the axes au < av; out3 (3 doubles, accumulated):
// <grad_au, dirs_au>, <grad_av, dirs_av>, <Z', K' P'>.  Zout may be NULL.
int tc_pair_grad_dir_axes(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int au, int av,
                          const float* Z, const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st,
                          const int64_t* h_lay, const PushDst* push) {
    using namespace ktc;
    Geom g;
    MapDims md;
    if (!make_geom(g, md, d, h_g, au, av, c)) return 3;
    const Lay lz = lay_of(h_lay, 0, c), lp = lay_of(h_lay, 1, c);
    Lay lo = lay_of(h_lay, 2, c);
    CUtensorMap tmPu, tmZv, tmO;
    PushMaps pm;
    GradParams p;
    p.push = Push{0, G, G};
    memset(&pm, 0, sizeof(pm));
    if (int rc = make_map5(&tmPu, P, g, md, lp, c, G, 8)) return rc;
    if (int rc = make_map5(&tmZv, Z, g, md, lz, c, 8, G)) return rc;
    if (push != nullptr) {                       // Zout (T_v T_u Z) goes to the peers' buffers, split along axis u
        if (push->mode != 2) return 3;
        if (int rc = fill_push(pm, p.push, g, md, *push, c, G, true)) return rc;
        tmO = tmPu;
        lo = Lay{c, c, 0};
        Zout = const_cast<float*>(P);            // (only its non-NULL-ness is used below: the STORE instantiation)
    } else if (int rc = make_map5(&tmO, Zout != nullptr ? Zout : P, g, md, Zout != nullptr ? lo : lp, c, G, 8)) {
        return rc;
    }
    p.cwz = (int)lz.cw;
    p.cwp = (int)lp.cw;
    p.cwo = (int)lo.cw;
    p.lo = lo;
    p.Odirect = (push == nullptr && Zout != nullptr && use_direct_store(g, lo)) ? Zout : nullptr;
    p.g = g;
    p.col_u = cols + (int64_t)au * gmax;
    p.col_v = cols + (int64_t)av * gmax;
    p.dir_u = dirs + (int64_t)au * gmax;
    p.dir_v = dirs + (int64_t)av * gmax;
    p.out3 = out3;
    const size_t smem = 1024 + 8 * TIMG_B + (size_t)kGradStages * STAGE_B + 2 * TILE_B + (2 * kGradStages + 9) * 8 + 3 * 8 * 4 + 16;
    const long long grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
    if (Zout != nullptr) {
        auto kfn = pair_grad_dir_tc_kernel<true, kGradStages>;
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_tc(attr)");
        kfn<<<(unsigned)grid, NTHREADS, smem, st>>>(tmPu, tmZv, tmO, pm, p);
    } else {
        auto kfn = pair_grad_dir_tc_kernel<false, kGradStages>;
        WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "kron_tc(attr)");
        kfn<<<(unsigned)grid, NTHREADS, smem, st>>>(tmPu, tmZv, tmO, pm, p);
    }
    WISKI_CHECK_LAUNCH("kron_tc(pair_grad_dir)");
    count_launches(1);
    return 0;
}

int tc_pair_grad_dir(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int pair, const float* Z,
                     const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st, const int64_t* h_lay) {
    return tc_pair_grad_dir_axes(cols, dirs, d, h_g, gmax, 2 * pair, 2 * pair + 1, Z, P, Zout, c, out3, st, h_lay, nullptr);
}

}  // namespace wiski
