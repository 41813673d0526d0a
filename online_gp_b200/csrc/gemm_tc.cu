// fp32 tensor-core versions of the two real contractions on the WISKI path (sm_100a: tcgen05 + TMEM + TMA):
//   Gram        G[r x r2]  = A^T Bm        A [m x r],  Bm [m x r2]   (Q - I = L^T (K L);  k10)
//   panel rmul  Out[m x r2] = P M          P [m x r],  M [r x r2]    (L <- L (U S~), grad_KL = L grad_Q;  k7)
//
// Precision: operands are fp32, tcgen05.mma kind::tf32 keeps 10 mantissa bits, so each operand tile is split in
// shared memory into big = x with the low 13 mantissa bits cleared and small = x - big (exact), and every k-step
// issues three MMAs  big*big + big*small + small*big  into the same fp32 TMEM accumulator ("3xTF32"; the dropped
// small*small term is 2^-22 relative).  K is additionally cut into short slices whose partial tiles are summed in
// double by reduce_parts_kernel, so accumulator rounding does not grow with m.
//
// Structure of one CTA (256 threads, 1 CTA / SM), output tile 128 x BN (BN <= 256, multiple of 32):
//   warp 0      TMA producer: cp.async.bulk.tensor boxes of 32 floats (128 B, SWIZZLE_128B) -> smem stage ring
//   warp 1      MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::tf32, tcgen05.commit -> mbarriers
//   warp 2      TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4-7   operand split (generic-proxy read/modify/write of the landed tiles + fence.proxy.async), then the
//               epilogue: tcgen05.ld 32x32b.x32 -> registers -> global
// Operand layouts in smem follow the canonical UMMA layouts (CUTLASS cute/atom/mma_traits_sm100.hpp):
//   MN-major tf32 : SWIZZLE_128B_BASE32B (the only MN-major layout the hardware accepts for 32-bit operands; TMA
//                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 32 floats (MN, contiguous) x 4 K rows of 128 B;
//                   LBO = stride between MN atoms, SBO = stride between 4-row K groups (512 B).
//   K-major  SW128: rows of 32 floats (K), 8-row groups of 1024 B (SBO); K advances by 32 B inside the row.
#include <cuda.h>
#include "common.cuh"

namespace wiski {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): SWIZZLE_128B, version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                             uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B (K-major), 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, M = 128, N = n.
__host__ __device__ inline uint32_t make_idesc(bool a_mn_major, bool b_mn_major, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                          // c_format = F32
    d |= 2u << 7;                          // a_format = TF32
    d |= 2u << 10;                         // b_format = TF32
    d |= (a_mn_major ? 1u : 0u) << 15;     // a_major
    d |= (b_mn_major ? 1u : 0u) << 16;     // b_major
    d |= (uint32_t)(n >> 3) << 17;         // n_dim
    d |= (uint32_t)(128 >> 4) << 24;       // m_dim
    return d;
}

// How a work item's slice index z maps to operand coordinates (all zero / false = plain K-sliced GEMM):
//   K coordinate of A (B) starts at z * a_kstride (z * b_kstride);  the slice covers min(k_per_slice, Kdim - z * a_kstride);
//   with nb > 0 a work item instead walks nb "batches" of kbpb k-blocks each: batch o = z * nb + j adds
//   o * a_mstride (o * b_nstride) to the MN coordinate of A (B) and restarts K at 0 (sum over batches of A_o^T B_o);
//   m_fastest orders the tiles of one slice m-tile-fastest (CTAs running together share the B tile through L2).
struct Batching {
    int64_t a_kstride, b_kstride;
    int64_t a_mstride, b_nstride;
    int64_t n_batches;      // total number of batches (nb > 0)
    int nb, kbpb;
    int m_fastest;
    // column-chunked operand / result (blocks of cw columns, block j a row-major [rows, cw] matrix; 0 = plain):
    //   MN-major B given as a 2-D map over the stacked blocks [nblk * b_crows, b_cw]: column n -> (n % b_cw, k + (n / b_cw) * b_crows)
    //   result written to out + (n / o_cw) * o_cstride + row * ld_out + n % o_cw
    int64_t b_cw, b_crows;
    int64_t o_cw, o_cstride;
    // terms = 3: 3xTF32 split products (fp32-grade result); 1: one kind::tf32 product of the raw operands (the hardware
    //   truncates them to tf32: a uniform relative bias of ~-2^-11 plus ~2^-12 noise per product — used for gradient
    //   quantities only, see settings.backward_gemm_tf32_passes)
    int terms;
    // n_list > 0: a slice's tiles are the listed (m-tile, n-tile) pairs only (symmetric result: tiles entirely below the
    //   diagonal are skipped and mirrored by the reduction)
    int n_list;
    unsigned char list_m[24], list_n[24];
    // a_kwrap > 0: the K coordinate of A wraps after a_kwrap k-blocks while B keeps advancing:  A [B_0 ; B_1] = A B_0 + A B_1
    //   (panel GEMM with only the small operand split: B_0 = M, B_1 = M - tf32(M), see tc_panel_rmul_f32 terms = 2)
    int a_kwrap;
    // round_a (SPLIT = false kernels): the A tile is rounded to NEAREST tf32 in shared memory before the products (the
    //   tensor core itself truncates, which biases every product by ~-2^-12: harmless per value, not for a gradient that
    //   drives 8 000 Adam steps)
    int round_a;
    // o_ptrs[0] != NULL (with o_cw > 0): column block j of the result is written at o_ptrs[j] (row-major, pitch ld_out)
    // instead of out + j * o_cstride: the blocks go straight into the peers' NVLink-mapped buffers (PushDst mode 1)
    float* o_ptrs[8];
};

// ------------------------------------------------------------------------------------------------ kernel
// Persistent: each CTA walks work items  w = blockIdx.x, blockIdx.x + gridDim.x, ...  with
//   w -> (K slice z, m-tile, n-tile), n-tile fastest (CTAs that share an A tile / a K slice run together and share
//   operands through L2: ncu shows DRAM reads == one pass over each operand).
// D[128 x BN] = sum_k opA[k][m] * opB[k][n] over the K slice.
//   A_KMAJOR = false: A global [Kdim rows][Mdim cols] (MN-major; Gram).   true: A global [Mdim rows][Kdim cols] (rmul).
//   B_KMAJOR likewise for B ([Kdim][Ndim] MN-major, or [Ndim][Kdim] K-major).
// out: slice z writes out + z * slice_stride, row-major with leading dimension ld_out.
// Warp roles (384 threads): 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 4-7 epilogue | 8-11 operand split.
// Two TMEM accumulators (2 x BN <= 512 columns) let the MMAs of tile i+1 overlap the epilogue of tile i.
// SPLIT = true: a stage holds each operand tile twice (raw | fp32 remainder) for the 3xTF32 products; false: raw tiles
// only (single tf32 pass, bt.terms == 1) — half the shared memory per stage, which pays for BK = 32 (128-byte rows per
// TMA request instead of 64-byte ones: the A panel of the panel GEMM is fetched as row pieces 1.7 KB apart).
template <bool A_KMAJOR, bool B_KMAJOR, int BK, int STAGES, bool SPLIT = true>
__global__ void __launch_bounds__(384, 1)
tc_gemm3x_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
                 int64_t Mdim, int64_t Ndim, int64_t Kdim, int BN, int n_tiles_n, int n_tiles_m, int64_t n_work,
                 int64_t k_per_slice, int64_t ld_out, int64_t slice_stride, const Batching bt) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int A_BYTES = 128 * BK * 4;
    const int B_BYTES = BN * BK * 4;
    constexpr int NCOPY = SPLIT ? 2 : 1;
    const int STAGE_BYTES = NCOPY * (A_BYTES + B_BYTES);
    constexpr int STG_PITCH = 36;                               // words per staged row (32 + 4: conflict-free 16 B rows)
    float* staging = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);        // [4 warps][32][36]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES + 4 * 32 * STG_PITCH * 4);
    uint64_t* full_bar = bars;                  // [STAGES] TMA landed
    uint64_t* ready_bar = bars + STAGES;        // [STAGES] operands split
    uint64_t* empty_bar = bars + 2 * STAGES;    // [STAGES] MMAs retired
    uint64_t* tmem_full_bar = bars + 3 * STAGES;       // [2]
    uint64_t* tmem_empty_bar = bars + 3 * STAGES + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tiles_per_slice = bt.n_list > 0 ? (int64_t)bt.n_list : (int64_t)n_tiles_n * n_tiles_m;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&ready_bar[s], 128);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;

    // work item -> (m0, n0, k range)
    auto decode = [&](int64_t w, int64_t& m0, int64_t& n0, int64_t& kb0, int& nkb, int64_t& z) {
        z = w / tiles_per_slice;
        const int64_t t = w - z * tiles_per_slice;
        if (bt.n_list > 0) {
            m0 = (int64_t)bt.list_m[t] * 128;
            n0 = (int64_t)bt.list_n[t] * (int64_t)BN;
        } else if (bt.m_fastest) {
            m0 = (t % n_tiles_m) * 128;
            n0 = (t / n_tiles_m) * (int64_t)BN;
        } else {
            m0 = (t / n_tiles_n) * 128;
            n0 = (t % n_tiles_n) * (int64_t)BN;
        }
        kb0 = z * bt.a_kstride;
        if (bt.nb > 0) {
            int64_t nbh = bt.n_batches - z * bt.nb;
            if (nbh > bt.nb) nbh = bt.nb;
            nkb = (int)nbh * bt.kbpb;
        } else {
            int64_t k_end = kb0 + k_per_slice;
            if (k_end > Kdim) k_end = Kdim;
            nkb = (int)((k_end - kb0 + BK - 1) / BK);
        }
    };

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
                int64_t m0, n0, kb0, z;
                int nkb;
                decode(w, m0, n0, kb0, nkb, z);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + (size_t)stage * STAGE_BYTES;
                    uint8_t* sB = sA + NCOPY * A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(A_BYTES + B_BYTES));
                    int kA, kB, mA = (int)m0, nB = (int)n0;
                    if (bt.nb > 0) {
                        const int j = kb / bt.kbpb;
                        const int64_t o = z * bt.nb + j;
                        kA = kB = (kb - j * bt.kbpb) * BK;
                        mA += (int)(o * bt.a_mstride);
                        nB += (int)(o * bt.b_nstride);
                    } else {
                        kA = (int)(kb0 + (int64_t)(bt.a_kwrap > 0 ? kb % bt.a_kwrap : kb) * BK);
                        kB = (int)(z * bt.b_kstride + (int64_t)kb * BK);
                    }
                    if (A_KMAJOR) {
                        tma_load_2d(sA, &tmA, &full_bar[stage], kA, mA);                     // box {BK k, 128 rows}
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)                                         // boxes {32 cols, BK rows}
                            tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], mA + 32 * j, kA);
                    }
                    if (B_KMAJOR) {
                        tma_load_2d(sB, &tmB, &full_bar[stage], kB, nB);                     // box {BK k, BN rows}
                    } else if (bt.b_cw > 0) {
                        for (int j = 0; j < BN / 32; ++j) {
                            const int n = nB + 32 * j;
                            const int blk = n / (int)bt.b_cw;
                            tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], n - blk * (int)bt.b_cw,
                                        kB + blk * (int)bt.b_crows);
                        }
                    } else {
                        for (int j = 0; j < BN / 32; ++j)
                            tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], nB + 32 * j, kB);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(!A_KMAJOR, !B_KMAJOR, BN);
            int stage = 0;
            uint32_t phase = 0;
            int64_t it = 0;
            for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                int64_t m0, n0, kb0, z;
                int nkb;
                decode(w, m0, n0, kb0, nkb, z);
                const int acc = (int)(it & 1);
                mbar_wait(&tmem_empty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&ready_bar[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t aB = smem_u32(smem + (size_t)stage * STAGE_BYTES);   // A big
                    const uint32_t aS = aB + A_BYTES;                                     // A small
                    const uint32_t bB = aB + NCOPY * A_BYTES;                             // B big
                    const uint32_t bS = bB + B_BYTES;                                     // B small
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        uint64_t dAb, dAs;
                        if (A_KMAJOR) {
                            // K-major rows of BK floats: BK = 32 -> SWIZZLE_128B (8-row groups of 1024 B),
                            //                            BK = 16 -> SWIZZLE_64B  (8-row groups of 512 B)
                            dAb = make_desc(aB + ks * 32, 0, BK * 32, BK == 32 ? 2 : 4);
                            dAs = make_desc(aS + ks * 32, 0, BK * 32, BK == 32 ? 2 : 4);
                        } else {
                            dAb = make_desc(aB + ks * 1024, BK * 128, 512, 1);
                            dAs = make_desc(aS + ks * 1024, BK * 128, 512, 1);
                        }
                        const uint64_t dBb = B_KMAJOR ? make_desc(bB + ks * 32, 0, BK * 32, BK == 32 ? 2 : 4)
                                                      : make_desc(bB + ks * 1024, BK * 128, 512, 1);
                        const uint64_t dBs = B_KMAJOR ? make_desc(bS + ks * 32, 0, BK * 32, BK == 32 ? 2 : 4)
                                                      : make_desc(bS + ks * 1024, BK * 128, 512, 1);
                        if (!SPLIT || bt.terms == 1) {
                            umma_tf32(tmem_d, dAb, dBb, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        } else {
                            umma_tf32(tmem_d, dAs, dBb, idesc, (kb > 0 || ks > 0) ? 1u : 0u);   // small terms first
                            umma_tf32(tmem_d, dAb, dBs, idesc, 1u);
                            umma_tf32(tmem_d, dAb, dBb, idesc, 1u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------ operand split: small = x - trunc_tf32(x)
        const int t = threadIdx.x - 256;
        int stage = 0;
        uint32_t phase = 0;
        const int nvec = (A_BYTES + B_BYTES) / 16;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x) {
            int64_t m0, n0, kb0, z;
            int nkb;
            decode(w, m0, n0, kb0, nkb, z);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                uint8_t* sA = smem + (size_t)stage * STAGE_BYTES;
                if (!SPLIT && bt.round_a) {
                    for (int i = t; i < A_BYTES / 16; i += 128) {
                        uint4* a = reinterpret_cast<uint4*>(sA) + i;
                        uint4 v = *a;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.x) : "f"(__uint_as_float(v.x)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.y) : "f"(__uint_as_float(v.y)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.z) : "f"(__uint_as_float(v.z)));
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v.w) : "f"(__uint_as_float(v.w)));
                        *a = v;
                    }
                }
#pragma unroll 4
                for (int i = t; i < ((!SPLIT || bt.terms == 1) ? 0 : nvec); i += 128) {
                    // vectors [0, A_BYTES/16) belong to A (big at sA, small at sA + A_BYTES), the rest to B
                    const bool isA = i < A_BYTES / 16;
                    uint4* big = reinterpret_cast<uint4*>(isA ? sA : sA + NCOPY * A_BYTES) + (isA ? i : i - A_BYTES / 16);
                    uint4* sml = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(big) + (isA ? A_BYTES : B_BYTES));
                    uint4 v = *big;
                    uint4 s;
                    // kind::tf32 ignores the low 13 mantissa bits of the raw tile (measured: results identical to an
                    // explicitly masked "big" tile), so only the remainder tile has to be produced
                    s.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xffffe000u));
                    s.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xffffe000u));
                    s.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xffffe000u));
                    s.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xffffe000u));
                    *sml = s;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&ready_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue: TMEM -> regs -> per-warp smem transpose
        // -> 128-byte row segments in global memory.  Warp wq owns TMEM lanes (= output rows) 32 wq .. 32 wq + 31.
        const int wq = warp & 3;
        float* stg = staging + wq * 32 * STG_PITCH;
        int64_t it = 0;
        for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            int64_t m0, n0, kb0, z;
            int nkb;
            decode(w, m0, n0, kb0, nkb, z);
            const int acc = (int)(it & 1);
            mbar_wait(&tmem_full_bar[acc], (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float* obase0 = out + z * slice_stride;
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float* obase = obase0;
                int64_t ncol0 = n0 + c0;                 // column of this 32-wide group inside its block
                if (bt.o_cw > 0) {
                    const int64_t blk = ncol0 / bt.o_cw;
                    obase = bt.o_ptrs[0] != nullptr ? bt.o_ptrs[blk] : obase + blk * bt.o_cstride;
                    ncol0 -= blk * bt.o_cw;
                }
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                uint4* dst = reinterpret_cast<uint4*>(stg + lane * STG_PITCH);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                // 8 lanes cover one 128-byte row segment; 4 rows per instruction
#pragma unroll
                for (int rr = 0; rr < 32; rr += 4) {
                    const int lr = rr + (lane >> 3), c4 = (lane & 7) * 4;
                    const int64_t row = m0 + wq * 32 + lr, col = n0 + c0 + c4;
                    if (row < Mdim && col < Ndim)                     // Ndim % 4 == 0 (tc_shape_ok)
                        *reinterpret_cast<uint4*>(obase + row * ld_out + ncol0 + c4) =
                            *reinterpret_cast<const uint4*>(stg + lr * STG_PITCH + c4);
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 row-major tensor [rows][cols], box {32 cols, box_rows}, SWIZZLE_128B, OOB -> zeros.
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int box_rows, bool mn_major,
                    int box_cols = 32) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) {
        set_error("tc gemm: cuTensorMapEncodeTiled unavailable");
        return 2;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE,
                      mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                               : (box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("tc gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld", (int)rc, (long long)rows,
                  (long long)cols);
        return 2;
    }
    return 0;
}

static inline int pick_bn(int64_t n, int* ntiles) {
    int64_t npad = ceil_div(n, 32) * 32;
    int nt = (int)ceil_div(npad, 256);
    int bn = (int)(ceil_div(ceil_div(npad, nt), 32) * 32);
    *ntiles = (int)ceil_div(npad, bn);
    return bn;
}

constexpr int kGramBK = 16, kGramStages = 4;
constexpr int kRmulBK = 16, kRmulStages = 4;
constexpr int64_t kGramSliceRows = 2048;

static inline bool tc_shape_ok(int64_t m, int64_t r, int64_t r2) {
    return m >= 2048 && r >= 64 && r2 >= 64 && (r % 4) == 0 && (r2 % 4) == 0 && r <= 65536 && r2 <= 65536;
}

template <bool A_KMAJOR, bool B_KMAJOR, int BK, int STAGES, bool SPLIT = true>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, float* out, int64_t Mdim, int64_t Ndim, int64_t Kdim,
                  int bn, int ntiles, int64_t k_per_slice, int64_t nslices, int64_t ld_out, int64_t slice_stride,
                  cudaStream_t st, const char* name, const Batching* btp = nullptr) {
    Batching bt = {};
    bt.a_kstride = bt.b_kstride = k_per_slice;
    bt.terms = 3;
    if (btp != nullptr) bt = *btp;
    if (bt.terms != 1) bt.terms = 3;
    size_t smem = (size_t)STAGES * (SPLIT ? 2 : 1) * ((size_t)128 * BK * 4 + (size_t)bn * BK * 4) + 4 * 32 * 36 * 4 +
                  (3 * STAGES + 5) * 8 + 1024;
    auto kfn = tc_gemm3x_kernel<A_KMAJOR, B_KMAJOR, BK, STAGES, SPLIT>;
    WISKI_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), name);
    const int n_tiles_m = (int)ceil_div(Mdim, 128);
    const int64_t n_work = (bt.n_list > 0 ? (int64_t)bt.n_list : (int64_t)n_tiles_m * ntiles) * nslices;
    const int64_t grid = n_work < kNumSMs ? n_work : kNumSMs;
    kfn<<<(unsigned)grid, 384, smem, st>>>(tmA, tmB, out, Mdim, Ndim, Kdim, bn, ntiles, n_tiles_m, n_work, k_per_slice,
                                           ld_out, slice_stride, bt);
    WISKI_CHECK_LAUNCH(name);
    count_launches(1);
    return 0;
}

template <typename T>
__global__ void reduce_slices_kernel(const T* __restrict__ part, int64_t nparts, int64_t n, T* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double s = 0.0;
    for (int64_t p = 0; p < nparts; ++p) s += (double)part[p * n + e];
    out[e] = (T)s;
}

int64_t tc_gram_work_elems(int64_t m, int64_t r, int64_t r2) {
    if (!tc_shape_ok(m, r, r2)) return 0;
    return ceil_div(m, kGramSliceRows) * r * r2;
}

// nblk > 1: Bm is column-chunked, nblk blocks [m, r2 / nblk] stacked (block j = columns [j r2/nblk, (j+1) r2/nblk)).
// symmetric: the caller guarantees A^T Bm is symmetric (Q - I = L^T (K L), K symmetric): tiles entirely below the
// diagonal are not computed; the reduction mirrors them from the upper triangle.
template <typename T>
__global__ void reduce_slices_sym_kernel(const T* __restrict__ part, int64_t nparts, int64_t r, int bn, T* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r * r) return;
    int64_t i = e / r, j = e - i * r;
    // tile (i / 128, j / bn) was skipped iff its first row is at or beyond its last column + 1
    const bool skipped = (i / 128) * 128 >= (j / bn + 1) * (int64_t)bn;
    const int64_t src = skipped ? j * r + i : e;
    double s = 0.0;
    for (int64_t p = 0; p < nparts; ++p) s += (double)part[p * r * r + src];
    out[e] = (T)s;
}

int tc_gram_f32(const float* A, const float* Bm, int64_t m, int64_t r, int64_t r2, float* G, float* work,
                cudaStream_t st, int64_t nblk, bool symmetric) {
    if (!tc_shape_ok(m, r, r2)) return 3;
    if (nblk > 1 && (r2 % nblk != 0 || (r2 / nblk) % 32 != 0 || nblk * m >= (int64_t)1 << 31)) return 3;
    if (symmetric && r != r2) symmetric = false;
    CUtensorMap tmA, tmB;
    if (int rc = make_map(&tmA, A, m, r, kGramBK, true)) return rc;
    if (int rc = (nblk > 1 ? make_map(&tmB, Bm, nblk * m, r2 / nblk, kGramBK, true) : make_map(&tmB, Bm, m, r2, kGramBK, true)))
        return rc;
    int ntiles;
    int bn = pick_bn(r2, &ntiles);
    int64_t nslices = ceil_div(m, kGramSliceRows);
    Batching bt = {kGramSliceRows, kGramSliceRows, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (nblk > 1) { bt.b_cw = r2 / nblk; bt.b_crows = m; }
    if (symmetric) {
        const int ntm = (int)ceil_div(r, 128);
        int n_list = 0;
        for (int mt = 0; mt < ntm && n_list < 24; ++mt)
            for (int nt = 0; nt < ntiles && n_list < 24; ++nt)
                if (!((int64_t)mt * 128 >= (int64_t)(nt + 1) * bn)) {
                    bt.list_m[n_list] = (unsigned char)mt;
                    bt.list_n[n_list] = (unsigned char)nt;
                    ++n_list;
                }
        if (n_list >= 24 || n_list == ntm * ntiles) symmetric = false;      // nothing to skip (or too many tiles to list)
        else bt.n_list = n_list;
    }
    if (int rc = launch<false, false, kGramBK, kGramStages>(tmA, tmB, work, r, r2, m, bn, ntiles, kGramSliceRows, nslices, r2,
                                                     r * r2, st, "tc_gram", &bt))
        return rc;
    int64_t n = r * r2;
    if (symmetric)
        reduce_slices_sym_kernel<float><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(work, nslices, r, bn, G);
    else
        reduce_slices_kernel<float><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(work, nslices, n, G);
    WISKI_CHECK_LAUNCH("tc_gram(reduce)");
    count_launches(1);
    return 0;
}

// [M ; 0 ; M - tf32(M) ; 0]: the small operand of the panel GEMM and its fp32 remainder stacked along K, each padded to
// a multiple of 32 rows (work: 2 * ceil32(r) * r2 floats)
__global__ void stack_split_kernel(const float* __restrict__ M, int64_t r, int64_t r2, int64_t rpad, float* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * rpad * r2) return;
    const int64_t row = e / r2, col = e - row * r2;
    const int64_t k = row >= rpad ? row - rpad : row;
    float v = 0.f;
    if (k < r) {
        const uint32_t x = __float_as_uint(M[k * r2 + col]);
        v = row >= rpad ? __uint_as_float(x) - __uint_as_float(x & 0xffffe000u) : __uint_as_float(x);
    }
    out[e] = v;
}
int64_t tc_rmul_work_elems(int64_t r, int64_t r2) { return 2 * ceil_div(r, 32) * 32 * r2; }

// nblk > 1: Out is written column-chunked, nblk blocks [m, r2 / nblk] stacked.
// terms: 3 = 3xTF32 (both operands split in shared memory); 1 = one tf32 pass of the raw operands; 2 = the small operand
// M exact (M and its remainder stacked along K, built in `work`), the panel truncated to tf32 by the tensor core: the
// panel's truncation error is independent from row to row, so it averages out of every sum over grid rows taken
// downstream (the hyper-parameter gradient), while an error in M would be shared by all rows.
int tc_panel_rmul_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, float* Out, cudaStream_t st,
                      int64_t nblk, int terms, float* work, const PushDst* push) {
    if (!tc_shape_ok(m, r, r2)) return 3;
    if (terms == 2 && work == nullptr) terms = 3;
    if (nblk > 1 && (r2 % nblk != 0 || (r2 / nblk) % 32 != 0)) return 3;
    if (push != nullptr && (push->mode != 1 || push->n_dst != nblk || nblk < 2 || nblk > 8)) return 3;
    CUtensorMap tmA, tmB;
    int ntiles;
    int bn = pick_bn(r2, &ntiles);
    Batching bt = {r, r, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    bt.terms = terms == 3 ? 3 : 1;
    int64_t ld_out = r2;
    if (nblk > 1) { bt.o_cw = r2 / nblk; bt.o_cstride = m * (r2 / nblk); ld_out = r2 / nblk; }
    if (push != nullptr) {
        for (int j = 0; j < push->n_dst; ++j) bt.o_ptrs[j] = push->dst[j];
        Out = push->dst[0];
    }
    if (terms == 1 || terms == 2) {
        constexpr int BK1 = 32, ST1 = 4;                                    // 4 x (16 KB + bn * 128 B) <= 176 KB
        if (int rc = make_map(&tmA, P, m, r, 128, false, BK1)) return rc;
        if (terms == 1) {
            if (int rc = make_map(&tmB, M, r, r2, BK1, true)) return rc;
            return launch<true, false, BK1, ST1, false>(tmA, tmB, Out, m, r2, r, bn, ntiles, r, 1, ld_out, 0, st,
                                                        "tc_panel_rmul(tf32x1)", &bt);
        }
        const int64_t rpad = ceil_div(r, 32) * 32;
        stack_split_kernel<<<(unsigned)ceil_div(2 * rpad * r2, 256), 256, 0, st>>>(M, r, r2, rpad, work);
        WISKI_CHECK_LAUNCH("tc_panel_rmul(stack)");
        count_launches(1);
        if (int rc = make_map(&tmB, work, 2 * rpad, r2, BK1, true)) return rc;
        bt.a_kstride = bt.b_kstride = 2 * rpad;
        bt.a_kwrap = (int)(rpad / BK1);
        bt.round_a = 1;
        return launch<true, false, BK1, ST1, false>(tmA, tmB, Out, m, r2, 2 * rpad, bn, ntiles, 2 * rpad, 1, ld_out, 0, st,
                                                    "tc_panel_rmul(tf32, M split)", &bt);
    }
    if (int rc = make_map(&tmA, P, m, r, 128, false, kRmulBK)) return rc;   // K-major A: box {BK k, 128 rows}
    if (int rc = make_map(&tmB, M, r, r2, kRmulBK, true)) return rc;   // MN-major B: box {32 cols, BK rows}
    return launch<true, false, kRmulBK, kRmulStages>(tmA, tmB, Out, m, r2, r, bn, ntiles, r, 1, ld_out, 0, st, "tc_panel_rmul",
                                                     &bt);
}

// debug: Out = P @ Mt^T with Mt [r2 x r] row-major (both operands K-major)
int tc_panel_rmul_nt_f32(const float* P, int64_t m, int64_t r, const float* Mt, int64_t r2, float* Out, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    int ntiles;
    int bn = pick_bn(r2, &ntiles);
    if (int rc = make_map(&tmA, P, m, r, 128, false, kRmulBK)) return rc;
    if (int rc = make_map(&tmB, Mt, r2, r, bn, false, kRmulBK)) return rc;
    return launch<true, true, kRmulBK, kRmulStages>(tmA, tmB, Out, m, r2, r, bn, ntiles, r, 1, r2, 0, st, "tc_panel_rmul_nt");
}

// ------------------------------------------------------------------------------------------------ Kronecker axes
// One Kronecker axis with g >= 64 points is a real contraction (2 g flop per element), so it runs on the tensor
// cores as a batched GEMM over the panel viewed as [outer, g, inner]:
//   apply     Y[o] = T X[o]                 A = T (dense symmetric Toeplitz, g x g, MN-major), B = X[o] (g x inner, MN-major)
//   contract  S = sum_o Z[o] P[o]^T (g x g)  A = Z[o], B = P[o] (both K-major, K = inner), then acc[k] += sum_{|a-b|=k} S[a][b]
__global__ void toeplitz_dense_kernel(const float* __restrict__ col, int g, float* __restrict__ T) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)g * g) return;
    int a = (int)(e / g), b = (int)(e % g);
    T[e] = col[a > b ? a - b : b - a];
}

// acc64[k] += sum_{a} S[a][a+k] (+ S[a+k][a] for k > 0); one block per diagonal
__global__ void fold_diagonals_kernel(const float* __restrict__ S, int g, double* __restrict__ acc64) {
    const int k = blockIdx.x;
    double s = 0.0;
    for (int a = threadIdx.x; a + k < g; a += blockDim.x) {
        s += (double)S[(int64_t)a * g + a + k];
        if (k > 0) s += (double)S[(int64_t)(a + k) * g + a];
    }
    __shared__ double red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        acc64[k] += t;
    }
}

bool tc_axis_ok(int64_t g, int64_t outer, int64_t inner) {
    return g >= 64 && g <= 65536 && (g % 4) == 0 && (inner % 4) == 0 && inner >= 64 && outer >= 1 &&
           outer * g < (int64_t)1 << 31 && inner < (int64_t)1 << 31;
}

static void contract_plan(int64_t g, int64_t outer, int64_t inner, int64_t* nslices, int64_t* kps, int* nb) {
    int nt;
    pick_bn(g, &nt);
    const int64_t tiles = ceil_div(g, 128) * nt;
    const int64_t want = ceil_div((int64_t)4 * kNumSMs, tiles);          // ~4 waves of work items
    if (outer == 1) {
        int64_t k = ceil_div(ceil_div(inner, want), 16) * 16;
        if (k < 2048) k = 2048;
        *kps = k;
        *nslices = ceil_div(inner, k);
        *nb = 0;
    } else {
        int64_t ns = outer < want ? outer : want;
        *nb = (int)ceil_div(outer, ns);
        *nslices = ceil_div(outer, (int64_t)*nb);
        *kps = inner;
    }
}

int64_t tc_axis_work_elems(int64_t g, int64_t outer, int64_t inner, int contract) {
    if (!tc_axis_ok(g, outer, inner)) return 0;
    if (!contract) return g * g;
    int64_t nslices, kps;
    int nb;
    contract_plan(g, outer, inner, &nslices, &kps, &nb);
    return (nslices + 1) * g * g;
}

int tc_axis_apply_f32(const float* X, float* Y, const float* col, int64_t g, int64_t outer, int64_t inner, float* work,
                      cudaStream_t st) {
    if (!tc_axis_ok(g, outer, inner)) return 3;
    toeplitz_dense_kernel<<<(unsigned)ceil_div(g * g, 256), 256, 0, st>>>(col, (int)g, work);
    WISKI_CHECK_LAUNCH("tc_axis_apply(toeplitz)");
    count_launches(1);
    CUtensorMap tmA, tmB;
    if (int rc = make_map(&tmA, work, g, g, kGramBK, true)) return rc;
    if (int rc = make_map(&tmB, X, outer * g, inner, kGramBK, true)) return rc;
    int ntiles;
    int bn = pick_bn(inner, &ntiles);
    Batching bt = {0, g, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0};
    return launch<false, false, kGramBK, kGramStages>(tmA, tmB, Y, g, inner, g, bn, ntiles, g, outer, inner, g * inner, st,
                                                      "tc_axis_apply", &bt);
}

int tc_axis_contract_f32(const float* Z, const float* P, int64_t g, int64_t outer, int64_t inner, double* acc64,
                         float* work, cudaStream_t st) {
    if (!tc_axis_ok(g, outer, inner)) return 3;
    int64_t nslices, kps;
    int nb;
    contract_plan(g, outer, inner, &nslices, &kps, &nb);
    int ntiles;
    int bn = pick_bn(g, &ntiles);
    CUtensorMap tmA, tmB;
    if (int rc = make_map(&tmA, Z, outer * g, inner, 128, false, kRmulBK)) return rc;
    if (int rc = make_map(&tmB, P, outer * g, inner, bn, false, kRmulBK)) return rc;
    Batching bt = {kps, kps, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (nb > 0) bt = Batching{0, 0, g, g, outer, nb, (int)ceil_div(inner, kRmulBK), 0, 0, 0, 0, 0};
    float* parts = work + g * g;
    if (int rc = launch<true, true, kRmulBK, kRmulStages>(tmA, tmB, parts, g, g, inner, bn, ntiles, kps, nslices, g, g * g,
                                                          st, "tc_axis_contract", &bt))
        return rc;
    const int64_t n = g * g;
    reduce_slices_kernel<float><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(parts, nslices, n, work);
    fold_diagonals_kernel<<<(unsigned)g, 256, 0, st>>>(work, (int)g, acc64);
    WISKI_CHECK_LAUNCH("tc_axis_contract(fold)");
    count_launches(2);
    return 0;
}

}  // namespace wiski
