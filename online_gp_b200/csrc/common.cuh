// Shared helpers for libwiski_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/wiski_b200.h"

namespace wiski {

void set_error(const char* fmt, ...);
void count_launches(int n);   // kernels enqueued by this library (bench.py reports it as gpu_launches)

#define WISKI_CHECK_ARG(cond, ...)                \
    do {                                          \
        if (!(cond)) {                            \
            ::wiski::set_error(__VA_ARGS__);      \
            return 1;                             \
        }                                         \
    } while (0)

#define WISKI_CHECK_LAUNCH(name)                                                          \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::wiski::set_error("%s: CUDA error: %s", name, cudaGetErrorString(_e));       \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

#define WISKI_CHECK_CUDA(expr, name)                                                      \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::wiski::set_error("%s: CUDA error: %s", name, cudaGetErrorString(_e));       \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// "Background" launches (wiski_set_background): HBM-bound passes the host layer puts on a side stream UNDER a tensor-bound
// kernel (settings.overlap_root_update).  They keep a small resident footprint (1 - 2 CTAs per SM) so that the persistent
// tcgen05 kernels and the latency-bound r x r chain still find block slots, registers and shared memory, and they ask for
// the largest shared-memory carve-out — the split the tensor-core kernels need: an SM cannot change its L1 / shared split
// while any CTA is resident, so with the default split every kernel that needs more shared memory would wait for the
// background launch to drain (measured: the main stream stalled for the full 0.8 ms).
int background_mode();
template <typename K>
static inline cudaError_t apply_background_carveout(K kfn) {
    return cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout,
                                background_mode() ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
}

// Destinations of a launch that pushes its result into the peers' (NVLink-mapped) buffers; see kron_tc.cu (PushMaps)
// and gemm_tc.cu (o_ptrs).  dst[j] = where this rank's part starts in rank j's buffer.
struct PushDst {
    float* const* dst;
    int n_dst;
    int mode;      // 1: column block j -> rank j;  2: axis-0 range j of the rows -> rank j
};

// kron_tc.cu: two Kronecker axes per pass on the tensor pipe (tcgen05 / TMEM / TMA); return 3 for shapes they do not take
int tc_pair_apply_axes(const float* cols, int d, const int64_t* h_g, int64_t gmax, int au, int av, const float* X, float* Y,
                       int64_t c, cudaStream_t st, const int64_t* h_lay, long long* prof, const PushDst* push = nullptr);
int tc_pair_grad_dir_axes(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax, int au, int av,
                          const float* Z, const float* P, float* Zout, int64_t c, double* out3, cudaStream_t st,
                          const int64_t* h_lay, const PushDst* push = nullptr);

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// round-to-nearest, never contracted into FMA: used where the reference's op order must be reproduced exactly
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace wiski
