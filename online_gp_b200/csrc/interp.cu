// k1/k2/k3/k13: cubic-convolution SKI interpolation stencils, W-gather and W^T-scatter (sm_100a).
//
// Replaces GPyTorch Interpolation.interpolate / left_interp / _sparse_left_interp_t reached
// from online_gp/models/batched_fixed_noise_online_gp.py:22-28,143,205-210,261.  The index arithmetic reproduces
// the reference's op order with explicit round-to-nearest intrinsics (subtract, true divide, floor — never a
// multiply by a reciprocal, never an FMA contraction), so indices are bit-exact against the oracle.
#include <stdarg.h>
#include "common.cuh"

namespace wiski {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }
static int g_background = 0;
int background_mode() { return g_background; }

template <typename T>
struct InterpParams {
    int d;
    int64_t g[WISKI_MAX_DIMS];
    int64_t coeff[WISKI_MAX_DIMS];  // prod_{j>i} g_j
    T lo[WISKI_MAX_DIMS];
    T delta[WISKI_MAX_DIMS];
    T first4[WISKI_MAX_DIMS * 4];
    T last4[WISKI_MAX_DIMS * 4];
    T gmin[WISKI_MAX_DIMS];
    T gmax[WISKI_MAX_DIMS];
};

// Keys cubic convolution weight for scaled distance s (a = -0.5), reference op order.
template <typename T>
__device__ __forceinline__ T cubic_weight(T s) {
    T U = fabs(s);
    if (U < T(1)) {
        // ((1.5*U - 2.5)*U)*U + 1
        return add_rn(mul_rn(mul_rn(sub_rn(mul_rn(T(1.5), U), T(2.5)), U), U), T(1));
    }
    // ((-0.5*U + 2.5)*U - 4)*U + 2
    return add_rn(mul_rn(sub_rn(mul_rn(add_rn(mul_rn(T(-0.5), U), T(2.5)), U), T(4)), U), T(2));
}

template <typename T>
__device__ __forceinline__ T cubic_weight_grad(T s) {
    T U = fabs(s);
    T sg = s < T(0) ? T(-1) : T(1);
    if (U < T(1)) return sg * (T(4.5) * U * U - T(5) * U);
    if (U < T(2)) return sg * (T(-1.5) * U * U + T(5) * U - T(4));
    return T(0);
}

// One thread per stencil entry (n, k).  digit_i(k) = (k / 4^(d-1-i)) % 4 selects the tap in dimension i.
template <typename T>
__global__ void interp_fwd_kernel(const T* __restrict__ x, int64_t q, int64_t s, InterpParams<T> p,
                                  int64_t* __restrict__ idx, T* __restrict__ val, int* __restrict__ oob) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= q * s) return;
    int64_t n = tid / s;
    int k = (int)(tid % s);
    int64_t flat = 0;
    T v = T(1);
    bool bad = false;
    for (int i = 0; i < p.d; ++i) {
        T xi = x[n * p.d + i];
        if (sub_rn(xi, p.gmin[i]) < T(-1e-7) || sub_rn(xi, p.gmax[i]) > T(1e-7)) bad = true;
        T u = div_rn(sub_rn(xi, p.lo[i]), p.delta[i]);
        T lower = floor(u);
        T frac = sub_rn(u, lower);
        lower = lower - T(1);
        int tap = (k >> (2 * (p.d - 1 - i))) & 3;
        T w;
        int64_t base;
        if (lower < T(0)) {  // left boundary: one-hot on the closest of the first four grid points
            int best = 0;
            T bd = fabs(sub_rn(p.first4[i * 4], xi));
            for (int t = 1; t < 4; ++t) {
                T dd = fabs(sub_rn(p.first4[i * 4 + t], xi));
                if (dd < bd) { bd = dd; best = t; }
            }
            w = (tap == best) ? T(1) : T(0);
            base = 0;
        } else if (lower > T(p.g[i] - 4)) {  // right boundary
            int best = 0;
            T bd = fabs(sub_rn(p.last4[i * 4], xi));
            for (int t = 1; t < 4; ++t) {
                T dd = fabs(sub_rn(p.last4[i * 4 + t], xi));
                if (dd < bd) { bd = dd; best = t; }
            }
            w = (tap == best) ? T(1) : T(0);
            base = p.g[i] - 4;
        } else {
            // scaled distance = frac + [1, 0, -1, -2][tap]
            w = cubic_weight(add_rn(frac, T(1 - tap)));
            base = (int64_t)lower;
        }
        flat += (base + tap) * p.coeff[i];
        v = mul_rn(v, w);
    }
    idx[tid] = flat;
    val[tid] = v;
    if (bad && oob != nullptr && k == 0) atomicOr(oob, 1);
}

// grad_x[n,i] = sum_k grad_val[n,k] * d val[n,k] / d x[n,i];  one warp per (n, i).
template <typename T>
__global__ void interp_bwd_kernel(const T* __restrict__ x, int64_t q, int64_t s, InterpParams<T> p,
                                  const T* __restrict__ gval, T* __restrict__ gx) {
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= q * p.d) return;
    int64_t n = warp / p.d;
    int di = (int)(warp % p.d);
    T acc = T(0);
    for (int k = lane; k < s; k += 32) {
        T v = T(1);
        for (int i = 0; i < p.d; ++i) {
            T xi = x[n * p.d + i];
            T u = div_rn(sub_rn(xi, p.lo[i]), p.delta[i]);
            T lower = floor(u);
            T frac = sub_rn(u, lower);
            lower = lower - T(1);
            int tap = (k >> (2 * (p.d - 1 - i))) & 3;
            bool boundary = (lower < T(0)) || (lower > T(p.g[i] - 4));
            T sd = add_rn(frac, T(1 - tap));
            T w;
            if (i == di) w = boundary ? T(0) : cubic_weight_grad(sd) / p.delta[i];
            else if (boundary) {
                const T* pts = (lower < T(0)) ? &p.first4[i * 4] : &p.last4[i * 4];
                int best = 0;
                T bd = fabs(sub_rn(pts[0], xi));
                for (int t = 1; t < 4; ++t) {
                    T dd = fabs(sub_rn(pts[t], xi));
                    if (dd < bd) { bd = dd; best = t; }
                }
                w = (tap == best) ? T(1) : T(0);
            } else w = cubic_weight(sd);
            v *= w;
        }
        acc += gval[n * s + k] * v;
    }
    acc = warp_sum(acc);
    if (lane == 0) gx[n * p.d + di] = acc;
}

// ---- gather: out[n, :] = sum_k val[n,k] * src[idx[n,k], :]
// Wide rows (c >= 32): block = 8 warps; warps split the stencil, lanes split a 128-column chunk; smem reduce.
template <typename T>
__global__ void gather_wide_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val, int64_t s,
                                   const T* __restrict__ src, int64_t c, T* __restrict__ out) {
    constexpr int kWarps = 8, kCols = 128;
    __shared__ T red[kWarps][kCols];
    int64_t n = blockIdx.x;
    int64_t c0 = (int64_t)blockIdx.y * kCols;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T acc[4] = {T(0), T(0), T(0), T(0)};
    for (int64_t k = warp; k < s; k += kWarps) {
        int64_t row = idx[n * s + k];
        T w = val[n * s + k];
        const T* rp = src + row * c + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t col = lane + 32 * j;
            if (c0 + col < c) acc[j] += w * rp[col];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp][lane + 32 * j] = acc[j];
    __syncthreads();
    if (threadIdx.x < kCols) {
        T t = T(0);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w][threadIdx.x];
        if (c0 + threadIdx.x < c) out[n * c + c0 + threadIdx.x] = t;
    }
}

// Narrow rows (c < 32): one warp per (n, column); lanes split the stencil.
template <typename T>
__global__ void gather_narrow_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val, int64_t q, int64_t s,
                                     const T* __restrict__ src, int64_t c, T* __restrict__ out) {
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= q * c) return;
    int64_t n = warp / c, col = warp % c;
    T acc = T(0);
    for (int64_t k = lane; k < s; k += 32) acc += val[n * s + k] * src[idx[n * s + k] * c + col];
    acc = warp_sum(acc);
    if (lane == 0) out[n * c + col] = acc;
}

// ---- scatter-add: dst[idx[n,k], :] += val[n,k] * src[n, :]
template <typename T>
__global__ void scatter_add_kernel(const int64_t* __restrict__ idx, const T* __restrict__ val, int64_t q, int64_t s,
                                   const T* __restrict__ src, int64_t c, T* __restrict__ dst) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= q * s * c) return;
    int64_t col = tid % c;
    int64_t nk = tid / c;
    int64_t n = nk / s;
    atomicAdd(&dst[idx[nk] * c + col], val[nk] * src[n * c + col]);
}

template <typename T>
static int fill_params(InterpParams<T>& p, int d, const int64_t* h_g, const T* h_lo, const T* h_delta,
                       const T* h_first4, const T* h_last4, const T* h_gmin, const T* h_gmax) {
    WISKI_CHECK_ARG(d >= 1 && d <= WISKI_MAX_DIMS, "interp: d=%d outside [1,%d]", d, WISKI_MAX_DIMS);
    p.d = d;
    int64_t coeff = 1;
    for (int i = d - 1; i >= 0; --i) {
        WISKI_CHECK_ARG(h_g[i] >= 4, "interp: grid size %lld < 4 in dim %d", (long long)h_g[i], i);
        p.g[i] = h_g[i];
        p.coeff[i] = coeff;
        coeff *= h_g[i];
        p.lo[i] = h_lo[i];
        p.delta[i] = h_delta[i];
        for (int t = 0; t < 4; ++t) {
            p.first4[i * 4 + t] = h_first4 ? h_first4[i * 4 + t] : T(0);
            p.last4[i * 4 + t] = h_last4 ? h_last4[i * 4 + t] : T(0);
        }
        p.gmin[i] = h_gmin ? h_gmin[i] : T(0);
        p.gmax[i] = h_gmax ? h_gmax[i] : T(0);
    }
    return 0;
}

template <typename T>
static int interp_fwd(const T* x, int64_t q, int d, const int64_t* h_g, const T* h_lo, const T* h_delta,
                      const T* h_first4, const T* h_last4, const T* h_gmin, const T* h_gmax, int64_t* idx, T* val,
                      int* oob, void* stream) {
    InterpParams<T> p;
    if (int rc = fill_params(p, d, h_g, h_lo, h_delta, h_first4, h_last4, h_gmin, h_gmax)) return rc;
    WISKI_CHECK_ARG(h_first4 && h_last4 && h_gmin && h_gmax, "interp_fwd: null grid description");
    if (q == 0) return 0;
    int64_t s = 1LL << (2 * d);
    int64_t total = q * s;
    interp_fwd_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(x, q, s, p, idx, val, oob);
    WISKI_CHECK_LAUNCH("interp_fwd");
    count_launches(1);
    return 0;
}

template <typename T>
static int interp_bwd(const T* x, int64_t q, int d, const int64_t* h_g, const T* h_lo, const T* h_delta,
                      const T* h_first4, const T* h_last4, const T* gval, T* gx, void* stream) {
    InterpParams<T> p;
    if (int rc = fill_params(p, d, h_g, h_lo, h_delta, h_first4, h_last4, (const T*)nullptr, (const T*)nullptr))
        return rc;
    if (q == 0) return 0;
    int64_t s = 1LL << (2 * d);
    int64_t warps = q * d;
    interp_bwd_kernel<T><<<(unsigned)ceil_div(warps * 32, 128), 128, 0, as_stream(stream)>>>(x, q, s, p, gval, gx);
    WISKI_CHECK_LAUNCH("interp_bwd");
    count_launches(1);
    return 0;
}

template <typename T>
static int gather(const int64_t* idx, const T* val, int64_t q, int64_t s, const T* src, int64_t m, int64_t c, T* out,
                  void* stream) {
    WISKI_CHECK_ARG(q >= 0 && s > 0 && c > 0 && m > 0, "gather: bad sizes q=%lld s=%lld c=%lld m=%lld", (long long)q,
                    (long long)s, (long long)c, (long long)m);
    if (q == 0) return 0;
    if (c >= 32) {
        dim3 grid((unsigned)q, (unsigned)ceil_div(c, 128));
        gather_wide_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(idx, val, s, src, c, out);
    } else {
        gather_narrow_kernel<T><<<(unsigned)ceil_div(q * c * 32, 128), 128, 0, as_stream(stream)>>>(idx, val, q, s,
                                                                                                   src, c, out);
    }
    WISKI_CHECK_LAUNCH("gather");
    count_launches(1);
    return 0;
}

template <typename T>
static int scatter_add(const int64_t* idx, const T* val, int64_t q, int64_t s, const T* src, int64_t m, int64_t c,
                       T* dst, void* stream) {
    WISKI_CHECK_ARG(q >= 0 && s > 0 && c > 0 && m > 0, "scatter_add: bad sizes");
    if (q == 0) return 0;
    int64_t total = q * s * c;
    scatter_add_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(idx, val, q, s, src, c, dst);
    WISKI_CHECK_LAUNCH("scatter_add");
    count_launches(1);
    return 0;
}

}  // namespace wiski

extern "C" {

const char* wiski_last_error(void) { return wiski::g_err; }
int wiski_abi_version(void) { return 1; }
long long wiski_launch_count(void) { return __atomic_load_n(&wiski::g_launches, __ATOMIC_RELAXED); }
int wiski_set_background(int on) {
    const int prev = wiski::g_background;
    if (on >= 0) wiski::g_background = on != 0;
    return prev;
}

int wiski_interp_fwd_f32(const float* x, int64_t q, int d, const int64_t* h_g, const float* h_lo,
                         const float* h_delta, const float* h_first4, const float* h_last4, const float* h_gmin,
                         const float* h_gmax, int64_t* idx, float* val, int* oob_flag, void* stream) {
    return wiski::interp_fwd<float>(x, q, d, h_g, h_lo, h_delta, h_first4, h_last4, h_gmin, h_gmax, idx, val,
                                    oob_flag, stream);
}
int wiski_interp_fwd_f64(const double* x, int64_t q, int d, const int64_t* h_g, const double* h_lo,
                         const double* h_delta, const double* h_first4, const double* h_last4, const double* h_gmin,
                         const double* h_gmax, int64_t* idx, double* val, int* oob_flag, void* stream) {
    return wiski::interp_fwd<double>(x, q, d, h_g, h_lo, h_delta, h_first4, h_last4, h_gmin, h_gmax, idx, val,
                                     oob_flag, stream);
}
int wiski_interp_bwd_f32(const float* x, int64_t q, int d, const int64_t* h_g, const float* h_lo,
                         const float* h_delta, const float* h_first4, const float* h_last4, const float* grad_val,
                         float* grad_x, void* stream) {
    return wiski::interp_bwd<float>(x, q, d, h_g, h_lo, h_delta, h_first4, h_last4, grad_val, grad_x, stream);
}
int wiski_interp_bwd_f64(const double* x, int64_t q, int d, const int64_t* h_g, const double* h_lo,
                         const double* h_delta, const double* h_first4, const double* h_last4, const double* grad_val,
                         double* grad_x, void* stream) {
    return wiski::interp_bwd<double>(x, q, d, h_g, h_lo, h_delta, h_first4, h_last4, grad_val, grad_x, stream);
}
int wiski_gather_f32(const int64_t* idx, const float* val, int64_t q, int64_t s, const float* src, int64_t m,
                     int64_t c, float* out, void* stream) {
    return wiski::gather<float>(idx, val, q, s, src, m, c, out, stream);
}
int wiski_gather_f64(const int64_t* idx, const double* val, int64_t q, int64_t s, const double* src, int64_t m,
                     int64_t c, double* out, void* stream) {
    return wiski::gather<double>(idx, val, q, s, src, m, c, out, stream);
}
int wiski_scatter_add_f32(const int64_t* idx, const float* val, int64_t q, int64_t s, const float* src, int64_t m,
                          int64_t c, float* dst, void* stream) {
    return wiski::scatter_add<float>(idx, val, q, s, src, m, c, dst, stream);
}
int wiski_scatter_add_f64(const int64_t* idx, const double* val, int64_t q, int64_t s, const double* src, int64_t m,
                          int64_t c, double* dst, void* stream) {
    return wiski::scatter_add<double>(idx, val, q, s, src, m, c, dst, stream);
}

}  // extern "C"
