// rg_common.cuh -- shared host/device helpers of libregrid_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/regrid_b200.h"

namespace rg {

// ---------------------------------------------------------------------------
// error plumbing (thread-local message, no global mutable state besides it)
// ---------------------------------------------------------------------------
extern thread_local char g_err[512];

inline int fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

inline int cuda_fail(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

#define RG_CUDA(call)                                                \
    do {                                                             \
        cudaError_t _e = (call);                                     \
        if (_e != cudaSuccess) return rg::cuda_fail(_e, #call);      \
    } while (0)

#define RG_LAUNCH_CHECK(name)                                        \
    do {                                                             \
        cudaError_t _e = cudaGetLastError();                         \
        if (_e != cudaSuccess) return rg::cuda_fail(_e, name);       \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

constexpr int kNumSM = 148;  // B200

// bump allocator over the caller-provided workspace
struct Carver {
    char* base;
    size_t off;
    explicit Carver(void* p) : base((char*)p), off(0) {}
    template <class T>
    T* take(size_t n)
    {
        off = align_up(off);
        T* r = (T*)(base ? base + off : nullptr);
        off += n * sizeof(T);
        return r;
    }
    size_t total() const { return align_up(off); }
};

// exclusive scan int32 -> int64, n elements in, n+1 out (out[n] = total).
// `block_sums` scratch: ceil(n / kScanTile) + 1 int64.
constexpr int kScanTile = 2048;
int exclusive_scan_i32_i64(cudaStream_t st, const int32_t* in, int64_t* out, int64_t n, int64_t* block_sums);
int exclusive_scan_i32_i32(cudaStream_t st, const int32_t* in, int32_t* out, int64_t n, int64_t* block_sums);
inline size_t scan_scratch_elems(int64_t n) { return (size_t)ceil_div(n, kScanTile) + 2; }
// one launch (decoupled look-back); `status_zeroed`: scan_status_elems(n) words, ZERO at entry; optional report of the
// total with a capacity check (see rg_util.cu)
int exclusive_scan_i32_i64_single(cudaStream_t st, const int32_t* in, int64_t* out, int64_t n, unsigned long long* status_zeroed,
                                  int64_t capacity, int64_t* report, int32_t* cap_flag);
inline size_t scan_status_elems(int64_t n) { return (size_t)ceil_div(n, kScanTile) + 1; }

// ---------------------------------------------------------------------------
// device arithmetic: every fused multiply-add is explicit (compile with -fmad=false)
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// NumPy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum),
// the association np.add.reduceat uses per segment (reference: _weights_arrays.py:72).
// `S` = stride of the operands in doubles (fragments are stored as 16-byte (key, value) records: S = 2).
template <int S = 1>
__device__ inline double np_pairwise_sum(const double* a, int64_t n)
{
    if (n < 8) {
        double res = -0.0;
        for (int64_t i = 0; i < n; i++) res = dadd(res, a[i * S]);
        return res;
    } else if (n <= 128) {
        double r[8];
#pragma unroll
        for (int q = 0; q < 8; q++) r[q] = a[q * S];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int q = 0; q < 8; q++) r[q] = dadd(r[q], a[(i + q) * S]);
        }
        double res = dadd(dadd(dadd(r[0], r[1]), dadd(r[2], r[3])), dadd(dadd(r[4], r[5]), dadd(r[6], r[7])));
        for (; i < n; i++) res = dadd(res, a[i * S]);
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return dadd(np_pairwise_sum<S>(a, n2), np_pairwise_sum<S>(a + n2 * S, n - n2));
    }
}

// one segment of np.add.reduceat: out = a[0]; out += pairwise_sum(a[1:])
template <int S = 1>
__device__ inline double np_reduceat_segment(const double* a, int64_t n)
{
    double out = a[0];
    if (n > 1) out = dadd(out, np_pairwise_sum<S>(a + S, n - 1));
    return out;
}

#endif  // __CUDACC__

}  // namespace rg
