// rg_util.cu -- error strings and the device-wide exclusive scan used by the
// count -> scan -> emit compaction steps (no atomics decide any output order).
#include "rg_common.cuh"

namespace rg {

thread_local char g_err[512] = "";

// ---- three-kernel scan: per-tile sums, scan of the tile sums, per-tile scan + offset ----

template <int THREADS, int ITEMS>
__global__ void k_scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums)
{
    __shared__ int64_t warp_sums[THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * (THREADS * ITEMS);
    int64_t acc = 0;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
        int64_t idx = base + (int64_t)q * THREADS + threadIdx.x;
        if (idx < n) acc += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t s = 0;
        for (int w = 0; w < THREADS / 32; w++) s += warp_sums[w];
        tile_sums[blockIdx.x] = s;
    }
}

// single block: exclusive scan of the tile sums in place; writes the grand total at [ntiles]
__global__ void k_scan_tile_offsets(int64_t* tile_sums, int64_t ntiles)
{
    __shared__ int64_t carry;
    __shared__ int64_t warp_tot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += blockDim.x) {
        int64_t idx = base + threadIdx.x;
        int64_t v = idx < ntiles ? tile_sums[idx] : 0;
        int64_t inc = v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int64_t w = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
            int64_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_tot[lane] = winc - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        int64_t excl = carry + warp_tot[wid] + (inc - v);
        if (idx < ntiles) tile_sums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}

template <int THREADS, int ITEMS, class OutT>
__global__ void k_scan_apply(const int32_t* __restrict__ in, OutT* __restrict__ out, int64_t n,
                             const int64_t* __restrict__ tile_offsets, int64_t ntiles)
{
    // blocked arrangement: thread t owns ITEMS consecutive elements
    __shared__ int64_t warp_tot[THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * (THREADS * ITEMS) + (int64_t)threadIdx.x * ITEMS;
    int32_t v[ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
        v[q] = (base + q < n) ? in[base + q] : 0;
        tsum += v[q];
    }
    int64_t inc = tsum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < wid; w++) woff += warp_tot[w];
    int64_t run = tile_offsets[blockIdx.x] + woff + (inc - tsum);
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
        if (base + q < n) out[base + q] = (OutT)run;
        run += v[q];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = (OutT)tile_offsets[ntiles];
}

template <class OutT>
static int scan_impl(cudaStream_t st, const int32_t* in, OutT* out, int64_t n, int64_t* block_sums)
{
    constexpr int THREADS = 256, ITEMS = kScanTile / THREADS;
    if (n <= 0) {
        RG_CUDA(cudaMemsetAsync(out, 0, sizeof(OutT), st));
        return RG_OK;
    }
    const int64_t ntiles = ceil_div(n, kScanTile);
    k_scan_tile_sums<THREADS, ITEMS><<<(unsigned)ntiles, THREADS, 0, st>>>(in, n, block_sums);
    RG_LAUNCH_CHECK("k_scan_tile_sums");
    k_scan_tile_offsets<<<1, 1024, 0, st>>>(block_sums, ntiles);
    RG_LAUNCH_CHECK("k_scan_tile_offsets");
    k_scan_apply<THREADS, ITEMS, OutT><<<(unsigned)ntiles, THREADS, 0, st>>>(in, out, n, block_sums, ntiles);
    RG_LAUNCH_CHECK("k_scan_apply");
    return RG_OK;
}

// ---------------------------------------------------------------------------
// Single-launch exclusive scan (int32 -> int64) with decoupled look-back: a tile publishes its aggregate, then its
// inclusive prefix, in one 64-bit status word (2 state bits + 62 value bits); warp 0 of a tile looks back over 32
// predecessors at a time until it meets an inclusive prefix.  Tiles are numbered by a ticket counter, so a tile's
// predecessors always started before it.  `status` (scan_status_elems(n) words, the ticket in the last one) must be
// ZERO at entry.  The last tile also stores the grand total at out[n] and, when `report` is given, at report[0],
// raising *cap_flag when it exceeds `capacity` (or the int32 range the fragment offsets are kept in).
// ---------------------------------------------------------------------------
constexpr unsigned long long kScanAgg = 1ull << 62, kScanInc = 2ull << 62, kScanMask = (1ull << 62) - 1;

template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS) k_scan_lookback(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n,
                                                           unsigned long long* status, int64_t ntiles, int64_t capacity,
                                                           int64_t* __restrict__ report, int32_t* __restrict__ cap_flag)
{
    __shared__ int64_t warp_tot[THREADS / 32];
    __shared__ int64_t s_prefix;
    __shared__ unsigned s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd((unsigned*)(status + ntiles), 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * (THREADS * ITEMS) + (int64_t)threadIdx.x * ITEMS;
    int32_t v[ITEMS];
    int64_t tsum = 0;
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
        v[q] = (base + q < n) ? in[base + q] : 0;
        tsum += v[q];
    }
    int64_t inc = tsum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int64_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        if (w < wid) woff += warp_tot[w];
        total += warp_tot[w];
    }
    if (wid == 0) {
        volatile unsigned long long* vs = status;
        if (lane == 0) vs[tile] = (tile == 0 ? kScanInc : kScanAgg) | (unsigned long long)total;
        int64_t prefix = 0;
        if (tile > 0) {
            for (int64_t j = tile - 1 - lane;; j -= 32) {   // (tiles before the first count as inclusive 0)
                unsigned long long sv = kScanInc;
                if (j >= 0) {
                    do { sv = vs[j]; } while ((sv >> 62) == 0ull);
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (sv >> 62) == 2ull);
                const int first = inc_mask ? __ffs(inc_mask) - 1 : 32;
                int64_t val = lane <= first ? (int64_t)(sv & kScanMask) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (inc_mask) break;
            }
            if (lane == 0) vs[tile] = kScanInc | (unsigned long long)(prefix + total);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == ntiles - 1) {
                const int64_t grand = prefix + total;
                out[n] = grand;
                if (report) {
                    report[0] = grand;
                    if (grand > capacity || grand >= INT32_MAX) *cap_flag = 1;
                }
            }
        }
    }
    __syncthreads();
    int64_t run = s_prefix + woff + (inc - tsum);
#pragma unroll
    for (int q = 0; q < ITEMS; q++) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
    }
}

int exclusive_scan_i32_i64_single(cudaStream_t st, const int32_t* in, int64_t* out, int64_t n, unsigned long long* status_zeroed,
                                  int64_t capacity, int64_t* report, int32_t* cap_flag)
{
    constexpr int THREADS = 256, ITEMS = kScanTile / THREADS;
    if (n <= 0) return RG_E_ARG;
    const int64_t ntiles = ceil_div(n, kScanTile);
    k_scan_lookback<THREADS, ITEMS><<<(unsigned)ntiles, THREADS, 0, st>>>(in, out, n, status_zeroed, ntiles, capacity, report, cap_flag);
    RG_LAUNCH_CHECK("k_scan_lookback");
    return RG_OK;
}

int exclusive_scan_i32_i64(cudaStream_t st, const int32_t* in, int64_t* out, int64_t n, int64_t* block_sums)
{
    return scan_impl<int64_t>(st, in, out, n, block_sums);
}

int exclusive_scan_i32_i32(cudaStream_t st, const int32_t* in, int32_t* out, int64_t n, int64_t* block_sums)
{
    return scan_impl<int32_t>(st, in, out, n, block_sums);
}

// ---------------------------------------------------------------------------
// transposed weights (regridding/_weights/_weights_transposed/_weights_transposed.py)
// ---------------------------------------------------------------------------

// cell_length (c1d/_grids.py:10-35) of S stacked edge arrays: out[s, i] = x[s, i+1] - x[s, i]
__global__ void k_cell_length(int64_t S, int64_t n, const double* __restrict__ x, double* __restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * (n - 1)) return;
    const int64_t s = e / (n - 1), i = e - s * (n - 1);
    out[e] = dsub(x[s * n + i + 1], x[s * n + i]);
}

// wT.py:236-249: values / square(w[ii]) (only with weights), then * vol_in[ii] / vol_out[io], NumPy's evaluation
// order (no fastmath); negative indices wrap like NumPy fancy indexing
__global__ void k_transpose_conservative(int64_t nnz, int64_t n_in, int64_t n_out,
                                         const int64_t* __restrict__ ii, const int64_t* __restrict__ io,
                                         const double* __restrict__ v, const double* __restrict__ vol_in,
                                         const double* __restrict__ vol_out, const double* __restrict__ w_in,
                                         double* __restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int64_t a = ii[e], b = io[e];
    if (a < 0) a += n_in;
    if (b < 0) b += n_out;
    double x = v[e];
    if (w_in) {
        const double w = w_in[a];
        x = ddiv(x, dmul(w, w));
    }
    out[e] = ddiv(dmul(x, vol_in[a]), vol_out[b]);
}

}  // namespace rg

extern "C" int rg_cell_length_1d(int device, void* stream, int64_t S, int64_t n, const double* x, double* length)
{
    if (S < 0 || n < 2 || !x || !length) return rg::fail(RG_E_ARG, "rg_cell_length_1d: bad argument");
    if (S == 0) return RG_OK;
    RG_CUDA(cudaSetDevice(device));
    rg::k_cell_length<<<(unsigned)rg::ceil_div(S * (n - 1), 256), 256, 0, (cudaStream_t)stream>>>(S, n, x, length);
    RG_LAUNCH_CHECK("k_cell_length");
    return RG_OK;
}

extern "C" int rg_transpose_conservative(int device, void* stream, int64_t nnz, int64_t n_in, int64_t n_out,
                                         const int64_t* indices_input, const int64_t* indices_output,
                                         const double* values, const double* volume_input,
                                         const double* volume_output, const double* weights_input,
                                         double* values_transposed)
{
    if (nnz < 0 || n_in <= 0 || n_out <= 0) return rg::fail(RG_E_ARG, "rg_transpose_conservative: bad sizes");
    if (nnz == 0) return RG_OK;
    if (!indices_input || !indices_output || !values || !volume_input || !volume_output || !values_transposed)
        return rg::fail(RG_E_ARG, "rg_transpose_conservative: null pointer");
    RG_CUDA(cudaSetDevice(device));
    rg::k_transpose_conservative<<<(unsigned)rg::ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
        nnz, n_in, n_out, indices_input, indices_output, values, volume_input, volume_output, weights_input,
        values_transposed);
    RG_LAUNCH_CHECK("k_transpose_conservative");
    return RG_OK;
}

// fp64 FMA-chain throughput of this device: the denominator of the build's fp64 roofline (the
// vector fp64 peak is not in MEASURED_PEAKS.json).  8 independent DFMA chains per thread, all SMs full.
namespace rg {
__global__ void __launch_bounds__(256) k_fp64_fma_chain(int iters, double seed, double* __restrict__ sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0,
           a7 = a0 + 7.0;
    const double m = 1.0 + 1e-9, c = 1e-9;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;  // keeps the chains alive
}
}  // namespace rg

extern "C" int rg_measure_fp64_peak(int device, void* stream, double* tflops_host)
{
    if (!tflops_host) return rg::fail(RG_E_ARG, "rg_measure_fp64_peak: null output");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = nullptr;
    RG_CUDA(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1;
    RG_CUDA(cudaEventCreate(&e0));
    RG_CUDA(cudaEventCreate(&e1));
    const int iters = 8192, blocks = rg::kNumSM * 8, threads = 256;
    rg::k_fp64_fma_chain<<<blocks, threads, 0, st>>>(256, 1.0, sink);  // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; rep++) {
        RG_CUDA(cudaEventRecord(e0, st));
        rg::k_fp64_fma_chain<<<blocks, threads, 0, st>>>(iters, 1.0, sink);
        RG_CUDA(cudaEventRecord(e1, st));
        RG_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        RG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops_host = best;
    return RG_OK;
}

extern "C" const char* rg_last_error_string(void) { return rg::g_err; }
extern "C" int rg_version(void) { return 100; }
