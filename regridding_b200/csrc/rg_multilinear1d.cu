// rg_multilinear1d.cu -- 1D multilinear weights and the saved-weights ordering of raw triplets.
//
// Replaces regridding/_weights/_weights_multilinear.py:
//   find_indices(method="searchsorted")            :19-25  (np.searchsorted(left) - 1 with the edge fix-ups of
//                                                           _find_indices_searchsorted.py:42-58)
//   clamp / "below" fix-up, bounds                 :105-119
//   _weights_from_indices_multilinear_1d           :142-206 (w1 = (x - x0) / (x1 - x0), w0 = 1 - w1, optional
//                                                           weights_input factors, NaN poisoning for bounds="nan")
// and the ordering half of _coalesce (regridding/_weights/_weights_arrays.py:44-73): weights() returns every element
// sorted by (indices_input, indices_output).  A multilinear element never holds a repeated pair (two distinct
// input indices per output point), so no merge is needed; the sort is stable like the reference's argsort.
//
// k_multilinear1d writes the raw triplets in the builder's emission order (2 i, 2 i + 1 per output point i) together
// with the 64-bit sort key ((spectrum * n_in + input) * n_out + output); a device radix sort (CUB -- plumbing, as a
// library sort would be) orders (key, position) and k_gather_triplets writes the public arrays.  The same
// rg_sort_triplets serves the 2D multilinear weights.
#include "rg_common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace rg {

constexpr int kBoundsExtrapolate1 = 0, kBoundsNan1 = 1, kBoundsRaise1 = 2;

__global__ void k_multilinear1d(int64_t D, int64_t n, int64_t m, const double* __restrict__ x_in,
                                const double* __restrict__ x_out, const double* __restrict__ w_in, int bounds,
                                int64_t* __restrict__ ii, int64_t* __restrict__ io, double* __restrict__ vv,
                                uint64_t* __restrict__ keys, unsigned long long* __restrict__ n_outside)
{
    const int64_t d = blockIdx.y;
    const double* xi = x_in + d * n;
    const int64_t index_max = n - 2;
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double p = x_out[d * m + i];
        // np.searchsorted(xi, p, side="left") - 1 (NaN sorts last), _find_indices_searchsorted.py:42-45
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo) / 2;
            const double a = xi[mid];
            const bool less = (a < p) || (p != p && a == a);
            if (less) lo = mid + 1;
            else hi = mid;
        }
        int64_t r = lo - 1;
        bool outside;
        if (p == xi[0]) { r = 0; outside = false; }           // :47-48
        else outside = (r < 0) || (r > index_max);             // fill_value (a huge positive index) :49-58, wml.py:108
        // wml.py:115-119: the side a point fell out on is recovered from the coordinate
        const bool below = p < xi[0];
        int64_t i0 = outside ? index_max : r;                  // np.clip(fill_value, 0, index_max) == index_max
        if (i0 < 0) i0 = 0;
        if (below) i0 = 0;
        const int64_t i1 = i0 + 1;
        const double x0 = xi[i0], x1 = xi[i1];
        double w1 = ddiv(dsub(p, x0), dsub(x1, x0));           // :185-186 (njit without fastmath: plain IEEE)
        double w0 = dsub(1.0, w1);
        if (w_in) {
            w0 = dmul(w0, w_in[d * n + i0]);
            w1 = dmul(w1, w_in[d * n + i1]);
        }
        if (outside) {
            bad++;
            if (bounds == kBoundsNan1) { w0 = __longlong_as_double(0x7ff8000000000000LL); w1 = w0; }
        }
        const int64_t e = (d * m + i) * 2;
        ii[e] = i0; io[e] = i; vv[e] = w0;
        ii[e + 1] = i1; io[e + 1] = i; vv[e + 1] = w1;
        const uint64_t base = ((uint64_t)d * (uint64_t)n) * (uint64_t)m + (uint64_t)i;
        keys[e] = base + (uint64_t)i0 * (uint64_t)m;
        keys[e + 1] = base + (uint64_t)i1 * (uint64_t)m;
    }
    if (bad) atomicAdd(n_outside, bad);
}

__global__ void k_triplet_keys(int64_t n, int64_t n_out, const int64_t* __restrict__ ii, const int64_t* __restrict__ io,
                               uint64_t* __restrict__ keys, uint32_t* __restrict__ pos)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (keys) keys[e] = (uint64_t)ii[e] * (uint64_t)n_out + (uint64_t)io[e];
    pos[e] = (uint32_t)e;
}

__global__ void k_gather_triplets(int64_t n, const uint32_t* __restrict__ order, const int64_t* __restrict__ ii,
                                  const int64_t* __restrict__ io, const double* __restrict__ vv,
                                  int64_t* __restrict__ oi, int64_t* __restrict__ oo, double* __restrict__ ov)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t s = order[e];
    oi[e] = ii[s];
    oo[e] = io[s];
    ov[e] = vv[s];
}

struct SortLayout {
    uint64_t *keys, *keys_alt;
    uint32_t *pos, *pos_alt;
    void* cub;
    size_t cub_bytes;
    int64_t *ii, *io;   // raw triplets of rg_multilinear1d_weights
    double* vv;
    unsigned long long* counter;
    size_t bytes;
};

static int key_bits(uint64_t key_max)
{
    int b = 1;
    while (b < 64 && (key_max >> b) != 0) b++;
    return b;
}

static SortLayout sort_layout(void* ws, int64_t n, bool raw)
{
    SortLayout l;
    Carver c(ws);
    l.keys = c.take<uint64_t>(n);
    l.keys_alt = c.take<uint64_t>(n);
    l.pos = c.take<uint32_t>(n);
    l.pos_alt = c.take<uint32_t>(n);
    l.cub_bytes = 0;
    cub::DoubleBuffer<uint64_t> kb(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, l.cub_bytes, kb, vb, (int)n, 0, 64, (cudaStream_t)0);
    l.cub = c.take<char>(l.cub_bytes + 16);
    l.ii = l.io = nullptr;
    l.vv = nullptr;
    if (raw) {
        l.ii = c.take<int64_t>(n);
        l.io = c.take<int64_t>(n);
        l.vv = c.take<double>(n);
    }
    l.counter = c.take<unsigned long long>(2);
    l.bytes = c.total();
    return l;
}

// stable sort of (key, position) and gather into the public arrays
static int sort_and_gather(cudaStream_t st, const SortLayout& l, int64_t n, uint64_t key_max, const int64_t* ii,
                           const int64_t* io, const double* vv, int64_t* oi, int64_t* oo, double* ov)
{
    cub::DoubleBuffer<uint64_t> kb(l.keys, l.keys_alt);
    cub::DoubleBuffer<uint32_t> vb(l.pos, l.pos_alt);
    size_t bytes = l.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(l.cub, bytes, kb, vb, (int)n, 0, key_bits(key_max), st);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
    k_gather_triplets<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, vb.Current(), ii, io, vv, oi, oo, ov);
    RG_LAUNCH_CHECK("k_gather_triplets");
    return RG_OK;
}

}  // namespace rg

using namespace rg;

extern "C" int rg_sort_triplets_workspace_bytes(int64_t n, size_t* bytes_host)
{
    if (n < 0 || n >= INT32_MAX || !bytes_host) return fail(RG_E_ARG, "rg_sort_triplets_workspace_bytes: bad argument");
    *bytes_host = sort_layout(nullptr, n > 0 ? n : 1, false).bytes;
    return RG_OK;
}

extern "C" int rg_sort_triplets(int device, void* stream, int64_t n, int64_t n_in, int64_t n_out,
                                const int64_t* ii, const int64_t* io, const double* vv,
                                int64_t* out_ii, int64_t* out_io, double* out_vv,
                                void* workspace, size_t workspace_bytes)
{
    if (n < 0 || n >= INT32_MAX || n_in <= 0 || n_out <= 0) return fail(RG_E_ARG, "rg_sort_triplets: bad argument");
    if (n == 0) return RG_OK;
    if (!ii || !io || !vv || !out_ii || !out_io || !out_vv || !workspace) return fail(RG_E_ARG, "rg_sort_triplets: null pointer");
    if ((double)n_in * (double)n_out >= 1.8e19) return fail(RG_E_TOO_LARGE, "rg_sort_triplets: key overflow");
    const SortLayout l = sort_layout(workspace, n, false);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_sort_triplets: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    k_triplet_keys<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, n_out, ii, io, l.keys, l.pos);
    RG_LAUNCH_CHECK("k_triplet_keys");
    return sort_and_gather(st, l, n, (uint64_t)n_in * (uint64_t)n_out, ii, io, vv, out_ii, out_io, out_vv);
}

extern "C" int rg_multilinear1d_workspace_bytes(int64_t D, int64_t m, size_t* bytes_host)
{
    if (D < 0 || m < 0 || !bytes_host || 2 * D * m >= INT32_MAX) return fail(RG_E_ARG, "rg_multilinear1d_workspace_bytes: bad argument");
    *bytes_host = sort_layout(nullptr, 2 * D * m > 0 ? 2 * D * m : 1, true).bytes;
    return RG_OK;
}

extern "C" int rg_multilinear1d_weights(int device, void* stream, int64_t D, int64_t n, int64_t m,
                                        const double* x_in, const double* x_out, const double* w_in_or_null, int bounds,
                                        int64_t* out_ii, int64_t* out_io, double* out_vv, int64_t* n_outside_host,
                                        void* workspace, size_t workspace_bytes)
{
    if (D < 0 || D > 65535 || n < 2 || m < 0 || bounds < 0 || bounds > 2)
        return fail(RG_E_ARG, "rg_multilinear1d_weights: bad argument (at most 65535 spectra per call)");
    if (n_outside_host) *n_outside_host = 0;
    const int64_t total = 2 * D * m;
    if (total == 0) return RG_OK;
    if (total >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_multilinear1d_weights: more than 2^31 triplets per call");
    if ((double)D * (double)n * (double)m >= 1.8e19) return fail(RG_E_TOO_LARGE, "rg_multilinear1d_weights: key overflow");
    if (!x_in || !x_out || !out_ii || !out_io || !out_vv || !workspace) return fail(RG_E_ARG, "rg_multilinear1d_weights: null pointer");
    const SortLayout l = sort_layout(workspace, total, true);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_multilinear1d_weights: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaMemsetAsync(l.counter, 0, sizeof(unsigned long long) * 2, st));
    {
        dim3 grid((unsigned)(ceil_div(m, 256) < 1024 ? ceil_div(m, 256) : 1024), (unsigned)D);
        k_multilinear1d<<<grid, 256, 0, st>>>(D, n, m, x_in, x_out, w_in_or_null, bounds, l.ii, l.io, l.vv, l.keys, l.counter);
        RG_LAUNCH_CHECK("k_multilinear1d");
    }
    k_triplet_keys<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(total, 1, l.ii, l.io, nullptr, l.pos);
    RG_LAUNCH_CHECK("k_triplet_keys");
    int rc = sort_and_gather(st, l, total, (uint64_t)D * (uint64_t)n * (uint64_t)m, l.ii, l.io, l.vv, out_ii, out_io, out_vv);
    if (rc) return rc;
    if (n_outside_host) {
        unsigned long long c = 0;
        RG_CUDA(cudaMemcpyAsync(&c, l.counter, sizeof(c), cudaMemcpyDeviceToHost, st));
        RG_CUDA(cudaStreamSynchronize(st));
        *n_outside_host = (int64_t)c;
    }
    return RG_OK;
}
