// rg_locate.cu -- cell location: the 2D walk (new 2D extension of find_indices built on
// the reference's internal locators) and the 1D brute / searchsorted kernels.
#include "rg_common.cuh"
#include "rg_geom.cuh"
#include "rg_boundary.cuh"

namespace rg {

// ---------------------------------------------------------------------------
// 2D: lowest-index containing cell (index_of_point_brute semantics,
// regridding/_weights/_weights_conservative_2d/_grids.py:223-279), found by Newton
// iteration + 3x3 lowest-index resolve (the role of index_of_point_secant,
// _grids.py:356-463).
// Pass 1: one thread per point, coalesced over the point arrays; the grid stays
// L2-resident.  A point Newton cannot place in a cell is classified EXACTLY: outside the
// vertex bounding box or boundary winding number 0 (the reference's own line-start
// test, c2d.py:308-317) => `fill`; otherwise it is queued for the exhaustive pass 2.
// ---------------------------------------------------------------------------
__global__ void k_locate_points(GridView g, Boundary bnd, const double* __restrict__ bbox,
                                int64_t n, const double* __restrict__ px, const double* __restrict__ py,
                                int64_t fill, int64_t* __restrict__ out, uint8_t* __restrict__ pending,
                                int32_t* __restrict__ n_pending)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double x = px[p], y = py[p];
    const int r = locate_newton(g, x, y, 0.5 * g.nx, 0.5 * g.ny);
    uint8_t pend = 0;
    if (r >= 0) {
        out[p] = r;
    } else {
        const bool in_box = bbox[0] <= x && x <= bbox[2] && bbox[1] <= y && y <= bbox[3];
        if (!in_box || boundary_winding(bnd, x, y) == 0.0) {
            out[p] = fill;
        } else {
            out[p] = fill;
            pend = 1;
            atomicAdd(n_pending, 1);
        }
    }
    pending[p] = pend;
}

// Pass 2: exhaustive and exact (index_of_point_brute).  One warp per pending point scans
// all cells in row-major order and keeps the first containing one.
__global__ void k_locate_brute(GridView g, int64_t n, const double* __restrict__ px, const double* __restrict__ py,
                               int64_t fill, int64_t* __restrict__ out, const uint8_t* __restrict__ pending)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ncy = g.ny - 1;
    const int64_t nc = (int64_t)(g.nx - 1) * ncy;
    for (int64_t p = warp; p < n; p += nwarps) {
        if (!pending[p]) continue;
        const double x = px[p], y = py[p];
        int64_t best = -1;
        for (int64_t base = 0; base < nc && best < 0; base += 32) {
            const int64_t c = base + lane;
            bool hit = false;
            if (c < nc) hit = cell_contains(g, (int)(c / ncy), (int)(c % ncy), x, y);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) best = base + (__ffs(m) - 1);
        }
        if (lane == 0) out[p] = best >= 0 ? best : fill;
    }
}

struct LocateLayout {
    Boundary bnd;
    double* bbox;
    uint8_t* pending;
    int32_t* counter;
    size_t bytes;
};

static LocateLayout locate_layout(void* ws, int64_t nx, int64_t ny, int64_t n_points)
{
    LocateLayout l;
    Carver c(ws);
    carve_boundary(c, l.bnd, nx, ny);
    l.bbox = c.take<double>(4);
    l.pending = c.take<uint8_t>((size_t)n_points + 1);
    l.counter = c.take<int32_t>(4);
    l.bytes = c.total();
    return l;
}

// ---------------------------------------------------------------------------
// 1D find_indices.  x_in (D, n) is staged through shared memory per row when it fits.
// brute:        first q with x[q] <= p <= x[q+1] (inclusive), _find_indices_brute.py:38-49
// searchsorted: np.searchsorted(left) - 1 with fix-ups, _find_indices_searchsorted.py:42-58
// ---------------------------------------------------------------------------
__global__ void k_find_1d(int method, int64_t D, int64_t n, int64_t m,
                          const double* __restrict__ x_in, const double* __restrict__ x_out,
                          int64_t fill, int64_t* __restrict__ out)
{
    const int64_t d = blockIdx.y;
    const double* xi = x_in + d * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double p = x_out[d * m + i];
        int64_t r;
        if (method == 0) {
            r = fill;
            for (int64_t q = 0; q < n - 1; q++) {
                if (xi[q] <= p && p <= xi[q + 1]) {
                    r = q;
                    break;
                }
            }
        } else {
            int64_t lo = 0, hi = n;  // first index with xi[idx] >= p (NaN sorts last)
            while (lo < hi) {
                const int64_t mid = lo + (hi - lo) / 2;
                const double a = xi[mid];
                const bool less = (a < p) || (p != p && a == a);
                if (less) lo = mid + 1;
                else hi = mid;
            }
            r = lo - 1;
            if (p == xi[0]) r = 0;
            else if (r < 0) r = fill;
            else if (r > n - 2) r = fill;
        }
        out[d * m + i] = r;
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_find_indices_2d_workspace_bytes(int64_t nx, int64_t ny, int64_t n_points, size_t* bytes_host)
{
    if (nx < 2 || ny < 2 || n_points < 0 || !bytes_host) return fail(RG_E_ARG, "rg_find_indices_2d_workspace_bytes: bad argument");
    *bytes_host = locate_layout(nullptr, nx, ny, n_points).bytes;
    return RG_OK;
}

extern "C" int rg_find_indices_2d(int device, void* stream, int64_t nx, int64_t ny,
                                  const double* x, const double* y,
                                  int64_t n_points, const double* px, const double* py,
                                  int64_t fill, int64_t* cell_flat,
                                  void* workspace, size_t workspace_bytes)
{
    if (nx < 2 || ny < 2 || !x || !y || n_points < 0) return fail(RG_E_ARG, "rg_find_indices_2d: bad argument");
    if (nx * ny >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_find_indices_2d: grid too large");
    if (n_points == 0) return RG_OK;
    if (!px || !py || !cell_flat || !workspace) return fail(RG_E_ARG, "rg_find_indices_2d: null pointer");
    LocateLayout l = locate_layout(workspace, nx, ny, n_points);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_find_indices_2d: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const GridView g{ x, y, (int)nx, (int)ny };
    int rc = build_boundary(st, g, l.bnd, l.bbox);
    if (rc) return rc;
    RG_CUDA(cudaMemsetAsync(l.counter, 0, sizeof(int32_t) * 4, st));
    k_locate_points<<<(unsigned)ceil_div(n_points, 256), 256, 0, st>>>(g, l.bnd, l.bbox, n_points, px, py, fill,
                                                                      cell_flat, l.pending, l.counter);
    RG_LAUNCH_CHECK("k_locate_points");
    int32_t n_pending = 0;
    RG_CUDA(cudaMemcpyAsync(&n_pending, l.counter, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    if (n_pending > 0) {
        int64_t warps = n_points < 148 * 64 ? n_points : 148 * 64;
        k_locate_brute<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, st>>>(g, n_points, px, py, fill, cell_flat, l.pending);
        RG_LAUNCH_CHECK("k_locate_brute");
    }
    return RG_OK;
}

extern "C" int rg_find_indices_1d(int device, void* stream, int method, int64_t D, int64_t n, int64_t m,
                                  const double* x_in, const double* x_out, int64_t fill, int64_t* out)
{
    if (D < 0 || n < 1 || m < 0 || (method != 0 && method != 1)) return fail(RG_E_ARG, "rg_find_indices_1d: bad argument");
    if (D == 0 || m == 0) return RG_OK;
    if (!x_in || !x_out || !out) return fail(RG_E_ARG, "rg_find_indices_1d: null pointer");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 128;
    for (int64_t d0 = 0; d0 < D; d0 += 65535) {
        const int64_t nd = D - d0 < 65535 ? D - d0 : 65535;
        int64_t gx = ceil_div(m, T);
        if (gx > 4096) gx = 4096;
        dim3 grid((unsigned)gx, (unsigned)nd);
        k_find_1d<<<grid, T, 0, st>>>(method, nd, n, m, x_in + d0 * n, x_out + d0 * m, fill, out + d0 * m);
        RG_LAUNCH_CHECK("k_find_1d");
    }
    return RG_OK;
}
