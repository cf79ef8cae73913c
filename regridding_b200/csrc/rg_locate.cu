// rg_locate.cu -- cell location: the 2D walk (new 2D extension of find_indices built on
// the reference's internal locators) and the 1D brute / searchsorted kernels.
#include "rg_common.cuh"
#include "rg_geom.cuh"
#include "rg_boundary.cuh"

namespace rg {

// ---------------------------------------------------------------------------
// 2D: lowest-index containing cell (index_of_point_brute semantics,
// regridding/_weights/_weights_conservative_2d/_grids.py:223-279), found by a WALK: every thread locates a run of
// consecutive points, each by a Newton iteration in index space seeded with its predecessor's solution (1-2
// iterations on regular point sets) + the exact containment predicate / 3x3 lowest-index resolve (the role of
// index_of_point_secant, _grids.py:356-463).  The grid stays L2-resident; the points stream through once.
// A point the iteration cannot place in a cell is classified EXACTLY (see k_locate_walk); the rare leftovers go
// to the exhaustive pass 2.
// ---------------------------------------------------------------------------
constexpr int kLocRaster = 1024;   // occupancy raster over the grid's bounding box
constexpr int kLocRun = 8;         // points per lane: a warp walks a strip of 32 x kLocRun consecutive points

__device__ __forceinline__ int loc_raster_index(double x, double lo, double scale)
{
    const double t = floor((x - lo) * scale);   // monotone in x
    return (int)fmin(fmax(t, 0.0), (double)(kLocRaster - 1));
}

// raster scales: ONE evaluation shared by the kernel that marks and the kernel that looks up
__global__ void k_locate_raster_scales(const double* __restrict__ bbox, double* __restrict__ scales)
{
    if (threadIdx.x == 0) {
        scales[0] = kLocRaster / (bbox[2] - bbox[0]);
        scales[1] = kLocRaster / (bbox[3] - bbox[1]);
    }
}

// Every block of kLocBlock x kLocBlock cells marks the raster cells the bounding box of its vertices touches.  The box
// contains the bounding box of every cell of the block, so a point in an UNMARKED raster cell lies in no cell's
// bounding box and no cell contains it -- exact, whatever the mesh looks like (marking by blocks only makes the
// "maybe" set a little larger).
constexpr int kLocBlock = 4;
__global__ void k_locate_raster(GridView g, const double* __restrict__ bbox, const double* __restrict__ scales,
                                uint8_t* __restrict__ raster)
{
    const int nbx = (g.nx - 1 + kLocBlock - 1) / kLocBlock, nby = (g.ny - 1 + kLocBlock - 1) / kLocBlock;
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (int64_t)nbx * nby) return;
    const int bi = (int)(b / nby) * kLocBlock, bj = (int)(b % nby) * kLocBlock;
    const int ie = min(bi + kLocBlock, g.nx - 1), je = min(bj + kLocBlock, g.ny - 1);
    double xlo = INFINITY, ylo = INFINITY, xhi = -INFINITY, yhi = -INFINITY;
    for (int i = bi; i <= ie; i++)
        for (int j = bj; j <= je; j++) {
            const double x = g.x[(int64_t)i * g.ny + j], y = g.y[(int64_t)i * g.ny + j];
            xlo = fmin(xlo, x); xhi = fmax(xhi, x);
            ylo = fmin(ylo, y); yhi = fmax(yhi, y);
        }
    const int ix0 = loc_raster_index(xlo, bbox[0], scales[0]), ix1 = loc_raster_index(xhi, bbox[0], scales[0]);
    const int iy0 = loc_raster_index(ylo, bbox[1], scales[1]), iy1 = loc_raster_index(yhi, bbox[1], scales[1]);
    for (int ix = ix0; ix <= ix1; ix++)
        for (int iy = iy0; iy <= iy1; iy++) raster[ix * kLocRaster + iy] = 1;
}

// Newton in index space from the seed (i, j) on the bilinear map of the current cell, written for instruction count:
// a whole warp runs as long as its slowest lane, so the COMMON cases must be cheap.
//   * the iterate is accepted as soon as it stays in the cell whose corners are in registers, the last step was short
//     (below 0.5 cells with the solution 5e-2 cells inside, or below 1e-6 cells with the solution 1e-5 cells inside)
//     and the point is strictly inside the quad by the signs of the four edge cross products.  That far from every
//     edge the reference's containment predicate (point_is_inside_polygon, geometry.py:737-829) cannot disagree,
//     and in a mesh without overlapping cells no lower-index cell contains an interior point;
//   * everything else -- a point on or within 1e-5 cells of an edge, a concave cell, a diverging iteration, a point
//     outside -- goes through locate_newton: exact predicate, 3x3 lowest-index resolve.
// The reciprocal of the Jacobian determinant is an approximation refined once: the iteration corrects itself.
__device__ inline int locate_seeded(const GridView& g, double px, double py, double& i, double& j)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    int i0 = min(max(__double2int_rd(i), 0), ncx - 1), j0 = min(max(__double2int_rd(j), 0), ncy - 1);
#pragma unroll 1
    for (int it = 0; it < 8; it++) {
        const double* gx = g.x + (i0 * g.ny + j0);
        const double* gy = g.y + (i0 * g.ny + j0);
        const double x00 = gx[0], x01 = gx[1], x10 = gx[g.ny], x11 = gx[g.ny + 1];
        const double y00 = gy[0], y01 = gy[1], y10 = gy[g.ny], y11 = gy[g.ny + 1];
        const double u = i - i0, v = j - j0, u1 = 1.0 - u, v1 = 1.0 - v;
        const double xa = dfma(x10, u, x00 * u1), xb = dfma(x11, u, x01 * u1);
        const double ya = dfma(y10, u, y00 * u1), yb = dfma(y11, u, y01 * u1);
        const double ex = dfma(xb, v, xa * v1) - px, ey = dfma(yb, v, ya * v1) - py;
        const double dxdi = dfma(x11 - x01, v, (x10 - x00) * v1);
        const double dxdj = dfma(x11 - x10, u, (x01 - x00) * u1);
        const double dydi = dfma(y11 - y01, v, (y10 - y00) * v1);
        const double dydj = dfma(y11 - y10, u, (y01 - y00) * u1);
        const double det = dfma(dxdi, dydj, -(dxdj * dydi));
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(det));
        const double rdet = dfma(r0, dfma(-det, r0, 1.0), r0);
        const double di = dfma(dydj, ex, -(dxdj * ey)) * rdet;
        const double dj = dfma(dxdi, ey, -(dydi * ex)) * rdet;
        const double big = fmax(fabs(di), fabs(dj));
        if (!(big < 4.0 * (g.nx + g.ny))) break;   // diverging (or NaN): the careful path decides
        i -= di;
        j -= dj;
        const int in0 = __double2int_rd(i), jn0 = __double2int_rd(j);
        if (in0 == i0 && jn0 == j0) {
            const double fu = i - i0, fv = j - j0;
            const double margin = fmin(fmin(fu, 1.0 - fu), fmin(fv, 1.0 - fv));
            const bool settled = big < 1e-6;
            if ((big < 0.5 && margin > 5e-2) || (settled && margin > 1e-5)) {
                // strictly inside the quad (i0,j0),(i0+1,j0),(i0+1,j0+1),(i0,j0+1)?  four edge cross products, one sign
                const double c0 = dfma(x10 - x00, py - y00, -((y10 - y00) * (px - x00)));
                const double c1 = dfma(x11 - x10, py - y10, -((y11 - y10) * (px - x10)));
                const double c2 = dfma(x01 - x11, py - y11, -((y01 - y11) * (px - x11)));
                const double c3 = dfma(x00 - x01, py - y01, -((y00 - y01) * (px - x01)));
                if ((c0 > 0.0 && c1 > 0.0 && c2 > 0.0 && c3 > 0.0) || (c0 < 0.0 && c1 < 0.0 && c2 < 0.0 && c3 < 0.0))
                    return i0 * ncy + j0;
                break;   // concave or degenerate cell: exact path
            }
            if (settled) break;   // on (or within 1e-5 cells of) an edge: exact path
        } else {
            if (big < 1e-6) break;   // settled outside the grid (clamped cell) or exactly on a cell border
            i0 = min(max(in0, 0), ncx - 1);
            j0 = min(max(jn0, 0), ncy - 1);
        }
    }
    if (!(i == i) || !(j == j)) { i = 0.5 * g.nx; j = 0.5 * g.ny; }
    const int r = locate_newton(g, px, py, i, j);
    i = fmin(fmax(i, -1.0), (double)g.nx);
    j = fmin(fmax(j, -1.0), (double)g.ny);
    return r;
}

// Pass 1: a warp owns a strip of 32 x kLocRun consecutive points; lane k takes the points k, k + 32, ... of the strip
// (coalesced loads and stores; neighbouring lanes work in neighbouring cells, so the grid loads of a warp share
// cache lines).  Every point is seeded with the affine map through three corners of the grid PLUS the error that
// map made at the lane's previous point (32 points earlier: the correction varies slowly), which lands within a
// fraction of a cell on regular point sets.  A point the iteration cannot place in a cell is classified EXACTLY:
// outside the vertex bounding box, in an unmarked cell of the occupancy raster, or boundary winding number 0 (the
// reference's own line-start test, c2d.py:308-317) => `fill`; otherwise it is queued for the exhaustive pass 2.
__global__ void __launch_bounds__(128)
k_locate_walk(GridView g, Boundary bnd, const double* __restrict__ bbox, const double* __restrict__ scales,
              const uint8_t* __restrict__ raster, int64_t n, const double* __restrict__ px, const double* __restrict__ py,
              int64_t fill, int64_t* __restrict__ out, uint8_t* __restrict__ pending, int32_t* __restrict__ n_pending)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t p0 = warp * (32 * kLocRun) + lane;
    if (p0 >= n) return;
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    // affine map (x, y) -> (i, j) through the corners (0, 0), (ncx, 0), (0, ncy)
    double m00 = 0.0, m01 = 0.0, m10 = 0.0, m11 = 0.0;
    const double x00 = g.x[0], y00 = g.y[0];
    {
        const double ax = (g.x[(int64_t)ncx * g.ny] - x00) / ncx, ay = (g.y[(int64_t)ncx * g.ny] - y00) / ncx;
        const double bx = (g.x[ncy] - x00) / ncy, by = (g.y[ncy] - y00) / ncy;
        const double det0 = ax * by - bx * ay;
        if (det0 != 0.0 && det0 == det0) {
            m00 = by / det0; m01 = -bx / det0;
            m10 = -ay / det0; m11 = ax / det0;
        }
    }
    const double bx0 = bbox[0], by0 = bbox[1], bx1 = bbox[2], by1 = bbox[3], sx = scales[0], sy = scales[1];
    double ci = 0.0, cj = 0.0;   // what the affine map missed at the previous point
#pragma unroll 1
    for (int q = 0; q < kLocRun; q++) {
        const int64_t p = p0 + 32 * q;
        if (p >= n) break;
        const double x = px[p], y = py[p];
        const double ai = m00 * (x - x00) + m01 * (y - y00), aj = m10 * (x - x00) + m11 * (y - y00);
        double i = fmin(fmax(ai + ci, -1.0), (double)ncx + 1.0), j = fmin(fmax(aj + cj, -1.0), (double)ncy + 1.0);
        if (m00 == 0.0 && m01 == 0.0) { i = 0.5 * g.nx; j = 0.5 * g.ny; }
        // a point outside the vertex bounding box or in an unmarked cell of the occupancy raster lies in no cell
        const bool maybe = bx0 <= x && x <= bx1 && by0 <= y && y <= by1 &&
                           raster[loc_raster_index(x, bx0, sx) * kLocRaster + loc_raster_index(y, by0, sy)];
        uint8_t pend = 0;
        int64_t res = fill;
        if (maybe) {
            const int r = locate_seeded(g, x, y, i, j);
            ci = i - ai;
            cj = j - aj;
            if (r >= 0) {
                res = r;
            } else if (boundary_winding(bnd, x, y) != 0.0) {
                pend = 1;
                atomicAdd(n_pending, 1);
            }
        }
        out[p] = res;
        pending[p] = pend;
    }
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
    uint8_t* raster;
    double* scales;
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
    l.raster = c.take<uint8_t>((size_t)kLocRaster * kLocRaster);
    l.scales = c.take<double>(2);
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
    RG_CUDA(cudaMemsetAsync(l.raster, 0, (size_t)kLocRaster * kLocRaster, st));
    k_locate_raster_scales<<<1, 32, 0, st>>>(l.bbox, l.scales);
    k_locate_raster<<<(unsigned)ceil_div(ceil_div(nx - 1, kLocBlock) * ceil_div(ny - 1, kLocBlock), 256), 256, 0, st>>>(
        g, l.bbox, l.scales, l.raster);
    RG_LAUNCH_CHECK("k_locate_raster");
    k_locate_walk<<<(unsigned)ceil_div(ceil_div(n_points, 32 * kLocRun) * 32, 128), 128, 0, st>>>(
        g, l.bnd, l.bbox, l.scales, l.raster, n_points, px, py, fill, cell_flat, l.pending, l.counter);
    RG_LAUNCH_CHECK("k_locate_walk");
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
