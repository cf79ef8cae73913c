// rg_locate.cu -- cell location: the 2D walk (new 2D extension of find_indices built on
// the reference's internal locators) and the 1D brute / searchsorted kernels.
#include "rg_common.cuh"
#include "rg_geom.cuh"
#include "rg_boundary.cuh"

namespace rg {

// ---------------------------------------------------------------------------
// 2D: lowest-index containing cell (index_of_point_brute semantics,
// regridding/_weights/_weights_conservative_2d/_grids.py:223-279), found by a WALK.  Three passes, enqueued back to back
// with no host round trip (the later ones read their work counts on the device):
//   fast  (k_locate_fast)  one thread per point: seed cell predicted from the lane's previous points, cell walk by the
//                          signs of the four edge cross products; points in no cell's bounding box (occupancy raster)
//                          are answered at once; whatever the walk does not settle is MARKED (sentinel + queue);
//   slow  (k_locate_slow)  one warp per marked point: boundary winding number for points whose walk left the grid,
//                          Newton + the exact containment predicate / 3x3 lowest-index resolve for the others (the role
//                          of index_of_point_secant, _grids.py:356-463);
//   brute (k_locate_brute) exhaustive scan for points inside the boundary polygon that no cell was found for.
// The grid stays L2-resident; the points stream through once.
// ---------------------------------------------------------------------------
constexpr int kLocRaster = 1024;   // occupancy raster over the grid's bounding box
constexpr int kLocRun = 8;         // points per lane: a warp walks a strip of 32 x kLocRun consecutive points

__device__ __forceinline__ int loc_raster_index(double x, double lo, float scale)
{
    // any NON-DECREASING function of x does (a point inside a box then gets an index inside the box's index range), as
    // long as marking and lookup use the same one: the difference is rounded to fp32 (monotone), scaled and floored
    const int t = __float2int_rd(__fmul_rn((float)(x - lo), scale));
    return min(max(t, 0), kLocRaster - 1);
}

// Raster frame = bounding box of the grid's BOUNDARY vertices (the union of the boundary group boxes: no pass over
// the whole grid), and its scales: ONE evaluation shared by the kernel that marks and the kernel that looks up.
// Raster indices are clamped, and a clamped index is still a monotone function of the coordinate: vertices of a
// folded mesh that leave the frame mark its border cells, points beyond it look those up -- the raster stays exact.
__global__ void k_locate_raster_scales(Boundary bnd, double* __restrict__ bbox, float* __restrict__ scales)
{
    if (threadIdx.x == 0) {
        double xlo = INFINITY, ylo = INFINITY, xhi = -INFINITY, yhi = -INFINITY;
        for (int q = 0; q < bnd.n_g2; q++) {
            const BBox B = bnd.bb2[q];
            xlo = fmin(xlo, B.xlo); ylo = fmin(ylo, B.ylo);
            xhi = fmax(xhi, B.xhi); yhi = fmax(yhi, B.yhi);
        }
        bbox[0] = xlo; bbox[1] = ylo; bbox[2] = xhi; bbox[3] = yhi;
        scales[0] = (float)(kLocRaster / (xhi - xlo));
        scales[1] = (float)(kLocRaster / (yhi - ylo));
    }
}

// Every block of kLocBlock x kLocBlock cells marks the raster cells the bounding box of its vertices touches.  The box
// contains the bounding box of every cell of the block, so a point in an UNMARKED raster cell lies in no cell's
// bounding box and no cell contains it -- exact, whatever the mesh looks like (marking by blocks only makes the
// "maybe" set a little larger).
constexpr int kLocBlock = 4;
__global__ void k_locate_raster(GridView g, const double* __restrict__ bbox, const float* __restrict__ scales,
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

// Cell walk: from the seed cell, the signs of the four edge cross products either accept the cell or say which edge
// to cross (1-2 cell tests per point on regular point sets; 24 fp64 operations per test against ~60 for a Newton
// step with its acceptance test).  With A = c0 + c1 + c2 + c3 (twice the signed area of the quad, whatever the
// point) and d_k = c_k / A:
//   * all d_k > 1e-5: the point is strictly inside the intersection of the four inner half-planes (the kernel of the
//     quad, inside it also for a concave cell), at least ~1e-5 cells from every edge -- eight orders of magnitude
//     above the rounding error of the products.  The reference's containment predicate (point_is_inside_polygon,
//     geometry.py:737-829) cannot disagree there, and in a mesh without overlapping cells no lower-index cell
//     contains an interior point: the cell is the answer;
//   * some d_k < -1e-5: the point is beyond that edge: step into the neighbour across it;
//   * anything else (on or within ~1e-5 cells of an edge line, a degenerate cell, a step off the grid, more than
//     kWalkSteps steps) returns -1 (-2: the walk left the grid) and the point goes to the slow pass
//     (boundary winding number, locate_seeded -> locate_newton: exact).
// (fi, fj) come back as the estimate d3 / (d3 + d1), d0 / (d0 + d2) of the point's index coordinates (exact in a
// parallelogram): the seed correction for the lane's next point.
constexpr int kWalkSteps = 6;
// (i0, j0): seed cell in, last cell out.  One exit: the lanes of a warp that settle after different numbers of steps
// run the code after the loop together.
__device__ __forceinline__ int locate_cellwalk(const GridView& g, double px, double py, int& i0, int& j0, double& fi, double& fj)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    int status = 1;   // 1 walking, 0 found, -1 not settled, -2 left the grid
    double d0 = 0.0, d1 = 1.0, d2 = 1.0, d3 = 0.0;
#pragma unroll 1
    for (int it = 0; it < kWalkSteps && status == 1; it++) {
        const double* gx = g.x + (i0 * g.ny + j0);
        const double* gy = g.y + (i0 * g.ny + j0);
        const double x00 = gx[0], x01 = gx[1], x10 = gx[g.ny], x11 = gx[g.ny + 1];
        const double y00 = gy[0], y01 = gy[1], y10 = gy[g.ny], y11 = gy[g.ny + 1];
        // edges in the order (i0,j0)->(i0+1,j0) [side j = j0], ->(i0+1,j0+1) [side i = i0+1], ->(i0,j0+1) [side j = j0+1],
        // ->(i0,j0) [side i = i0]
        const double c0 = dfma(x10 - x00, py - y00, -((y10 - y00) * (px - x00)));
        const double c1 = dfma(x11 - x10, py - y10, -((y11 - y10) * (px - x10)));
        const double c2 = dfma(x01 - x11, py - y11, -((y01 - y11) * (px - x11)));
        const double c3 = dfma(x00 - x01, py - y01, -((y00 - y01) * (px - x01)));
        const double A = (c0 + c1) + (c2 + c3);
        const double tol = 1e-5 * fabs(A);
        const int flip = __double2hiint(A) & (int)0x80000000;   // d_k = sign(A) c_k
        d0 = __hiloint2double(__double2hiint(c0) ^ flip, __double2loint(c0));
        d1 = __hiloint2double(__double2hiint(c1) ^ flip, __double2loint(c1));
        d2 = __hiloint2double(__double2hiint(c2) ^ flip, __double2loint(c2));
        d3 = __hiloint2double(__double2hiint(c3) ^ flip, __double2loint(c3));
        if (!(tol > 0.0)) {   // degenerate cell or NaN
            status = -1;
        } else if (d0 > tol && d1 > tol && d2 > tol && d3 > tol) {
            status = 0;
        } else {
            const double mt = -tol;
            const int di = (d1 < mt) ? 1 : ((d3 < mt) ? -1 : 0);
            const int dj = (d2 < mt) ? 1 : ((d0 < mt) ? -1 : 0);
            const int i1 = i0 + di, j1 = j0 + dj;
            if (di == 0 && dj == 0) status = -1;       // within the tolerance of an edge line: slow pass
            else if (i1 < 0 || j1 < 0 || i1 >= ncx || j1 >= ncy) status = -2;   // off the grid: the slow pass classifies the point
            else { i0 = i1; j0 = j1; }
        }
    }
    if (status == 0) {
        double r0, r1;
        const double si = d3 + d1, sj = d0 + d2;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(si));
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r1) : "d"(sj));
        fi = (double)i0 + d3 * r0;
        fj = (double)j0 + d0 * r1;
        return i0 * ncy + j0;
    }
    return status == 1 ? -1 : status;
}

// affine map (x, y) -> (i, j) through the corners (0, 0), (ncx, 0), (0, ncy) of the grid: rows (m00 m01), (m10 m11)
struct LocAffine {
    double x00, y00, m00, m01, m10, m11;
};
__device__ __forceinline__ LocAffine loc_affine(const GridView& g)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    LocAffine a = { g.x[0], g.y[0], 0.0, 0.0, 0.0, 0.0 };
    const double ax = (g.x[(int64_t)ncx * g.ny] - a.x00) / ncx, ay = (g.y[(int64_t)ncx * g.ny] - a.y00) / ncx;
    const double bx = (g.x[ncy] - a.x00) / ncy, by = (g.y[ncy] - a.y00) / ncy;
    const double det0 = ax * by - bx * ay;
    if (det0 != 0.0 && det0 == det0) {
        a.m00 = by / det0; a.m01 = -bx / det0;
        a.m10 = -ay / det0; a.m11 = ax / det0;
    }
    return a;
}

// Seed table: the affine map misses the true index coordinates by the grid's distortion (tens of cells); the miss
// varies slowly, so it is tabulated at a kSeedTab^2 lattice of grid vertices.  A lane's FIRST point has no
// predecessor to take a correction from: it looks the miss up at its affine index and once more at the corrected
// one, which leaves an error of a cell or two for the walk.
constexpr int kSeedTab = 256;
__global__ void k_locate_seed_table(GridView g, double2* __restrict__ tab, LocAffine* __restrict__ affine)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= kSeedTab * kSeedTab) return;
    if (q == 0) *affine = loc_affine(g);
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    const int a = q / kSeedTab, b = q % kSeedTab;
    const int i = (int)(((int64_t)a * ncx + (kSeedTab - 1) / 2) / (kSeedTab - 1)), j = (int)(((int64_t)b * ncy + (kSeedTab - 1) / 2) / (kSeedTab - 1));
    const LocAffine M = loc_affine(g);
    const double x = g.x[(int64_t)i * g.ny + j], y = g.y[(int64_t)i * g.ny + j];
    const double ai = M.m00 * (x - M.x00) + M.m01 * (y - M.y00), aj = M.m10 * (x - M.x00) + M.m11 * (y - M.y00);
    double2 c = make_double2((double)i - ai, (double)j - aj);
    if (!(c.x == c.x) || !(c.y == c.y)) c = make_double2(0.0, 0.0);
    tab[q] = c;
}
__device__ __forceinline__ double2 loc_seed_lookup(const double2* __restrict__ tab, double i, double j, double si, double sj)
{
    const int a = min(max(__double2int_rn(i * si), 0), kSeedTab - 1), b = min(max(__double2int_rn(j * sj), 0), kSeedTab - 1);
    return tab[a * kSeedTab + b];
}

// Boundary winding number for the slow pass.  boundary_winding walks two levels of boxes for every point (~150
// dependent box loads before the first edge); the slow pass asks for thousands of points next to the boundary, so the
// edge groups are also indexed by RASTER ROW: bit g1 of row r is set iff the y-range of group g1 meets the row's
// y-interval (same monotone index function as the occupancy raster).  A point only visits the groups of its row --
// a superset of the groups whose y-range contains it; every other edge contributes exactly 0.
constexpr int kRowMaskMaxWords = 64;   // up to 2048 edge groups (65536 boundary edges); larger grids walk the boxes
__global__ void k_locate_rowmask(Boundary bnd, const double* __restrict__ bbox, const float* __restrict__ scales,
                                 uint32_t* __restrict__ rowmask, int words)
{
    const int g1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (g1 >= bnd.n_g1) return;
    const BBox B = bnd.bb1[g1];
    const int r0 = loc_raster_index(B.ylo, bbox[1], scales[1]), r1 = loc_raster_index(B.yhi, bbox[1], scales[1]);
    for (int r = r0; r <= r1; r++) atomicOr(&rowmask[r * words + (g1 >> 5)], 1u << (g1 & 31));
}
// One WARP per point: lane = edge of a group (coalesced edge loads); the contributions are multiples of 1/2, so the
// lanes' partial sums add up exactly in any order.  Every lane returns the sum.
__device__ inline double loc_boundary_winding_warp(const Boundary& b, const double* __restrict__ bbox,
                                                   const float* __restrict__ scales, const uint32_t* __restrict__ rowmask,
                                                   int words, double px, double py)
{
    const int lane = threadIdx.x & 31;
    double w = 0.0;
    auto edge = [&](int s) {
        if (s < b.n_edges) {
            const double x0 = dsub(b.x3[s], px), y0 = dsub(b.y3[s], py);
            const double x1 = dsub(b.x4[s], px), y1 = dsub(b.y4[s], py);
            const bool reversed = (s < b.ne_a0) || (s >= 2 * b.ne_a0 + b.ne_a1);   // see boundary_winding_group
            w += reversed ? winding_edge(x1, y1, x0, y0) : winding_edge(x0, y0, x1, y1);
        }
    };
    if (words > 0 && py == py) {
        const uint32_t* row = rowmask + loc_raster_index(py, bbox[1], scales[1]) * words;
        for (int k0 = 0; k0 < words; k0 += 32) {
            const uint32_t mine = (k0 + lane < words) ? row[k0 + lane] : 0u;
            unsigned nonzero = __ballot_sync(0xffffffffu, mine != 0u);
            while (nonzero) {
                const int k = __ffs(nonzero) - 1;
                nonzero &= nonzero - 1;
                uint32_t m = __shfl_sync(0xffffffffu, mine, k);
                while (m) {
                    const int g1 = (k0 + k) * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    edge(g1 * 32 + lane);
                }
            }
        }
    } else if (py == py) {   // no row index (very long boundaries): every group whose box contains py
        for (int g1 = 0; g1 < b.n_g1; g1++) {
            const BBox B1 = b.bb1[g1];
            if (B1.ylo <= py && py <= B1.yhi) edge(g1 * 32 + lane);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    return w;
}

// Points the fast pass does not settle are marked in the OUTPUT array by sentinels (three values below every cell
// index and different from `fill`) and, when they fit, listed in a queue: no per-point flag array, no host round trip.
struct LocSentinels {
    int64_t off;     // the cell walk left the grid               -> k_locate_slow (boundary winding number first)
    int64_t unres;   // the cell walk did not settle the point    -> k_locate_slow
    int64_t brute;   // inside the boundary polygon, not placed   -> k_locate_brute
};
static LocSentinels loc_sentinels(int64_t fill)
{
    int64_t v[3];
    int64_t c = INT64_MIN;
    for (int q = 0; q < 3; q++, c++) {
        if (c == fill) c++;
        v[q] = c;
    }
    return { v[0], v[1], v[2] };
}
constexpr uint32_t kQueueOff = 0x80000000u;   // queue entry: point index | kQueueOff when the walk left the grid

struct LocTables {
    const float* scales;      // [2] raster scales
    const double* bbox;       // [4] raster frame
    const uint8_t* raster;    // [kLocRaster^2]
    const double2* seed_tab;  // [kSeedTab^2]
    const LocAffine* affine;
    const uint32_t* rowmask;  // [kLocRaster][rowmask_words] boundary edge groups by raster row (0 words: not built)
    int rowmask_words;
};

__device__ __forceinline__ void loc_seed_cell(const GridView& g, const LocTables& T, double x, double y, double& i, double& j)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    const LocAffine M = *T.affine;
    const double tsi = (double)(kSeedTab - 1) / ncx, tsj = (double)(kSeedTab - 1) / ncy;
    const double ai = M.m00 * (x - M.x00) + M.m01 * (y - M.y00), aj = M.m10 * (x - M.x00) + M.m11 * (y - M.y00);
    const double2 c1 = loc_seed_lookup(T.seed_tab, ai, aj, tsi, tsj);
    const double2 c2 = loc_seed_lookup(T.seed_tab, ai + c1.x, aj + c1.y, tsi, tsj);
    i = fmin(fmax(ai + c2.x, -1.0), (double)ncx + 1.0);
    j = fmin(fmax(aj + c2.y, -1.0), (double)ncy + 1.0);
    if (M.m00 == 0.0 && M.m01 == 0.0) { i = 0.5 * g.nx; j = 0.5 * g.ny; }
}

// Pass 1 (FAST): a warp owns a strip of 32 x kLocRun consecutive points; lane k takes the points k, k + 32, ... of
// the strip (coalesced loads and stores; neighbouring lanes work in neighbouring cells, so the grid loads of a warp
// share cache lines).  Every point is seeded with the affine map through three corners of the grid PLUS a prediction
// of what that map misses there: the miss at the lane's previous located point, extrapolated linearly through the
// last two (the miss varies slowly: the prediction lands within a tenth of a cell or so on regular point sets, so
// the first cell tested is usually the right one), and located by the cell walk.  A non-finite point or a point in
// an unmarked cell of the occupancy raster lies in no cell => `fill`.  Whatever the walk does not settle is only
// MARKED (sentinel in the output + queue): the Newton iteration, the exact predicate and the boundary winding number
// live in k_locate_slow, so that this kernel needs few registers and runs at full occupancy.
#ifndef RG_LOC_MINB
#define RG_LOC_MINB 9
#endif
__global__ void __launch_bounds__(128, RG_LOC_MINB)
k_locate_fast(GridView g, const LocTables T, int64_t n, const double* __restrict__ px, const double* __restrict__ py,
              int64_t fill, LocSentinels sent, int64_t* __restrict__ out, int32_t* __restrict__ counter,
              uint32_t* __restrict__ queue, int64_t queue_cap)
{
    // per-launch constants live in shared memory and are re-read every round (volatile): registers are what bounds
    // the occupancy of this kernel
    __shared__ double s_c[8];   // affine map x00 y00 m00 m01 m10 m11, raster origin bx0 by0
    __shared__ float s_f[2];    // raster scales
    if (threadIdx.x < 6) s_c[threadIdx.x] = ((const double*)T.affine)[threadIdx.x];
    else if (threadIdx.x < 8) s_c[threadIdx.x] = T.bbox[threadIdx.x - 6];
    else if (threadIdx.x < 10) s_f[threadIdx.x - 8] = T.scales[threadIdx.x - 8];
    __syncthreads();
    const volatile double* vc = s_c;
    const volatile float* vf = s_f;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t p0 = warp * (32 * kLocRun) + lane;
    if (p0 >= n) return;
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    int have = 0;                   // located predecessors the prediction can use (0, 1, 2+)
    double ci = 0.0, cj = 0.0;      // what the affine map missed at the last located point
    double si = 0.0, sj = 0.0;      // ... and its change per point of the lane
    double gap = 1.0;               // points of the lane since then
#pragma unroll 1
    for (int q = 0; q < kLocRun; q++) {
        const int64_t p = p0 + 32 * q;
        if (p >= n) break;
        const double x = px[p], y = py[p];
        const bool maybe = fabs(x) < INFINITY && fabs(y) < INFINITY &&   // (false for NaN)
                           T.raster[loc_raster_index(x, vc[6], vf[0]) * kLocRaster + loc_raster_index(y, vc[7], vf[1])];
        int64_t res = fill;
        if (maybe) {
            const double rx = x - vc[0], ry = y - vc[1];
            const double ai = vc[2] * rx + vc[3] * ry, aj = vc[4] * rx + vc[5] * ry;
            if (have == 0) {
                const double tsi = (double)(kSeedTab - 1) / ncx, tsj = (double)(kSeedTab - 1) / ncy;
                const double2 c1 = loc_seed_lookup(T.seed_tab, ai, aj, tsi, tsj);
                const double2 c2 = loc_seed_lookup(T.seed_tab, ai + c1.x, aj + c1.y, tsi, tsj);
                ci = c2.x;
                cj = c2.y;
            }
            // (the conversion saturates, NaN -> 0: no clamp in fp64; a useless seed only sends the point to the slow pass)
            int i0 = min(max(__double2int_rd(ai + dfma(gap, si, ci)), 0), ncx - 1);
            int j0 = min(max(__double2int_rd(aj + dfma(gap, sj, cj)), 0), ncy - 1);
            double i, j;
            const int r = locate_cellwalk(g, x, y, i0, j0, i, j);
            if (r >= 0) {
                res = r;
                const double ni = i - ai, nj = j - aj;
                if (have > 0) {
                    double rg_;
                    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rg_) : "d"(gap));
                    si = (ni - ci) * rg_;
                    sj = (nj - cj) * rg_;
                }
                ci = ni;
                cj = nj;
                have = have < 2 ? have + 1 : 2;
                gap = 0.0;
            } else {
                res = r == -2 ? sent.off : sent.unres;
                const int at = atomicAdd(&counter[1], 1);
                if (at < queue_cap) queue[at] = (uint32_t)p | (r == -2 ? kQueueOff : 0u);
            }
        }
        gap += 1.0;
        out[p] = res;
    }
}

// Pass 1b (SLOW): the points the cell walk marked (under 1 % at config 5: the band around the grid's boundary and
// points within ~1e-5 cells of an edge), taken from the queue the fast pass filled -- or, if they did not fit, found
// by scanning the output for the sentinels.  One WARP per point.  A point whose walk left the grid is first
// classified EXACTLY by the boundary winding number (the reference's own line-start test, c2d.py:308-317), its lanes
// sharing the boundary edges: 0 => outside the boundary polygon => `fill` (nearly all of them).  The others: Newton
// from the table seed + exact predicate / 3x3 lowest-index resolve (locate_seeded -> locate_newton; every lane
// computes the same); a point inside the boundary polygon that no cell is found for is marked for the exhaustive
// pass 2.  Grid-stride: the number of marked points is only known on the device.
// (the rare part of the slow pass, kept out of line so that the winding loop gets by with few registers)
__device__ __noinline__ int loc_slow_newton(const GridView& g, const LocTables& T, double x, double y)
{
    double i, j;
    loc_seed_cell(g, T, x, y, i, j);
    return locate_seeded(g, x, y, i, j);
}

#ifndef RG_LOCSLOW_MINB
#define RG_LOCSLOW_MINB 12
#endif
__global__ void __launch_bounds__(128, RG_LOCSLOW_MINB)
k_locate_slow(GridView g, Boundary bnd, const LocTables T, int64_t n,
              const double* __restrict__ px, const double* __restrict__ py,
              int64_t fill, LocSentinels sent, int64_t* __restrict__ out, int32_t* __restrict__ counter,
              const uint32_t* __restrict__ queue, int64_t queue_cap)
{
    const int64_t n_marked = counter[1];
    if (n_marked == 0) return;
    const bool queued = n_marked > 0 && n_marked <= queue_cap;   // (a wrapped 32-bit count is negative: scan)
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    auto one = [&](int64_t p, bool off) {   // all lanes of the warp, same arguments
        const double x = px[p], y = py[p];
        int64_t res = fill;
        bool inside_boundary = false, have_w = false;
        if (off) {
            have_w = true;
            inside_boundary = loc_boundary_winding_warp(bnd, T.bbox, T.scales, T.rowmask, T.rowmask_words, x, y) != 0.0;
        }
        if (!have_w || inside_boundary) {
            const int r = loc_slow_newton(g, T, x, y);
            if (r >= 0) {
                res = r;
            } else {
                if (!have_w) inside_boundary = loc_boundary_winding_warp(bnd, T.bbox, T.scales, T.rowmask, T.rowmask_words, x, y) != 0.0;
                if (inside_boundary) {
                    res = sent.brute;
                    if (lane == 0) atomicAdd(&counter[0], 1);
                }
            }
        }
        __syncwarp();
        if (lane == 0) out[p] = res;
    };
    if (queued) {
        for (int64_t t = warp; t < n_marked; t += nwarps) {
            const uint32_t e = queue[t];
            one((int64_t)(e & ~kQueueOff), (e & kQueueOff) != 0u);
        }
    } else {
        // the marks did not fit the queue: every warp scans a share of the output, 32 points at a time
        for (int64_t base = warp * 32; base < n; base += nwarps * 32) {
            const int64_t v = base + lane < n ? out[base + lane] : fill;
            unsigned todo = __ballot_sync(0xffffffffu, v == sent.off || v == sent.unres);
            const unsigned offs = __ballot_sync(0xffffffffu, v == sent.off);
            while (todo) {
                const int k = __ffs(todo) - 1;
                todo &= todo - 1;
                one(base + k, (offs >> k) & 1u);
            }
        }
    }
}

// Pass 2: exhaustive and exact (index_of_point_brute).  One warp per marked point scans
// all cells in row-major order and keeps the first containing one.
__global__ void k_locate_brute(GridView g, int64_t n, const double* __restrict__ px, const double* __restrict__ py,
                               int64_t fill, LocSentinels sent, int64_t* __restrict__ out, const int32_t* __restrict__ counter)
{
    if (counter[0] == 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ncy = g.ny - 1;
    const int64_t nc = (int64_t)(g.nx - 1) * ncy;
    for (int64_t p = warp; p < n; p += nwarps) {
        if (out[p] != sent.brute) continue;
        const double x = px[p], y = py[p];
        int64_t best = -1;
        for (int64_t base = 0; base < nc && best < 0; base += 32) {
            const int64_t c = base + lane;
            bool hit = false;
            if (c < nc) hit = cell_contains(g, (int)(c / ncy), (int)(c % ncy), x, y);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) best = base + (__ffs(m) - 1);
        }
        __syncwarp();
        if (lane == 0) out[p] = best >= 0 ? best : fill;
    }
}

struct LocateLayout {
    Boundary bnd;
    double* bbox;
    int32_t* counter;    // [0] points marked for the exhaustive pass, [1] points the cell walk did not settle
    uint8_t* raster;
    float* scales;
    double2* seed_tab;   // [kSeedTab^2] what the affine map misses at a coarse lattice of grid vertices
    LocAffine* affine;
    uint32_t* rowmask;   // [kLocRaster][rowmask_words]
    int rowmask_words;
    uint32_t* queue;     // [queue_cap] points the cell walk did not settle (when they fit and n_points < 2^31)
    int64_t queue_cap;
    size_t bytes;
};

static LocateLayout locate_layout(void* ws, int64_t nx, int64_t ny, int64_t n_points)
{
    LocateLayout l;
    Carver c(ws);
    carve_boundary(c, l.bnd, nx, ny);
    l.bbox = c.take<double>(4);
    l.counter = c.take<int32_t>(4);
    l.raster = c.take<uint8_t>((size_t)kLocRaster * kLocRaster);
    l.scales = c.take<float>(2);
    l.seed_tab = c.take<double2>((size_t)kSeedTab * kSeedTab);
    l.affine = c.take<LocAffine>(1);
    l.rowmask_words = (int)ceil_div((int64_t)l.bnd.n_g1, 32);
    if (l.rowmask_words > kRowMaskMaxWords || getenv("RG_LOC_NO_ROWMASK")) l.rowmask_words = 0;   // (env: test hook)
    l.rowmask = c.take<uint32_t>((size_t)kLocRaster * (l.rowmask_words > 0 ? l.rowmask_words : 1));
    l.queue_cap = n_points < ((int64_t)1 << 31) ? n_points / 16 + 1024 : 0;
    if (const char* e = getenv("RG_LOC_QUEUE_CAP")) l.queue_cap = atoll(e) < l.queue_cap ? atoll(e) : l.queue_cap;   // test hook: forces the scan
    l.queue = c.take<uint32_t>((size_t)(l.queue_cap > 0 ? l.queue_cap : 1));
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
    int rc = build_boundaries(st, 1, &g, &l.bnd, &l.bbox, 0);   // (the raster frame comes from the boundary boxes)
    if (rc) return rc;
    RG_CUDA(cudaMemsetAsync(l.counter, 0, sizeof(int32_t) * 4, st));
    RG_CUDA(cudaMemsetAsync(l.raster, 0, (size_t)kLocRaster * kLocRaster, st));
    k_locate_raster_scales<<<1, 32, 0, st>>>(l.bnd, l.bbox, l.scales);
    k_locate_raster<<<(unsigned)ceil_div(ceil_div(nx - 1, kLocBlock) * ceil_div(ny - 1, kLocBlock), 256), 256, 0, st>>>(
        g, l.bbox, l.scales, l.raster);
    RG_LAUNCH_CHECK("k_locate_raster");
    k_locate_seed_table<<<kSeedTab * kSeedTab / 256, 256, 0, st>>>(g, l.seed_tab, l.affine);
    if (l.rowmask_words > 0) {
        RG_CUDA(cudaMemsetAsync(l.rowmask, 0, sizeof(uint32_t) * (size_t)kLocRaster * l.rowmask_words, st));
        k_locate_rowmask<<<(unsigned)ceil_div((int64_t)l.bnd.n_g1, 128), 128, 0, st>>>(l.bnd, l.bbox, l.scales, l.rowmask, l.rowmask_words);
    }
    const LocTables T{ l.scales, l.bbox, l.raster, l.seed_tab, l.affine, l.rowmask, l.rowmask_words };
    const LocSentinels sent = loc_sentinels(fill);
    // no host round trip: the slow and the exhaustive pass read their work counts on the device (and leave at once
    // when there is none)
    k_locate_fast<<<(unsigned)ceil_div(ceil_div(n_points, (int64_t)32 * kLocRun) * 32, 128), 128, 0, st>>>(
        g, T, n_points, px, py, fill, sent, cell_flat, l.counter, l.queue, l.queue_cap);
    RG_LAUNCH_CHECK("k_locate_fast");
    {
        const int64_t want = ceil_div(l.queue_cap > 0 ? l.queue_cap : n_points, 128);
        k_locate_slow<<<(unsigned)(want < kNumSM * 16 ? want : kNumSM * 16), 128, 0, st>>>(
            g, l.bnd, T, n_points, px, py, fill, sent, cell_flat, l.counter, l.queue, l.queue_cap);
        RG_LAUNCH_CHECK("k_locate_slow");
    }
    {
        const int64_t warps = n_points < kNumSM * 64 ? n_points : kNumSM * 64;
        k_locate_brute<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, st>>>(g, n_points, px, py, fill, sent, cell_flat, l.counter);
        RG_LAUNCH_CHECK("k_locate_brute");
    }
    if (getenv("RG_LOC_DEBUG")) {
        int32_t counters[2] = { 0, 0 };
        RG_CUDA(cudaMemcpyAsync(counters, l.counter, sizeof(counters), cudaMemcpyDeviceToHost, st));
        RG_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "rg_find_indices_2d: %lld points, %d to the slow pass, %d to the exhaustive pass\n",
                (long long)n_points, counters[1], counters[0]);
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
