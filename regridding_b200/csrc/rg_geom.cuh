// rg_geom.cuh -- fp64 device primitives of the conservative sweep.
//
// Each primitive reproduces, operation for operation, what the reference's
// Numba/LLVM build evaluates (fastmath=True kernels on an x86-64 FMA host; the
// contraction pattern was measured against the JIT and is pinned by the oracle and
// the golden vectors, see oracle/oracle_regrid.c and SURVEY.md Appendix B).  All
// fused operations are explicit; the translation unit is compiled with -fmad=false.
#pragma once
#include "rg_common.cuh"

namespace rg {

struct GridView {
    const double* x;
    const double* y;
    int nx, ny;  // vertex counts, row-major (nx, ny)
};

// ---------------------------------------------------------------------------
// two_line_segment_intersection_parameters + two_line_segments_intersect
// regridding/geometry.py:370-445, 448-475 (bbox pre-check :153-284, :422-433).
// Returns true on a hit (0 <= t < 1 and 0 <= u < 1, half-open); t is valid then.
// JIT form: tdet/det/-udet with one rounded product and one fused, t = tdet * (1/det).
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool seg_hit(double x1, double y1, double x2, double y2,
                                        double x3, double y3, double x4, double y4, double& t)
{
    if (!(fmin(x1, x2) <= fmax(x3, x4) && fmin(x3, x4) <= fmax(x1, x2))) return false;
    if (!(fmin(y1, y2) <= fmax(y3, y4) && fmin(y3, y4) <= fmax(y1, y2))) return false;
    const double x13 = dsub(x1, x3), y13 = dsub(y1, y3);
    const double x12 = dsub(x1, x2), y12 = dsub(y1, y2);
    const double x34 = dsub(x3, x4), y34 = dsub(y3, y4);
    const double tdet = dfma(x13, y34, -dmul(y13, x34));
    const double det = dfma(x12, y34, -dmul(y12, x34));
    const double nudet = dfma(y12, x13, -dmul(y13, x12));
    const double rinv = ddiv(1.0, det);
    t = dmul(tdet, rinv);
    const double u = dmul(nudet, rinv);
    return (0.0 <= t) && (t < 1.0) && (0.0 <= u) && (u < 1.0);
}

// two_line_segment_intersection, regridding/geometry.py:478-556 (JIT: fma(t, x2-x1, x1)).
__device__ __forceinline__ void seg_point(double x1, double y1, double x2, double y2, double t,
                                          double& x, double& y)
{
    x = dfma(t, dsub(x2, x1), x1);
    y = dfma(t, dsub(y2, y1), y1);
}

// One edge (x0,y0)->(x1,y1), already relative to the query point, of the extended
// winding number of point_is_inside_polygon, regridding/geometry.py:737-829.
__device__ __forceinline__ double winding_edge(double x0, double y0, double x1, double y1)
{
    if (dmul(y0, y1) < 0.0) {
        const double r = dadd(x0, ddiv(dmul(y0, dsub(x1, x0)), dsub(y0, y1)));
        if (r > 0.0) return (y0 < 0.0) ? 1.0 : -1.0;
        return (y0 < 0.0) ? -1.0 : 1.0;
    } else if (y0 == 0.0) {
        if (x0 > 0.0) return (y1 > 0.0) ? 0.5 : ((y1 < 0.0) ? -0.5 : 0.0);
        if (x0 < 0.0) return (y1 < 0.0) ? 0.5 : ((y1 > 0.0) ? -0.5 : 0.0);
        return 0.0;
    } else if (y1 == 0.0) {
        if (x1 > 0.0) return (y0 < 0.0) ? 0.5 : ((y0 > 0.0) ? -0.5 : 0.0);
        if (x1 < 0.0) return (y0 > 0.0) ? 0.5 : ((y0 < 0.0) ? -0.5 : 0.0);
        return 0.0;
    }
    return 0.0;
}

// Containment of (px,py) in the quad (i,j),(i+1,j),(i+1,j+1),(i,j+1):
// regridding/_weights/_weights_conservative_2d/_grids.py:256-277, 331-347.
__device__ __forceinline__ bool cell_contains(const GridView& g, int i, int j, double px, double py)
{
    const int64_t a = (int64_t)i * g.ny + j;
    const double ax = dsub(g.x[a], px), ay = dsub(g.y[a], py);                            // (i, j)
    const double bx = dsub(g.x[a + g.ny], px), by = dsub(g.y[a + g.ny], py);              // (i+1, j)
    const double cx = dsub(g.x[a + g.ny + 1], px), cy = dsub(g.y[a + g.ny + 1], py);      // (i+1, j+1)
    const double dx = dsub(g.x[a + 1], px), dy = dsub(g.y[a + 1], py);                    // (i, j+1)
    // edge v runs from vertex v-1 (wrapping) to vertex v
    double w = winding_edge(dx, dy, ax, ay);
    w += winding_edge(ax, ay, bx, by);
    w += winding_edge(bx, by, cx, cy);
    w += winding_edge(cx, cy, dx, dy);
    return w != 0.0;
}

// the same predicate on corner coordinates the caller already holds (vertex order (i,j), (i+1,j), (i+1,j+1), (i,j+1))
__device__ __forceinline__ bool quad_contains(double x00, double y00, double x10, double y10, double x11, double y11,
                                              double x01, double y01, double px, double py)
{
    const double ax = dsub(x00, px), ay = dsub(y00, py);
    const double bx = dsub(x10, px), by = dsub(y10, py);
    const double cx = dsub(x11, px), cy = dsub(y11, py);
    const double dx = dsub(x01, px), dy = dsub(y01, py);
    double w = winding_edge(dx, dy, ax, ay);
    w += winding_edge(ax, ay, bx, by);
    w += winding_edge(bx, by, cx, cy);
    w += winding_edge(cx, cy, dx, dy);
    return w != 0.0;
}

// Lowest-index containing cell among the 3x3 neighbourhood of (i0, j0):
// _index_of_point_local, _grids.py:286-349.  Returns the flat cell index or -1.
__device__ inline int locate_local(const GridView& g, double px, double py, int i0, int j0)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    const int ilo = max(i0 - 1, 0), ihi = min(i0 + 2, ncx);
    const int jlo = max(j0 - 1, 0), jhi = min(j0 + 2, ncy);
    for (int i = ilo; i < ihi; i++)
        for (int j = jlo; j < jhi; j++)
            if (cell_contains(g, i, j, px, py)) return i * ncy + j;
    return -1;
}

constexpr int kLocOutside = -1;  // the point is outside the grid
constexpr int kLocUnknown = -2;  // the iteration gave no verdict

// Newton iteration in index space on the piecewise-bilinear map (the role of
// index_of_point_secant, _grids.py:356-463; analytic Jacobian instead of forward
// differences -- any iteration that lands in a containing cell gives the same
// answer, because the result is always the lowest-index containing cell of the
// 3x3 neighbourhood).  Returns flat cell >= 0, kLocOutside or kLocUnknown.
__device__ inline int locate_newton(const GridView& g, double px, double py, double i, double j)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    bool converged = false;
    for (int it = 0; it < 48; it++) {
        int i0 = (int)floor(i), j0 = (int)floor(j);
        i0 = min(max(i0, 0), ncx - 1);
        j0 = min(max(j0, 0), ncy - 1);
        const int64_t a = (int64_t)i0 * g.ny + j0;
        const double x00 = g.x[a], x01 = g.x[a + 1], x10 = g.x[a + g.ny], x11 = g.x[a + g.ny + 1];
        const double y00 = g.y[a], y01 = g.y[a + 1], y10 = g.y[a + g.ny], y11 = g.y[a + g.ny + 1];
        const double u = i - i0, v = j - j0;
        const double X = (x00 * (1 - u) + x10 * u) * (1 - v) + (x01 * (1 - u) + x11 * u) * v;
        const double Y = (y00 * (1 - u) + y10 * u) * (1 - v) + (y01 * (1 - u) + y11 * u) * v;
        const double ex = X - px, ey = Y - py;
        const double dxdi = (x10 - x00) * (1 - v) + (x11 - x01) * v;
        const double dxdj = (x01 - x00) * (1 - u) + (x11 - x10) * u;
        const double dydi = (y10 - y00) * (1 - v) + (y11 - y01) * v;
        const double dydj = (y01 - y00) * (1 - u) + (y11 - y10) * u;
        const double det = dxdi * dydj - dxdj * dydi;
        if (det == 0.0 || !(det == det)) return kLocUnknown;
        double di = (dydj * ex - dxdj * ey) / det;
        double dj = (-dydi * ex + dxdi * ey) / det;
        // keep a diverging step bounded by the grid size
        const double lim_i = (double)g.nx, lim_j = (double)g.ny;
        di = fmin(fmax(di, -lim_i), lim_i);
        dj = fmin(fmax(dj, -lim_j), lim_j);
        i -= di;
        j -= dj;
        if (fabs(di) < 1e-9 && fabs(dj) < 1e-9) {
            converged = true;
            break;
        }
    }
    if (!converged) return kLocUnknown;
    int ic = (int)floor(fmin(fmax(i, -2.0), (double)ncx + 2.0));
    int jc = (int)floor(fmin(fmax(j, -2.0), (double)ncy + 2.0));
    // Fast path: the solution sits well inside cell (ic, jc) in index space (the Newton step was < 1e-9) and the
    // exact containment test agrees.  In a mesh without overlapping cells no other cell contains the point then,
    // so the lowest-index scan of the 3x3 neighbourhood below would return the same cell.
    if (ic >= 0 && jc >= 0 && ic < ncx && jc < ncy) {
        const double fu = i - ic, fv = j - jc;
        constexpr double kMargin = 1e-6;
        if (fu > kMargin && fu < 1.0 - kMargin && fv > kMargin && fv < 1.0 - kMargin && cell_contains(g, ic, jc, px, py))
            return ic * ncy + jc;
    }
    int found = locate_local(g, px, py, min(max(ic, 0), ncx - 1), min(max(jc, 0), ncy - 1));
    if (found >= 0) return found;
    if (i < 0.0 || j < 0.0 || i > (double)ncx || j > (double)ncy) return kLocOutside;
    return kLocUnknown;
}

// index-space estimate of (px, py) by the affine map through the corners (0, 0), (ncx, 0), (0, ncy) of the grid
// (clamped to one cell beyond the grid; the grid centre for a degenerate map)
__device__ __forceinline__ void affine_seed(const GridView& g, double px, double py, double& i, double& j)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    const double x00 = g.x[0], y00 = g.y[0];
    const double ax = (g.x[(int64_t)ncx * g.ny] - x00) / ncx, ay = (g.y[(int64_t)ncx * g.ny] - y00) / ncx;
    const double bx = (g.x[ncy] - x00) / ncy, by = (g.y[ncy] - y00) / ncy;
    const double det0 = ax * by - bx * ay;
    i = 0.5 * g.nx;
    j = 0.5 * g.ny;
    if (det0 != 0.0 && det0 == det0) {
        const double r0 = 1.0 / det0;
        const double ti = ((px - x00) * by - (py - y00) * bx) * r0, tj = ((py - y00) * ax - (px - x00) * ay) * r0;
        if (ti == ti && tj == tj) {
            i = fmin(fmax(ti, -1.0), (double)ncx + 1.0);
            j = fmin(fmax(tj, -1.0), (double)ncy + 1.0);
        }
    }
}

// Cheap variant for the walk-state GUESSES of the 2D build (k_vertex_guess*): wrong guesses only cost a repair,
// so the iteration starts from the affine map through three corners of the grid, stops as soon as the Newton step
// is below 1e-4 cells and accepts the cell when the point is 1e-3 cells away from its edges (or "outside" when it
// is 1e-3 cells beyond the border); anything else continues with the full-precision locate_newton from where it
// stands.
__device__ inline int locate_guess(const GridView& g, double px, double py)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    const double x00 = g.x[0], y00 = g.y[0];
    const double ax = (g.x[(int64_t)ncx * g.ny] - x00) / ncx, ay = (g.y[(int64_t)ncx * g.ny] - y00) / ncx;
    const double bx = (g.x[ncy] - x00) / ncy, by = (g.y[ncy] - y00) / ncy;
    const double det0 = ax * by - bx * ay;
    double i = 0.5 * g.nx, j = 0.5 * g.ny;
    if (det0 != 0.0 && det0 == det0) {
        const double r0 = 1.0 / det0;
        i = ((px - x00) * by - (py - y00) * bx) * r0;
        j = ((py - y00) * ax - (px - x00) * ay) * r0;
        i = fmin(fmax(i, -1.0), (double)ncx + 1.0);
        j = fmin(fmax(j, -1.0), (double)ncy + 1.0);
    }
    // Newton on the bilinear map of the current cell, written for instruction count (a whole warp runs as long as its
    // slowest lane): fused arithmetic and an approximate reciprocal refined once -- the iteration corrects itself, and
    // a guess is verified by the chain check anyway
#pragma unroll 1
    for (int it = 0; it < 12; it++) {
        const int i0 = min(max(__double2int_rd(i), 0), ncx - 1), j0 = min(max(__double2int_rd(j), 0), ncy - 1);
        const double* gx = g.x + ((int64_t)i0 * g.ny + j0);
        const double* gy = g.y + ((int64_t)i0 * g.ny + j0);
        const double x00c = gx[0], x01 = gx[1], x10 = gx[g.ny], x11 = gx[g.ny + 1];
        const double y00c = gy[0], y01 = gy[1], y10 = gy[g.ny], y11 = gy[g.ny + 1];
        const double u = i - i0, v = j - j0, u1 = 1.0 - u, v1 = 1.0 - v;
        const double xa = dfma(x10, u, x00c * u1), xb = dfma(x11, u, x01 * u1);
        const double ya = dfma(y10, u, y00c * u1), yb = dfma(y11, u, y01 * u1);
        const double ex = dfma(xb, v, xa * v1) - px, ey = dfma(yb, v, ya * v1) - py;
        const double dxdi = dfma(x11 - x01, v, (x10 - x00c) * v1);
        const double dxdj = dfma(x11 - x10, u, (x01 - x00c) * u1);
        const double dydi = dfma(y11 - y01, v, (y10 - y00c) * v1);
        const double dydj = dfma(y11 - y10, u, (y01 - y00c) * u1);
        const double det = dfma(dxdi, dydj, -(dxdj * dydi));
        double rr;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rr) : "d"(det));
        const double rdet = dfma(rr, dfma(-det, rr, 1.0), rr);
        const double di = dfma(dydj, ex, -(dxdj * ey)) * rdet;
        const double dj = dfma(dxdi, ey, -(dydi * ex)) * rdet;
        const double big = fmax(fabs(di), fabs(dj));
        if (!(big < 4.0 * (g.nx + g.ny))) break;   // diverging or NaN (degenerate cell): the careful path decides
        i -= di;
        j -= dj;
        if (big < 1e-4) {
            const int ic = __double2int_rd(i), jc = __double2int_rd(j);
            if (ic >= 0 && jc >= 0 && ic < ncx && jc < ncy) {
                const double fu = i - ic, fv = j - jc;
                // (no exact containment test here: with the step below 1e-4 the solution is ~1e-8 cells from the
                // root, so 1e-3 cells inside the unit square of a convex cell is inside the cell)
                if (fu > 1e-3 && fu < 1.0 - 1e-3 && fv > 1e-3 && fv < 1.0 - 1e-3) return ic * ncy + jc;
            } else if (i < -1e-3 || j < -1e-3 || i > ncx + 1e-3 || j > ncy + 1e-3) {
                return kLocOutside;  // clearly beyond the border cells: a guess needs no exact verification
            }
            break;
        }
    }
    if (!(i == i) || !(j == j)) { i = 0.5 * g.nx; j = 0.5 * g.ny; }
    return locate_newton(g, px, py, i, j);
}

}  // namespace rg
