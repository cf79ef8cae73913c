/*
 * oracle_regrid.c -- CPU ORACLE for the first-order conservative regridding path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (regridding_b200/) may
 * import, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference leg,
 * and there only as the checker / the CPU baseline.
 *
 * It is a plain-C restatement of the algorithm of sun-data/regridding (the
 * reference is Python + Numba; it has no native sources, so there is nothing to
 * compile into oracle/_ref).  Every function cites the reference file:line it
 * follows.  Paths are relative to the reference root:
 *   c2d.py    = regridding/_weights/_weights_conservative_2d/_weights_conservative_2d.py
 *   grids2.py = regridding/_weights/_weights_conservative_2d/_grids.py
 *   geom.py   = regridding/geometry.py
 *   c1d.py    = regridding/_weights/_weights_conservative_1d/_weights_conservative_1d.py
 *   grids1.py = regridding/_weights/_weights_conservative_1d/_grids.py
 *   warr.py   = regridding/_weights/_weights_arrays.py
 *   rfw.py    = regridding/_regrid/_regrid_from_weights.py
 *   fib.py    = regridding/_find_indices/_find_indices_brute.py
 *   fis.py    = regridding/_find_indices/_find_indices_searchsorted.py
 *   interp.py = regridding/_interp_ndarray.py
 *
 * PARITY PINNING: pinned.  tests/golden/ holds outputs of the reference itself
 * (run in the build container through Numba, script tests/golden/make_golden.py)
 * and tests/test_oracle_golden.py checks this file against them.
 *
 * Arithmetic modes (orc_set_mode):
 *   1 (default) "jit":   reproduces the floating-point contraction pattern that
 *                        Numba/LLVM emits for the fastmath=True kernels on an
 *                        x86-64 FMA host (which product of a*b-c*d is fused,
 *                        reciprocal-multiply for t and u, fma for the
 *                        intersection point).  This is what users of the
 *                        reference get.
 *   0           "strict": plain IEEE evaluation in source order, what the
 *                        reference computes under NUMBA_DISABLE_JIT=1.
 * Compile with -ffp-contract=off so that only the explicit fma() calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXSIZE INT64_MAX

static int g_mode = 1;

void orc_set_mode(int mode) { g_mode = mode; }
int orc_get_mode(void) { return g_mode; }
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* bench.py's reference arm: torchrun exports OMP_NUM_THREADS=1, which would starve the CPU baseline */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------ */
/* geometry primitives                                                       */
/* ------------------------------------------------------------------------ */

/* geom.py:968-995  area_triangle: signed area of (origin, v1, v2). */
static inline double area_triangle(double x1, double y1, double x2, double y2)
{
    if (g_mode)
        return 0.5 * fma(x1, y2, -(x2 * y1));
    return (x1 * y2 - x2 * y1) / 2;
}

/* geom.py:64-105  point_is_inside_box_2d (inclusive). */
static inline int point_in_box(double x, double y, double xlo, double ylo, double xhi, double yhi)
{
    if (!(xlo <= x && x <= xhi))
        return 0;
    if (!(ylo <= y && y <= yhi))
        return 0;
    return 1;
}

/* geom.py:153-284  bounding_boxes_intersect_2d (touching counts). */
static inline int bboxes_intersect(double xp1, double yp1, double xp2, double yp2,
                                   double xq1, double yq1, double xq2, double yq2)
{
    double t;
    if (xp1 > xp2) { t = xp1; xp1 = xp2; xp2 = t; }
    if (xq1 > xq2) { t = xq1; xq1 = xq2; xq2 = t; }
    if (!(xp1 <= xq2 && xq1 <= xp2))
        return 0;
    if (yp1 > yp2) { t = yp1; yp1 = yp2; yp2 = t; }
    if (yq1 > yq2) { t = yq1; yq1 = yq2; yq2 = t; }
    if (!(yp1 <= yq2 && yq1 <= yp2))
        return 0;
    return 1;
}

/* geom.py:370-445  two_line_segment_intersection_parameters. */
static inline void seg_params(double x1, double y1, double x2, double y2,
                              double x3, double y3, double x4, double y4,
                              double* t, double* u)
{
    if (!bboxes_intersect(x1, y1, x2, y2, x3, y3, x4, y4)) {
        *t = INFINITY;
        *u = INFINITY;
        return;
    }
    if (g_mode) {
        double tdet = fma(x1 - x3, y3 - y4, -((y1 - y3) * (x3 - x4)));
        double det = fma(x1 - x2, y3 - y4, -((y1 - y2) * (x3 - x4)));
        double nudet = fma(y1 - y2, x1 - x3, -((y1 - y3) * (x1 - x2)));
        double rinv = 1.0 / det;
        *t = tdet * rinv;
        *u = nudet * rinv;
    } else {
        double tdet = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4);
        double udet = (x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3);
        double det = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4);
        *t = tdet / det;
        *u = -udet / det;
    }
}

/* geom.py:448-475  two_line_segments_intersect: half-open on both parameters. */
static inline int seg_hit(double t, double u)
{
    return (0 <= t && t < 1) && (0 <= u && u < 1);
}

/* geom.py:478-556  two_line_segment_intersection. */
static inline void seg_point(double x1, double y1, double x2, double y2, double t,
                             double* x, double* y)
{
    if (g_mode) {
        *x = fma(t, x2 - x1, x1);
        *y = fma(t, y2 - y1, y1);
    } else {
        *x = x1 + t * (x2 - x1);
        *y = y1 + t * (y2 - y1);
    }
}

/* geom.py:737-829  point_is_inside_polygon: extended winding number with
 * half-integer increments for vertices on the horizontal through the point.
 * Edge v runs from vertex v-1 (wrapping) to vertex v. */
int orc_point_in_polygon(double x, double y, const double* vx, const double* vy, int64_t n)
{
    double w = 0;
    for (int64_t v = 0; v < n; v++) {
        int64_t i = (v == 0) ? n - 1 : v - 1;
        double x0 = vx[i] - x, y0 = vy[i] - y;
        double x1 = vx[v] - x, y1 = vy[v] - y;
        if (y0 * y1 < 0) {
            double r = x0 + y0 * (x1 - x0) / (y0 - y1);
            if (r > 0)
                w += (y0 < 0) ? 1 : -1;
            else
                w += (y0 < 0) ? -1 : 1;
        } else if (y0 == 0) {
            if (x0 > 0) {
                if (y1 > 0) w += 0.5;
                else if (y1 < 0) w -= 0.5;
            } else if (x0 < 0) {
                if (y1 < 0) w += 0.5;
                else if (y1 > 0) w -= 0.5;
            }
        } else if (y1 == 0) {
            if (x1 > 0) {
                if (y0 < 0) w += 0.5;
                else if (y0 > 0) w -= 0.5;
            } else if (x1 < 0) {
                if (y0 > 0) w += 0.5;
                else if (y0 < 0) w -= 0.5;
            }
        }
    }
    return w != 0;
}

/* ------------------------------------------------------------------------ */
/* 2D grids                                                                  */
/* ------------------------------------------------------------------------ */

/* A strided 2D view so that the transposed ("axis aligned right",
 * c2d/_arrays.py:123-149) grid needs no copy. */
typedef struct {
    const double* x;
    const double* y;
    int64_t n0, n1; /* shape */
    int64_t s0, s1; /* strides in elements */
} grid_t;

#define GX(g, i, j) ((g).x[(i) * (g).s0 + (j) * (g).s1])
#define GY(g, i, j) ((g).y[(i) * (g).s0 + (j) * (g).s1])

static grid_t grid_make(const double* x, const double* y, int64_t nx, int64_t ny)
{
    grid_t g = { x, y, nx, ny, ny, 1 };
    return g;
}

static grid_t grid_transposed(grid_t g)
{
    grid_t t = { g.x, g.y, g.n1, g.n0, g.s1, g.s0 };
    return t;
}

/* grids2.py:50-140  grid_volume: signed cell areas.  Axis-0 pass first (x and y
 * swapped on the transposed view), then axis-1; per edge: left cell += a, right
 * cell -= a, visited in ascending i, so each cell accumulates
 * (((0 - a_i) + a_{i+1}) - b_j) + b_{j+1}. */
void orc_grid_volume(const double* x, const double* y, int64_t nx, int64_t ny, double* out)
{
    int64_t ncx = nx - 1, ncy = ny - 1;
    for (int64_t c = 0; c < ncx * ncy; c++)
        out[c] = 0.0;
    for (int axis = 0; axis < 2; axis++) {
        grid_t g = grid_make(x, y, nx, ny);
        int64_t o0 = ncy, o1 = 1; /* strides of out */
        if (axis == 0) {
            g = grid_transposed(g);
            const double* tmp = g.x; g.x = g.y; g.y = tmp;
            o0 = 1; o1 = ncy;
        }
        int64_t num_i = g.n0, num_j = g.n1;
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < num_j - 1; j++) {
            for (int64_t i = 0; i < num_i; i++) {
                double a;
                if (g_mode && i == 0) {
                    /* LLVM peels the first iteration of this loop and there fuses the OTHER
                     * product: RN(x1*y2) - x2*y1 (measured against the JIT). */
                    a = 0.5 * fma(-GX(g, i, j + 1), GY(g, i, j), GX(g, i, j) * GY(g, i, j + 1));
                } else {
                    a = area_triangle(GX(g, i, j), GY(g, i, j), GX(g, i, j + 1), GY(g, i, j + 1));
                }
                if (i - 1 >= 0)
                    out[(i - 1) * o0 + j * o1] += a;
                if (i < num_i - 1)
                    out[i * o0 + j * o1] -= a;
            }
        }
    }
}

/* grids2.py:167-215  grid_boundary: 2(nx-1)+2(ny-1) vertices, counter-clockwise
 * in index space. */
static int64_t grid_boundary(grid_t g, double** bx_out, double** by_out)
{
    int64_t nx = g.n0, ny = g.n1;
    int64_t nb = 2 * (nx - 1) + 2 * (ny - 1);
    double* bx = (double*)malloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1));
    double* by = (double*)malloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1));
    int64_t n = 0;
    for (int64_t i = 0; i < nx - 1; i++) { bx[n] = GX(g, i, 0); by[n] = GY(g, i, 0); n++; }
    for (int64_t j = 0; j < ny - 1; j++) { bx[n] = GX(g, nx - 1, j); by[n] = GY(g, nx - 1, j); n++; }
    for (int64_t i = 0; i < nx - 1; i++) { bx[n] = GX(g, nx - 1 - i, ny - 1); by[n] = GY(g, nx - 1 - i, ny - 1); n++; }
    for (int64_t j = 0; j < ny - 1; j++) { bx[n] = GX(g, 0, ny - 1 - j); by[n] = GY(g, 0, ny - 1 - j); n++; }
    *bx_out = bx;
    *by_out = by;
    return nb;
}

static inline int cell_contains(grid_t g, int64_t i, int64_t j, double px, double py)
{
    double vx[4], vy[4];
    vx[0] = GX(g, i, j);         vy[0] = GY(g, i, j);
    vx[1] = GX(g, i + 1, j);     vy[1] = GY(g, i + 1, j);
    vx[2] = GX(g, i + 1, j + 1); vy[2] = GY(g, i + 1, j + 1);
    vx[3] = GX(g, i, j + 1);     vy[3] = GY(g, i, j + 1);
    return orc_point_in_polygon(px, py, vx, vy, 4);
}

/* grids2.py:223-279  index_of_point_brute: first containing cell in row-major
 * order, else (maxsize, maxsize). */
static void index_of_point_brute(grid_t g, double px, double py, int64_t* oi, int64_t* oj)
{
    for (int64_t i = 0; i < g.n0 - 1; i++)
        for (int64_t j = 0; j < g.n1 - 1; j++)
            if (cell_contains(g, i, j, px, py)) {
                *oi = i; *oj = j;
                return;
            }
    *oi = ORC_MAXSIZE; *oj = ORC_MAXSIZE;
}

/* grids2.py:286-349  _index_of_point_local: lowest-index containing cell among
 * the 3x3 neighbourhood of (i0, j0). */
static void index_of_point_local(grid_t g, double px, double py, int64_t i0, int64_t j0,
                                 int64_t* oi, int64_t* oj)
{
    int64_t ncx = g.n0 - 1, ncy = g.n1 - 1;
    int64_t ilo = i0 - 1 > 0 ? i0 - 1 : 0, ihi = i0 + 2 < ncx ? i0 + 2 : ncx;
    int64_t jlo = j0 - 1 > 0 ? j0 - 1 : 0, jhi = j0 + 2 < ncy ? j0 + 2 : ncy;
    for (int64_t i = ilo; i < ihi; i++)
        for (int64_t j = jlo; j < jhi; j++)
            if (cell_contains(g, i, j, px, py)) {
                *oi = i; *oj = j;
                return;
            }
    *oi = ORC_MAXSIZE; *oj = ORC_MAXSIZE;
}

/* interp.py:226-269  _bilinear_interpolation in index space; the cell index is
 * clamped so the map extrapolates linearly outside the grid. */
static double bilerp(const double* a, int64_t n0, int64_t n1, int64_t s0, int64_t s1, double x, double y)
{
    int64_t x0 = (int64_t)floor(x), y0 = (int64_t)floor(y);
    if (x0 < 0) x0 = 0; else if (x0 > n0 - 2) x0 = n0 - 2;
    if (y0 < 0) y0 = 0; else if (y0 > n1 - 2) y0 = n1 - 2;
    double a00 = a[x0 * s0 + y0 * s1];
    double a01 = a[x0 * s0 + (y0 + 1) * s1];
    double a10 = a[(x0 + 1) * s0 + y0 * s1];
    double a11 = a[(x0 + 1) * s0 + (y0 + 1) * s1];
    double dx = x - (double)x0, dy = y - (double)y0;
    double w00 = (1 - dx) * (1 - dy), w01 = (1 - dx) * dy, w10 = dx * (1 - dy), w11 = dx * dy;
    return (a00 * w00) + (a01 * w01) + (a10 * w10) + (a11 * w11);
}

/* grids2.py:356-463  index_of_point_secant: Newton iteration in index space on
 * the bilinear map (forward-difference Jacobian, h = 1e-3), resolved to the
 * lowest-index containing cell by the local search; falls back to brute. */
void orc_index_of_point_secant_g(grid_t g, double px, double py, int64_t* oi, int64_t* oj)
{
    const double h = 1e-3;
    int64_t ncx = g.n0 - 1, ncy = g.n1 - 1;
    double i = (double)g.n0 / 2, j = (double)g.n1 / 2;
    for (int it = 0; it < 100; it++) {
        double X = bilerp(g.x, g.n0, g.n1, g.s0, g.s1, i, j);
        double Y = bilerp(g.y, g.n0, g.n1, g.s0, g.s1, i, j);
        double fi = floor(i), fj = floor(j);
        double ex = X - px, ey = Y - py;
        if (fi >= 0 && fj >= 0 && fi < (double)ncx && fj < (double)ncy) {
            int64_t i0 = (int64_t)fi, j0 = (int64_t)fj;
            if (cell_contains(g, i0, j0, px, py)) {
                index_of_point_local(g, px, py, i0, j0, oi, oj);
                return;
            }
        }
        if (fabs(ex) < 1e-10 && fabs(ey) < 1e-10) {
            /* converged on a face or outside the grid */
            if (!(fabs(fi) < 9e18 && fabs(fj) < 9e18)) { *oi = ORC_MAXSIZE; *oj = ORC_MAXSIZE; return; }
            index_of_point_local(g, px, py, (int64_t)fi, (int64_t)fj, oi, oj);
            return;
        }
        double dx_di = (bilerp(g.x, g.n0, g.n1, g.s0, g.s1, i + h, j) - X) / h;
        double dx_dj = (bilerp(g.x, g.n0, g.n1, g.s0, g.s1, i, j + h) - X) / h;
        double dy_di = (bilerp(g.y, g.n0, g.n1, g.s0, g.s1, i + h, j) - Y) / h;
        double dy_dj = (bilerp(g.y, g.n0, g.n1, g.s0, g.s1, i, j + h) - Y) / h;
        double det = dx_di * dy_dj - dx_dj * dy_di;
        if (det == 0) {
            index_of_point_brute(g, px, py, oi, oj);
            return;
        }
        double di = (+dy_dj * ex - dx_dj * ey) / det;
        double dj = (-dy_di * ex + dx_di * ey) / det;
        i -= di;
        j -= dj;
        if (!(fabs(i) < 1e18 && fabs(j) < 1e18)) /* diverged: the reference would fall through to brute eventually */
            break;
    }
    index_of_point_brute(g, px, py, oi, oj);
}

void orc_index_of_point_secant(const double* x, const double* y, int64_t nx, int64_t ny,
                               double px, double py, int64_t* out)
{
    orc_index_of_point_secant_g(grid_make(x, y, nx, ny), px, py, &out[0], &out[1]);
}

void orc_index_of_point_brute(const double* x, const double* y, int64_t nx, int64_t ny,
                              double px, double py, int64_t* out)
{
    index_of_point_brute(grid_make(x, y, nx, ny), px, py, &out[0], &out[1]);
}

/* Many points against one grid (test harness for the 2D find_indices extension): flat cell index
 * i*(ny-1)+j per point, `fill` when no cell contains it.  method 0 = secant (c2d/_grids.py:356-463),
 * 1 = brute (c2d/_grids.py:223-279). */
void orc_index_of_points(int method, const double* x, const double* y, int64_t nx, int64_t ny,
                         int64_t npts, const double* px, const double* py, int64_t fill, int64_t* out)
{
    const grid_t g = grid_make(x, y, nx, ny);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = 0; p < npts; p++) {
        int64_t i, j;
        if (method == 0)
            orc_index_of_point_secant_g(g, px[p], py[p], &i, &j);
        else
            index_of_point_brute(g, px[p], py[p], &i, &j);
        out[p] = (i == ORC_MAXSIZE || j == ORC_MAXSIZE) ? fill : i * (ny - 1) + j;
    }
}

/* ------------------------------------------------------------------------ */
/* 2D conservative sweep (Ramshaw 1985)                                      */
/* ------------------------------------------------------------------------ */

typedef struct {
    int64_t* ii;
    int64_t* io;
    double* v;
    int64_t n, cap;
} tripvec_t;

static void tv_push(tripvec_t* t, int64_t ii, int64_t io, double v)
{
    if (t->n == t->cap) {
        int64_t cap = t->cap ? 2 * t->cap : 64;
        t->ii = (int64_t*)realloc(t->ii, sizeof(int64_t) * (size_t)cap);
        t->io = (int64_t*)realloc(t->io, sizeof(int64_t) * (size_t)cap);
        t->v = (double*)realloc(t->v, sizeof(double) * (size_t)cap);
        t->cap = cap;
    }
    t->ii[t->n] = ii;
    t->io[t->n] = io;
    t->v[t->n] = v;
    t->n++;
}

typedef struct {
    grid_t sweep;   /* aligned: last axis is the walking direction */
    grid_t stat;    /* static grid, never transposed */
    double bb_lo_x, bb_lo_y, bb_hi_x, bb_hi_y;
    const double* bnd_x;
    const double* bnd_y;
    int64_t nbnd;
    const double* vol_in;   /* (ncx_in, ncy_in) */
    const double* w_in;     /* optional, same shape */
    int64_t ncx_in, ncy_in, ncx_out, ncy_out;
    int sweep_input;
    int axis_sweep;
} sweep_ctx_t;

typedef struct {
    double x1, y1, x2, y2;
    int64_t k;           /* index_sweep_y */
    int64_t ci, cj;      /* static cell */
    int64_t last;        /* index_edge_last */
    int outside;
} walk_t;

/* c2d.py:401-552  _step_outside_static: scan ALL boundary edges of the static
 * grid, keep the hit with the smallest t (strict <, so ties keep the first in
 * scan order), skipping the cell just left. */
static void step_outside(const sweep_ctx_t* c, walk_t* s)
{
    grid_t g = c->stat;
    int64_t nc[2] = { g.n0 - 1, g.n1 - 1 };
    double px1 = s->x1, py1 = s->y1, px2 = s->x2, py2 = s->y2;
    int64_t last_cell = s->last;
    int found = 0;
    double t_min = INFINITY;
    int64_t new_last = ORC_MAXSIZE;
    int outside = 1;
    for (int axis = 0; axis < 2; axis++) {
        grid_t a = (axis == 0) ? grid_transposed(g) : g;
        int64_t num_i = a.n0, num_j = a.n1;
        for (int face = 0; face < 2; face++) {
            int64_t j_edge = face ? num_j - 1 : 0;
            for (int64_t i0 = 0; i0 < num_i - 1; i0++) {
                double x3 = GX(a, i0, j_edge), y3 = GY(a, i0, j_edge);
                double x4 = GX(a, i0 + 1, j_edge), y4 = GY(a, i0 + 1, j_edge);
                double t, u;
                seg_params(px1, py1, px2, py2, x3, y3, x4, y4, &t, &u);
                if (!seg_hit(t, u))
                    continue;
                if (!(t < t_min))
                    continue;
                int64_t m = i0;
                int64_t n = face ? nc[axis] - 1 : 0;
                int64_t cx, cy;
                if (axis == 0) { cx = n; cy = m; } else { cx = m; cy = n; }
                int64_t flat = cx * nc[1] + cy;
                if (flat == last_cell)
                    continue;
                found = 1;
                t_min = t;
                s->ci = cx;
                s->cj = cy;
                seg_point(px1, py1, px2, py2, t, &s->x2, &s->y2);
                new_last = face * 2 + axis;
                outside = 0;
            }
        }
    }
    if (!found)
        s->k += 1;
    s->last = new_last;
    s->outside = outside;
}

/* c2d.py:757-830 + 876-912  _calc_and_save_weights / _index_input_output. */
static void emit_piece(const sweep_ctx_t* c, tripvec_t* out, double x1, double y1, double x2, double y2,
                       int64_t L, int64_t k, int64_t ci, int64_t cj)
{
    double area;
    if (g_mode) {
        /* The JIT evaluates the NEGATED form RN(x2*y1) - x1*y2 (fused) and folds the
         * axis_sweep sign into a select on +-0.5; bit-symmetric with geom.py:993 +
         * c2d.py:783-784 except for the sign of an exact zero. */
        double neg2 = fma(-x1, y2, x2 * y1);
        area = neg2 * (c->axis_sweep == 0 ? 0.5 : -0.5);
    } else {
        area = area_triangle(x1, y1, x2, y2);
        if (c->axis_sweep == 0)
            area = -area;
    }
    int64_t nlines = c->sweep.n0;
    for (int side = 0; side < 2; side++) {
        int64_t i = side ? L : L - 1;
        if (side == 0 && !(i >= 0))
            continue;
        if (side == 1 && !(i < nlines - 1))
            continue;
        int64_t si = i, sj = k;
        if (c->axis_sweep == 0) { si = k; sj = i; }
        int64_t in_i, in_j, out_i, out_j;
        if (c->sweep_input) { in_i = si; in_j = sj; out_i = ci; out_j = cj; }
        else                { in_i = ci; in_j = cj; out_i = si; out_j = sj; }
        int64_t flat_in = in_i * c->ncy_in + in_j;
        int64_t flat_out = out_i * c->ncy_out + out_j;
        double a = side ? -area : area;
        double w;
        if (c->w_in && g_mode) {
            /* fastmath reassociation seen in the JIT: (area * w_in) / volume */
            w = (a * c->w_in[flat_in]) / c->vol_in[flat_in];
        } else {
            w = a / c->vol_in[flat_in];
            if (c->w_in)
                w *= c->w_in[flat_in];
        }
        tv_push(out, flat_in, flat_out, w);
    }
}

/* c2d.py:561-749  _step_inside_static: test the (up to) four edges of the
 * current static cell in the order v = 0..3, skipping the edge just crossed;
 * FIRST hit wins; a weight is emitted for the (possibly clipped) piece always. */
static void step_inside(const sweep_ctx_t* c, walk_t* s, int64_t L, tripvec_t* out)
{
    static const int64_t vert[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } }; /* grids2.py:153-158 */
    static const int64_t norm[4][2] = { { -1, 0 }, { 0, -1 }, { 1, 0 }, { 0, 1 } }; /* grids2.py:143-148 */
    grid_t g = c->stat;
    int64_t ncs0 = c->sweep_input ? c->ncx_out : c->ncx_in;
    int64_t ncs1 = c->sweep_input ? c->ncy_out : c->ncy_in;
    double x1 = s->x1, y1 = s->y1, x2 = s->x2, y2 = s->y2;
    int64_t nk, ni, nj, nlast;
    int hit = 0;
    for (int v = 0; v < 4; v++) {
        if ((int64_t)v == s->last)
            continue;
        int pv = (v + 3) & 3;
        int64_t i3 = vert[pv][0] + s->ci, j3 = vert[pv][1] + s->cj;
        int64_t i4 = vert[v][0] + s->ci, j4 = vert[v][1] + s->cj;
        double t, u;
        seg_params(x1, y1, x2, y2, GX(g, i3, j3), GY(g, i3, j3), GX(g, i4, j4), GY(g, i4, j4), &t, &u);
        if (seg_hit(t, u)) {
            seg_point(x1, y1, s->x2, s->y2, t, &x2, &y2);
            nk = s->k;
            ni = s->ci + norm[v][0];
            nj = s->cj + norm[v][1];
            nlast = (v + 2) % 4;
            hit = 1;
            break;
        }
    }
    if (!hit) {
        nk = s->k + 1;
        ni = s->ci;
        nj = s->cj;
        nlast = ORC_MAXSIZE;
    }
    emit_piece(c, out, x1, y1, x2, y2, L, s->k, s->ci, s->cj);
    int outside = 0;
    if (ni < 0 || nj < 0 || ni >= ncs0 || nj >= ncs1) {
        nlast = s->ci * ncs1 + s->cj;
        ni = ORC_MAXSIZE;
        nj = ORC_MAXSIZE;
        outside = 1;
    }
    s->x2 = x2;
    s->y2 = y2;
    s->k = nk;
    s->ci = ni;
    s->cj = nj;
    s->last = nlast;
    s->outside = outside;
}

/* c2d.py:286-387  body of the prange loop of _sweep_along_axis: one sweep line. */
static void sweep_line(const sweep_ctx_t* c, int64_t L, tripvec_t* out)
{
    grid_t sw = c->sweep;
    int64_t n = sw.n1;
    walk_t s;
    s.k = 0;
    s.ci = ORC_MAXSIZE;
    s.cj = ORC_MAXSIZE;
    s.outside = 1;
    s.last = ORC_MAXSIZE;
    s.x1 = GX(sw, L, 0);
    s.y1 = GY(sw, L, 0);
    if (point_in_box(s.x1, s.y1, c->bb_lo_x, c->bb_lo_y, c->bb_hi_x, c->bb_hi_y)) {
        if (orc_point_in_polygon(s.x1, s.y1, c->bnd_x, c->bnd_y, c->nbnd)) {
            orc_index_of_point_secant_g(c->stat, s.x1, s.y1, &s.ci, &s.cj);
            s.outside = 0;
            if (s.ci == ORC_MAXSIZE) /* the reference would index out of range here (undefined) */
                s.outside = 1;
        }
    }
    while (s.k < n - 1) {
        s.x2 = GX(sw, L, s.k + 1);
        s.y2 = GY(sw, L, s.k + 1);
        if (s.outside)
            step_outside(c, &s);
        else
            step_inside(c, &s, L, out);
        s.x1 = s.x2;
        s.y1 = s.y2;
    }
}

/* c2d.py:80-126 weights_conservative_2d + c2d.py:133-203 _sweep_grid +
 * c2d.py:27-73 _compact_rows.  Returns the raw (uncoalesced) triplets in the
 * reference's emission order: pass (sweep OUTPUT axis 0, axis 1; sweep INPUT
 * axis 0, axis 1), then line, then walk order, left before right.
 * Arrays are malloc'ed; release with orc_free. */
int64_t orc_weights_conservative_2d(const double* xin, const double* yin, int64_t nxi, int64_t nyi,
                                    const double* xout, const double* yout, int64_t nxo, int64_t nyo,
                                    const double* w_in,
                                    int64_t** ii_out, int64_t** io_out, double** v_out)
{
    int64_t ncxi = nxi - 1, ncyi = nyi - 1;
    double* vol = (double*)malloc(sizeof(double) * (size_t)(ncxi * ncyi > 0 ? ncxi * ncyi : 1));
    orc_grid_volume(xin, yin, nxi, nyi, vol);

    grid_t gin = grid_make(xin, yin, nxi, nyi);
    grid_t gout = grid_make(xout, yout, nxo, nyo);

    tripvec_t* pass_rows[4];
    int64_t pass_nlines[4];
    int p = 0;
    for (int sweep_input = 0; sweep_input < 2; sweep_input++) {
        grid_t gs = sweep_input ? gin : gout;
        grid_t gt = sweep_input ? gout : gin;
        sweep_ctx_t c;
        memset(&c, 0, sizeof(c));
        c.stat = gt;
        c.bb_lo_x = c.bb_lo_y = INFINITY;
        c.bb_hi_x = c.bb_hi_y = -INFINITY;
        for (int64_t q = 0; q < gt.n0 * gt.n1; q++) {
            if (gt.x[q] < c.bb_lo_x) c.bb_lo_x = gt.x[q];
            if (gt.x[q] > c.bb_hi_x) c.bb_hi_x = gt.x[q];
            if (gt.y[q] < c.bb_lo_y) c.bb_lo_y = gt.y[q];
            if (gt.y[q] > c.bb_hi_y) c.bb_hi_y = gt.y[q];
        }
        double *bx, *by;
        c.nbnd = grid_boundary(gt, &bx, &by);
        c.bnd_x = bx;
        c.bnd_y = by;
        c.vol_in = vol;
        c.w_in = w_in;
        c.ncx_in = ncxi; c.ncy_in = ncyi;
        c.ncx_out = nxo - 1; c.ncy_out = nyo - 1;
        c.sweep_input = sweep_input;
        for (int axis = 0; axis < 2; axis++, p++) {
            c.sweep = (axis == 0) ? grid_transposed(gs) : gs;
            c.axis_sweep = axis;
            int64_t nlines = c.sweep.n0;
            tripvec_t* rows = (tripvec_t*)calloc((size_t)(nlines > 0 ? nlines : 1), sizeof(tripvec_t));
#pragma omp parallel for schedule(dynamic, 1)
            for (int64_t L = 0; L < nlines; L++)
                sweep_line(&c, L, &rows[L]);
            pass_rows[p] = rows;
            pass_nlines[p] = nlines;
        }
        free(bx);
        free(by);
    }
    int64_t total = 0;
    for (p = 0; p < 4; p++)
        for (int64_t L = 0; L < pass_nlines[p]; L++)
            total += pass_rows[p][L].n;
    size_t alloc = (size_t)(total > 0 ? total : 1);
    int64_t* ii = (int64_t*)malloc(sizeof(int64_t) * alloc);
    int64_t* io = (int64_t*)malloc(sizeof(int64_t) * alloc);
    double* v = (double*)malloc(sizeof(double) * alloc);
    int64_t w = 0;
    for (p = 0; p < 4; p++) {
        for (int64_t L = 0; L < pass_nlines[p]; L++) {
            tripvec_t* r = &pass_rows[p][L];
            if (r->n) {
                memcpy(ii + w, r->ii, sizeof(int64_t) * (size_t)r->n);
                memcpy(io + w, r->io, sizeof(int64_t) * (size_t)r->n);
                memcpy(v + w, r->v, sizeof(double) * (size_t)r->n);
                w += r->n;
            }
            free(r->ii); free(r->io); free(r->v);
        }
        free(pass_rows[p]);
    }
    free(vol);
    *ii_out = ii;
    *io_out = io;
    *v_out = v;
    return total;
}

/* ------------------------------------------------------------------------ */
/* coalesce (NumPy semantics)                                                */
/* ------------------------------------------------------------------------ */

/* NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
 * @TYPE@_pairwise_sum; numpy 2.3.x as pinned by the image, `numpy>2` in the
 * reference's pyproject.toml:25), which np.add.reduceat uses per segment. */
static double np_pairwise_sum(const double* a, int64_t n)
{
    if (n < 8) {
        double res = -0.0;
        for (int64_t i = 0; i < n; i++)
            res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int q = 0; q < 8; q++) r[q] = a[q];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int q = 0; q < 8; q++) r[q] += a[i + q];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++)
            res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

/* One segment of np.add.reduceat: out = a[0]; out += pairwise_sum(a[1:]). */
double orc_reduceat_segment(const double* a, int64_t n)
{
    double out = a[0];
    if (n > 1)
        out += np_pairwise_sum(a + 1, n - 1);
    return out;
}

typedef struct { int64_t key; int64_t idx; } keyidx_t;

static int cmp_keyidx(const void* a, const void* b)
{
    const keyidx_t* p = (const keyidx_t*)a;
    const keyidx_t* q = (const keyidx_t*)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    if (p->idx != q->idx) return p->idx < q->idx ? -1 : 1;
    return 0;
}

/* warr.py:44-73  _coalesce: key relative to the minima, STABLE sort, merge
 * duplicate (input, output) pairs with np.add.reduceat.  Output arrays must hold
 * n elements; returns the number of unique pairs. */
int64_t orc_coalesce(int64_t n, const int64_t* ii, const int64_t* io, const double* v,
                     int64_t* ii_o, int64_t* io_o, double* v_o)
{
    if (n == 0)
        return 0;
    int64_t base_in = ii[0], base_out = io[0], max_out = io[0];
    for (int64_t q = 1; q < n; q++) {
        if (ii[q] < base_in) base_in = ii[q];
        if (io[q] < base_out) base_out = io[q];
        if (io[q] > max_out) max_out = io[q];
    }
    int64_t span = max_out - base_out + 1;
    keyidx_t* k = (keyidx_t*)malloc(sizeof(keyidx_t) * (size_t)n);
    for (int64_t q = 0; q < n; q++) {
        k[q].key = (ii[q] - base_in) * span + (io[q] - base_out);
        k[q].idx = q;
    }
    qsort(k, (size_t)n, sizeof(keyidx_t), cmp_keyidx); /* (key, original index) order == stable sort by key */
    double* tmp = (double*)malloc(sizeof(double) * 256);
    int64_t tmp_cap = 256;
    int64_t m = 0;
    int64_t s = 0;
    while (s < n) {
        int64_t e = s + 1;
        while (e < n && k[e].key == k[s].key) e++;
        int64_t len = e - s;
        if (len > tmp_cap) {
            tmp_cap = 2 * len;
            tmp = (double*)realloc(tmp, sizeof(double) * (size_t)tmp_cap);
        }
        for (int64_t q = 0; q < len; q++) tmp[q] = v[k[s + q].idx];
        /* floor division / modulo on non-negative keys */
        ii_o[m] = k[s].key / span + base_in;
        io_o[m] = k[s].key % span + base_out;
        v_o[m] = orc_reduceat_segment(tmp, len);
        m++;
        s = e;
    }
    free(tmp);
    free(k);
    return m;
}

/* ------------------------------------------------------------------------ */
/* apply                                                                     */
/* ------------------------------------------------------------------------ */

/* rfw.py:165-182  _regrid_from_weights for D slices that SHARE one set of
 * weights (the broadcast case of rfw.py:108).  Sequential scatter-add per slice
 * with separately rounded multiply and add; negative indices wrap.
 * values_out must be zero-filled by the caller (rfw.py:111-118). */
void orc_regrid_from_weights(int64_t nnz, const int64_t* ii, const int64_t* io, const double* v,
                             int64_t D, int64_t n_in, int64_t n_out,
                             const double* values_in, double* values_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < D; d++) {
        const double* vin = values_in + d * n_in;
        double* vout = values_out + d * n_out;
        for (int64_t w = 0; w < nnz; w++) {
            int64_t a = ii[w] < 0 ? ii[w] + n_in : ii[w];
            int64_t b = io[w] < 0 ? io[w] + n_out : io[w];
            double prod = v[w] * vin[a];
            vout[b] = vout[b] + prod;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* 1D conservative                                                           */
/* ------------------------------------------------------------------------ */

/* grids1.py:38-73  index_of_point: bisection, returns the right bracket. */
static int64_t index_of_point_1d(const double* g, int64_t sg, int64_t num, double point)
{
    int64_t lo = 0, hi = num;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) / 2;
        if (g[mid * sg] > point) hi = mid; else lo = mid;
    }
    return hi;
}

/* c1d.py:60-189 _weights_conservative_1d, c1d.py:193-236 _step_outside_static,
 * c1d.py:240-318 _step_inside_static.  One spectrum.  Outputs must hold
 * n_in + n_out elements (an upper bound on the pieces).  Reproduces the
 * reference's treatment of descending grids, including that length_input is
 * taken on the REVERSED array but indexed with the complemented (negative,
 * wrap-around) index. */
int64_t orc_weights_conservative_1d(const double* x_in, int64_t n_in, const double* x_out, int64_t n_out,
                                    const double* w_in, int64_t* ii, int64_t* io, double* v)
{
    int rev_sw = !(x_in[0] < x_in[n_in - 1]);
    int rev_st = !(x_out[0] < x_out[n_out - 1]);
    /* reversed views: element q of the view is base[q * stride] */
    const double* sw = rev_sw ? x_in + (n_in - 1) : x_in;
    int64_t ssw = rev_sw ? -1 : 1;
    const double* st = rev_st ? x_out + (n_out - 1) : x_out;
    int64_t sst = rev_st ? -1 : 1;
#define SW(q) (sw[(q) * ssw])
#define ST(q) (st[(q) * sst])
    int64_t ncell = n_in - 1;
    double st_left = ST(0), st_right = ST(n_out - 1);
    int64_t k = 0, s;
    int outside;
    double p1 = SW(0);
    if (st_left == p1) { outside = 0; s = 0; }
    else if (st_left < p1 && p1 < st_right) { outside = 0; s = index_of_point_1d(st, sst, n_out, p1) - 1; }
    else { outside = 1; s = ORC_MAXSIZE; }
    int64_t n = 0;
    while (k < n_in - 1) {
        double p2 = SW(k + 1);
        if (outside) {
            double e = ST(0);
            if (p1 < e && e < p2) { s = 0; p2 = e; }
            else if (e == p2) { k += 1; s = 0; }
            else { k += 1; }
            if (s < ORC_MAXSIZE) outside = 0;
        } else {
            int64_t i_in = rev_sw ? ~k : k;
            int64_t i_out = rev_st ? ~s : s;
            double e = ST(s + 1);
            if (p1 < e && e < p2) { s += 1; p2 = e; }
            else if (e == p2) { s += 1; k += 1; }
            else { k += 1; }
            /* length_input = diff(x_sweep) on the (possibly reversed) view, indexed with wrap-around */
            int64_t li = i_in < 0 ? i_in + ncell : i_in;
            double length = SW(li + 1) - SW(li);
            double ratio = (p2 - p1) / length;
            if (w_in) {
                int64_t wi = i_in < 0 ? i_in + ncell : i_in;
                ratio = ratio * w_in[wi];
            }
            ii[n] = i_in; io[n] = i_out; v[n] = ratio;
            n++;
            if (!(0 <= s && s < n_out - 1)) break;
        }
        p1 = p2;
    }
#undef SW
#undef ST
    return n;
}

/* wcons.py:59-106: the per-spectrum loop over a stack of S independent grids.
 * counts[s] receives the number of triplets of spectrum s, written at
 * offset s * (n_in + n_out) of the output arrays. */
void orc_weights_conservative_1d_batched(int64_t S, const double* x_in, int64_t n_in,
                                         const double* x_out, int64_t n_out, const double* w_in,
                                         int64_t* ii, int64_t* io, double* v, int64_t* counts)
{
    int64_t cap = n_in + n_out;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; s++)
        counts[s] = orc_weights_conservative_1d(x_in + s * n_in, n_in, x_out + s * n_out, n_out,
                                                w_in ? w_in + s * (n_in - 1) : NULL,
                                                ii + s * cap, io + s * cap, v + s * cap);
}

/* ------------------------------------------------------------------------ */
/* find_indices (1D)                                                         */
/* ------------------------------------------------------------------------ */

/* fib.py:24-51  first m with x[m] <= p <= x[m+1] (inclusive), else fill. */
void orc_find_indices_brute_1d(int64_t D, int64_t n, int64_t m, const double* x_in, const double* x_out,
                               int64_t fill, int64_t* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < D; d++) {
        const double* xi = x_in + d * n;
        for (int64_t i = 0; i < m; i++) {
            double p = x_out[d * m + i];
            int64_t r = fill;
            for (int64_t q = 0; q < n - 1; q++)
                if (xi[q] <= p && p <= xi[q + 1]) { r = q; break; }
            out[d * m + i] = r;
        }
    }
}

/* fis.py:24-62  np.searchsorted(side="left") - 1 with the edge fix-ups. */
void orc_find_indices_searchsorted_1d(int64_t D, int64_t n, int64_t m, const double* x_in, const double* x_out,
                                      int64_t fill, int64_t* out)
{
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < D; d++) {
        const double* xi = x_in + d * n;
        for (int64_t i = 0; i < m; i++) {
            double p = x_out[d * m + i];
            int64_t lo = 0, hi = n; /* first index with xi[idx] >= p  (NaN sorts last) */
            while (lo < hi) {
                int64_t mid = lo + (hi - lo) / 2;
                int less = (xi[mid] < p) || (p != p && xi[mid] == xi[mid]);
                if (less) lo = mid + 1; else hi = mid;
            }
            int64_t r = lo - 1;
            if (p == xi[0]) r = 0;
            else if (r < 0) r = fill;
            else if (r > n - 2) r = fill;
            out[d * m + i] = r;
        }
    }
}

/* _fill/_gauss_seidel.py:83-139  red-black Gauss-Seidel relaxation of the cells flagged in `where`
 * (uint8, 1 = missing) on a periodic (num_y, num_x) grid, num_t independent frames, in place.
 * The sweep is sequential (j outer, i inner) inside a colour exactly like the reference: with an odd
 * size the periodic wrap joins two cells of the SAME colour, so the order decides which of them sees
 * the other's new value.  The JIT (fastmath) folds dcent into the two constants and contracts
 *     (dxxinv * (a_w + a_e) + dyyinv * (a_s + a_n)) * dcent
 * into fma(dyyinv * dcent, a_s + a_n, (dxxinv * dcent) * (a_w + a_e)); measured against the reference
 * (tests/golden/golden_fill.npz) -- strict mode evaluates the source expression as written. */
void orc_fill_gauss_seidel_2d(double* a, const unsigned char* where, int64_t num_t, int64_t num_y, int64_t num_x,
                              int64_t num_iterations)
{
    const double dx = (1.0 - -1.0) / (double)(num_x - 1);
    const double dy = (1.0 - -1.0) / (double)(num_y - 1);
    const double dxxinv = 1.0 / (dx * dx);
    const double dyyinv = 1.0 / (dy * dy);
    const double dcent = 1.0 / (2.0 * (dxxinv + dyyinv));
    const double cx = dxxinv * dcent, cy = dyyinv * dcent;
    const int strict = orc_get_mode() == 0;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < num_t; t++) {
        double* at = a + t * num_y * num_x;
        const unsigned char* wt = where + t * num_y * num_x;
        for (int64_t k = 0; k < num_iterations; k++) {
            for (int odd = 0; odd < 2; odd++) {
                for (int64_t j = 0; j < num_y; j++) {
                    for (int64_t i = 0; i < num_x; i++) {
                        if (((i + j) & 1) != odd || !wt[j * num_x + i]) continue;
                        const int64_t i9 = (i - 1 + num_x) % num_x, i1 = (i + 1) % num_x;
                        const int64_t j9 = (j - 1 + num_y) % num_y, j1 = (j + 1) % num_y;
                        const double sx = at[j * num_x + i9] + at[j * num_x + i1];
                        const double sy = at[j9 * num_x + i] + at[j1 * num_x + i];
                        at[j * num_x + i] = strict ? (dxxinv * sx + dyyinv * sy) * dcent : fma(cy, sy, cx * sx);
                    }
                }
            }
        }
    }
}
