// rg_boundary.cuh -- boundary of a static grid: edges in the reference's scan order, two
// levels of bounding boxes over them, the entry query of _step_outside_static and the
// exact boundary winding number.  Shared by the 2D build and the 2D cell location.
#pragma once
#include "rg_common.cuh"
#include "rg_geom.cuh"

namespace rg {

struct BBox { double xlo, ylo, xhi, yhi; };

struct Boundary {
    int n_edges, n_g1, n_g2;
    int ne_a0;  // ny-1 edges on each axis-0 face (i = 0, i = nx-1)
    int ne_a1;  // nx-1 edges on each axis-1 face (j = 0, j = ny-1)
    double *x3, *y3, *x4, *y4;  // endpoints in the reference's scan order (c2d.py:462-483)
    int32_t* cell;              // flat index of the cell entered through the edge (c2d.py:499-515)
    BBox* bb1;                  // bbox of 32 consecutive edges
    BBox* bb2;                  // bbox of 32 consecutive groups
};

// ---------------------------------------------------------------------------
// bounding box of a grid (x.min(), y.min(), x.max(), y.max(); c2d.py:179-180)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_double(double* addr, double v)
{
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (!(v < __longlong_as_double((long long)assumed))) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v)
{
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (!(v > __longlong_as_double((long long)assumed))) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

// ---------------------------------------------------------------------------
// K2: boundary edges in the scan order of _step_outside_static (c2d.py:462-483):
//   axis 0: faces i = 0 then i = nx-1, edge m joins (i_face, m)-(i_face, m+1), enters cell (0 | ncx-1, m)
//   axis 1: faces j = 0 then j = ny-1, edge m joins (m, j_face)-(m+1, j_face), enters cell (m, 0 | ncy-1)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void boundary_edge(const GridView& g, const Boundary& b, int s,
                                              double& x3, double& y3, double& x4, double& y4, int& cell)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    int axis, face, m;
    if (s < 2 * b.ne_a0) {
        axis = 0; face = s >= b.ne_a0; m = s - face * b.ne_a0;
    } else {
        const int r = s - 2 * b.ne_a0;
        axis = 1; face = r >= b.ne_a1; m = r - face * b.ne_a1;
    }
    int64_t v3, v4;
    int ci, cj;
    if (axis == 0) {
        const int i = face ? g.nx - 1 : 0;
        v3 = (int64_t)i * g.ny + m;
        v4 = v3 + 1;
        ci = face ? ncx - 1 : 0;
        cj = m;
    } else {
        const int j = face ? g.ny - 1 : 0;
        v3 = (int64_t)m * g.ny + j;
        v4 = v3 + g.ny;
        ci = m;
        cj = face ? ncy - 1 : 0;
    }
    x3 = g.x[v3]; y3 = g.y[v3];
    x4 = g.x[v4]; y4 = g.y[v4];
    cell = ci * ncy + cj;
}

// Up to two grids per launch (blockIdx.y): the build prepares the input and the output grid together.
struct BoundarySet {
    GridView g[2];
    Boundary b[2];
    double* bbox[2];
    int n;
};

static __global__ void k_bbox_init(BoundarySet S)
{
    if (threadIdx.x < 4 * S.n) S.bbox[threadIdx.x >> 2][threadIdx.x & 3] = ((threadIdx.x & 3) < 2) ? INFINITY : -INFINITY;
}

static __global__ void k_bbox(const __grid_constant__ BoundarySet S)
{
    const GridView& g = S.g[blockIdx.y];
    double* bbox = S.bbox[blockIdx.y];  // xlo, ylo, xhi, yhi
    const int64_t n = (int64_t)g.nx * g.ny;
    double xlo = INFINITY, ylo = INFINITY, xhi = -INFINITY, yhi = -INFINITY;
    // eight (then four) independent 16-byte loads of each array in flight per thread (the scan is latency bound otherwise)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t done = 0;   // elements covered by the vector loop
    if (((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.y)) & 15) == 0) {
        const double2* x2 = reinterpret_cast<const double2*>(g.x);
        const double2* y2 = reinterpret_cast<const double2*>(g.y);
        const int64_t n2 = n / 2;
        int64_t r = q;
        for (; r + 7 * stride < n2; r += 8 * stride) {
            double2 x[8], y[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { x[u] = x2[r + u * stride]; y[u] = y2[r + u * stride]; }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                xlo = fmin(xlo, fmin(x[u].x, x[u].y)); xhi = fmax(xhi, fmax(x[u].x, x[u].y));
                ylo = fmin(ylo, fmin(y[u].x, y[u].y)); yhi = fmax(yhi, fmax(y[u].x, y[u].y));
            }
        }
        for (; r + 3 * stride < n2; r += 4 * stride) {
            double2 x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { x[u] = x2[r + u * stride]; y[u] = y2[r + u * stride]; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                xlo = fmin(xlo, fmin(x[u].x, x[u].y)); xhi = fmax(xhi, fmax(x[u].x, x[u].y));
                ylo = fmin(ylo, fmin(y[u].x, y[u].y)); yhi = fmax(yhi, fmax(y[u].x, y[u].y));
            }
        }
        for (; r < n2; r += stride) {
            const double2 x = x2[r], y = y2[r];
            xlo = fmin(xlo, fmin(x.x, x.y)); xhi = fmax(xhi, fmax(x.x, x.y));
            ylo = fmin(ylo, fmin(y.x, y.y)); yhi = fmax(yhi, fmax(y.x, y.y));
        }
        done = 2 * n2;
    }
    for (q += done; q < n; q += stride) {
        const double x = g.x[q], y = g.y[q];
        xlo = fmin(xlo, x); xhi = fmax(xhi, x);
        ylo = fmin(ylo, y); yhi = fmax(yhi, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xlo = fmin(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
        ylo = fmin(ylo, __shfl_xor_sync(0xffffffffu, ylo, o));
        xhi = fmax(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
        yhi = fmax(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_min_double(&bbox[0], xlo);
        atomic_min_double(&bbox[1], ylo);
        atomic_max_double(&bbox[2], xhi);
        atomic_max_double(&bbox[3], yhi);
    }
}

// one warp per group of 32 edges: lane = edge; the group's bounding box by shuffles
static __global__ void k_boundary_edges_bb1(const __grid_constant__ BoundarySet S)
{
    const GridView& g = S.g[blockIdx.y];
    const Boundary& b = S.b[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int g1 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g1 >= b.n_g1) return;
    const int s = g1 * 32 + lane;
    BBox r = { INFINITY, INFINITY, -INFINITY, -INFINITY };
    if (s < b.n_edges) {
        double x3, y3, x4, y4;
        int cell;
        boundary_edge(g, b, s, x3, y3, x4, y4, cell);
        b.x3[s] = x3; b.y3[s] = y3;
        b.x4[s] = x4; b.y4[s] = y4;
        b.cell[s] = cell;
        r.xlo = fmin(x3, x4); r.xhi = fmax(x3, x4);
        r.ylo = fmin(y3, y4); r.yhi = fmax(y3, y4);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r.xlo = fmin(r.xlo, __shfl_xor_sync(0xffffffffu, r.xlo, o));
        r.ylo = fmin(r.ylo, __shfl_xor_sync(0xffffffffu, r.ylo, o));
        r.xhi = fmax(r.xhi, __shfl_xor_sync(0xffffffffu, r.xhi, o));
        r.yhi = fmax(r.yhi, __shfl_xor_sync(0xffffffffu, r.yhi, o));
    }
    if (lane == 0) b.bb1[g1] = r;
}

static __global__ void k_boundary_bb2(const __grid_constant__ BoundarySet S)
{
    const Boundary& b = S.b[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int g2 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g2 >= b.n_g2) return;
    const int s = g2 * 32 + lane;
    BBox r = { INFINITY, INFINITY, -INFINITY, -INFINITY };
    if (s < b.n_g1) r = b.bb1[s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r.xlo = fmin(r.xlo, __shfl_xor_sync(0xffffffffu, r.xlo, o));
        r.ylo = fmin(r.ylo, __shfl_xor_sync(0xffffffffu, r.ylo, o));
        r.xhi = fmax(r.xhi, __shfl_xor_sync(0xffffffffu, r.xhi, o));
        r.yhi = fmax(r.yhi, __shfl_xor_sync(0xffffffffu, r.yhi, o));
    }
    if (lane == 0) b.bb2[g2] = r;
}

__device__ __forceinline__ int edge_local_id(const Boundary& b, int s)
{
    // index_flat((face, axis), (2, 2)) = face*2 + axis  (c2d.py:532-535)
    if (s < b.ne_a0) return 0;
    if (s < 2 * b.ne_a0) return 2;
    if (s < 2 * b.ne_a0 + b.ne_a1) return 1;
    return 3;
}

// _step_outside_static (c2d.py:401-552): among ALL boundary edges hit by p->q keep the
// smallest t (strict <: ties keep the first in scan order), skipping the cell just left.
// The two bbox levels only skip edges whose own bbox test (geometry.py:422-433) fails.
__device__ inline int boundary_entry(const Boundary& b, double px, double py, double qx, double qy,
                                     int last_cell, double& t_best)
{
    const double sxlo = fmin(px, qx), sxhi = fmax(px, qx);
    const double sylo = fmin(py, qy), syhi = fmax(py, qy);
    int best = -1;
    t_best = INFINITY;
    for (int g2 = 0; g2 < b.n_g2; g2++) {
        const BBox B2 = b.bb2[g2];
        if (!(sxlo <= B2.xhi && B2.xlo <= sxhi && sylo <= B2.yhi && B2.ylo <= syhi)) continue;
        const int g1e = min(b.n_g1, (g2 + 1) * 32);
        for (int g1 = g2 * 32; g1 < g1e; g1++) {
            const BBox B1 = b.bb1[g1];
            if (!(sxlo <= B1.xhi && B1.xlo <= sxhi && sylo <= B1.yhi && B1.ylo <= syhi)) continue;
            const int se = min(b.n_edges, (g1 + 1) * 32);
            for (int s = g1 * 32; s < se; s++) {
                double t;
                if (!seg_hit(px, py, qx, qy, b.x3[s], b.y3[s], b.x4[s], b.y4[s], t)) continue;
                if (!(t < t_best)) continue;
                if (b.cell[s] == last_cell) continue;
                t_best = t;
                best = s;
            }
        }
    }
    return best;
}


// Extended winding number of the boundary polygon (grid_boundary, _grids.py:167-215, under
// point_is_inside_polygon, geometry.py:737-829) around (px, py), evaluated only on the
// edges whose y-range contains py: every other edge contributes exactly 0, and the
// contributions are multiples of 1/2, so the sum is exact in any order.
// contribution of the 32 edges of group g1
__device__ __forceinline__ double boundary_winding_group(const Boundary& b, int g1, double px, double py)
{
    double w = 0.0;
    const int se = min(b.n_edges, (g1 + 1) * 32);
    for (int s = g1 * 32; s < se; s++) {
        const double x0 = dsub(b.x3[s], px), y0 = dsub(b.y3[s], py);
        const double x1 = dsub(b.x4[s], px), y1 = dsub(b.y4[s], py);
        // the polygon runs counter-clockwise in index space; the scan-order edges of the
        // i = 0 face and of the j = ny-1 face run the other way round
        const bool reversed = (s < b.ne_a0) || (s >= 2 * b.ne_a0 + b.ne_a1);
        w += reversed ? winding_edge(x1, y1, x0, y0) : winding_edge(x0, y0, x1, y1);
    }
    return w;
}

__device__ inline double boundary_winding(const Boundary& b, double px, double py)
{
    double w = 0.0;
    for (int g2 = 0; g2 < b.n_g2; g2++) {
        const BBox B2 = b.bb2[g2];
        if (!(B2.ylo <= py && py <= B2.yhi)) continue;
        const int g1e = min(b.n_g1, (g2 + 1) * 32);
        for (int g1 = g2 * 32; g1 < g1e; g1++) {
            const BBox B1 = b.bb1[g1];
            if (!(B1.ylo <= py && py <= B1.yhi)) continue;
            w += boundary_winding_group(b, g1, px, py);
        }
    }
    return w;
}

static void carve_boundary(Carver& c, Boundary& b, int64_t nx, int64_t ny)
{
    b.ne_a0 = (int)(ny - 1);
    b.ne_a1 = (int)(nx - 1);
    b.n_edges = 2 * b.ne_a0 + 2 * b.ne_a1;
    b.n_g1 = (int)ceil_div(b.n_edges, 32);
    b.n_g2 = (int)ceil_div(b.n_g1, 32);
    b.x3 = c.take<double>(b.n_edges);
    b.y3 = c.take<double>(b.n_edges);
    b.x4 = c.take<double>(b.n_edges);
    b.y4 = c.take<double>(b.n_edges);
    b.cell = c.take<int32_t>(b.n_edges);
    b.bb1 = c.take<BBox>(b.n_g1);
    b.bb2 = c.take<BBox>(b.n_g2);
}


// builds the boundary structures and the bboxes (xlo, ylo, xhi, yhi) of one or two grids: 4 launches
// (`n_bbox` < n: the bboxes of the grids n_bbox .. n-1 are only initialised; the caller reduces them elsewhere)
static inline int build_boundaries(cudaStream_t st, int n, const GridView* g, const Boundary* b, double* const* bbox,
                                   int n_bbox = -1)
{
    if (n_bbox < 0) n_bbox = n;
    const int T = 256;
    BoundarySet S;
    memset(&S, 0, sizeof(S));
    S.n = n;
    int g1max = 0, g2max = 0;
    for (int q = 0; q < n; q++) {
        S.g[q] = g[q]; S.b[q] = b[q]; S.bbox[q] = bbox[q];
        g1max = b[q].n_g1 > g1max ? b[q].n_g1 : g1max;
        g2max = b[q].n_g2 > g2max ? b[q].n_g2 : g2max;
    }
    k_bbox_init<<<1, 32, 0, st>>>(S);
    if (n_bbox > 0) k_bbox<<<dim3(kNumSM * 4, n_bbox), T, 0, st>>>(S);
    RG_LAUNCH_CHECK("k_bbox");
    k_boundary_edges_bb1<<<dim3((unsigned)ceil_div((int64_t)g1max * 32, T), n), T, 0, st>>>(S);
    k_boundary_bb2<<<dim3((unsigned)ceil_div((int64_t)g2max * 32, T), n), T, 0, st>>>(S);
    RG_LAUNCH_CHECK("k_boundary");
    return RG_OK;
}

static inline int build_boundary(cudaStream_t st, const GridView& g, const Boundary& b, double* bbox)
{
    return build_boundaries(st, 1, &g, &b, &bbox);
}

}  // namespace rg
