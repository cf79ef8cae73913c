// rg_multilinear2d.cu -- 2D multilinear (bilinear) weights on a curvilinear vertex grid and their
// application: BASELINE config 5 ("find_indices + multilinear regrid of 4096x4096 curvilinear vertices
// onto 8192x8192 output points").
//
// The reference has NO 2D multilinear (regridding/_weights/_weights_multilinear.py:128-131 raises); this is
// the 2D extension of its 1D rule (wml.py:105-119, 185-202):
//   * the containing cell comes from rg_find_indices_2d (the reference's internal locators);
//   * inside the cell the point is (u, v) of the cell's bilinear map X(u, v) = x00 + u (x10 - x00) +
//     v (x01 - x00) + u v (x00 - x10 - x01 + x11) (u along axis 0), solved by Newton from (1/2, 1/2);
//   * weights (1-u)(1-v), (1-u) v, u (1-v), u v at the vertices (i,j), (i,j+1), (i+1,j), (i+1,j+1) --
//     ascending flat vertex index, the order the reference's (input, output)-sorted layout gives;
//   * points outside the grid: bounds "extrapolate" continues the bilinear map of the nearest border cell
//     (u, v outside [0, 1]: the 1D rule clamps the cell index and keeps the unclamped ratio, wml.py:105-119),
//     "nan" poisons the four weights (wml.py:192-194), "raise" is reported to the host (wml.py:110-114).
// Parity: unpinned by the reference (no behaviour to compare with); pinned by the NumPy restatement in
// oracle/oracle.py (multilinear2d_weights) and by exactness on functions linear in (x, y).
#include "rg_common.cuh"
#include "rg_geom.cuh"

namespace rg {

constexpr int kBoundsExtrapolate = 0;
constexpr int kBoundsNan = 1;
constexpr int kBoundsRaise = 2;

// Newton in index space on the piecewise-bilinear map, continuous result (the iteration of locate_newton
// without the final containment resolve).  Returns false if it did not converge.
__device__ inline bool newton_index(const GridView& g, double px, double py, double& i, double& j)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
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
        if (det == 0.0 || !(det == det)) return false;
        double di = (dydj * ex - dxdj * ey) / det;
        double dj = (-dydi * ex + dxdi * ey) / det;
        di = fmin(fmax(di, -(double)g.nx), (double)g.nx);
        dj = fmin(fmax(dj, -(double)g.ny), (double)g.ny);
        i -= di;
        j -= dj;
        if (fabs(di) < 1e-9 && fabs(dj) < 1e-9) return true;
    }
    return false;
}

__global__ void k_bilinear_weights(GridView g, int64_t n, const double* __restrict__ px, const double* __restrict__ py,
                                   const int64_t* __restrict__ cell, int64_t fill, int mode,
                                   int64_t* __restrict__ idx4, double* __restrict__ w4, int32_t* __restrict__ n_outside)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < n;
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    int64_t c = valid ? cell[p] : 0;
    const bool outside = valid && (c == fill);
    const unsigned m = __ballot_sync(0xffffffffu, outside);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(n_outside, __popc(m));
    if (!valid) return;
    const double x = px[p], y = py[p];
    int ci, cj;
    double u = 0.5, v = 0.5;
    bool poison = false;
    if (outside) {
        ci = cj = 0;
        if (mode != kBoundsExtrapolate) {
            poison = true;
        } else {
            double i = 0.5 * g.nx, j = 0.5 * g.ny;
            if (newton_index(g, x, y, i, j)) {
                ci = min(max((int)floor(fmin(fmax(i, -2.0), (double)ncx + 2.0)), 0), ncx - 1);
                cj = min(max((int)floor(fmin(fmax(j, -2.0), (double)ncy + 2.0)), 0), ncy - 1);
                u = i - ci;
                v = j - cj;
            } else {
                poison = true;  // no nearest cell could be determined
            }
        }
    } else {
        ci = (int)(c / ncy);
        cj = (int)(c - (int64_t)ci * ncy);
    }
    const int64_t a = (int64_t)ci * g.ny + cj;
    if (!poison) {
        const double x00 = g.x[a], x01 = g.x[a + 1], x10 = g.x[a + g.ny], x11 = g.x[a + g.ny + 1];
        const double y00 = g.y[a], y01 = g.y[a + 1], y10 = g.y[a + g.ny], y11 = g.y[a + g.ny + 1];
        const double ax = dsub(x10, x00), bx = dsub(x01, x00), cx = dadd(dsub(dsub(x00, x10), x01), x11);
        const double ay = dsub(y10, y00), by = dsub(y01, y00), cy = dadd(dsub(dsub(y00, y10), y01), y11);
        double prev = INFINITY;
        for (int it = 0; it < 24; it++) {
            const double ex = dsub(dadd(dadd(dadd(x00, dmul(u, ax)), dmul(v, bx)), dmul(dmul(u, v), cx)), x);
            const double ey = dsub(dadd(dadd(dadd(y00, dmul(u, ay)), dmul(v, by)), dmul(dmul(u, v), cy)), y);
            const double xu = dadd(ax, dmul(v, cx)), xv = dadd(bx, dmul(u, cx));
            const double yu = dadd(ay, dmul(v, cy)), yv = dadd(by, dmul(u, cy));
            const double det = dsub(dmul(xu, yv), dmul(xv, yu));
            if (det == 0.0 || !(det == det)) {
                poison = true;
                break;
            }
            const double du = ddiv(dsub(dmul(yv, ex), dmul(xv, ey)), det);
            const double dv = ddiv(dsub(dmul(xu, ey), dmul(yu, ex)), det);
            u = dsub(u, du);
            v = dsub(v, dv);
            // converged, or stagnating at the rounding noise of the residual (~ eps * |x| / cell size)
            const double step = fmax(fabs(du), fabs(dv));
            if (step < 1e-14 || (step < 1e-6 && step >= prev)) break;
            prev = step;
        }
    }
    double w00, w01, w10, w11;
    if (poison) {
        w00 = w01 = w10 = w11 = __longlong_as_double(0x7ff8000000000000LL);
    } else {
        const double u1 = dsub(1.0, u), v1 = dsub(1.0, v);
        w00 = dmul(u1, v1);
        w01 = dmul(u1, v);
        w10 = dmul(u, v1);
        w11 = dmul(u, v);
    }
    longlong4 iq;
    iq.x = a; iq.y = a + 1; iq.z = a + g.ny; iq.w = a + g.ny + 1;
    *reinterpret_cast<longlong4*>(idx4 + 4 * p) = iq;
    *reinterpret_cast<double4*>(w4 + 4 * p) = make_double4(w00, w01, w10, w11);
}

// values_out[f][p] = sum_k w4[p][k] * values_in[f][idx4[p][k]], k ascending (ascending input index), separately
// rounded multiply and add from +0.0: what the reference's apply loop does with (input, output)-sorted weights.
template <int FT>
__global__ void __launch_bounds__(256)
k_ell4_apply(int64_t n_frames, int64_t n_in, int64_t n_points, const int64_t* __restrict__ idx4,
             const double* __restrict__ w4, const double* __restrict__ vin, double* __restrict__ vout)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    const longlong4 iq = *reinterpret_cast<const longlong4*>(idx4 + 4 * p);
    const double4 wq = *reinterpret_cast<const double4*>(w4 + 4 * p);
    const int64_t f0 = (int64_t)blockIdx.y * FT;
#pragma unroll
    for (int t = 0; t < FT; t++) {
        const int64_t f = f0 + t;
        if (f >= n_frames) break;
        const double* in = vin + f * n_in;
        double acc = 0.0;
        acc = dadd(acc, dmul(wq.x, __ldg(in + iq.x)));
        acc = dadd(acc, dmul(wq.y, __ldg(in + iq.y)));
        acc = dadd(acc, dmul(wq.z, __ldg(in + iq.z)));
        acc = dadd(acc, dmul(wq.w, __ldg(in + iq.w)));
        vout[f * n_points + p] = acc;
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_multilinear2d_weights(int device, void* stream, int64_t nx, int64_t ny,
                                        const double* x, const double* y,
                                        int64_t n_points, const double* px, const double* py,
                                        const int64_t* cell_flat, int64_t fill, int bounds_mode,
                                        int64_t* idx4, double* w4, int32_t* n_outside_dev)
{
    if (nx < 2 || ny < 2 || !x || !y || n_points < 0 || bounds_mode < 0 || bounds_mode > 2)
        return fail(RG_E_ARG, "rg_multilinear2d_weights: bad argument");
    if (nx * ny >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_multilinear2d_weights: grid too large");
    if (!n_outside_dev) return fail(RG_E_ARG, "rg_multilinear2d_weights: null counter");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaMemsetAsync(n_outside_dev, 0, sizeof(int32_t), st));
    if (n_points == 0) return RG_OK;
    if (!px || !py || !cell_flat || !idx4 || !w4) return fail(RG_E_ARG, "rg_multilinear2d_weights: null pointer");
    if ((((uintptr_t)idx4) & 31) || (((uintptr_t)w4) & 31))
        return fail(RG_E_ARG, "rg_multilinear2d_weights: idx4 / w4 must be 32-byte aligned");
    const GridView g{ x, y, (int)nx, (int)ny };
    k_bilinear_weights<<<(unsigned)ceil_div(n_points, 256), 256, 0, st>>>(g, n_points, px, py, cell_flat, fill, bounds_mode,
                                                                         idx4, w4, n_outside_dev);
    RG_LAUNCH_CHECK("k_bilinear_weights");
    return RG_OK;
}

extern "C" int rg_ell4_apply(int device, void* stream, int64_t n_frames, int64_t n_in, int64_t n_points,
                             const int64_t* idx4, const double* w4, const double* values_in, double* values_out)
{
    if (n_frames < 0 || n_in < 1 || n_points < 0) return fail(RG_E_ARG, "rg_ell4_apply: bad argument");
    if (n_frames == 0 || n_points == 0) return RG_OK;
    if (!idx4 || !w4 || !values_in || !values_out) return fail(RG_E_ARG, "rg_ell4_apply: null pointer");
    if ((((uintptr_t)idx4) & 31) || (((uintptr_t)w4) & 31)) return fail(RG_E_ARG, "rg_ell4_apply: idx4 / w4 must be 32-byte aligned");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int FT = 8;
    for (int64_t f0 = 0; f0 < n_frames; f0 += (int64_t)65535 * FT) {
        const int64_t nf = n_frames - f0 < (int64_t)65535 * FT ? n_frames - f0 : (int64_t)65535 * FT;
        dim3 grid((unsigned)ceil_div(n_points, 256), (unsigned)ceil_div(nf, FT));
        k_ell4_apply<FT><<<grid, 256, 0, st>>>(nf, n_in, n_points, idx4, w4, values_in + f0 * n_in,
                                               values_out + f0 * n_points);
        RG_LAUNCH_CHECK("k_ell4_apply");
    }
    return RG_OK;
}
