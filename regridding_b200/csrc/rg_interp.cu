// rg_interp.cu -- ndarray_linear_interpolation on the device.
//
// Replaces the compiled kernels of regridding/_interp_ndarray.py:
//   _ndarray_linear_interpolation_1d / _linear_interpolation   :192-209, 226-247  (not fastmath: plain IEEE
//       a0 * (1 - dx) + a1 * dx with the cell index clamped to [0, n - 2] => linear extrapolation)
//   _ndarray_linear_interpolation_2d / _bilinear_interpolation :211-223, 250-297  (fastmath=True; the parfor body
//       LLVM emits on the reference's x86-64 FMA host, measured with inspect_asm and pinned by goldens, is
//           p = fma(1 - dy, a00, RN(dy * a01));  q = fma(1 - dy, a10, RN(dy * a11));  result = fma(dx, q - p, p))
// The orthogonal-axis bookkeeping (`axis`, `axis_indices`, broadcasting) stays on the host (_interp.py): here every
// slice d of D interpolates its own array a[d] at its own indices.
#include "rg_common.cuh"

namespace rg {

__global__ void k_interp_linear_1d(int64_t D, int64_t n, int64_t m, int64_t a_stride, int64_t x_stride,
                                   const double* __restrict__ a, const double* __restrict__ x, double* __restrict__ out)
{
    const int64_t d = blockIdx.y;
    const double* ad = a + d * a_stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const double xx = x[d * x_stride + i];
        double f = floor(xx);
        int64_t x0 = (f < 0.0) ? 0 : ((f > (double)(n - 2)) ? n - 2 : (int64_t)f);   // NaN falls to the cast like int(nan)
        if (!(xx == xx)) x0 = 0;
        const double dx = dsub(xx, (double)x0);
        out[d * m + i] = dadd(dmul(ad[x0], dsub(1.0, dx)), dmul(ad[x0 + 1], dx));
    }
}

__global__ void k_interp_bilinear_2d(int64_t D, int64_t nx, int64_t ny, int64_t P, int64_t a_stride, int64_t xy_stride,
                                     const double* __restrict__ a, const double* __restrict__ x,
                                     const double* __restrict__ y, double* __restrict__ out)
{
    const int64_t d = blockIdx.y;
    const double* ad = a + d * a_stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
        const double xx = x[d * xy_stride + i], yy = y[d * xy_stride + i];
        const double fx = floor(xx), fy = floor(yy);
        int64_t x0 = (fx < 0.0) ? 0 : ((fx > (double)(nx - 2)) ? nx - 2 : (int64_t)fx);
        int64_t y0 = (fy < 0.0) ? 0 : ((fy > (double)(ny - 2)) ? ny - 2 : (int64_t)fy);
        if (!(xx == xx)) x0 = 0;
        if (!(yy == yy)) y0 = 0;
        const double dx = dsub(xx, (double)x0), dy = dsub(yy, (double)y0);
        const double a00 = ad[x0 * ny + y0], a01 = ad[x0 * ny + y0 + 1];
        const double a10 = ad[(x0 + 1) * ny + y0], a11 = ad[(x0 + 1) * ny + y0 + 1];
        const double omy = dsub(1.0, dy);
        const double p = dfma(omy, a00, dmul(dy, a01));
        const double q = dfma(omy, a10, dmul(dy, a11));
        out[d * P + i] = dfma(dx, dsub(q, p), p);
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_interp_linear_1d(int device, void* stream, int64_t D, int64_t n, int64_t m,
                                   int64_t a_stride, int64_t x_stride,
                                   const double* a, const double* x, double* out)
{
    if (D < 0 || n < 2 || m < 0 || a_stride < 0 || x_stride < 0) return fail(RG_E_ARG, "rg_interp_linear_1d: bad argument");
    if (D == 0 || m == 0) return RG_OK;
    if (!a || !x || !out) return fail(RG_E_ARG, "rg_interp_linear_1d: null pointer");
    RG_CUDA(cudaSetDevice(device));
    const int T = 256;
    int64_t gx = ceil_div(m, T);
    if (gx > 4096) gx = 4096;
    for (int64_t d0 = 0; d0 < D; d0 += 65535) {
        const int64_t nd = D - d0 < 65535 ? D - d0 : 65535;
        k_interp_linear_1d<<<dim3((unsigned)gx, (unsigned)nd), T, 0, (cudaStream_t)stream>>>(
            nd, n, m, a_stride, x_stride, a + d0 * a_stride, x + d0 * x_stride, out + d0 * m);
        RG_LAUNCH_CHECK("k_interp_linear_1d");
    }
    return RG_OK;
}

extern "C" int rg_interp_bilinear_2d(int device, void* stream, int64_t D, int64_t nx, int64_t ny, int64_t P,
                                     int64_t a_stride, int64_t xy_stride,
                                     const double* a, const double* x, const double* y, double* out)
{
    if (D < 0 || nx < 2 || ny < 2 || P < 0 || a_stride < 0 || xy_stride < 0)
        return fail(RG_E_ARG, "rg_interp_bilinear_2d: bad argument");
    if (D == 0 || P == 0) return RG_OK;
    if (!a || !x || !y || !out) return fail(RG_E_ARG, "rg_interp_bilinear_2d: null pointer");
    RG_CUDA(cudaSetDevice(device));
    const int T = 256;
    int64_t gx = ceil_div(P, T);
    if (gx > 8192) gx = 8192;
    for (int64_t d0 = 0; d0 < D; d0 += 65535) {
        const int64_t nd = D - d0 < 65535 ? D - d0 : 65535;
        k_interp_bilinear_2d<<<dim3((unsigned)gx, (unsigned)nd), T, 0, (cudaStream_t)stream>>>(
            nd, nx, ny, P, a_stride, xy_stride, a + d0 * a_stride, x + d0 * xy_stride, y + d0 * xy_stride, out + d0 * P);
        RG_LAUNCH_CHECK("k_interp_bilinear_2d");
    }
    return RG_OK;
}
