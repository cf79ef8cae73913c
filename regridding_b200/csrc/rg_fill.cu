// rg_fill.cu -- fill(method="gauss_seidel"): red-black Gauss-Seidel relaxation of missing cells.
//
// Replaces _fill_gauss_seidel_2d / _iteration_gauss_seidel_2d
// (regridding/_fill/_gauss_seidel.py:83-139; called from fill_gauss_seidel, :13-59).
//
// The reference sweeps a frame sequentially (j outer, i inner) once per colour.  Cells of one colour only read
// cells of the other colour -- except across the periodic wrap when a size is odd: then (j, 0) and (j, nx-1)
// (or (0, i) and (ny-1, i)) have the SAME colour, and the sweep order decides who sees whose new value: the
// first row / column reads the old value of the last one, the last one reads the new value of the first.  That
// order is reproduced with up to three dependency LEVELS per colour (level = [i == nx-1 and nx odd] +
// [j == ny-1 and ny odd]) separated by grid-wide barriers; inside a level every update is independent, so the
// result equals the sequential sweep bit for bit.  The caller passes, per (colour, level), the flat indices of
// the missing cells (built once with device-side compaction), so an iteration touches only the cells it updates.
// One cooperative launch runs all iterations of all frames (grid.sync() between levels).
//
// Arithmetic: the JIT (fastmath) evaluates (dxxinv (a_w + a_e) + dyyinv (a_s + a_n)) dcent as
// fma(dyyinv dcent, a_s + a_n, (dxxinv dcent) (a_w + a_e)) -- measured against the reference, see
// oracle_regrid.c:orc_fill_gauss_seidel_2d and tests/golden/golden_fill.npz.
#include <cooperative_groups.h>

#include "rg_common.cuh"

namespace cg = cooperative_groups;

namespace rg {

struct FillLists {
    const int32_t* idx[2][3];  // [colour][level] flat cell indices t * ny * nx + j * nx + i
    int64_t n[2][3];
};

__global__ void __launch_bounds__(256)
k_fill_gauss_seidel(double* __restrict__ a, const FillLists L, int ny, int nx, int num_iterations, double cx, double cy)
{
    cg::grid_group grid = cg::this_grid();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    const int plane = ny * nx;
    for (int k = 0; k < num_iterations; k++) {
        for (int odd = 0; odd < 2; odd++) {
            for (int lvl = 0; lvl < 3; lvl++) {
                const int64_t n = L.n[odd][lvl];
                if (n == 0 && lvl > 0) continue;  // uniform: no barrier for an empty dependency level
                const int32_t* __restrict__ idx = L.idx[odd][lvl];
                for (int64_t q = gtid; q < n; q += gsize) {
                    const int c = idx[q];
                    const int t0 = (c / plane) * plane;
                    const int r = c - t0;
                    const int j = r / nx, i = r - j * nx;
                    const int i9 = i == 0 ? nx - 1 : i - 1, i1 = i == nx - 1 ? 0 : i + 1;
                    const int j9 = j == 0 ? ny - 1 : j - 1, j1 = j == ny - 1 ? 0 : j + 1;
                    const double* at = a + t0;
                    // L2 loads (ld.global.cg): the neighbours were written by other SMs before the last barrier
                    const double sx = dadd(__ldcg(at + j * nx + i9), __ldcg(at + j * nx + i1));
                    const double sy = dadd(__ldcg(at + j9 * nx + i), __ldcg(at + j1 * nx + i));
                    a[c] = dfma(cy, sy, dmul(cx, sx));
                }
                grid.sync();
            }
        }
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_fill_gauss_seidel_2d(int device, void* stream, double* a, int64_t num_t, int64_t num_y, int64_t num_x,
                                       const int32_t* const* idx_lists_host /* 6 device pointers [colour][level] */,
                                       const int64_t* counts_host /* 6 */, int64_t num_iterations)
{
    if (num_t < 0 || num_y < 2 || num_x < 2 || num_iterations < 0 || !idx_lists_host || !counts_host)
        return fail(RG_E_ARG, "rg_fill_gauss_seidel_2d: bad argument (needs at least 2 cells along both axes)");
    if (num_t * num_y * num_x >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_fill_gauss_seidel_2d: more than 2^31 cells per call");
    if (num_t == 0 || num_iterations == 0) return RG_OK;
    if (!a) return fail(RG_E_ARG, "rg_fill_gauss_seidel_2d: null array");
    FillLists L;
    int64_t most = 0, total = 0;
    for (int q = 0; q < 6; q++) {
        const int64_t n = counts_host[q];
        if (n < 0 || (n > 0 && !idx_lists_host[q])) return fail(RG_E_ARG, "rg_fill_gauss_seidel_2d: bad index list");
        L.idx[q / 3][q % 3] = idx_lists_host[q];
        L.n[q / 3][q % 3] = n;
        most = n > most ? n : most;
        total += n;
    }
    if (total == 0) return RG_OK;
    RG_CUDA(cudaSetDevice(device));
    // _gauss_seidel.py:118-127 (Python floats: plain IEEE)
    const double dx = (1.0 - -1.0) / (double)(num_x - 1);
    const double dy = (1.0 - -1.0) / (double)(num_y - 1);
    const double dxxinv = 1.0 / (dx * dx);
    const double dyyinv = 1.0 / (dy * dy);
    const double dcent = 1.0 / (2.0 * (dxxinv + dyyinv));
    double cx = dxxinv * dcent, cy = dyyinv * dcent;
    int per_sm = 0;
    RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fill_gauss_seidel, 256, 0));
    int sms = 0;
    RG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int64_t blocks = ceil_div(most, 256);
    const int64_t cap = (int64_t)per_sm * sms;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    int ny = (int)num_y, nx = (int)num_x, iters = (int)num_iterations;
    void* args[] = { &a, &L, &ny, &nx, &iters, &cx, &cy };
    RG_CUDA(cudaLaunchCooperativeKernel((const void*)k_fill_gauss_seidel, dim3((unsigned)blocks), dim3(256), args, 0,
                                        (cudaStream_t)stream));
    return RG_OK;
}
