// rg_cons1d.cu -- 1D first-order conservative resampling, batched over S spectra.
//
// Replaces _weights_conservative_1d and its per-spectrum Python driver
// (regridding/_weights/_weights_conservative_1d/_weights_conservative_1d.py:12-56, 60-318;
//  regridding/_weights/_weights_conservative.py:59-106).  Not fastmath in the reference
// (c1d.py:59), so plain IEEE subtraction / division reproduces it bit for bit.
#include "rg_common.cuh"

namespace rg {

// A possibly reversed view of one edge array: element q is base[q * stride].
struct View1D {
    const double* base;
    int64_t stride;
    __device__ __forceinline__ double operator()(int64_t q) const { return base[q * stride]; }
};

__device__ __forceinline__ View1D ascending_view(const double* x, int64_t n, bool& reversed)
{
    reversed = !(x[0] < x[n - 1]);  // c1d.py:100-110
    return reversed ? View1D{ x + (n - 1), -1 } : View1D{ x, 1 };
}

// ---------------------------------------------------------------------------
// weights materialisation: the sequential walk of c1d.py:138-175, one thread per
// spectrum, triplets in the reference's emission order.
// ---------------------------------------------------------------------------
__global__ void k_cons1d_walk(int64_t S, int64_t n, int64_t m,
                              const double* __restrict__ x_in, const double* __restrict__ x_out,
                              const double* __restrict__ w_in,
                              int64_t* __restrict__ ii, int64_t* __restrict__ io, double* __restrict__ vv,
                              int64_t* __restrict__ counts)
{
    const int64_t sp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sp >= S) return;
    bool rev_sw, rev_st;
    const View1D sw = ascending_view(x_in + sp * n, n, rev_sw);
    const View1D st = ascending_view(x_out + sp * m, m, rev_st);
    const double* w = w_in ? w_in + sp * (n - 1) : nullptr;
    const int64_t cap = n + m;
    int64_t* oii = ii + sp * cap;
    int64_t* oio = io + sp * cap;
    double* ov = vv + sp * cap;
    const int64_t ncell = n - 1;

    const double st_left = st(0), st_right = st(m - 1);
    int64_t k = 0, s;
    bool outside;
    double p1 = sw(0);
    if (st_left == p1) {  // c1d.py:124-126
        outside = false;
        s = 0;
    } else if (st_left < p1 && p1 < st_right) {  // c1d.py:127-133 + _grids.py:38-73 (bisection)
        int64_t lo = 0, hi = m;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) / 2;
            if (st(mid) > p1) hi = mid;
            else lo = mid;
        }
        outside = false;
        s = hi - 1;
    } else {  // c1d.py:134-136
        outside = true;
        s = INT64_MAX;
    }
    int64_t cnt = 0;
    while (k < n - 1) {
        double p2 = sw(k + 1);
        if (outside) {  // c1d.py:193-236
            const double e = st(0);
            if (p1 < e && e < p2) { s = 0; p2 = e; }
            else if (e == p2) { k += 1; s = 0; }
            else { k += 1; }
            if (s < INT64_MAX) outside = false;
        } else {  // c1d.py:240-318
            const int64_t i_in = rev_sw ? ~k : k;
            const int64_t i_out = rev_st ? ~s : s;
            const double e = st(s + 1);
            if (p1 < e && e < p2) { s += 1; p2 = e; }
            else if (e == p2) { s += 1; k += 1; }
            else { k += 1; }
            // length_input is diff() of the (possibly reversed) view, indexed with the
            // complemented index -- the reference's behaviour for descending grids (c1d.py:118, 305-307)
            const int64_t li = i_in < 0 ? i_in + ncell : i_in;
            const double length = dsub(sw(li + 1), sw(li));
            double ratio = ddiv(dsub(p2, p1), length);
            if (w) ratio = dmul(ratio, w[li]);
            oii[cnt] = i_in;
            oio[cnt] = i_out;
            ov[cnt] = ratio;
            cnt++;
            if (!(0 <= s && s < m - 1)) break;  // c1d.py:172-173
        }
        p1 = p2;
    }
    counts[sp] = cnt;
}

// ---------------------------------------------------------------------------
// fused regrid: one thread per (spectrum, output cell).  The pieces the walk emits are
// exactly the (input cell, output cell) pairs with a positive-length overlap, with
// p1 = the larger left edge and p2 = the smaller right edge (copies of grid values, so
// no rounding is involved in choosing them); the reference's apply
// (rfw.py:179-182 on the (input, output)-sorted triplets) accumulates each output cell
// in ascending wrapped input index with separately rounded multiply and add.
// Lanes take consecutive output cells of one spectrum: coalesced edge loads and stores.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_regrid1d(int64_t S, int64_t n, int64_t m,
           const double* __restrict__ x_in, const double* __restrict__ x_out,
           const double* __restrict__ w_in,
           const double* __restrict__ vin, double* __restrict__ vout)
{
    const int64_t sp = blockIdx.y;
    bool rev_sw, rev_st;
    const View1D sw = ascending_view(x_in + sp * n, n, rev_sw);
    const View1D st = ascending_view(x_out + sp * m, m, rev_st);
    const double* w = w_in ? w_in + sp * (n - 1) : nullptr;
    const double* vi = vin + sp * (n - 1);
    double* vo = vout + sp * (m - 1);
    const int64_t ncell = n - 1;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < m - 1; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = rev_st ? (m - 2 - o) : o;  // position in the ascending view
        const double a = st(s), b = st(s + 1);
        // first sweep cell k with sw(k+1) > a
        int64_t lo = 0, hi = n - 1;  // answer in [0, n-1]
        while (lo < hi) {
            const int64_t mid = (lo + hi) / 2;
            if (sw(mid + 1) > a) hi = mid;
            else lo = mid + 1;
        }
        const int64_t k0 = lo;
        // last sweep cell k with sw(k) < b  (k1 < k0 when there is no overlap)
        lo = -1; hi = n - 2;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) / 2;
            if (sw(mid) < b) lo = mid;
            else hi = mid - 1;
        }
        const int64_t k1 = lo;
        double acc = 0.0;
        const int64_t cnt = k1 - k0 + 1;
        for (int64_t q = 0; q < cnt; q++) {
            const int64_t k = rev_sw ? (k1 - q) : (k0 + q);  // ascending wrapped input index
            const double l = sw(k), r = sw(k + 1);
            const double p1 = l > a ? l : a;
            const double p2 = r < b ? r : b;
            if (!(p1 < p2)) continue;
            const int64_t li = rev_sw ? (ncell - 1 - k) : k;  // wrapped index (= the reference's ~k + ncell)
            const double length = dsub(sw(li + 1), sw(li));
            double ratio = ddiv(dsub(p2, p1), length);
            if (w) ratio = dmul(ratio, w[li]);
            acc = dadd(acc, dmul(ratio, vi[li]));
        }
        vo[o] = acc;
    }
}

}  // namespace rg

using namespace rg;

extern "C" int rg_cons1d_batched(int device, void* stream, int64_t S, int64_t n, int64_t m,
                                 const double* x_in, const double* x_out, const double* w_in,
                                 int64_t* ii, int64_t* io, double* v, int64_t* counts)
{
    if (S < 0 || n < 2 || m < 2) return fail(RG_E_ARG, "rg_cons1d_batched: bad sizes");
    if (S == 0) return RG_OK;
    if (!x_in || !x_out || !ii || !io || !v || !counts) return fail(RG_E_ARG, "rg_cons1d_batched: null pointer");
    RG_CUDA(cudaSetDevice(device));
    k_cons1d_walk<<<(unsigned)ceil_div(S, 64), 64, 0, (cudaStream_t)stream>>>(S, n, m, x_in, x_out, w_in, ii, io, v, counts);
    RG_LAUNCH_CHECK("k_cons1d_walk");
    return RG_OK;
}

extern "C" int rg_regrid1d_conservative(int device, void* stream, int64_t S, int64_t n, int64_t m,
                                        const double* x_in, const double* x_out, const double* w_in,
                                        const double* values_in, double* values_out)
{
    if (S < 0 || n < 2 || m < 2) return fail(RG_E_ARG, "rg_regrid1d_conservative: bad sizes");
    if (S == 0) return RG_OK;
    if (!x_in || !x_out || !values_in || !values_out) return fail(RG_E_ARG, "rg_regrid1d_conservative: null pointer");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 256;
    int64_t gx = ceil_div(m - 1, T);
    if (gx > 64) gx = 64;
    for (int64_t s0 = 0; s0 < S; s0 += 65535) {
        const int64_t ns = S - s0 < 65535 ? S - s0 : 65535;
        dim3 grid((unsigned)gx, (unsigned)ns);
        k_regrid1d<<<grid, T, 0, st>>>(ns, n, m, x_in + s0 * n, x_out + s0 * m,
                                       w_in ? w_in + s0 * (n - 1) : nullptr,
                                       values_in + s0 * (n - 1), values_out + s0 * (m - 1));
        RG_LAUNCH_CHECK("k_regrid1d");
    }
    return RG_OK;
}
