// rg_cons1d.cu -- 1D first-order conservative resampling, batched over S spectra.
//
// Replaces _weights_conservative_1d and its per-spectrum Python driver
// (regridding/_weights/_weights_conservative_1d/_weights_conservative_1d.py:12-56, 60-318;
//  regridding/_weights/_weights_conservative.py:59-106).  Not fastmath in the reference
// (c1d.py:59), so plain IEEE subtraction / division reproduces it bit for bit.
#include "rg_common.cuh"
#include "rg_async.cuh"

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

// ---------------------------------------------------------------------------
// fused regrid, streamed (the config-2 path: 1M spectra x 4096 bins, no input weights).
//
// One persistent CTA per SM, two stages of shared memory.  Thread 0 brings the three rows of the NEXT spectrum (sweep
// edges, static edges, values: 98 KB at 4096 bins = the algorithmic minimum of HBM traffic) with three bulk copies
// (cp.async.bulk, the TMA engine; completion on an mbarrier with expect-tx) while all warps compute the CURRENT one,
// so loads stay in flight during the whole compute phase and no thread spends instructions on staging.  Rows of odd
// length start 8 bytes off every other spectrum: the copy then starts one double earlier ("skew") and the readers
// add the skew; the first / last spectrum of a call, whose over-read could leave the caller's buffers, are staged by
// plain loads instead.
//
// k_regrid1d_staged (below) was issue bound (ncu: 71 % of the issue slots, 41 k warp instructions per spectrum:
// 35 % searches, 40 % pieces, 15 % staging).  Here a thread takes one static cell: the sweep cell of its left edge
// comes from a linear-interpolation guess corrected by neighbour steps whose exit conditions DEFINE the cell (lanes
// that run out of steps take the plain binary search); the right edge needs no search at all, the piece loop runs
// while the sweep cell's left edge is below it and carries every right edge over as the next left edge.  Same
// pieces, same operations in the same order as k_regrid1d: bit-identical.
// ---------------------------------------------------------------------------
constexpr int kStreamThreads = 1024;

struct Stream1DStage {
    double* sw;   // staged rows: element q of the row lives at [skew + q]
    double* st;
    double* vi;
};

template <bool GEN>
__device__ __forceinline__ void regrid1d_stream_cells(const double* __restrict__ sw_raw, const double* __restrict__ st_raw,
                                                      const double* __restrict__ vi, int n, int m, bool rev_sw, bool rev_st,
                                                      double* __restrict__ vo, int iters)
{
    // ascending views (c1d.py:100-110); GEN = false: both rows ascend, the views are the rows
    auto SW = [&](int k) -> double { return GEN ? sw_raw[rev_sw ? n - 1 - k : k] : sw_raw[k]; };
    auto ST = [&](int e) -> double { return GEN ? st_raw[rev_st ? m - 1 - e : e] : st_raw[e]; };
    const int ncell = n - 1;
    const double sw0 = SW(0);
    const float scale = (float)((double)(n - 1) / (SW(n - 1) - sw0));
    // K(a) = first sweep cell in [0, n-1] whose right edge is beyond a (n-1: none): a linear-interpolation guess moved
    // down while the cell's left edge is beyond a, then up while its right edge is not; a loop that ends by its
    // condition leaves (K == 0 || SW(K) <= a) && (K == n-1 || SW(K+1) > a), which defines K for sorted edges.  Lanes
    // that run out of steps (grids far from uniform) take the plain binary search.
    auto locate = [&](double a) -> int {
        int c = __float2int_rd(fminf(fmaxf((float)(a - sw0) * scale, 0.0f), (float)(n - 1)));
        int t = 0;
#pragma unroll 1
        for (; t < 4 && c > 0 && SW(c) > a; t++) c--;
        bool ok = t < 4;
#pragma unroll 1
        for (t = 0; t < 4 && c < n - 1 && SW(c + 1) <= a; t++) c++;
        ok = ok && t < 4;
        if (!ok) {
            int lo = 0, hi = n - 1;
            for (int it = 0; it < iters && lo < hi; it++) {
                const int mid = (lo + hi) >> 1;
                if (SW(mid + 1) > a) hi = mid; else lo = mid + 1;
            }
            c = lo;
        }
        return c;
    };
    for (int cell = threadIdx.x; cell < m - 1; cell += blockDim.x) {
        const double a = ST(cell), b = ST(cell + 1);
        const int K = locate(a);
        double acc = 0.0;
        if (!GEN) {
            // cells K, K+1, ... while their left edge is below b (a cell that only touches b has no piece)
            double l = sw_raw[K];
#pragma unroll 1
            for (int k = K; k <= n - 2 && l < b; k++) {
                const double rr = sw_raw[k + 1];
                const double p1 = l > a ? l : a;
                const double p2 = rr < b ? rr : b;
                if (p1 < p2) {
                    const double ratio = ddiv(dsub(p2, p1), dsub(rr, l));
                    acc = dadd(acc, dmul(ratio, vi[k]));
                }
                l = rr;
            }
            vo[cell] = acc;
        } else {
            const int Kb = locate(b);
            const int k0 = K;
            const int k1 = min(Kb - (SW(Kb) == b ? 1 : 0), n - 2);
            const int cnt = k1 - k0 + 1;
            for (int q = 0; q < cnt; q++) {
                const int k = rev_sw ? (k1 - q) : (k0 + q);  // ascending wrapped input index
                const double l = SW(k), rr = SW(k + 1);
                const double p1 = l > a ? l : a;
                const double p2 = rr < b ? rr : b;
                if (!(p1 < p2)) continue;
                const int li = rev_sw ? (ncell - 1 - k) : k;  // wrapped index (= the reference's ~k + ncell)
                const double length = dsub(SW(li + 1), SW(li));   // c1d.py:118, 305-307 (indexed in the view)
                const double ratio = ddiv(dsub(p2, p1), length);
                acc = dadd(acc, dmul(ratio, vi[li]));
            }
            vo[rev_st ? (m - 2 - cell) : cell] = acc;
        }
    }
}

__global__ void __launch_bounds__(kStreamThreads, 1)
k_regrid1d_stream(int64_t S, int n, int m, int cap_n, int cap_m,
                  const double* __restrict__ x_in, const double* __restrict__ x_out,
                  const double* __restrict__ vin, double* __restrict__ vout)
{
    extern __shared__ __align__(16) double smem_d[];
    __shared__ __align__(8) uint64_t full[2];
    // stage layout: [cap_n] sweep edges, [cap_m] static edges, [cap_n] values (capacities even, >= length + 2)
    const int stage_doubles = 2 * cap_n + cap_m;
    auto stage = [&](int s) -> Stream1DStage {
        double* b = smem_d + (size_t)s * stage_doubles;
        return Stream1DStage{ b, b + cap_n, b + cap_n + cap_m };
    };
    auto skew_of = [](const double* p) -> int { return (int)(((uintptr_t)p >> 3) & 1); };
    // a spectrum may be bulk-copied when the 16-byte aligned copies stay inside the caller's buffers
    auto bulk_ok = [&](int64_t sp) -> bool {
        const double* r0 = x_in + sp * n;
        const double* r1 = x_out + sp * m;
        const double* r2 = vin + sp * (n - 1);
        const int s0 = skew_of(r0), s1 = skew_of(r1), s2 = skew_of(r2);
        const bool front = sp > 0 || (s0 | s1 | s2) == 0;
        const bool back = sp < S - 1 || (((s0 + n) | (s1 + m) | (s2 + n - 1)) & 1) == 0;
        return front && back;
    };
    auto issue = [&](int64_t sp, int s) {   // thread 0 only
        const Stream1DStage g = stage(s);
        const double* r0 = x_in + sp * n;
        const double* r1 = x_out + sp * m;
        const double* r2 = vin + sp * (n - 1);
        const int s0 = skew_of(r0), s1 = skew_of(r1), s2 = skew_of(r2);
        const unsigned b0 = (unsigned)(((s0 + n) * 8 + 15) & ~15), b1 = (unsigned)(((s1 + m) * 8 + 15) & ~15),
                       b2 = (unsigned)(((s2 + n - 1) * 8 + 15) & ~15);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the stage was last touched by ordinary loads / stores
        mbar_arrive_expect_tx(&full[s], b0 + b1 + b2);
        bulk_load(smem_u32(g.sw), r0 - s0, b0, &full[s]);
        bulk_load(smem_u32(g.st), r1 - s1, b1, &full[s]);
        bulk_load(smem_u32(g.vi), r2 - s2, b2, &full[s]);
    };
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    int iters = 0;
    while ((1 << iters) < n) iters++;
    unsigned phase[2] = { 0u, 0u };
    int64_t sp = blockIdx.x;
    if (threadIdx.x == 0 && sp < S && bulk_ok(sp)) issue(sp, 0);
    for (int it = 0; sp < S; sp += gridDim.x, it++) {
        const int s = it & 1;
        const int64_t nxt = sp + gridDim.x;
        // (stage s^1 was released by the __syncthreads that ended the previous iteration)
        if (threadIdx.x == 0 && nxt < S && bulk_ok(nxt)) issue(nxt, s ^ 1);
        const Stream1DStage g = stage(s);
        const double* r0 = x_in + sp * n;
        const double* r1 = x_out + sp * m;
        const double* r2 = vin + sp * (n - 1);
        int s0 = skew_of(r0), s1 = skew_of(r1), s2 = skew_of(r2);
        if (bulk_ok(sp)) {
            mbar_wait(&full[s], phase[s]);
            phase[s] ^= 1u;
        } else {   // first / last spectrum of the call
            s0 = s1 = s2 = 0;
            for (int q = threadIdx.x; q < n; q += blockDim.x) g.sw[q] = r0[q];
            for (int q = threadIdx.x; q < m; q += blockDim.x) g.st[q] = r1[q];
            for (int q = threadIdx.x; q < n - 1; q += blockDim.x) g.vi[q] = r2[q];
            __syncthreads();
        }
        const double* sw = g.sw + s0;
        const double* st = g.st + s1;
        const double* vi = g.vi + s2;
        const bool rev_sw = !(sw[0] < sw[n - 1]);  // c1d.py:100-110
        const bool rev_st = !(st[0] < st[m - 1]);
        double* vo = vout + sp * (m - 1);
        if (!rev_sw && !rev_st) regrid1d_stream_cells<false>(sw, st, vi, n, m, false, false, vo, iters);
        else regrid1d_stream_cells<true>(sw, st, vi, n, m, rev_sw, rev_st, vo, iters);
        __syncthreads();   // every warp is done with stage s before it is refilled
    }
}

// ---------------------------------------------------------------------------
// fused regrid, shared-memory staged (the config-2 path: 1M spectra x 4096 bins).
// One CTA per spectrum: the ascending views of both edge arrays and the spectrum's values are
// staged in shared memory with coalesced loads (131 kB of HBM traffic per spectrum = the
// algorithmic minimum), then a warp takes 64 consecutive output EDGES: every lane locates its two
// edges in the sweep grid with ONE binary search each in shared memory (K = first sweep cell whose
// right edge lies beyond it), the neighbouring lane's K bounds the cell's input range, and the
// accumulation is the same as k_regrid1d's (ascending wrapped input index, the reference's
// rounding sequence), so the result is bit-identical.
// ---------------------------------------------------------------------------
template <bool HAS_W>
__global__ void __launch_bounds__(512, 2)
k_regrid1d_staged(int64_t S, int n, int m,
                  const double* __restrict__ x_in, const double* __restrict__ x_out,
                  const double* __restrict__ w_in,
                  const double* __restrict__ vin, double* __restrict__ vout)
{
    extern __shared__ __align__(16) double smem_d[];
    double* sw = smem_d;           // [n]   sweep (input) edges, ascending view
    double* st = sw + n;           // [m]   static (output) edges, ascending view
    double* vi = st + m;           // [n-1] input values, original order
    double* ws = vi + (n - 1);     // [n-1] input weights, original order (HAS_W)
    const int64_t sp = blockIdx.x;
    const double* xi = x_in + sp * n;
    const double* xo = x_out + sp * m;
    const bool rev_sw = !(xi[0] < xi[n - 1]);  // c1d.py:100-110
    const bool rev_st = !(xo[0] < xo[m - 1]);
    // asynchronous 8-byte copies (all in flight at once; rows of odd length are only 8-byte aligned)
    auto cp8 = [](double* dst, const double* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
    };
    for (int q = threadIdx.x; q < n; q += blockDim.x) cp8(sw + q, rev_sw ? xi + (n - 1 - q) : xi + q);
    for (int q = threadIdx.x; q < m; q += blockDim.x) cp8(st + q, rev_st ? xo + (m - 1 - q) : xo + q);
    {
        const double* v = vin + sp * (n - 1);
        for (int q = threadIdx.x; q < n - 1; q += blockDim.x) cp8(vi + q, v + q);
        if (HAS_W) {
            const double* w = w_in + sp * (n - 1);
            for (int q = threadIdx.x; q < n - 1; q += blockDim.x) cp8(ws + q, w + q);
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    double* vo = vout + sp * (m - 1);
    const int ncell = n - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    int iters = 0;
    while ((1 << iters) < n) iters++;  // binary-search steps that always suffice for [0, n-1]
    constexpr int kBracket = 7, kBracketSteps = 4;  // a verified bracket holds <= 15 candidates: 4 steps
    const double sw0 = sw[0];
    const double scale = (double)(n - 1) / (sw[n - 1] - sw0);
    // one view cell (edges a, b) of the static grid, input cells k0 .. k1 may overlap it
    auto cell = [&](int ecell, double a, double b, int Ka, int Kb) {
        // k0 = first cell whose right edge is beyond a, k1 = last cell whose left edge is below b (= Kb, or Kb - 1
        // when b coincides with a sweep edge; a too large k1 only adds pieces that the p1 < p2 test rejects)
        const int k0 = Ka;
        const int k1 = min(Kb - (sw[Kb] == b ? 1 : 0), n - 2);
        double acc = 0.0;
        const int cnt = k1 - k0 + 1;
        for (int q = 0; q < cnt; q++) {
            const int k = rev_sw ? (k1 - q) : (k0 + q);  // ascending wrapped input index
            const double l = sw[k], r = sw[k + 1];
            const double p1 = l > a ? l : a;
            const double p2 = r < b ? r : b;
            if (!(p1 < p2)) continue;
            const int li = rev_sw ? (ncell - 1 - k) : k;  // wrapped index (= the reference's ~k + ncell)
            const double length = dsub(sw[li + 1], sw[li]);
            double ratio = ddiv(dsub(p2, p1), length);
            if (HAS_W) ratio = dmul(ratio, ws[li]);
            acc = dadd(acc, dmul(ratio, vi[li]));
        }
        vo[rev_st ? (m - 2 - ecell) : ecell] = acc;
    };
    // a warp takes 64 consecutive view edges e0 .. e0+63 = 63 view cells; a lane locates edges e0+lane and
    // e0+32+lane with two interleaved binary searches (K = first sweep cell in [0, n-1] whose right edge is beyond
    // the edge; n-1: none)
    for (int e0 = warp * 63; e0 < m - 1; e0 += nwarps * 63) {
        const double a1 = st[min(e0 + lane, m - 1)], a2 = st[min(e0 + 32 + lane, m - 1)];
        // bracket from linear interpolation between the end edges (spectral grids are close to uniform); a bracket
        // that does not verify (K >= lo: sw[lo] <= a;  K <= hi: sw[hi+1] > a) falls back to the whole range
        int lo1, hi1, lo2, hi2;
        {
            const double g1 = (a1 - sw0) * scale, g2 = (a2 - sw0) * scale;
            const int c1 = (int)fmin(fmax(g1, 0.0), (double)(n - 1)), c2 = (int)fmin(fmax(g2, 0.0), (double)(n - 1));
            lo1 = max(c1 - kBracket, 0); hi1 = min(c1 + kBracket, n - 1);
            lo2 = max(c2 - kBracket, 0); hi2 = min(c2 + kBracket, n - 1);
            if (!((lo1 == 0 || sw[lo1] <= a1) && (hi1 == n - 1 || sw[hi1 + 1] > a1))) { lo1 = 0; hi1 = n - 1; }
            if (!((lo2 == 0 || sw[lo2] <= a2) && (hi2 == n - 1 || sw[hi2 + 1] > a2))) { lo2 = 0; hi2 = n - 1; }
        }
        const int span = max(hi1 - lo1, hi2 - lo2);
        const int steps = __reduce_max_sync(0xffffffffu, span > 2 * kBracket ? iters : kBracketSteps);
        for (int it = 0; it < steps; it++) {
            const int mid1 = (lo1 + hi1) >> 1, mid2 = (lo2 + hi2) >> 1;
            const bool c1 = sw[mid1 + 1] > a1, c2 = sw[mid2 + 1] > a2;  // (mid = n-1 reads one element past sw: unused)
            if (lo1 < hi1) { if (c1) hi1 = mid1; else lo1 = mid1 + 1; }
            if (lo2 < hi2) { if (c2) hi2 = mid2; else lo2 = mid2 + 1; }
        }
        const int K1 = lo1, K2 = lo2;
        // right edge of my two cells: the next lane's edge (lane 31 of the first group: lane 0 of the second)
        int Kb1 = __shfl_down_sync(0xffffffffu, K1, 1);
        double b1 = __shfl_down_sync(0xffffffffu, a1, 1);
        const int Kw = __shfl_sync(0xffffffffu, K2, 0);
        const double bw = __shfl_sync(0xffffffffu, a2, 0);
        if (lane == 31) { Kb1 = Kw; b1 = bw; }
        const int Kb2 = __shfl_down_sync(0xffffffffu, K2, 1);
        const double b2 = __shfl_down_sync(0xffffffffu, a2, 1);
        if (e0 + lane < m - 1) cell(e0 + lane, a1, b1, K1, Kb1);
        if (lane < 31 && e0 + 32 + lane < m - 1) cell(e0 + 32 + lane, a2, b2, K2, Kb2);
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
    // shared-memory staged kernel whenever one spectrum (both edge arrays + values [+ weights]) fits on chip
    // streamed kernel (persistent CTAs, two stages filled by bulk copies) when two spectra fit on chip
    if (!w_in && n < (1 << 24) && m < (1 << 24)) {
        const int cap_n = (int)((n + 2 + 1) & ~(int64_t)1), cap_m = (int)((m + 2 + 1) & ~(int64_t)1);
        const size_t smem2 = (size_t)2 * (2 * cap_n + cap_m) * sizeof(double);
        if (smem2 <= 227 * 1024 && !getenv("RG_NO_STREAM1D")) {
            RG_CUDA(cudaFuncSetAttribute(k_regrid1d_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            const unsigned grid = (unsigned)(S < kNumSM ? S : kNumSM);
            k_regrid1d_stream<<<grid, kStreamThreads, smem2, st>>>(S, (int)n, (int)m, cap_n, cap_m, x_in, x_out, values_in, values_out);
            RG_LAUNCH_CHECK("k_regrid1d_stream");
            return RG_OK;
        }
    }
    const size_t smem = (size_t)(n + m + (n - 1) + (w_in ? n - 1 : 0)) * sizeof(double);
    if (smem <= 200 * 1024 && n < (1 << 30) && m < (1 << 30)) {
        auto kern = w_in ? k_regrid1d_staged<true> : k_regrid1d_staged<false>;
        RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t chunk = 1 << 30;
        for (int64_t s0 = 0; s0 < S; s0 += chunk) {
            const int64_t ns = S - s0 < chunk ? S - s0 : chunk;
            kern<<<(unsigned)ns, 512, smem, st>>>(ns, (int)n, (int)m, x_in + s0 * n, x_out + s0 * m,
                                                w_in ? w_in + s0 * (n - 1) : nullptr,
                                                values_in + s0 * (n - 1), values_out + s0 * (m - 1));
            RG_LAUNCH_CHECK("k_regrid1d_staged");
        }
        return RG_OK;
    }
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
