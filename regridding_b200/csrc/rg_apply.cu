// rg_apply.cu -- shared-weights application (CSR x dense over frames) and the
// COO(input-major) -> CSR(output-major) conversion.
//
// Replaces _regrid_from_weights, regridding/_regrid/_regrid_from_weights.py:165-182:
//     for w in range(nnz): out[io[w]] += v[w] * in[ii[w]]        (per orthogonal slice)
// The public COO is sorted by (input, output), so each output cell accumulates its
// contributions in ascending input index, multiply and add rounded separately (no
// fastmath in the reference).  A CSR row that lists its entries in ascending input
// index and accumulates from +0.0 with __dmul_rn/__dadd_rn reproduces those bits.
#include "rg_common.cuh"

namespace rg {

// ---------------------------------------------------------------------------
// COO -> CSR
// ---------------------------------------------------------------------------

__global__ void k_row_hist(const int64_t* __restrict__ io, int64_t nnz, int32_t* __restrict__ hist)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nnz) atomicAdd(&hist[io[w]], 1);
}

// Scatter the COO positions into the row buckets.  The slot inside a bucket comes
// from an atomic cursor, so the order inside a bucket is arbitrary here; k_row_rank
// below restores the original COO order inside each row (positions are unique, so
// that order is total and the result does not depend on the atomics).  For the
// public layout, sorted by (input, output), COO order inside a row IS ascending
// input index; for any other COO it is still the reference's accumulation order.
__global__ void k_row_fill(const int64_t* __restrict__ io, int64_t nnz,
                           const int32_t* __restrict__ row_ptr, int32_t* __restrict__ cursor,
                           int32_t* __restrict__ tmp_pos)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nnz) return;
    int64_t r = io[w];
    int32_t slot = row_ptr[r] + atomicAdd(&cursor[r], 1);
    tmp_pos[slot] = (int32_t)w;
}

// Rank sort inside each row: position p goes to slot #{q in row : q < p}.
// One lane per row for short rows; rows longer than 32 are ranked by the whole warp.
__global__ void k_row_rank(const int32_t* __restrict__ row_ptr, int64_t n_out,
                           const int32_t* __restrict__ tmp_pos,
                           const int64_t* __restrict__ ii, const double* __restrict__ v,
                           int32_t* __restrict__ col, double* __restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t r = warp * 32 + lane;
    int32_t beg = 0, end = 0;
    if (r < n_out) {
        beg = row_ptr[r];
        end = row_ptr[r + 1];
    }
    const int32_t len = end - beg;
    if (len <= 32) {
        for (int32_t e = beg; e < end; e++) {
            int32_t p = tmp_pos[e];
            int32_t rank = 0;
            for (int32_t f = beg; f < end; f++) rank += (tmp_pos[f] < p);
            col[beg + rank] = (int32_t)ii[p];
            val[beg + rank] = v[p];
        }
    }
    unsigned longmask = __ballot_sync(0xffffffffu, len > 32);
    while (longmask) {
        int src = __ffs(longmask) - 1;
        longmask &= longmask - 1;
        int32_t b = __shfl_sync(0xffffffffu, beg, src);
        int32_t e_ = __shfl_sync(0xffffffffu, end, src);
        for (int32_t e = b + lane; e < e_; e += 32) {
            int32_t p = tmp_pos[e];
            int32_t rank = 0;
            for (int32_t f = b; f < e_; f++) rank += (tmp_pos[f] < p);
            col[b + rank] = (int32_t)ii[p];
            val[b + rank] = v[p];
        }
    }
}

// ---------------------------------------------------------------------------
// apply: one thread per output cell, FT frames in registers.
// Lanes map to consecutive output cells so stores are coalesced; the weights of a
// row are read once per FT frames.
// ---------------------------------------------------------------------------
template <int FT>
__global__ void __launch_bounds__(256)
k_apply_csr(int64_t n_frames, int64_t n_in, int64_t n_out,
            const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
            const double* __restrict__ val,
            const double* __restrict__ vin, double* __restrict__ vout)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t f0 = (int64_t)blockIdx.y * FT;
    if (o >= n_out) return;
    const int32_t beg = row_ptr[o], end = row_ptr[o + 1];
    double acc[FT];
#pragma unroll
    for (int t = 0; t < FT; t++) acc[t] = 0.0;
    const double* in0 = vin + f0 * n_in;
    const int nf = (int)((n_frames - f0) < FT ? (n_frames - f0) : FT);
    if (nf == FT) {
        for (int32_t w = beg; w < end; w++) {
            const int32_t c = col[w];
            const double a = val[w];
#pragma unroll
            for (int t = 0; t < FT; t++) acc[t] = dadd(acc[t], dmul(a, __ldg(in0 + (int64_t)t * n_in + c)));
        }
    } else {
        for (int32_t w = beg; w < end; w++) {
            const int32_t c = col[w];
            const double a = val[w];
#pragma unroll
            for (int t = 0; t < FT; t++)
                if (t < nf) acc[t] = dadd(acc[t], dmul(a, __ldg(in0 + (int64_t)t * n_in + c)));
        }
    }
    double* out0 = vout + f0 * n_out + o;
#pragma unroll
    for (int t = 0; t < FT; t++)
        if (t < nf) out0[(int64_t)t * n_out] = acc[t];
}

}  // namespace rg

using namespace rg;

extern "C" int rg_csr_workspace_bytes(int64_t nnz, int64_t n_out, size_t* bytes_host)
{
    if (!bytes_host || nnz < 0 || n_out < 0) return fail(RG_E_ARG, "rg_csr_workspace_bytes: bad argument");
    Carver c(nullptr);
    c.take<int32_t>((size_t)n_out + 1);               // hist / cursor
    c.take<int32_t>((size_t)nnz + 1);                 // tmp_pos
    c.take<int64_t>(scan_scratch_elems(n_out));       // scan scratch
    *bytes_host = c.total();
    return RG_OK;
}

extern "C" int rg_csr_from_coo(int device, void* stream, int64_t nnz, int64_t n_in, int64_t n_out,
                               const int64_t* ii, const int64_t* io, const double* v,
                               int32_t* row_ptr, int32_t* col, double* val,
                               void* workspace, size_t workspace_bytes)
{
    if (nnz < 0 || n_in <= 0 || n_out <= 0 || !row_ptr) return fail(RG_E_ARG, "rg_csr_from_coo: bad argument");
    if (nnz >= INT32_MAX || n_in >= INT32_MAX || n_out >= INT32_MAX)
        return fail(RG_E_TOO_LARGE, "rg_csr_from_coo: sizes exceed the int32 CSR index range");
    size_t need = 0;
    rg_csr_workspace_bytes(nnz, n_out, &need);
    if (!workspace || workspace_bytes < need) return fail(RG_E_WORKSPACE, "rg_csr_from_coo: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    Carver c(workspace);
    int32_t* hist = c.take<int32_t>((size_t)n_out + 1);
    int32_t* tmp_pos = c.take<int32_t>((size_t)nnz + 1);
    int64_t* scratch = c.take<int64_t>(scan_scratch_elems(n_out));

    RG_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * ((size_t)n_out + 1), st));
    const int T = 256;
    if (nnz > 0) {
        k_row_hist<<<(unsigned)ceil_div(nnz, T), T, 0, st>>>(io, nnz, hist);
        RG_LAUNCH_CHECK("k_row_hist");
    }
    int rc = exclusive_scan_i32_i32(st, hist, row_ptr, n_out, scratch);
    if (rc) return rc;
    if (nnz > 0) {
        RG_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * ((size_t)n_out + 1), st));
        k_row_fill<<<(unsigned)ceil_div(nnz, T), T, 0, st>>>(io, nnz, row_ptr, hist, tmp_pos);
        RG_LAUNCH_CHECK("k_row_fill");
        k_row_rank<<<(unsigned)ceil_div(n_out, T), T, 0, st>>>(row_ptr, n_out, tmp_pos, ii, v, col, val);
        RG_LAUNCH_CHECK("k_row_rank");
    }
    return RG_OK;
}

extern "C" int rg_apply_csr(int device, void* stream, int64_t n_frames, int64_t n_in, int64_t n_out,
                            const int32_t* row_ptr, const int32_t* col, const double* val,
                            const double* values_in, double* values_out)
{
    if (n_frames == 0) return RG_OK;  // nothing to do (empty tensors have null data pointers)
    if (n_frames < 0 || n_in <= 0 || n_out <= 0 || !row_ptr || !values_in || !values_out)
        return fail(RG_E_ARG, "rg_apply_csr: bad argument");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int FT = 8;
    const int T = 256;
    const int64_t fy = ceil_div(n_frames, FT);
    if (fy > 65535) {
        // split the frame range so that gridDim.y stays legal
        const int64_t chunk = 65535LL * FT;
        for (int64_t f = 0; f < n_frames; f += chunk) {
            int64_t nf = n_frames - f < chunk ? n_frames - f : chunk;
            dim3 grid((unsigned)ceil_div(n_out, T), (unsigned)ceil_div(nf, FT));
            k_apply_csr<FT><<<grid, T, 0, st>>>(nf, n_in, n_out, row_ptr, col, val,
                                                values_in + f * n_in, values_out + f * n_out);
            RG_LAUNCH_CHECK("k_apply_csr");
        }
        return RG_OK;
    }
    dim3 grid((unsigned)ceil_div(n_out, T), (unsigned)fy);
    k_apply_csr<FT><<<grid, T, 0, st>>>(n_frames, n_in, n_out, row_ptr, col, val, values_in, values_out);
    RG_LAUNCH_CHECK("k_apply_csr");
    return RG_OK;
}
