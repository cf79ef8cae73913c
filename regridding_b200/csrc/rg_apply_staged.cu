// rg_apply_staged.cu -- shared-weights apply, shared-memory staged (the HBM-roofline path).
//
// Same arithmetic as k_apply_csr (rg_apply.cu): each output cell accumulates
// val * in[col] over its CSR row in ascending input index from +0.0 with separately
// rounded multiply and add (regridding/_regrid/_regrid_from_weights.py:179-182), so the
// result is bit-identical; only the data movement differs.
//
// Why: with one thread per output cell the gathers in[f][col] of a warp touch ~12 cache
// lines per instruction and the kernel is bound by L1 wavefronts (0.25 of HBM peak
// measured).  Here a CTA owns a TILE of TH x TW output cells and walks FB frames in
// sub-blocks of T = 8 frames:
//   * the input cells the tile references (its FOOTPRINT: per input row one contiguous
//     span, precomputed once per weights by rg_apply_plan_build) are copied for 8 frames
//     with cp.async into shared memory as in_s[cell][frame] (row stride 9 doubles), double
//     buffered so the next sub-block streams in from HBM while this one is computed;
//   * a quarter-warp owns one output cell, LANE = FRAME: every gather in_s[lidx][lane] is a
//     contiguous 64-byte shared-memory read, the CSR entry (local cell
//     index u16 + weight) is a broadcast read, and there is no divergence inside a cell;
//   * results are staged in out_s[frame][cell] (odd stride) and written with full 256-byte
//     coalesced rows.
// The tile-local CSR (weights + u16 local indices) is loaded once per CTA and reused for
// all FB frames, so weights traffic is nnz * 10 B per FB frames.
// Tiles whose footprint does not fit (very different resolutions, scattered weights) are
// flagged by the plan and handled by the generic per-cell kernel.
#include "rg_common.cuh"

namespace rg {

constexpr int kTH = 8;           // tile height (output rows)
constexpr int kTW = 32;          // tile width  (output cols) = one full coalesced row of 256 B
constexpr int kTileCells = kTH * kTW;
constexpr int kT = 16;           // frames per sub-block: 8 lanes x 2 frames per lane
constexpr int kFB = 256;         // frames per CTA (the tile-local CSR is reread every kFB frames)
constexpr int kRMAX = 64;        // max input rows in a footprint
constexpr int kCP = 770;         // staged cells per frame (capacity); kCP/2 odd => the 8 frame lanes of a
                                 // quarter-warp hit 8 distinct 16-byte bank groups
constexpr int kCellsMax = kCP;
constexpr int kNnzMax = 2560;    // max CSR entries per tile
constexpr int kOutStride = kTileCells + 2;  // doubles per staged output frame (= 2 mod 16)
constexpr int kStagedThreads = 1024;
constexpr int kPatch = 12;       // tiles are issued in 12 x 12 patches (~ one wave of 148 CTAs) so that
                                 // footprint halos are shared through L2
static_assert((kCP / 2) % 2 == 1 && kCP % 2 == 0, "kCP/2 must be odd");
static_assert(kT * kOutStride <= kT * kCP, "output staging aliases one input buffer");

constexpr int kTileInfoInts = 4;  // r0, nrows, cells, nnz (nnz < 0: tile handled by the generic kernel)

struct StagedSmem {
    double in_s[2][kT * kCP];     // [buffer][frame][cell]; the consumed buffer doubles as out_s[frame][kOutStride]
    double val[kNnzMax];
    uint16_t lidx[kNnzMax];       // BYTE offset of the referenced cell inside a staged frame
    uint16_t rowptr[kTileCells + 2];
    int32_t row_src[kRMAX];       // per footprint row: offset of its span inside one input frame (doubles)
    int32_t row_off[kRMAX];       //                    offset of its span inside one staged frame
    int32_t row_len[kRMAX];       //                    span length
    alignas(8) uint64_t full[2];  // mbarriers: "buffer filled"
};

__host__ __device__ inline void tile_of_block(int64_t b, int tiles_x, int tiles_y, int& ty, int& tx)
{
    // patch-major order; patches and the tiles inside a patch are row-major
    const int px_count = (tiles_x + kPatch - 1) / kPatch;
    const int64_t full_rows = tiles_y / kPatch;                       // complete patch rows
    const int64_t per_patch_row = (int64_t)kPatch * tiles_x;           // tiles in a complete patch row
    int prow, ph;
    int64_t rem;
    if (b < full_rows * per_patch_row) {
        prow = (int)(b / per_patch_row);
        rem = b - (int64_t)prow * per_patch_row;
        ph = kPatch;
    } else {
        prow = (int)full_rows;
        rem = b - full_rows * per_patch_row;
        ph = tiles_y - prow * kPatch;
    }
    // inside a patch row: patches of width kPatch (last one narrower), each ph x pw tiles
    const int64_t per_full_patch = (int64_t)ph * kPatch;
    int pcol = (int)(rem / per_full_patch);
    if (pcol >= px_count) pcol = px_count - 1;
    const int64_t rem2 = rem - (int64_t)pcol * per_full_patch;
    const int pw = min(kPatch, tiles_x - pcol * kPatch);
    ty = prow * kPatch + (int)(rem2 / pw);
    tx = pcol * kPatch + (int)(rem2 % pw);
}

// ---------------------------------------------------------------------------
// plan: footprint of every tile + tile-local indices
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_plan_tiles(int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out, int tiles_x, int pad_even,
             const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
             int32_t* __restrict__ tile_info, int32_t* __restrict__ tile_rows, uint16_t* __restrict__ lidx,
             int32_t* __restrict__ n_generic)
{
    __shared__ int s_rmin, s_rmax, s_nnz;
    __shared__ int s_clo[kRMAX], s_chi[kRMAX], s_off[kRMAX + 1];
    const int tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile % tiles_x;
    const int th = (int)min((int64_t)kTH, h_out - (int64_t)ty * kTH);
    const int tw = (int)min((int64_t)kTW, w_out - (int64_t)tx * kTW);
    if (threadIdx.x == 0) { s_rmin = INT32_MAX; s_rmax = -1; s_nnz = 0; }
    for (int r = threadIdx.x; r < kRMAX; r += blockDim.x) { s_clo[r] = INT32_MAX; s_chi[r] = -1; }
    __syncthreads();
    // pass 1: input row range and entry count
    int lmin = INT32_MAX, lmax = -1, lcnt = 0;
    for (int tr = 0; tr < th; tr++) {
        const int64_t o0 = ((int64_t)ty * kTH + tr) * w_out + (int64_t)tx * kTW;
        const int32_t b = row_ptr[o0], e = row_ptr[o0 + tw];
        for (int32_t w = b + threadIdx.x; w < e; w += blockDim.x) {
            const int ci = (int)(col[w] / w_in);
            lmin = min(lmin, ci);
            lmax = max(lmax, ci);
            lcnt++;
        }
    }
    if (lcnt) { atomicMin(&s_rmin, lmin); atomicMax(&s_rmax, lmax); atomicAdd(&s_nnz, lcnt); }
    __syncthreads();
    const int rmin = s_rmin, nnz = s_nnz;
    const int nrows = nnz ? s_rmax - rmin + 1 : 0;
    bool generic = nrows > kRMAX || nnz > kNnzMax;
    if (!generic && nnz) {
        // pass 2: column span of every input row
        for (int tr = 0; tr < th; tr++) {
            const int64_t o0 = ((int64_t)ty * kTH + tr) * w_out + (int64_t)tx * kTW;
            const int32_t b = row_ptr[o0], e = row_ptr[o0 + tw];
            for (int32_t w = b + threadIdx.x; w < e; w += blockDim.x) {
                const int c = col[w];
                const int ci = (int)(c / w_in), cj = (int)(c - (int64_t)ci * w_in);
                atomicMin(&s_clo[ci - rmin], cj);
                atomicMax(&s_chi[ci - rmin], cj);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int off = 0;
            for (int r = 0; r < nrows; r++) {
                s_off[r] = off;
                if (s_chi[r] >= s_clo[r]) {
                    if (pad_even) {  // spans start on even columns and have even length: 16-byte copies
                        s_clo[r] &= ~1;
                        s_chi[r] |= 1;
                    }
                    off += s_chi[r] - s_clo[r] + 1;
                } else {
                    s_clo[r] = 0;
                }
            }
            s_off[nrows] = off;
        }
        __syncthreads();
        if (s_off[nrows] > kCellsMax) generic = true;
    }
    if (!generic && nnz) {
        // pass 3: tile-local cell index of every entry
        for (int tr = 0; tr < th; tr++) {
            const int64_t o0 = ((int64_t)ty * kTH + tr) * w_out + (int64_t)tx * kTW;
            const int32_t b = row_ptr[o0], e = row_ptr[o0 + tw];
            for (int32_t w = b + threadIdx.x; w < e; w += blockDim.x) {
                const int c = col[w];
                const int ci = (int)(c / w_in), cj = (int)(c - (int64_t)ci * w_in);
                lidx[w] = (uint16_t)(s_off[ci - rmin] + cj - s_clo[ci - rmin]);
            }
        }
        for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
            tile_rows[((int64_t)tile * kRMAX + r) * 2 + 0] = s_clo[r];
            tile_rows[((int64_t)tile * kRMAX + r) * 2 + 1] = s_off[r];
        }
    }
    if (threadIdx.x == 0) {
        int32_t* info = tile_info + (int64_t)tile * kTileInfoInts;
        info[0] = nnz ? rmin : 0;
        info[1] = generic ? 0 : nrows;
        info[2] = (generic || !nnz) ? 0 : s_off[nrows];
        info[3] = generic ? -1 : nnz;
        if (generic) atomicAdd(n_generic, 1);
    }
    (void)h_in;
}

// ---------------------------------------------------------------------------
// staged apply
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// mbarrier + bulk (TMA engine) copies: one instruction moves a whole footprint row span
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// WIDE: spans are even-aligned and the frame pitch is even, so every span is a 16-byte aligned,
// 16-byte multiple run and moves as ONE bulk copy; otherwise 8-byte cp.async per element.
template <bool WIDE>
__global__ void __launch_bounds__(kStagedThreads, 1)
k_apply_staged(int64_t n_frames, int64_t w_in, int64_t n_in, int64_t h_out, int64_t w_out, int tiles_x, int tiles_y,
               const int32_t* __restrict__ row_ptr, const double* __restrict__ val,
               const int32_t* __restrict__ tile_info, const int32_t* __restrict__ tile_rows,
               const uint16_t* __restrict__ lidx,
               const double* __restrict__ vin, double* __restrict__ vout)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StagedSmem& S = *reinterpret_cast<StagedSmem*>(smem_raw);
    int ty, tx;
    tile_of_block(blockIdx.x, tiles_x, tiles_y, ty, tx);
    const int tile = ty * tiles_x + tx;
    const int32_t* info = tile_info + (int64_t)tile * kTileInfoInts;
    const int tile_nnz = info[3];
    if (tile_nnz < 0) return;  // handled by the generic kernel
    const int r0 = info[0], nrows = info[1], cells = info[2];
    const int th = (int)min((int64_t)kTH, h_out - (int64_t)ty * kTH);
    const int tw = (int)min((int64_t)kTW, w_out - (int64_t)tx * kTW);
    const int64_t n_out = h_out * w_out;
    const int64_t f_begin = (int64_t)blockIdx.y * kFB;
    const int64_t f_end = min(n_frames, f_begin + kFB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = kStagedThreads / 32;
    const int64_t out_base = ((int64_t)ty * kTH) * w_out + (int64_t)tx * kTW;

    if (tile_nnz == 0) {
        // no weights reach this tile: the reference leaves zeros (rfw.py:111-118)
        if (lane < tw) {
            for (int64_t f = f_begin + warp; f < f_end; f += NW) {
                double* o = vout + f * n_out + out_base + lane;
                for (int tr = 0; tr < th; tr++) o[(int64_t)tr * w_out] = 0.0;
            }
        }
        return;
    }

    // ---- tile-local CSR and footprint table: loaded once, reused for every frame of this CTA ----
    {
        int base = 0;
        for (int tr = 0; tr < kTH; tr++) {
            if (tr < th) {
                const int64_t o0 = out_base + (int64_t)tr * w_out;
                const int32_t b = row_ptr[o0], e = row_ptr[o0 + tw];
                if (threadIdx.x < kTW)
                    S.rowptr[tr * kTW + threadIdx.x] = (uint16_t)(base + (row_ptr[o0 + min((int)threadIdx.x, tw)] - b));
                for (int32_t w = b + threadIdx.x; w < e; w += kStagedThreads) {
                    S.val[base + (w - b)] = val[w];
                    S.lidx[base + (w - b)] = (uint16_t)(lidx[w] * 8u);  // byte offset inside a staged frame
                }
                base += e - b;
            } else if (threadIdx.x < kTW) {
                S.rowptr[tr * kTW + threadIdx.x] = (uint16_t)base;
            }
        }
        if (threadIdx.x == 0) S.rowptr[kTileCells] = (uint16_t)base;
        for (int r = threadIdx.x; r < nrows; r += kStagedThreads) {
            const int clo = tile_rows[((int64_t)tile * kRMAX + r) * 2 + 0];
            const int off = tile_rows[((int64_t)tile * kRMAX + r) * 2 + 1];
            const int end = (r + 1 < nrows) ? tile_rows[((int64_t)tile * kRMAX + r + 1) * 2 + 1] : cells;
            S.row_src[r] = (int32_t)((int64_t)(r0 + r) * w_in + clo);
            S.row_off[r] = off;
            S.row_len[r] = end - off;
        }
        if (threadIdx.x == 0) {
            mbar_init(&S.full[0], kT);
            mbar_init(&S.full[1], kT);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
    }
    __syncthreads();

    // ---- footprint copy of one 16-frame sub-block: warp w (< 16) moves frame w, lane r moves input row r ----
    auto prefetch = [&](int64_t f0, int buf) {
        if (warp >= kT) return;
        const int64_t f = f0 + warp;
        double* dst = S.in_s[buf] + warp * kCP;
        if (WIDE) {
            const bool valid = f < f_end;
            if (lane == 0) mbar_arrive_expect_tx(&S.full[buf], valid ? (unsigned)cells * 8u : 0u);
            __syncwarp();
            if (valid) {
                const double* src = vin + f * n_in;
                for (int r = lane; r < nrows; r += 32) {
                    const int len = S.row_len[r];
                    if (len > 0) bulk_g2s(dst + S.row_off[r], src + S.row_src[r], (unsigned)len * 8u, &S.full[buf]);
                }
            }
        } else {
            if (f < f_end) {
                const double* src = vin + f * n_in;
                for (int r = 0; r < nrows; r++) {
                    const double* rs = src + S.row_src[r];
                    double* rd = dst + S.row_off[r];
                    for (int c = lane; c < S.row_len[r]; c += 32) cp_async<8>(rd + c, rs + c);
                }
            }
        }
    };

    const int nsub = (int)((f_end - f_begin + kT - 1) / kT);
    prefetch(f_begin, 0);
    if (!WIDE) cp_async_commit();
    if (nsub > 1) prefetch(f_begin + kT, 1);
    if (!WIDE) cp_async_commit();
    const int q = lane >> 3, t = lane & 7;  // quarter-warp = one output cell; lane owns frames t and t + 8
    for (int s = 0; s < nsub; s++) {
        const int64_t f0 = f_begin + (int64_t)s * kT;
        const int buf = s & 1;
        if (WIDE) {
            mbar_wait(&S.full[buf], (unsigned)((s >> 1) & 1));
        } else {
            cp_async_wait<1>();
            __syncthreads();
        }
        double* in = S.in_s[buf];
        const char* in0 = reinterpret_cast<const char*>(in + t * kCP);
        // ---- compute: quarter-warp per output cell; the trip count is made warp-uniform ----
        double acc[2][2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int o_local = 4 * (warp + k * NW) + q;
            const int beg = S.rowptr[o_local], n = (int)S.rowptr[o_local + 1] - beg;
            int nmax = n;
            nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
            nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
            const int last = max(beg + n - 1, 0);
            double a0 = 0.0, a1 = 0.0;
#pragma unroll 2
            for (int w = 0; w < nmax; w++) {
                const int wi = min(beg + w, last);
                const unsigned lo = S.lidx[wi];
                const double v = S.val[wi];
                const double x0 = *reinterpret_cast<const double*>(in0 + lo);
                const double x1 = *reinterpret_cast<const double*>(in0 + lo + 8 * kCP * 8);
                const double p0 = dmul(v, x0), p1 = dmul(v, x1);
                if (w < n) {
                    a0 = dadd(a0, p0);
                    a1 = dadd(a1, p1);
                }
            }
            acc[k][0] = a0;
            acc[k][1] = a1;
        }
        __syncthreads();  // everyone is done reading in_s[buf]: reuse it as out_s[frame][cell]
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int o_local = 4 * (warp + k * NW) + q;
            in[t * kOutStride + o_local] = acc[k][0];
            in[(t + 8) * kOutStride + o_local] = acc[k][1];
        }
        __syncthreads();
        // ---- write-out: warp w stores frame (w & 15), tile rows (w >> 4), +2, ...; 256 B per instruction ----
        {
            const int tt = warp & (kT - 1);
            const int64_t f = f0 + tt;
            if (f < f_end && lane < tw) {
                double* o = vout + f * n_out + out_base + lane;
                const double* si = in + tt * kOutStride + lane;
                for (int tr = warp >> 4; tr < th; tr += 2) o[(int64_t)tr * w_out] = si[tr * kTW];
            }
        }
        __syncthreads();  // out_s consumed: the buffer may be refilled
        if (s + 2 < nsub) {
            if (WIDE) fence_proxy_async();  // order our generic-proxy accesses before the async-proxy refill
            prefetch(f0 + 2 * kT, buf);
        }
        if (!WIDE) cp_async_commit();
    }
    if (!WIDE) cp_async_wait<0>();
}

// generic per-cell kernel restricted to the tiles the plan flagged
template <int FT>
__global__ void __launch_bounds__(256)
k_apply_generic_tiles(int64_t n_frames, int64_t n_in, int64_t h_out, int64_t w_out, int tiles_x,
                      const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                      const double* __restrict__ val, const int32_t* __restrict__ tile_info,
                      const double* __restrict__ vin, double* __restrict__ vout)
{
    const int64_t n_out = h_out * w_out;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t f0 = (int64_t)blockIdx.y * FT;
    if (o >= n_out) return;
    const int64_t orow = o / w_out, ocol = o - orow * w_out;
    const int64_t tile = (orow / kTH) * tiles_x + ocol / kTW;
    if (tile_info[tile * kTileInfoInts + 3] >= 0) return;
    const int32_t beg = row_ptr[o], end = row_ptr[o + 1];
    double acc[FT];
#pragma unroll
    for (int t = 0; t < FT; t++) acc[t] = 0.0;
    const double* in0 = vin + f0 * n_in;
    const int nf = (int)((n_frames - f0) < FT ? (n_frames - f0) : FT);
    for (int32_t w = beg; w < end; w++) {
        const int32_t c = col[w];
        const double a = val[w];
#pragma unroll
        for (int t = 0; t < FT; t++)
            if (t < nf) acc[t] = dadd(acc[t], dmul(a, __ldg(in0 + (int64_t)t * n_in + c)));
    }
    double* out0 = vout + f0 * n_out + o;
#pragma unroll
    for (int t = 0; t < FT; t++)
        if (t < nf) out0[(int64_t)t * n_out] = acc[t];
}

}  // namespace rg

using namespace rg;

static int64_t tiles_of(int64_t h_out, int64_t w_out, int* tiles_x)
{
    const int64_t tx = ceil_div(w_out, kTW), ty = ceil_div(h_out, kTH);
    if (tiles_x) *tiles_x = (int)tx;
    return tx * ty;
}

extern "C" int rg_apply_plan_sizes(int64_t h_out, int64_t w_out, int64_t* n_tiles_host,
                                   int64_t* tile_info_ints_host, int64_t* tile_rows_ints_host)
{
    if (h_out <= 0 || w_out <= 0 || !n_tiles_host || !tile_info_ints_host || !tile_rows_ints_host)
        return fail(RG_E_ARG, "rg_apply_plan_sizes: bad argument");
    const int64_t n = tiles_of(h_out, w_out, nullptr);
    *n_tiles_host = n;
    *tile_info_ints_host = n * kTileInfoInts + 4;  // + counter
    *tile_rows_ints_host = n * kRMAX * 2;
    return RG_OK;
}

extern "C" int rg_apply_plan_build(int device, void* stream, int64_t nnz,
                                   int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out,
                                   const int32_t* row_ptr, const int32_t* col,
                                   int32_t* tile_info, int32_t* tile_rows, uint16_t* lidx,
                                   int64_t* n_generic_tiles_host)
{
    if (h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0 || !row_ptr || !tile_info || !tile_rows || !n_generic_tiles_host)
        return fail(RG_E_ARG, "rg_apply_plan_build: bad argument");
    if (nnz > 0 && (!col || !lidx)) return fail(RG_E_ARG, "rg_apply_plan_build: null pointer");
    if (h_in * w_in >= INT32_MAX || h_out * w_out >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_apply_plan_build: too large");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x;
    const int64_t n_tiles = tiles_of(h_out, w_out, &tiles_x);
    int32_t* counter = tile_info + n_tiles * kTileInfoInts;
    RG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t) * 4, st));
    const int pad_even = (w_in % 2 == 0) && ((h_in * w_in) % 2 == 0);
    k_plan_tiles<<<(unsigned)n_tiles, 128, 0, st>>>(h_in, w_in, h_out, w_out, tiles_x, pad_even, row_ptr, col,
                                                    tile_info, tile_rows, lidx, counter);
    RG_LAUNCH_CHECK("k_plan_tiles");
    int32_t n_generic = 0;
    RG_CUDA(cudaMemcpyAsync(&n_generic, counter, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    *n_generic_tiles_host = n_generic;
    return RG_OK;
}

extern "C" int rg_apply_planned(int device, void* stream, int64_t n_frames,
                                int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out,
                                const int32_t* row_ptr, const int32_t* col, const double* val,
                                const int32_t* tile_info, const int32_t* tile_rows, const uint16_t* lidx,
                                int64_t n_generic_tiles,
                                const double* values_in, double* values_out)
{
    if (n_frames < 0 || h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0 || !row_ptr || !tile_info || !tile_rows ||
        !values_in || !values_out)
        return fail(RG_E_ARG, "rg_apply_planned: bad argument");
    if (n_frames == 0) return RG_OK;
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x;
    const int64_t n_tiles = tiles_of(h_out, w_out, &tiles_x);
    const int64_t n_in = h_in * w_in, n_out = h_out * w_out;
    static_assert(sizeof(StagedSmem) <= 227 * 1024, "staged tile does not fit in shared memory");
    const bool wide = (w_in % 2 == 0) && (n_in % 2 == 0) && ((uintptr_t)values_in % 16 == 0);
    const int tiles_y = (int)(n_tiles / tiles_x);
    auto kern = wide ? k_apply_staged<true> : k_apply_staged<false>;
    RG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StagedSmem)));
    if (n_generic_tiles < n_tiles) {
        const int64_t chunk = 65535LL * kFB;
        for (int64_t f = 0; f < n_frames; f += chunk) {
            const int64_t nf = n_frames - f < chunk ? n_frames - f : chunk;
            dim3 grid((unsigned)n_tiles, (unsigned)ceil_div(nf, kFB));
            kern<<<grid, kStagedThreads, sizeof(StagedSmem), st>>>(
                nf, w_in, n_in, h_out, w_out, tiles_x, tiles_y, row_ptr, val, tile_info, tile_rows, lidx,
                values_in + f * n_in, values_out + f * n_out);
            RG_LAUNCH_CHECK("k_apply_staged");
        }
    }
    if (n_generic_tiles > 0) {
        constexpr int FT = 8;
        const int64_t chunk = 65535LL * FT;
        for (int64_t f = 0; f < n_frames; f += chunk) {
            const int64_t nf = n_frames - f < chunk ? n_frames - f : chunk;
            dim3 grid((unsigned)ceil_div(n_out, 256), (unsigned)ceil_div(nf, FT));
            k_apply_generic_tiles<FT><<<grid, 256, 0, st>>>(nf, n_in, h_out, w_out, tiles_x, row_ptr, col, val,
                                                            tile_info, values_in + f * n_in, values_out + f * n_out);
            RG_LAUNCH_CHECK("k_apply_generic_tiles");
        }
    }
    return RG_OK;
}
