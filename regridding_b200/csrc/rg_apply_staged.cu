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

#ifndef RG_QUAD_INTERLEAVE
#define RG_QUAD_INTERLEAVE 0
#endif
#ifndef RG_PATCH_ROWS
#define RG_PATCH_ROWS 24
#endif
#ifndef RG_SKIP_COMPUTE
#define RG_SKIP_COMPUTE 0
#endif
#ifndef RG_SKIP_LOAD
#define RG_SKIP_LOAD 0
#endif
#ifndef RG_SKIP_STORE
#define RG_SKIP_STORE 0
#endif
#ifndef RG_L2_AHEAD
#define RG_L2_AHEAD 0
#endif
#ifndef RG_BULK_FILL
#define RG_BULK_FILL 0
#endif
#ifndef RG_PACKED_ENTRIES
#define RG_PACKED_ENTRIES 0
#endif
#ifndef RG_FB
#define RG_FB 512
#endif

namespace rg {

constexpr int kTH = 4;           // tile height (output rows)
constexpr int kTW = 32;          // tile width  (output cols) = one full coalesced row of 256 B
constexpr int kTileCells = kTH * kTW;
constexpr int kQuads = kTileCells / 4;  // a quad = 4 consecutive output cells = the 4 quarter-warps of a warp
constexpr int kT = 16;           // frames per sub-block: 8 lanes x 2 frames per lane
constexpr int kFB = RG_FB;         // frames per CTA (the tile-local CSR is reread every kFB frames)
constexpr int kRMAX = 64;        // max input rows in a footprint
#if RG_PACKED_ENTRIES
constexpr int kCP = 370;
constexpr int kPadMax = 1216;
#else
constexpr int kCP = 386;
constexpr int kPadMax = 1536;
#endif
                                 // kCP = doubles per staged frame; kCP/2 odd => the 8 frame lanes of a quarter-warp
                                 // hit 8 distinct 16-byte bank groups.  Slot kCP-1 of every frame holds 0.0.
constexpr int kCellsMax = kCP - 2;
constexpr int kZeroSlot = kCP - 1;
constexpr int kNnzMax = 1280;    // max CSR entries per tile
// kPadMax: max entries after padding the 4 rows of every quad to a common length
constexpr int kOutStride = kCP;  // staged output frame t lives in slots [0, 128) of staged input frame t
constexpr int kStagedThreads = 512;   // 2 CTAs per SM: their load / compute / store phases overlap
constexpr int kPairsPerLane = (kCP / 2 + 31) / 32;  // 16-byte pairs of one frame a lane copies per sub-block
constexpr int kPatch = 12;       // tiles are issued in patches of kPatchRows x kPatch tiles (~ one wave of 2 x 148
constexpr int kPatchRows = RG_PATCH_ROWS;   // CTAs) so that footprint halos are shared through L2
static_assert((kCP / 2) % 2 == 1 && kCP % 2 == 0, "kCP/2 must be odd");

constexpr int kTileInfoInts = 4;  // r0, nrows, cells, nnz (nnz < 0: tile handled by the generic kernel)

struct StagedSmem {
    double in_s[2][kT * kCP];     // [buffer][frame][cell]; after compute, frame t's slots [0,256) hold its outputs
#if RG_PACKED_ENTRIES
    struct alignas(16) Entry { double v; unsigned lo; unsigned pad; };
    Entry ent[kPadMax];           // padded tile-local CSR, interleaved per quad: entry (quad, w, q);
                                  // lo = BYTE offset of the referenced cell inside a staged frame
#else
    double val[kPadMax];          // padded tile-local CSR, interleaved per quad: entry (quad, w, q)
    uint16_t lidx[kPadMax];       // BYTE offset of the referenced cell inside a staged frame
#endif
    uint16_t quad_beg[kQuads + 1];
    uint16_t rowptr[kTileCells + 2];
    int32_t row_src[kRMAX];       // per footprint row: offset of its span inside one input frame (doubles)
    int32_t row_off[kRMAX + 1];   //                    offset of its span inside one staged frame
    alignas(8) uint64_t full[2];  // mbarriers: "buffer filled"
};

__host__ __device__ inline void tile_of_block(int64_t b, int tiles_x, int tiles_y, int& ty, int& tx)
{
    // patch-major order; patches (kPatchRows x kPatch tiles) and the tiles inside a patch are row-major
    const int px_count = (tiles_x + kPatch - 1) / kPatch;
    const int64_t full_rows = tiles_y / kPatchRows;                    // complete patch rows
    const int64_t per_patch_row = (int64_t)kPatchRows * tiles_x;       // tiles in a complete patch row
    int prow, ph;
    int64_t rem;
    if (b < full_rows * per_patch_row) {
        prow = (int)(b / per_patch_row);
        rem = b - (int64_t)prow * per_patch_row;
        ph = kPatchRows;
    } else {
        prow = (int)full_rows;
        rem = b - full_rows * per_patch_row;
        ph = tiles_y - prow * kPatchRows;
    }
    const int64_t per_full_patch = (int64_t)ph * kPatch;
    int pcol = (int)(rem / per_full_patch);
    if (pcol >= px_count) pcol = px_count - 1;
    const int64_t rem2 = rem - (int64_t)pcol * per_full_patch;
    const int pw = min(kPatch, tiles_x - pcol * kPatch);
    ty = prow * kPatchRows + (int)(rem2 / pw);
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
    __shared__ int s_rmin, s_rmax, s_nnz, s_pad;
    __shared__ int s_clo[kRMAX], s_chi[kRMAX], s_off[kRMAX + 1];
    const int tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile % tiles_x;
    const int th = (int)min((int64_t)kTH, h_out - (int64_t)ty * kTH);
    const int tw = (int)min((int64_t)kTW, w_out - (int64_t)tx * kTW);
    if (threadIdx.x == 0) { s_rmin = INT32_MAX; s_rmax = -1; s_nnz = 0; s_pad = 0; }
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
    // entries after padding the 4 rows of every quad (4 consecutive cells of a tile row) to a common length
    for (int qd = threadIdx.x; qd < kQuads; qd += blockDim.x) {
        const int tr = (4 * qd) / kTW, c0 = (4 * qd) % kTW;
        int m = 0;
        if (tr < th) {
            const int64_t o0 = ((int64_t)ty * kTH + tr) * w_out + (int64_t)tx * kTW;
            for (int c = c0; c < c0 + 4 && c < tw; c++) m = max(m, row_ptr[o0 + c + 1] - row_ptr[o0 + c]);
        }
        if (m) atomicAdd(&s_pad, 4 * m);
    }
    __syncthreads();
    bool generic = nrows > kRMAX || nnz > kNnzMax || s_pad > kPadMax || !pad_even;
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
                    s_clo[r] &= ~1;  // spans start on even columns and have even length: 16-byte copies
                    s_chi[r] |= 1;
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

__device__ __forceinline__ void cp_async_16(unsigned smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Requires even w_in / n_in and a 16-byte aligned values_in (the plan pads every footprint span to an even
// start and even length), so the footprint moves in 16-byte pieces.
__global__ void __launch_bounds__(kStagedThreads, 2)
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

    // ---- once per CTA: footprint table, tile-local CSR (padded per quad), zero slots, barriers ----
    {
        int base = 0;
        for (int tr = 0; tr < kTH; tr++) {
            int32_t b = 0, e = 0;
            if (tr < th) {
                const int64_t o0 = out_base + (int64_t)tr * w_out;
                b = row_ptr[o0];
                e = row_ptr[o0 + tw];
                if (threadIdx.x < kTW) S.rowptr[tr * kTW + threadIdx.x] = (uint16_t)(base + (row_ptr[o0 + min((int)threadIdx.x, tw)] - b));
            } else if (threadIdx.x < kTW) {
                S.rowptr[tr * kTW + threadIdx.x] = (uint16_t)base;
            }
            base += e - b;
        }
        if (threadIdx.x == 0) S.rowptr[kTileCells] = (uint16_t)base;
        for (int r = threadIdx.x; r < nrows; r += kStagedThreads) {
            S.row_src[r] = (int32_t)((int64_t)(r0 + r) * w_in + tile_rows[((int64_t)tile * kRMAX + r) * 2 + 0]);
            S.row_off[r] = tile_rows[((int64_t)tile * kRMAX + r) * 2 + 1];
        }
        if (threadIdx.x == 0) {
            S.row_off[nrows] = cells;
            mbar_init(&S.full[0], RG_BULK_FILL ? kT : kStagedThreads);
            mbar_init(&S.full[1], RG_BULK_FILL ? kT : kStagedThreads);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        if (threadIdx.x < 2 * kT) S.in_s[threadIdx.x / kT][(threadIdx.x % kT) * kCP + kZeroSlot] = 0.0;
    }
    __syncthreads();
    {
        // quad q4 holds cells 4*q4 .. 4*q4+3; all four rows are padded to the longest with (zero slot, 0.0)
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int qd = 0; qd < kQuads; qd++) {
                S.quad_beg[qd] = (uint16_t)acc;
                int m = 0;
                for (int c = 0; c < 4; c++) m = max(m, (int)S.rowptr[4 * qd + c + 1] - (int)S.rowptr[4 * qd + c]);
                acc += 4 * m;
            }
            S.quad_beg[kQuads] = (uint16_t)acc;
        }
    }
    __syncthreads();
    if (S.quad_beg[kQuads] > kPadMax) __trap();  // cannot happen: the plan routes such tiles to the generic kernel
    {
        // source position of local CSR entry j of cell c: the tile's rows are contiguous runs of the global CSR
        for (int qd = warp; qd < kQuads; qd += NW) {
            const int qb = S.quad_beg[qd], qn = (S.quad_beg[qd + 1] - qb) >> 2;
            const int c = lane & 3;  // cell of the quad
            const int cell = 4 * qd + c;
            const int tr = cell / kTW;
            const int lb = S.rowptr[cell], ln = (int)S.rowptr[cell + 1] - lb;
            const int64_t o0 = out_base + (int64_t)tr * w_out;
            const int32_t gb = (tr < th) ? row_ptr[o0] - (int32_t)S.rowptr[tr * kTW] : 0;  // global = gb + local
            for (int w = lane >> 2; w < qn; w += 8) {
                double v = 0.0;
                unsigned lo = kZeroSlot * 8u;
                if (w < ln) {
                    v = val[gb + lb + w];
                    lo = (unsigned)lidx[gb + lb + w] * 8u;
                }
#if RG_PACKED_ENTRIES
                S.ent[qb + 4 * w + c].v = v;
                S.ent[qb + 4 * w + c].lo = lo;
#else
                S.val[qb + 4 * w + c] = v;
                S.lidx[qb + 4 * w + c] = (uint16_t)lo;
#endif
            }
        }
    }
    __syncthreads();

    // ---- footprint copy: warp -> frame, lane -> 16-byte pairs lane + 32 j ----
    static_assert(kStagedThreads / 32 == kT, "one warp per frame of a sub-block");
    const int tt_p = warp;
    unsigned pf_mask = 0;
    int32_t pair_off[kPairsPerLane];  // source offset (doubles, inside a frame) of each of this lane's pairs; -1: none
    {
        const int npairs = cells >> 1;
#pragma unroll
        for (int j = 0; j < kPairsPerLane; j++) {
            const int p = lane + 32 * j;
            int32_t off = -1;
            bool first_of_row = false;
            if (p < npairs) {
                // row of staged cell 2p: last r with row_off[r] <= 2p
                int lo = 0, hi = nrows - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (S.row_off[mid] <= 2 * p) lo = mid; else hi = mid - 1;
                }
                off = S.row_src[lo] + (2 * p - S.row_off[lo]);
                first_of_row = (2 * p == S.row_off[lo]);
            }
            pair_off[j] = off;
            // this lane touches L2 for the pair if it starts a 128-byte line or a footprint row
            if (off >= 0 && ((off & 15) == 0 || first_of_row)) pf_mask |= 1u << j;
        }
    }
    const unsigned dst_lane = (unsigned)((tt_p * kCP + 2 * lane) * 8);
    auto prefetch = [&](int64_t f0, int buf) {
        const int64_t f = f0 + tt_p;
#if RG_BULK_FILL
        // one bulk (TMA engine) copy per footprint row of this warp's frame: no LSU / MIO traffic
        const bool valid = f < f_end && !RG_SKIP_LOAD;
        if (lane == 0) mbar_arrive_expect_tx(&S.full[buf], valid ? (unsigned)cells * 8u : 0u);
        __syncwarp();
        if (valid) {
            const double* src = vin + f * n_in;
            double* dst = S.in_s[buf] + tt_p * kCP;
            for (int r = lane; r < nrows; r += 32) {
                const int off = S.row_off[r], len = S.row_off[r + 1] - off;
                if (len > 0) bulk_g2s(dst + off, src + S.row_src[r], (unsigned)len * 8u, &S.full[buf]);
            }
        }
#else
        if (f < f_end) {
            const double* src = vin + f * n_in;
            const unsigned dst = smem_u32(S.in_s[buf]) + dst_lane;
#pragma unroll
            for (int j = 0; j < kPairsPerLane; j++)
                if (pair_off[j] >= 0 && !RG_SKIP_LOAD) cp_async_16(dst + j * 32 * 16, src + pair_off[j]);
        }
        cp_async_mbar_arrive(&S.full[buf]);
#endif
    };

    // HBM -> L2 a few sub-blocks ahead of the shared-memory copy: the copy then sees L2 latency, and DRAM
    // requests are spread over the whole iteration instead of arriving in bursts
    auto l2_prefetch = [&](int64_t f0) {
        const int64_t f = f0 + tt_p;
        if (RG_L2_AHEAD > 0 && f < f_end) {
            const double* src = vin + f * n_in;
#pragma unroll
            for (int j = 0; j < kPairsPerLane; j++)
                if (pf_mask & (1u << j)) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(src + pair_off[j]));
        }
    };

    const int nsub = (int)((f_end - f_begin + kT - 1) / kT);
    prefetch(f_begin, 0);
    if (nsub > 1) prefetch(f_begin + kT, 1);
    for (int a = 2; a < 2 + RG_L2_AHEAD; a++) l2_prefetch(f_begin + (int64_t)a * kT);
    const int q = lane >> 3, t = lane & 7;  // quarter-warp = one output cell; lane owns frames t and t + 8
    for (int s = 0; s < nsub; s++) {
        const int64_t f0 = f_begin + (int64_t)s * kT;
        const int buf = s & 1;
        mbar_wait(&S.full[buf], (unsigned)((s >> 1) & 1));
        double* in = S.in_s[buf];
        const char* in0 = reinterpret_cast<const char*>(in + t * kCP);
        // ---- compute: quarter-warp per output cell, rows of a quad share one (padded) trip count ----
        double acc[2][2];
#if RG_QUAD_INTERLEAVE
        {
            // the warp's two quads are walked together: two independent dependency chains per lane
            const int qbA = S.quad_beg[warp], qeA = S.quad_beg[warp + 1];
            const int qbB = S.quad_beg[warp + NW], qeB = S.quad_beg[warp + NW + 1];
            const int nA = (qeA - qbA) >> 2, nB = (qeB - qbB) >> 2;
            const int nmin = min(nA, nB);
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            int eA = qbA + q, eB = qbB + q;
#pragma unroll 2
            for (int w = 0; w < nmin; w++, eA += 4, eB += 4) {
                const unsigned loA = S.lidx[eA], loB = S.lidx[eB];
                const double vA = S.val[eA], vB = S.val[eB];
                const double xA0 = *reinterpret_cast<const double*>(in0 + loA);
                const double xA1 = *reinterpret_cast<const double*>(in0 + loA + 8 * kCP * 8);
                const double xB0 = *reinterpret_cast<const double*>(in0 + loB);
                const double xB1 = *reinterpret_cast<const double*>(in0 + loB + 8 * kCP * 8);
                a0 = dadd(a0, dmul(vA, xA0));
                a1 = dadd(a1, dmul(vA, xA1));
                b0 = dadd(b0, dmul(vB, xB0));
                b1 = dadd(b1, dmul(vB, xB1));
            }
            for (int w = nmin; w < nA; w++, eA += 4) {
                const unsigned lo = S.lidx[eA];
                const double v = S.val[eA];
                a0 = dadd(a0, dmul(v, *reinterpret_cast<const double*>(in0 + lo)));
                a1 = dadd(a1, dmul(v, *reinterpret_cast<const double*>(in0 + lo + 8 * kCP * 8)));
            }
            for (int w = nmin; w < nB; w++, eB += 4) {
                const unsigned lo = S.lidx[eB];
                const double v = S.val[eB];
                b0 = dadd(b0, dmul(v, *reinterpret_cast<const double*>(in0 + lo)));
                b1 = dadd(b1, dmul(v, *reinterpret_cast<const double*>(in0 + lo + 8 * kCP * 8)));
            }
            acc[0][0] = a0; acc[0][1] = a1; acc[1][0] = b0; acc[1][1] = b1;
        }
#else
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int qd = warp + k * NW;
            const int qb = S.quad_beg[qd], qe = S.quad_beg[qd + 1];
            double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
            for (int e = qb + q; e < (RG_SKIP_COMPUTE ? qb : qe); e += 4) {
#if RG_PACKED_ENTRIES
                const int4 raw = *reinterpret_cast<const int4*>(&S.ent[e]);  // one 16-byte shared load per entry
                const double v = __hiloint2double(raw.y, raw.x);
                const unsigned lo = (unsigned)raw.z;
#else
                const unsigned lo = S.lidx[e];
                const double v = S.val[e];
#endif
                const double x0 = *reinterpret_cast<const double*>(in0 + lo);
                const double x1 = *reinterpret_cast<const double*>(in0 + lo + 8 * kCP * 8);
                a0 = dadd(a0, dmul(v, x0));
                a1 = dadd(a1, dmul(v, x1));
            }
            acc[k][0] = a0;
            acc[k][1] = a1;
        }
#endif
        __syncthreads();  // everyone is done reading in_s[buf]: frame t's slots [0,256) now take its outputs
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int o_local = 4 * (warp + k * NW) + q;
            in[t * kOutStride + o_local] = acc[k][0];
            in[(t + 8) * kOutStride + o_local] = acc[k][1];
        }
        __syncthreads();
        // ---- write-out: warp w stores frame w, all tile rows; 256 B per instruction ----
        {
            const int64_t f = f0 + tt_p;
            if (f < f_end && lane < tw) {
                double* o = vout + f * n_out + out_base + lane;
                const double* si = in + tt_p * kOutStride + lane;
                for (int tr = 0; tr < (RG_SKIP_STORE ? 0 : th); tr++) o[(int64_t)tr * w_out] = si[tr * kTW];
            }
        }
        __syncthreads();  // outputs consumed: the buffer may be refilled
        if (s + 2 < nsub) {
            if (RG_BULK_FILL) fence_proxy_async();  // our generic-proxy accesses of the buffer precede the async refill
            prefetch(f0 + 2 * kT, buf);
        }
        l2_prefetch(f0 + (int64_t)(2 + RG_L2_AHEAD) * kT);
    }
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
    const bool aligned = ((uintptr_t)values_in % 16 == 0);
    const int tiles_y = (int)(n_tiles / tiles_x);
    if (!aligned) {
        // the staged kernel moves 16-byte pieces; a misaligned values pointer takes the generic kernel
        return rg_apply_csr(device, stream, n_frames, n_in, n_out, row_ptr, col, val, values_in, values_out);
    }
    RG_CUDA(cudaFuncSetAttribute(k_apply_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StagedSmem)));
    if (n_generic_tiles < n_tiles) {
        const int64_t chunk = 65535LL * kFB;
        for (int64_t f = 0; f < n_frames; f += chunk) {
            const int64_t nf = n_frames - f < chunk ? n_frames - f : chunk;
            dim3 grid((unsigned)n_tiles, (unsigned)ceil_div(nf, kFB));
            k_apply_staged<<<grid, kStagedThreads, sizeof(StagedSmem), st>>>(
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
