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

// development switches (ablations for profiles/r1_apply_tuning.md); all 0 in the product build
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
#ifndef RG_NO_ALIGN
#define RG_NO_ALIGN 0
#endif
#ifndef RG_FB
#define RG_FB 512
#endif
#ifndef RG_NBUF
#define RG_NBUF 2   // footprint buffers per CTA (experiment: 1 buffer, 3 CTAs per SM)
#endif
#ifndef RG_CTAS
#define RG_CTAS 2
#endif
#ifndef RG_CP_MODE
#define RG_CP_MODE "cg.shared.global"   // footprint copies: L2 -> shared memory, bypassing L1
#endif
#ifndef RG_WARP_REFILL
#define RG_WARP_REFILL 1   // a warp refills its own frame row right after writing it out (no third block barrier)
#endif
#ifndef RG_TAIL_OUT
#define RG_TAIL_OUT 1   // stage the outputs in the unused tail of the footprint buffer (one barrier per sub-block)
#endif


namespace rg {

constexpr int kTH = 4;           // tile height (output rows)
constexpr int kTW = 32;          // tile width  (output cols) = one full coalesced row of 256 B
constexpr int kTileCells = kTH * kTW;
constexpr int kQuads = kTileCells / 4;  // a quad = 4 consecutive output cells = the 4 quarter-warps of a warp
constexpr int kPairs = kTileCells / 2;  // a pair = 2 consecutive output cells = the 2 quarter-warps of a half-warp
constexpr int kT = 16;           // frames per sub-block: 8 lanes x 2 frames per lane
constexpr int kFB = RG_FB;       // frames per CTA (the tile-local CSR is rebuilt every kFB frames)
constexpr int kRMAX = 64;        // max input rows in a footprint
constexpr int kCP = 386;         // doubles per staged frame; kCP/2 odd => the 8 frame lanes of a quarter-warp
                                 // hit 8 distinct 16-byte bank groups
constexpr int kCellsMax = kCP - 2;
constexpr int kZeroEven = kCP - 2;   // slots kCP-2 (even) and kCP-1 (odd) of every staged frame hold 0.0:
constexpr int kZeroOdd = kCP - 1;    // the targets of padding entries, one per bank parity
constexpr int kNnzMax = 1280;    // max CSR entries per tile
constexpr int kPadMax = 1536;    // max SLOT entries per tile (4 cells x padded, aligned slots of every quad)
constexpr int kAlignMax = 16;    // rows up to this length take part in the bank-parity alignment
constexpr int kOutStride = kCP;  // staged output frame t lives in slots [0, 128) of staged input frame t ...
constexpr int kTailOut = kCellsMax - kTileCells;  // ... or, when the footprint has at most this many cells, in slots
                                 // [kTailOut, kTailOut + 128): the footprint part of the buffer is then free for the
                                 // next fill as soon as the compute is over, and two of three block barriers go
constexpr int kStagedThreads = 512;   // 2 CTAs per SM: their load / compute / store phases overlap
constexpr int kCopyPairs = kCellsMax / 2;            // 16-byte pieces of one staged frame
constexpr int kPairsPerLane = (kCopyPairs + 31) / 32;  // ... of which a lane copies up to this many per sub-block
constexpr int kPatch = 12;       // tiles are issued in patches of kPatchRows x kPatch tiles (~ one wave of 2 x 148
constexpr int kPatchRows = RG_PATCH_ROWS;   // CTAs) so that footprint halos are shared through L2
static_assert((kCP / 2) % 2 == 1 && kCP % 2 == 0, "kCP/2 must be odd");


// per-tile plan record (int32): r0, nrows, cells, nnz (nnz < 0: tile handled by the generic kernel), plain layout?,
// slot entries, slot base (int64 in two words), then quad_beg[kQuads + 1] as uint16, then pair_of[kPairs] as uint8:
// quad g is made of the pairs pair_of[2g], pair_of[2g+1] (pairs sorted by slot count, so that the two pairs of a
// quad -- which share one trip count -- have nearly equal lengths)
constexpr int kQuadWords = (kQuads + 2) / 2;
constexpr int kTileInfoInts = 8 + kQuadWords + kPairs / 4;

// Slot layout of a quad (4 cells c, L slots w, L even):  entry(w, c) = qb + (w >> 1) * 8 + c * 2 + (w & 1),
// so the two values (16 B) and the two offsets (4 B) of slots (w, w+1) of a cell are one shared-memory load each.
struct StagedSmem {
    double in_s[RG_NBUF][kT * kCP];     // [buffer][frame][cell]; after compute, frame t's slots [0,128) hold its outputs
    alignas(16) double val[kPadMax];
    alignas(16) uint16_t lidx[kPadMax];   // BYTE offset of the referenced cell inside a staged frame
    uint16_t quad_beg[kQuads + 2];
    uint8_t pair_of[kPairs];      // pair (2 consecutive output cells) handled by each half-warp slot of the tile
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

// Bank-parity alignment of the entry lists of two cells that share a half-warp.
//
// In the compute loop a half-warp reads, per slot, entry w of cell A (8 lanes = 8 frames) and entry w of cell B.
// With the frame-major staging (frame stride 16 B-odd) the 8 frame lanes of one cell cover all eight 16-byte
// bank groups at the 8-byte half selected by the PARITY of the staged cell index, so the two cells collide
// (2 wavefronts instead of 1) exactly when their indices have equal parity and differ.  Padding entries
// (weight 0.0 at an always-zero slot) may be inserted anywhere in a cell's list without changing its sum
// (acc + 0.0 * 0.0 == acc bit for bit; acc is never -0.0), so the two lists are aligned like an LCS:
// minimise 3 * slots + 2 * conflicts.  ops: 2 bits per slot, bit 0 = A advances, bit 1 = B advances.
__device__ void align_pair(const uint16_t* la, int a, const uint16_t* lb, int b, uint64_t& ops, int& L)
{
    uint8_t ch[kAlignMax + 1][kAlignMax + 1];
    int prev[kAlignMax + 1], cur[kAlignMax + 1];
    prev[0] = 0;
    for (int j = 1; j <= b; j++) { prev[j] = prev[j - 1] + 3; ch[0][j] = 2; }
    for (int i = 1; i <= a; i++) {
        cur[0] = prev[0] + 3;
        ch[i][0] = 1;
        const unsigned x = la[i - 1];
        for (int j = 1; j <= b; j++) {
            const unsigned y = lb[j - 1];
            const int conflict = (((x ^ y) & 1u) == 0u && x != y) ? 2 : 0;
            int best = prev[j - 1] + 3 + conflict, c = 3;
            if (prev[j] + 3 < best) { best = prev[j] + 3; c = 1; }
            if (cur[j - 1] + 3 < best) { best = cur[j - 1] + 3; c = 2; }
            cur[j] = best;
            ch[i][j] = (uint8_t)c;
        }
        for (int j = 0; j <= b; j++) prev[j] = cur[j];
    }
    int i = a, j = b, n = 0;
    uint64_t r = 0;
    while (i > 0 || j > 0) {   // walks from the last slot to the first: slot 0 ends up in the lowest bits
        const unsigned c = ch[i][j];
        r = (r << 2) | c;
        n++;
        i -= (int)(c & 1u);
        j -= (int)(c >> 1);
    }
    ops = r;
    L = n;
}

// Entry lists of the pair of cells (2p, 2p+1) of a tile and their slot layout.
struct PairLayout {
    int a, b;          // row lengths
    int32_t gA, gB;    // global CSR position of their first entries
    uint64_t ops;      // aligned: 2 bits per slot (bit 0: A advances, bit 1: B advances)
    int L;             // slots
    bool aligned;
};

__device__ PairLayout pair_layout(int p, int th, int tw, int64_t out_base, int64_t w_out,
                                  const int32_t* __restrict__ row_ptr, const uint16_t* __restrict__ lidx, bool align)
{
    PairLayout P;
    P.a = P.b = 0;
    P.gA = P.gB = 0;
    P.ops = 0;
    P.aligned = false;
    const int cA = 2 * p, tr = cA / kTW, col = cA % kTW;
    if (tr < th && col < tw) {
        const int64_t o = out_base + (int64_t)tr * w_out + col;
        P.gA = row_ptr[o];
        P.gB = row_ptr[o + 1];
        P.a = P.gB - P.gA;
        if (col + 1 < tw) P.b = row_ptr[o + 2] - P.gB;
    }
    P.L = max(P.a, P.b);
    if (align && P.a > 0 && P.b > 0 && P.a <= kAlignMax && P.b <= kAlignMax) {
        uint16_t la[kAlignMax], lb[kAlignMax];
        for (int k = 0; k < P.a; k++) la[k] = lidx[P.gA + k];
        for (int k = 0; k < P.b; k++) lb[k] = lidx[P.gB + k];
        align_pair(la, P.a, lb, P.b, P.ops, P.L);
        P.aligned = true;
    }
    return P;
}

// ---------------------------------------------------------------------------
// plan: footprint of every tile + tile-local indices
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_plan_tiles(int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out, int tiles_x, int pad_even,
             const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
             int32_t* __restrict__ tile_info, int32_t* __restrict__ tile_rows, uint16_t* __restrict__ lidx,
             int32_t* __restrict__ n_generic, int32_t* __restrict__ tile_slots)
{
    __shared__ int s_rmin, s_rmax, s_nnz, s_pad, s_plain;
    __shared__ uint16_t s_pair_len[kPairs], s_quad_beg[kQuads + 2];
    __shared__ uint8_t s_pair_of[kPairs];  // rank -> pair
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
    // slot entries of the UNALIGNED layout (the apply kernel's fall-back): the 4 rows of every quad padded to
    // their longest, rounded up to even
    for (int qd = threadIdx.x; qd < kQuads; qd += blockDim.x) {
        const int tr = (4 * qd) / kTW, c0 = (4 * qd) % kTW;
        int m = 0;
        if (tr < th) {
            const int64_t o0 = ((int64_t)ty * kTH + tr) * w_out + (int64_t)tx * kTW;
            for (int c = c0; c < c0 + 4 && c < tw; c++) m = max(m, row_ptr[o0 + c + 1] - row_ptr[o0 + c]);
        }
        if (m) atomicAdd(&s_pad, 4 * ((m + 1) & ~1));
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
        // copy table: source offset (doubles, inside one input frame) of every 16-byte piece of the staged frame
        const int npairs = s_off[nrows] >> 1;
        for (int p = threadIdx.x; p < kCopyPairs; p += blockDim.x) {
            int32_t off = -1;
            if (p < npairs) {
                int lo = 0, hi = nrows - 1;  // row of staged cell 2p: last r with s_off[r] <= 2p
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_off[mid] <= 2 * p) lo = mid; else hi = mid - 1;
                }
                off = (int32_t)((int64_t)(rmin + lo) * w_in + s_clo[lo] + (2 * p - s_off[lo]));
            }
            tile_rows[(int64_t)tile * kCopyPairs + p] = off;
        }
    }
    // slot layout: the two cells of a half-warp are aligned for bank parity (align_pair); the two pairs of a quad
    // share one even slot count.  If the aligned layout is too long for the tile, fall back to the unaligned one.
    int slots = 0, plain = RG_NO_ALIGN;
    if (!generic && nnz) {
        __syncthreads();  // lidx of this tile is complete
        const int64_t out_base = ((int64_t)ty * kTH) * w_out + (int64_t)tx * kTW;
        if (threadIdx.x < kPairs) {
            const PairLayout P = pair_layout(threadIdx.x, th, tw, out_base, w_out, row_ptr, lidx, !RG_NO_ALIGN);
            s_pair_len[threadIdx.x] = (uint16_t)min(P.L, 65535);
        }
        __syncthreads();
        // pairs sorted by slot count (stable): rank -> pair.  Quad g = ranks 2g, 2g+1.
        auto rank_pairs = [&]() {
            if (threadIdx.x < kPairs) {
                const int me = s_pair_len[threadIdx.x];
                int r = 0;
                for (int k = 0; k < kPairs; k++) {
                    const int o = s_pair_len[k];
                    r += (o < me) || (o == me && k < (int)threadIdx.x);
                }
                s_pair_of[r] = (uint8_t)threadIdx.x;
            }
            __syncthreads();
        };
        auto quad_len = [&](int qd) { return (max((int)s_pair_len[s_pair_of[2 * qd]], (int)s_pair_len[s_pair_of[2 * qd + 1]]) + 1) & ~1; };
        rank_pairs();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int qd = 0; qd < kQuads; qd++) acc += 4 * quad_len(qd);
            s_plain = (acc > kPadMax) ? 1 : RG_NO_ALIGN;
        }
        __syncthreads();
        plain = s_plain;
        if (plain) {
            if (threadIdx.x < kPairs) {
                const PairLayout P = pair_layout(threadIdx.x, th, tw, out_base, w_out, row_ptr, lidx, false);
                s_pair_len[threadIdx.x] = (uint16_t)min(P.L, 65535);
            }
            __syncthreads();
            rank_pairs();
        }
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int qd = 0; qd < kQuads; qd++) {
                // offsets are multiples of 8: bit 0 flags an odd slot count (the last, padding slot is not processed)
                const int odd = max((int)s_pair_len[s_pair_of[2 * qd]], (int)s_pair_len[s_pair_of[2 * qd + 1]]) & 1;
                s_quad_beg[qd] = (uint16_t)(min(acc, 65528) | odd);
                acc += 4 * quad_len(qd);
            }
            s_quad_beg[kQuads] = (uint16_t)min(acc, 65535);
            s_quad_beg[kQuads + 1] = 0;
            s_pad = acc;
        }
        __syncthreads();
        if (s_pad > kPadMax) generic = true;  // (sorted quads never need more slots than the positional ones checked above)
        slots = generic ? 0 : s_pad;
        uint16_t* qdst = reinterpret_cast<uint16_t*>(tile_info + (int64_t)tile * kTileInfoInts + 8);
        for (int k = threadIdx.x; k < kQuads + 2; k += blockDim.x) qdst[k] = s_quad_beg[k];
        uint8_t* pdst = reinterpret_cast<uint8_t*>(tile_info + (int64_t)tile * kTileInfoInts + 8 + kQuadWords);
        for (int k = threadIdx.x; k < kPairs; k += blockDim.x) pdst[k] = s_pair_of[k];
    }
    if (threadIdx.x == 0) {
        int32_t* info = tile_info + (int64_t)tile * kTileInfoInts;
        info[0] = nnz ? rmin : 0;
        info[1] = generic ? 0 : nrows;
        info[2] = (generic || !nnz) ? 0 : s_off[nrows];
        info[3] = generic ? -1 : nnz;
        info[4] = plain;
        info[5] = slots;
        info[6] = info[7] = 0;  // slot base: k_plan_base
        tile_slots[tile] = slots;
        if (generic) atomicAdd(n_generic, 1);
    }
    (void)h_in;
}

__global__ void k_plan_base(int64_t n_tiles, const int64_t* __restrict__ base, int32_t* __restrict__ tile_info)
{
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    memcpy(tile_info + tile * kTileInfoInts + 6, base + tile, sizeof(int64_t));
}

// slot arrays of every staged tile: entry(w, c) of quad qd at base + quad_beg[qd] + (w >> 1) * 8 + c * 2 + (w & 1)
__global__ void __launch_bounds__(kPairs)
k_plan_slots(int64_t h_out, int64_t w_out, int tiles_x, const int32_t* __restrict__ row_ptr, const double* __restrict__ val,
             const int32_t* __restrict__ tile_info, const uint16_t* __restrict__ lidx,
             double* __restrict__ slot_val, uint16_t* __restrict__ slot_lidx)
{
    const int tile = blockIdx.x;
    const int32_t* info = tile_info + (int64_t)tile * kTileInfoInts;
    if (info[3] <= 0) return;
    const int ty = tile / tiles_x, tx = tile % tiles_x;
    const int th = (int)min((int64_t)kTH, h_out - (int64_t)ty * kTH);
    const int tw = (int)min((int64_t)kTW, w_out - (int64_t)tx * kTW);
    const int64_t out_base = ((int64_t)ty * kTH) * w_out + (int64_t)tx * kTW;
    int64_t base;
    memcpy(&base, info + 6, sizeof(int64_t));
    const uint16_t* quad_beg = reinterpret_cast<const uint16_t*>(info + 8);
    const PairLayout P = pair_layout(threadIdx.x, th, tw, out_base, w_out, row_ptr, lidx, info[4] == 0);
    const uint8_t* pair_of = reinterpret_cast<const uint8_t*>(info + 8 + kQuadWords);
    int rank = 0;
    for (int k = 0; k < kPairs; k++)
        if (pair_of[k] == threadIdx.x) rank = k;
    const int qd = rank >> 1, cA = (rank & 1) * 2;  // my cells are cells cA, cA + 1 of quad qd
    const int qb = quad_beg[qd] & ~7, Lq = (((int)quad_beg[qd + 1] & ~7) - qb) >> 2;
    int ia = 0, ib = 0;
    for (int w = 0; w < Lq; w++) {
        bool hasA, hasB;
        if (!P.aligned) {
            hasA = w < P.a;
            hasB = w < P.b;
        } else {
            const unsigned op = (w < P.L) ? (unsigned)((P.ops >> (2 * w)) & 3u) : 0u;
            hasA = op & 1u;
            hasB = op & 2u;
        }
        double vA = 0.0, vB = 0.0;
        unsigned lA = 0, lB = 0;
        if (hasA) { vA = val[P.gA + ia]; lA = lidx[P.gA + ia]; ia++; }
        if (hasB) { vB = val[P.gB + ib]; lB = lidx[P.gB + ib]; ib++; }
        // a padding entry reads the always-zero slot of the bank parity its partner does NOT use
        if (!hasA) lA = (hasB && (lB & 1u)) ? kZeroEven : kZeroOdd;
        if (!hasB) lB = (lA & 1u) ? kZeroEven : kZeroOdd;
        const int64_t e = base + qb + (w >> 1) * 8 + cA * 2 + (w & 1);
        slot_val[e] = vA;
        slot_val[e + 2] = vB;
        slot_lidx[e] = (uint16_t)(lA * 8u);
        slot_lidx[e + 2] = (uint16_t)(lB * 8u);
    }
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

// Requires even w_in / n_in and a 16-byte aligned values_in (the plan pads every footprint span to an even
// start and even length), so the footprint moves in 16-byte pieces.
__global__ void __launch_bounds__(kStagedThreads, RG_CTAS)
k_apply_staged(int64_t n_frames, int64_t w_in, int64_t n_in, int64_t h_out, int64_t w_out, int tiles_x, int tiles_y,
               const int32_t* __restrict__ tile_info, const int32_t* __restrict__ tile_rows,
               const double* __restrict__ slot_val, const uint16_t* __restrict__ slot_lidx,
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
    const int cells = info[2];
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
        if (tw == kTW && th == kTH && (w_out & 1) == 0 && ((uintptr_t)vout & 15) == 0) {
            const int hr = lane >> 4, c2 = (lane & 15) * 2;
            double2* o = reinterpret_cast<double2*>(vout + (f_begin + warp) * n_out + out_base + (int64_t)hr * w_out + c2);
            const int64_t frame_step = NW * (n_out >> 1), row_step = w_out;  // in double2
            const double2 z = make_double2(0.0, 0.0);
            for (int64_t f = f_begin + warp; f < f_end; f += NW, o += frame_step) {
                o[0] = z;
                o[row_step] = z;
            }
        } else if (lane < tw) {
            for (int64_t f = f_begin + warp; f < f_end; f += NW) {
                double* o = vout + f * n_out + out_base + lane;
                for (int tr = 0; tr < th; tr++) o[(int64_t)tr * w_out] = 0.0;
            }
        }
        return;
    }

    // ---- once per CTA: footprint table, the tile's slot arrays (built by the plan), zero slots, barriers ----
    {
        const int nslots = info[5];
        int64_t base;
        memcpy(&base, info + 6, sizeof(int64_t));
        const uint16_t* qsrc = reinterpret_cast<const uint16_t*>(info + 8);
        if (threadIdx.x < kQuads + 2) S.quad_beg[threadIdx.x] = qsrc[threadIdx.x];
        if (threadIdx.x < kPairs / 4)
            reinterpret_cast<int32_t*>(S.pair_of)[threadIdx.x] = info[8 + kQuadWords + threadIdx.x];
        // slot base and count are multiples of 8 entries: 16-byte pieces
        const int4* sv = reinterpret_cast<const int4*>(slot_val + base);
        int4* dv = reinterpret_cast<int4*>(S.val);
        for (int k = threadIdx.x; k < nslots / 2; k += kStagedThreads) dv[k] = sv[k];
        const int4* sl = reinterpret_cast<const int4*>(slot_lidx + base);
        int4* dl = reinterpret_cast<int4*>(S.lidx);
        for (int k = threadIdx.x; k < nslots / 8; k += kStagedThreads) dl[k] = sl[k];
        if (threadIdx.x == 0) {
            mbar_init(&S.full[0], kStagedThreads);
            mbar_init(&S.full[1], kStagedThreads);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        if (threadIdx.x < 2 * RG_NBUF * kT) {
            const int b = threadIdx.x / (2 * kT), t = (threadIdx.x / 2) % kT;
            S.in_s[b][t * kCP + kZeroEven + (threadIdx.x & 1)] = 0.0;
        }
    }
    __syncthreads();

    // ---- footprint copy: warp -> frame, lane -> 16-byte pieces lane + 32 j (source offsets from the plan) ----
    static_assert(kStagedThreads / 32 == kT, "one warp per frame of a sub-block");
    const int tt_p = warp;
    const int npairs = cells >> 1;
    const int nops = (npairs + 31) >> 5;
    int32_t pair_off[kPairsPerLane];  // source offset (doubles, inside a frame) of each of this lane's pieces; -1: none
#pragma unroll
    for (int j = 0; j < kPairsPerLane; j++) {
        const int p = lane + 32 * j;
        pair_off[j] = (p < npairs) ? tile_rows[(int64_t)tile * kCopyPairs + p] : -1;
    }
    const unsigned dst_lane = (unsigned)((tt_p * kCP + 2 * lane) * 8);
    auto prefetch = [&](int64_t f0, int buf) {
        const int64_t f = f0 + tt_p;
        if (f < f_end && !RG_SKIP_LOAD) {
            const unsigned long long src = (unsigned long long)(vin + f * n_in);
            const unsigned dst = smem_u32(S.in_s[buf]) + dst_lane;
#pragma unroll
            for (int j = 0; j < kPairsPerLane; j++) {
                if (j < nops)  // uniform: no copy instruction is issued beyond the footprint
                    asm volatile(
                        "{\n.reg .pred p;\n.reg .u64 a;\n"
                        "setp.ge.s32 p, %2, 0;\n"
                        "mad.wide.u32 a, %2, 8, %1;\n"
                        "@p cp.async." RG_CP_MODE " [%0], [a], 16;\n}\n" ::"r"(dst + j * 32 * 16),
                        "l"(src), "r"(pair_off[j]));
            }
        }
        cp_async_mbar_arrive(&S.full[buf]);
    };

    const int nsub = (int)((f_end - f_begin + kT - 1) / kT);
    prefetch(f_begin, 0);
    if (RG_NBUF > 1 && nsub > 1) prefetch(f_begin + kT, 1);
    const int q = lane >> 3, t = lane & 7;  // quarter-warp = one output cell; lane owns frames t and t + 8
    // output cell (index in the tile) of this quarter-warp in its two quads: half-warp h of quad g owns pair pair_of[2g+h]
    int o_local[2];
#pragma unroll
    for (int k = 0; k < 2; k++) o_local[k] = 2 * (int)S.pair_of[2 * (2 * warp + k) + (q >> 1)] + (q & 1);
    // full-width tiles of 16-byte aligned output rows are written with 16-byte stores
    const bool vec_store = tw == kTW && th >= 2 && (w_out & 1) == 0 && ((uintptr_t)vout & 15) == 0;
    const int out_off = (RG_TAIL_OUT && RG_NBUF > 1 && cells <= kTailOut) ? kTailOut : 0;  // uniform over the CTA
    for (int s = 0; s < nsub; s++) {
        const int64_t f0 = f_begin + (int64_t)s * kT;
        const int buf = (RG_NBUF > 1) ? (s & 1) : 0;
        mbar_wait(&S.full[buf], (unsigned)((RG_NBUF > 1 ? (s >> 1) : s) & 1));
        double* in = S.in_s[buf];
        const char* in0 = reinterpret_cast<const char*>(in + t * kCP);
        // ---- compute: quarter-warp per output cell, all cells of a quad share one slot count.  The warp's two
        // quads (neighbours in the length-sorted order) are walked together: two independent chains per lane ----
        double acc[2][2];
        {
            const int qvA = S.quad_beg[2 * warp], qvB = S.quad_beg[2 * warp + 1], qvC = S.quad_beg[2 * warp + 2];
            const int qbA = qvA & ~7, qbB = qvB & ~7, qeB = qvC & ~7;
            const double2* vpA = reinterpret_cast<const double2*>(S.val + qbA) + q;
            const uint32_t* lpA = reinterpret_cast<const uint32_t*>(S.lidx + qbA) + q;
            const double2* vpB = reinterpret_cast<const double2*>(S.val + qbB) + q;
            const uint32_t* lpB = reinterpret_cast<const uint32_t*>(S.lidx + qbB) + q;
            const int oddA = qvA & 1, oddB = qvB & 1;
            const int nA = RG_SKIP_COMPUTE ? 0 : ((qbB - qbA) >> 3) - oddA;  // complete slot pairs
            const int nB = RG_SKIP_COMPUTE ? 0 : ((qeB - qbB) >> 3) - oddB;
            const int nmin = min(nA, nB);
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            auto two_slots = [&](const uint32_t* lp, const double2* vp, double& c0, double& c1) {
                const uint32_t l2 = *lp;   // offsets of slots 2j, 2j+1: one 4-byte load
                const double2 v2 = *vp;    // weights of slots 2j, 2j+1: one 16-byte load
                const unsigned lo0 = l2 & 0xffffu, lo1 = l2 >> 16;
                const double x00 = *reinterpret_cast<const double*>(in0 + lo0);
                const double x01 = *reinterpret_cast<const double*>(in0 + lo0 + 8 * kCP * 8);
                const double x10 = *reinterpret_cast<const double*>(in0 + lo1);
                const double x11 = *reinterpret_cast<const double*>(in0 + lo1 + 8 * kCP * 8);
                c0 = dadd(c0, dmul(v2.x, x00));
                c1 = dadd(c1, dmul(v2.x, x01));
                c0 = dadd(c0, dmul(v2.y, x10));
                c1 = dadd(c1, dmul(v2.y, x11));
            };
            auto one_slot = [&](const uint32_t* lp, const double2* vp, double& c0, double& c1) {
                const unsigned lo0 = *reinterpret_cast<const uint16_t*>(lp);
                const double v = *reinterpret_cast<const double*>(vp);
                c0 = dadd(c0, dmul(v, *reinterpret_cast<const double*>(in0 + lo0)));
                c1 = dadd(c1, dmul(v, *reinterpret_cast<const double*>(in0 + lo0 + 8 * kCP * 8)));
            };
            int j = 0;
#pragma unroll 1
            for (; j < nmin; j++) {
                two_slots(lpA + 4 * j, vpA + 4 * j, a0, a1);
                two_slots(lpB + 4 * j, vpB + 4 * j, b0, b1);
            }
#pragma unroll 1
            for (int i = j; i < nA; i++) two_slots(lpA + 4 * i, vpA + 4 * i, a0, a1);
#pragma unroll 1
            for (int i = j; i < nB; i++) two_slots(lpB + 4 * i, vpB + 4 * i, b0, b1);
            if (!RG_SKIP_COMPUTE) {
                if (oddA) one_slot(lpA + 4 * nA, vpA + 4 * nA, a0, a1);  // odd slot count: the last slot on its own
                if (oddB) one_slot(lpB + 4 * nB, vpB + 4 * nB, b0, b1);
            }
            acc[0][0] = a0; acc[0][1] = a1; acc[1][0] = b0; acc[1][1] = b1;
        }
        if (out_off == 0) __syncthreads();  // everyone is done reading in_s[buf]: slots [0,128) now take the outputs
#pragma unroll
        for (int k = 0; k < 2; k++) {
            in[t * kOutStride + out_off + o_local[k]] = acc[k][0];
            in[(t + 8) * kOutStride + out_off + o_local[k]] = acc[k][1];
        }
        __syncthreads();
        // tail staging: the footprint slots are free now, the outputs sit beyond them -> refill at once.  (The
        // outputs of this sub-block are read below; the next write to this tail happens two sub-blocks later,
        // after the barrier of the next sub-block, which every warp reaches only after its write-out here.)
        if (out_off != 0 && s + RG_NBUF < nsub) prefetch(f0 + RG_NBUF * kT, buf);
        // ---- write-out: warp w stores frame w; 16 B per lane, two tile rows (2 x 256 B) per instruction ----
        {
            const int64_t f = f0 + tt_p;
            if (f < f_end && !RG_SKIP_STORE) {
                if (vec_store) {
                    const int hr = lane >> 4, c2 = (lane & 15) * 2;  // half-warp -> tile row, lane -> 2 cells
                    double* o = vout + f * n_out + out_base + (int64_t)hr * w_out + c2;
                    const double* si = in + tt_p * kOutStride + out_off + hr * kTW + c2;
                    const double2 r0 = *reinterpret_cast<const double2*>(si);
                    const double2 r1 = *reinterpret_cast<const double2*>(si + 2 * kTW);
                    *reinterpret_cast<double2*>(o) = r0;
                    if (hr + 2 < th) *reinterpret_cast<double2*>(o + 2 * w_out) = r1;
                } else if (lane < tw) {
                    double* o = vout + f * n_out + out_base + lane;
                    const double* si = in + tt_p * kOutStride + out_off + lane;
                    for (int tr = 0; tr < th; tr++) o[(int64_t)tr * w_out] = si[tr * kTW];
                }
            }
        }
        if (out_off == 0) {
            // Frame row w of this buffer holds the outputs of frame w, which only warp w reads (just above), and warp
            // w is also the one that refills row w: program order inside the warp is enough, no block barrier.  (The
            // loads above have completed -- their values were stored -- before the asynchronous copies are issued;
            // nobody else touches the buffer before the "filled" mbarrier of sub-block s + 2.)
            if (!RG_WARP_REFILL) __syncthreads();  // outputs consumed: the buffer may be refilled
            if (s + RG_NBUF < nsub) prefetch(f0 + RG_NBUF * kT, buf);
        }
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

// layout of the caller's tile_info buffer: [n_tiles records][4 counters][n_tiles slot counts]
// [n_tiles + 1 slot bases (int64)][scan scratch (int64)]
struct PlanLayout {
    int64_t n_tiles, counter, counts, base, scratch, total_ints;
};
static PlanLayout plan_layout(int64_t n_tiles)
{
    PlanLayout L;
    L.n_tiles = n_tiles;
    L.counter = n_tiles * kTileInfoInts;
    L.counts = L.counter + 4;
    L.base = (L.counts + n_tiles + 1) / 2 * 2;  // int64-aligned (the buffer itself is >= 8-byte aligned)
    L.scratch = L.base + 2 * (n_tiles + 1);
    L.total_ints = L.scratch + 2 * (int64_t)scan_scratch_elems(n_tiles);
    return L;
}

extern "C" int rg_apply_plan_sizes(int64_t h_out, int64_t w_out, int64_t* n_tiles_host,
                                   int64_t* tile_info_ints_host, int64_t* tile_rows_ints_host)
{
    if (h_out <= 0 || w_out <= 0 || !n_tiles_host || !tile_info_ints_host || !tile_rows_ints_host)
        return fail(RG_E_ARG, "rg_apply_plan_sizes: bad argument");
    const int64_t n = tiles_of(h_out, w_out, nullptr);
    *n_tiles_host = n;
    *tile_info_ints_host = plan_layout(n).total_ints;
    *tile_rows_ints_host = n * kCopyPairs;
    return RG_OK;
}

extern "C" int rg_apply_plan_build(int device, void* stream, int64_t nnz,
                                   int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out,
                                   const int32_t* row_ptr, const int32_t* col,
                                   int32_t* tile_info, int32_t* tile_rows, uint16_t* lidx,
                                   int64_t* n_generic_tiles_host, int64_t* n_slots_host)
{
    if (h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0 || !row_ptr || !tile_info || !tile_rows || !n_generic_tiles_host ||
        !n_slots_host)
        return fail(RG_E_ARG, "rg_apply_plan_build: bad argument");
    if (nnz > 0 && (!col || !lidx)) return fail(RG_E_ARG, "rg_apply_plan_build: null pointer");
    if (h_in * w_in >= INT32_MAX || h_out * w_out >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_apply_plan_build: too large");
    if ((uintptr_t)tile_info % 8 != 0) return fail(RG_E_ARG, "rg_apply_plan_build: tile_info must be 8-byte aligned");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x;
    const int64_t n_tiles = tiles_of(h_out, w_out, &tiles_x);
    const PlanLayout L = plan_layout(n_tiles);
    int32_t* counter = tile_info + L.counter;
    int32_t* counts = tile_info + L.counts;
    int64_t* base = reinterpret_cast<int64_t*>(tile_info + L.base);
    int64_t* scratch = reinterpret_cast<int64_t*>(tile_info + L.scratch);
    RG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t) * 4, st));
    const int pad_even = (w_in % 2 == 0) && ((h_in * w_in) % 2 == 0);
    k_plan_tiles<<<(unsigned)n_tiles, 128, 0, st>>>(h_in, w_in, h_out, w_out, tiles_x, pad_even, row_ptr, col,
                                                    tile_info, tile_rows, lidx, counter, counts);
    RG_LAUNCH_CHECK("k_plan_tiles");
    const int rc = exclusive_scan_i32_i64(st, counts, base, n_tiles, scratch);
    if (rc != RG_OK) return rc;
    k_plan_base<<<(unsigned)ceil_div(n_tiles, 256), 256, 0, st>>>(n_tiles, base, tile_info);
    RG_LAUNCH_CHECK("k_plan_base");
    int32_t n_generic = 0;
    int64_t n_slots = 0;
    RG_CUDA(cudaMemcpyAsync(&n_generic, counter, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(&n_slots, base + n_tiles, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    *n_generic_tiles_host = n_generic;
    *n_slots_host = n_slots;
    return RG_OK;
}

extern "C" int rg_apply_plan_slots(int device, void* stream, int64_t h_out, int64_t w_out,
                                   const int32_t* row_ptr, const double* val,
                                   const int32_t* tile_info, const uint16_t* lidx,
                                   int64_t n_slots, double* slot_val, uint16_t* slot_lidx)
{
    if (h_out <= 0 || w_out <= 0 || n_slots < 0 || !row_ptr || !tile_info)
        return fail(RG_E_ARG, "rg_apply_plan_slots: bad argument");
    if (n_slots == 0) return RG_OK;
    if (!val || !lidx || !slot_val || !slot_lidx) return fail(RG_E_ARG, "rg_apply_plan_slots: null pointer");
    if ((uintptr_t)slot_val % 16 != 0 || (uintptr_t)slot_lidx % 16 != 0)
        return fail(RG_E_ARG, "rg_apply_plan_slots: slot arrays must be 16-byte aligned");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x;
    const int64_t n_tiles = tiles_of(h_out, w_out, &tiles_x);
    k_plan_slots<<<(unsigned)n_tiles, kPairs, 0, st>>>(h_out, w_out, tiles_x, row_ptr, val, tile_info, lidx, slot_val, slot_lidx);
    RG_LAUNCH_CHECK("k_plan_slots");
    return RG_OK;
}

extern "C" int rg_apply_planned(int device, void* stream, int64_t n_frames,
                                int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out,
                                const int32_t* row_ptr, const int32_t* col, const double* val,
                                const int32_t* tile_info, const int32_t* tile_rows,
                                const double* slot_val, const uint16_t* slot_lidx,
                                int64_t n_generic_tiles,
                                const double* values_in, double* values_out)
{
    if (n_frames == 0) return RG_OK;  // nothing to do (empty tensors have null data pointers)
    if (n_frames < 0 || h_in <= 0 || w_in <= 0 || h_out <= 0 || w_out <= 0 || !row_ptr || !tile_info || !tile_rows ||
        !values_in || !values_out)
        return fail(RG_E_ARG, "rg_apply_planned: bad argument");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int tiles_x;
    const int64_t n_tiles = tiles_of(h_out, w_out, &tiles_x);
    const int64_t n_in = h_in * w_in, n_out = h_out * w_out;
    static_assert(sizeof(StagedSmem) <= (228 / RG_CTAS - 1) * 1024, "the staged CTAs must fit in one SM's shared memory");
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
                nf, w_in, n_in, h_out, w_out, tiles_x, tiles_y, tile_info, tile_rows, slot_val, slot_lidx,
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
