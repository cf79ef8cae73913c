// rg_apply_bulk.cu -- shared-weights apply, staged through shared memory with bulk (TMA) copies: the HBM-roofline path.
//
// Same arithmetic as k_apply_csr (rg_apply.cu): each output cell accumulates val * in[col] over its CSR row in
// ascending input index from +0.0 with separately rounded multiply and add
// (regridding/_regrid/_regrid_from_weights.py:179-182), so the result is bit-identical; only the data movement differs.
//
// A CTA owns a TILE of kTH x 32 output cells and walks up to kFB frames in sub-blocks of kT = 8 frames:
//   * the input cells the tile references (its FOOTPRINT: per input row one contiguous span, computed once per
//     weights by rg_apply_plan_build) are brought into shared memory by kPW PRODUCER warps with one
//     cp.async.bulk (SASS UBLKCP, the TMA engine) per (frame, input row), completion on an mbarrier with
//     expect-tx; kNST stages, released by the consumers through "empty" mbarriers.  The copies never touch the
//     LSU / shared-memory store path, which the old per-lane cp.async (LDGSTS) fill saturated
//     (profiles/r1_apply_tuning.md: 15 shared-memory wavefronts per LDGSTS against an ideal 4);
//   * kTH CONSUMER warps, one per tile row: a quarter-warp owns one output cell, LANE = FRAME, so every gather
//     in_s[frame][cell] is a conflict-free shared-memory read (frame stride 16 B-odd), the CSR entry (byte offset +
//     weight, laid out by the plan) is a broadcast read, and there is no divergence inside a cell;
//   * a warp stages the 8 x 32 results of its row in its private slice of shared memory and writes them with
//     16-byte coalesced stores -- no block-wide barrier anywhere in the main loop: warps drift freely between the
//     full / empty mbarriers.
// The tile's slot arrays (weights + u16 byte offsets in consumption order) come in with two bulk copies.
// Tiles whose footprint does not fit (very different resolutions, scattered weights) are flagged by the plan and
// handled by the generic per-cell kernel.
//
// Odd sizes: bulk copies move 16-byte pieces from 16-byte aligned addresses.  Footprint spans are therefore taken
// in FLAT cell index space (even start, even length).  When n_in is odd, every odd frame starts 8 bytes off: its
// copies start one cell earlier (spans carry 2 spare cells for that) and the frame's lanes add 8 bytes to their
// gather addresses -- no fallback to the generic kernel for odd widths.
#include "rg_common.cuh"
#include "rg_async.cuh"

// development switches (ablations for profiles/r2_apply_tuning.md); all 0 in the product build
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
#ifndef RG_TH
#define RG_TH 16
#endif
#ifndef RG_PW
#define RG_PW 8
#endif
#ifndef RG_CP
#define RG_CP 1218
#endif
#ifndef RG_NST
#define RG_NST 2
#endif
#ifndef RG_L2PF
#define RG_L2PF 0   // producer warps prefetch the footprint of the sub-block this many ahead into L2 (0: off)
#endif
#ifndef RG_CHAINS
#define RG_CHAINS 4   // quads of a row walked together: independent accumulation chains per lane (2 or 4)
#endif
#ifndef RG_REGS_CONSUMER
#define RG_REGS_CONSUMER 96   // setmaxnreg: the consumer warps take the registers the producer warps do not need
#endif
#ifndef RG_REGS_PRODUCER
#define RG_REGS_PRODUCER 40
#endif
// the CTA's register pool is what the launch allocated (kBulkThreads x the launch-time count): the two budgets must
// fit in it or the consumers' setmaxnreg.inc waits forever
#ifndef RG_STORE_CS
#define RG_STORE_CS 0   // results are written with streaming (evict-first) stores: they are never read again, the footprint
#endif                  // halos the neighbouring tiles share should own the L2
#ifndef RG_FAKE_STORE
#define RG_FAKE_STORE 0   // ablation: full-width global stores of register values, no shared-memory staging (WRONG results)
#endif
#ifndef RG_DIRECT_STORE
#define RG_DIRECT_STORE 0   // experiment: 8-byte stores straight from the accumulators (no shared-memory staging)
#endif
#ifndef RG_PATCH_ROWS
#define RG_PATCH_ROWS 12
#endif
#ifndef RG_PATCH_COLS
#define RG_PATCH_COLS 12
#endif

namespace rg {

constexpr int kTH = RG_TH;       // tile height (output rows) = consumer warps
constexpr int kTW = 32;          // tile width  (output cols) = one full coalesced row of 256 B
constexpr int kTileCells = kTH * kTW;
constexpr int kQuads = kTileCells / 4;   // a quad = 4 output cells = the 4 quarter-warps of a warp instruction
constexpr int kPairs = kTileCells / 2;   // a pair = 2 consecutive output cells = the 2 quarter-warps of a half-warp
constexpr int kPairsRow = kTW / 2, kQuadsRow = kTW / 4;
constexpr int kT = 8;            // frames per stage: one per lane of a quarter-warp
constexpr int kNST = RG_NST;     // stages
constexpr int kCW = kTH;         // consumer warps (one per tile row)
constexpr int kPW = RG_PW;       // producer warps (a UBLKCP costs its warp ~60 cycles: the issue rate needs several)
constexpr int kBulkThreads = (kCW + kPW) * 32;
constexpr int kFB = RG_FB;       // frames per CTA
constexpr int kRMAX = 64;        // max input rows in a footprint (2 per producer lane)
constexpr int kCP = RG_CP;       // doubles per staged frame; kCP/2 odd => the 8 frame lanes of a quarter-warp hit 8
                                 // distinct 16-byte bank groups
constexpr int kCellsMax = kCP - 4;
constexpr int kZeroEven = kCP - 4;   // slots kCP-4 .. kCP-1 of every staged frame hold 0.0: the targets of padding
constexpr int kZeroOdd = kCP - 3;    // entries, one per bank parity (+1 for frames staged one cell later)
constexpr int kPadMax = kTH * 288;   // max SLOT entries per tile (4 cells x padded, aligned slots of every quad)
constexpr int kNnzMax = kPadMax;
constexpr int kAlignMax = 16;    // rows up to this length take part in the bank-parity alignment
constexpr int kOS = kTW;         // doubles per staged output frame row (warp-private staging, XOR-swizzled columns)
constexpr int kPatch = RG_PATCH_COLS;      // tiles are issued in patches of kPatchRows x kPatch tiles (~ one wave of
constexpr int kPatchRows = RG_PATCH_ROWS;  // CTAs) so that footprint halos are shared through L2
static_assert((kCP / 2) % 2 == 1 && kCP % 2 == 0, "kCP/2 must be odd");
static_assert(kT == 8, "lane & 7 = frame");
constexpr int kLaunchRegs = (65536 / ((kCW + kPW) * 32)) / 8 * 8;   // what __launch_bounds__(threads, 1) lets ptxas allocate
static_assert(kCW * RG_REGS_CONSUMER + kPW * RG_REGS_PRODUCER <= (kCW + kPW) * kLaunchRegs, "setmaxnreg budgets exceed the CTA's pool");
static_assert(kCW % 4 == 0 && kPW % 4 == 0, "setmaxnreg works on warpgroups (4 warps)");
static_assert(kTW == 32 && kPadMax % 8 == 0, "layout");

// per-tile plan record (int32): [0] input rows in the footprint (0: no weights reach the tile, -1: generic kernel),
// [1] staged cells per frame, [2] slot entries, [3] plain layout?, [4..5] slot base (int64), [6..7] spare; then
// quad_beg[kQuads + 2] as uint16, then pair_of[kPairs] as uint8: quad k of tile row r is made of the pairs
// pair_of[16 r + 2k], pair_of[16 r + 2k + 1] of that row (pairs sorted by slot count inside a row, so that the two
// pairs of a quad -- which share one trip count -- have nearly equal lengths)
constexpr int kQuadWords = (kQuads + 2) / 2;
constexpr int kTileInfoInts = 8 + kQuadWords + kPairs / 4;
// per-tile row table (int32 pairs): source offset (doubles, inside a frame, even) and dst | len << 16 (doubles, even)
constexpr int kRowInts = 2 * kRMAX;

// Slot layout of a quad (4 cells c, L slots w, L even):  entry(w, c) = qb + (w >> 1) * 8 + c * 2 + (w & 1),
// so the two values (16 B) and the two offsets (4 B) of slots (w, w+1) of a cell are one shared-memory load each.
struct BulkSmem {
    double in_s[kNST][kT * kCP];          // [stage][frame][cell]
    double out_s[kCW][kT / 2 * kOS];      // [consumer warp][frame of a half sub-block][cell of the warp's tile row]
    alignas(16) double val[kPadMax];
    alignas(16) uint16_t lidx[kPadMax];   // BYTE offset of the referenced cell inside a staged frame
    uint16_t quad_beg[kQuads + 2];
    uint8_t pair_of[kPairs];
    alignas(8) uint64_t full[kNST];       // "stage filled": kPW arrivals (each with the bytes of its frame)
    uint64_t ready[kNST];                 // "stage filled", relayed by consumer warp 0: what the other consumers wait on
    uint64_t empty[kNST];                 // "stage consumed": kCW arrivals
    uint64_t slots_ready;                 // the tile's slot arrays have landed
};
static_assert(sizeof(BulkSmem) <= 227 * 1024, "the CTA must fit in one SM's shared memory");

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
__global__ void __launch_bounds__(kPairs)
k_plan_tiles(int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out, int tiles_x, int spare,
             const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
             int32_t* __restrict__ tile_info, int32_t* __restrict__ tile_rows, uint16_t* __restrict__ lidx,
             int32_t* __restrict__ n_generic, int32_t* __restrict__ tile_slots)
{
    __shared__ int s_rmin, s_rmax, s_nnz, s_pad, s_plain;
    __shared__ uint16_t s_pair_len[kPairs], s_quad_beg[kQuads + 2];
    __shared__ uint8_t s_pair_of[kPairs];  // (tile row, rank) -> pair inside that row
    __shared__ int s_clo[kRMAX], s_chi[kRMAX], s_off[kRMAX + 1], s_start[kRMAX];
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
    bool generic = nrows > kRMAX || nnz > kNnzMax || s_pad > kPadMax;
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
                    // spans in FLAT cell index space: even start, even length (16-byte pieces from 16-byte aligned
                    // addresses); `spare` = 2 more cells when odd frames are staged one cell later (odd n_in)
                    const int64_t a = ((int64_t)(rmin + r) * w_in + s_clo[r]) & ~(int64_t)1;
                    const int64_t b = ((int64_t)(rmin + r) * w_in + s_chi[r]) | 1;
                    s_start[r] = (int)a;
                    off += (int)(b - a + 1) + spare;
                } else {
                    s_start[r] = 0;
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
                const int ci = (int)(c / w_in);
                lidx[w] = (uint16_t)(s_off[ci - rmin] + (c - s_start[ci - rmin]));
            }
        }
        // row table of the producer warps
        for (int r = threadIdx.x; r < kRMAX; r += blockDim.x) {
            int32_t src = 0, dl = 0;
            if (r < nrows && s_off[r + 1] > s_off[r]) {
                src = s_start[r];
                dl = s_off[r] | ((s_off[r + 1] - s_off[r]) << 16);
            }
            tile_rows[(int64_t)tile * kRowInts + 2 * r] = src;
            tile_rows[(int64_t)tile * kRowInts + 2 * r + 1] = dl;
        }
    }
    // slot layout: the two cells of a half-warp are aligned for bank parity (align_pair); the two pairs of a quad
    // share one even slot count.  If the aligned layout is too long for the tile, fall back to the unaligned one.
    int slots = 0, plain = RG_NO_ALIGN;
    if (!generic && nnz) {
        __syncthreads();  // lidx of this tile is complete
        const int64_t out_base = ((int64_t)ty * kTH) * w_out + (int64_t)tx * kTW;
        {
            const PairLayout P = pair_layout(threadIdx.x, th, tw, out_base, w_out, row_ptr, lidx, !RG_NO_ALIGN);
            s_pair_len[threadIdx.x] = (uint16_t)min(P.L, 65535);
        }
        __syncthreads();
        // pairs of one tile row sorted by slot count (stable): (row, rank) -> pair.  Quad k of a row = ranks 2k, 2k+1.
        auto rank_pairs = [&]() {
            const int row0 = (threadIdx.x / kPairsRow) * kPairsRow;
            const int me = s_pair_len[threadIdx.x];
            int r = 0;
            for (int k = row0; k < row0 + kPairsRow; k++) {
                const int o = s_pair_len[k];
                r += (o < me) || (o == me && k < (int)threadIdx.x);
            }
            s_pair_of[row0 + r] = (uint8_t)(threadIdx.x - row0);
            __syncthreads();
        };
        auto quad_max = [&](int qd) {
            const int row0 = (qd / kQuadsRow) * kPairsRow, k = qd % kQuadsRow;
            return max((int)s_pair_len[row0 + s_pair_of[row0 + 2 * k]], (int)s_pair_len[row0 + s_pair_of[row0 + 2 * k + 1]]);
        };
        rank_pairs();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int qd = 0; qd < kQuads; qd++) acc += 4 * ((quad_max(qd) + 1) & ~1);
            s_plain = (acc > kPadMax) ? 1 : RG_NO_ALIGN;
        }
        __syncthreads();
        plain = s_plain;
        if (plain) {
            const PairLayout P = pair_layout(threadIdx.x, th, tw, out_base, w_out, row_ptr, lidx, false);
            s_pair_len[threadIdx.x] = (uint16_t)min(P.L, 65535);
            __syncthreads();
            rank_pairs();
        }
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int qd = 0; qd < kQuads; qd++) {
                // offsets are multiples of 8; an odd slot count is rounded up with one padding slot
                const int m = quad_max(qd);
                s_quad_beg[qd] = (uint16_t)min(acc, 65528);
                acc += 4 * ((m + 1) & ~1);
            }
            s_quad_beg[kQuads] = (uint16_t)min(acc, 65528);
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
        info[0] = generic ? -1 : nrows;
        info[1] = (generic || !nnz) ? 0 : s_off[nrows];
        info[2] = slots;
        info[3] = plain;
        info[4] = info[5] = 0;  // slot base: k_plan_base
        info[6] = nnz;
        info[7] = 0;
        tile_slots[tile] = slots;
        if (generic) atomicAdd(n_generic, 1);
    }
    (void)h_in;
}

__global__ void k_plan_base(int64_t n_tiles, const int64_t* __restrict__ base, int32_t* __restrict__ tile_info)
{
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    memcpy(tile_info + tile * kTileInfoInts + 4, base + tile, sizeof(int64_t));
}

// slot arrays of every staged tile: entry(w, c) of quad qd at base + quad_beg[qd] + (w >> 1) * 8 + c * 2 + (w & 1)
__global__ void __launch_bounds__(kPairs)
k_plan_slots(int64_t h_out, int64_t w_out, int tiles_x, const int32_t* __restrict__ row_ptr, const double* __restrict__ val,
             const int32_t* __restrict__ tile_info, const uint16_t* __restrict__ lidx,
             double* __restrict__ slot_val, uint16_t* __restrict__ slot_lidx)
{
    const int tile = blockIdx.x;
    const int32_t* info = tile_info + (int64_t)tile * kTileInfoInts;
    if (info[0] <= 0 || info[2] <= 0) return;
    const int ty = tile / tiles_x, tx = tile % tiles_x;
    const int th = (int)min((int64_t)kTH, h_out - (int64_t)ty * kTH);
    const int tw = (int)min((int64_t)kTW, w_out - (int64_t)tx * kTW);
    const int64_t out_base = ((int64_t)ty * kTH) * w_out + (int64_t)tx * kTW;
    int64_t base;
    memcpy(&base, info + 4, sizeof(int64_t));
    const uint16_t* quad_beg = reinterpret_cast<const uint16_t*>(info + 8);
    const PairLayout P = pair_layout(threadIdx.x, th, tw, out_base, w_out, row_ptr, lidx, info[3] == 0);
    const uint8_t* pair_of = reinterpret_cast<const uint8_t*>(info + 8 + kQuadWords);
    const int row0 = (threadIdx.x / kPairsRow) * kPairsRow, me = threadIdx.x - row0;
    int rank = 0;
    for (int k = 0; k < kPairsRow; k++)
        if (pair_of[row0 + k] == me) rank = k;
    const int qd = (row0 / kPairsRow) * kQuadsRow + (rank >> 1), cA = (rank & 1) * 2;  // my cells are cells cA, cA + 1 of quad qd
    const int qb = quad_beg[qd], Lq = ((int)quad_beg[qd + 1] - qb) >> 2;
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
// bulk-copy staged apply
// ---------------------------------------------------------------------------
// vin must be 16-byte aligned.  `cells_left` = doubles from vin to the end of the caller's values_in buffer: copies
// are clipped to it (only the spare cells of the very last rows can reach beyond).
__global__ void __launch_bounds__(kBulkThreads, 1)
k_apply_bulk(int64_t n_frames, int64_t n_in, int64_t cells_left, int64_t h_out, int64_t w_out, int tiles_x, int tiles_y,
             const int32_t* __restrict__ tile_info, const int32_t* __restrict__ tile_rows,
             const double* __restrict__ slot_val, const uint16_t* __restrict__ slot_lidx,
             const double* __restrict__ vin, double* __restrict__ vout)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BulkSmem& S = *reinterpret_cast<BulkSmem*>(smem_raw);
    int ty, tx;
    tile_of_block(blockIdx.x, tiles_x, tiles_y, ty, tx);
    const int tile = ty * tiles_x + tx;
    const int32_t* info = tile_info + (int64_t)tile * kTileInfoInts;
    const int nrows = info[0];
    if (nrows < 0) return;  // handled by the generic kernel
    const int th = (int)min((int64_t)kTH, h_out - (int64_t)ty * kTH);
    const int tw = (int)min((int64_t)kTW, w_out - (int64_t)tx * kTW);
    const int64_t n_out = h_out * w_out;
    const int64_t f_begin = (int64_t)blockIdx.y * kFB;
    const int64_t f_end = min(n_frames, f_begin + kFB);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t out_base = ((int64_t)ty * kTH) * w_out + (int64_t)tx * kTW;
    const bool out_aligned = ((uintptr_t)vout & 15) == 0;

    if (nrows == 0) {
        // no weights reach this tile: the reference leaves zeros (rfw.py:111-118)
        constexpr int NW = kBulkThreads / 32;
        for (int64_t f = f_begin + warp; f < f_end; f += NW) {
            for (int tr = 0; tr < th; tr++) {
                const int64_t o = f * n_out + out_base + (int64_t)tr * w_out;
                if (tw == kTW && out_aligned && (o & 1) == 0) {
                    if (lane < 16) {
                        if (RG_STORE_CS) __stcs(reinterpret_cast<double2*>(vout + o + 2 * lane), make_double2(0.0, 0.0));
                        else *reinterpret_cast<double2*>(vout + o + 2 * lane) = make_double2(0.0, 0.0);
                    }
                } else if (lane < tw) {
                    vout[o + lane] = 0.0;
                }
            }
        }
        return;
    }

    const int nslots = info[2];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kNST; s++) { mbar_init(&S.full[s], kPW); mbar_init(&S.empty[s], kCW); mbar_init(&S.ready[s], 1); }
        mbar_init(&S.slots_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // the always-zero slots at the end of every staged frame (the bulk copies never reach them)
    if (threadIdx.x < kNST * kT * 4) {
        const int st = threadIdx.x / (kT * 4), t = (threadIdx.x / 4) % kT;
        S.in_s[st][t * kCP + kCellsMax + (threadIdx.x & 3)] = 0.0;
    }
    {
        const uint16_t* qsrc = reinterpret_cast<const uint16_t*>(info + 8);
        for (int k = threadIdx.x; k < kQuads + 2; k += kBulkThreads) S.quad_beg[k] = qsrc[k];
        for (int k = threadIdx.x; k < kPairs / 4; k += kBulkThreads)
            reinterpret_cast<int32_t*>(S.pair_of)[k] = info[8 + kQuadWords + k];
    }
    __syncthreads();

    const int nsub = (int)((f_end - f_begin + kT - 1) / kT);
    const int odd_in = (int)(n_in & 1);   // odd frames start 8 bytes off: staged one cell later

    if (warp >= kCW) {
        // =============================== producers ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(RG_REGS_PRODUCER));
        const int pw = warp - kCW;
        if (pw == 0 && lane == 0) {
            int64_t base;
            memcpy(&base, info + 4, sizeof(int64_t));
            mbar_arrive_expect_tx(&S.slots_ready, (unsigned)nslots * 10u);
            bulk_load(smem_u32(S.val), slot_val + base, (unsigned)nslots * 8u, &S.slots_ready);
            bulk_load(smem_u32(S.lidx), slot_lidx + base, (unsigned)nslots * 2u, &S.slots_ready);
        }
        // kPW <= kT: a warp owns kT / kPW frames of a stage; kPW > kT: kPW / kT warps share one frame (rows dealt out)
        constexpr int kFW = kPW > kT ? kT : kPW;     // frame groups
        constexpr int kPartW = kPW / kFW;            // warps per frame
        static_assert(kT % kFW == 0 && kPW % kFW == 0 && (kRMAX / 32) % kPartW == 0, "producer warps must tile frames x rows");
        constexpr int RPL = kRMAX / 32 / kPartW;
        const int fw = pw % kFW, part = pw / kFW;
        int32_t src[RPL];
        unsigned dst[RPL], len[RPL];
        unsigned row_bytes = 0;
#pragma unroll
        for (int j = 0; j < RPL; j++) {
            const int r = (lane + 32 * j) * kPartW + part;
            const int32_t a = tile_rows[(int64_t)tile * kRowInts + 2 * r];
            const unsigned dl = (r < nrows) ? (unsigned)tile_rows[(int64_t)tile * kRowInts + 2 * r + 1] : 0u;
            src[j] = a;
            dst[j] = (dl & 0xffffu) * 8u;
            len[j] = (dl >> 16) * 8u;
            row_bytes += len[j];
        }
        const unsigned frame_bytes = __reduce_add_sync(0xffffffffu, row_bytes);   // bytes this warp stages per frame
        // this lane's rows in the first frame of this producer warp; advanced by kT frames per sub-block
        const double* p0 = vin + (f_begin + fw) * n_in;
        const int64_t stage_step = (int64_t)kT * n_in;
        for (int s = 0; s < nsub; s++, p0 += stage_step) {
            const int st = s % kNST;
            if (s >= kNST) mbar_wait(&S.empty[st], (unsigned)((s / kNST - 1) & 1));
            const unsigned sbase = smem_u32(S.in_s[st]) + fw * (kCP * 8);
            const int64_t f0 = f_begin + (int64_t)s * kT + fw;
            if (!odd_in) {
                // even n_in: every frame is 16-byte aligned and no span reaches beyond its frame
                int nfr = 0;
#pragma unroll
                for (int i = 0; i < kT / kFW; i++) nfr += (f0 + i * kFW < f_end && !RG_SKIP_LOAD) ? 1 : 0;
                if (lane == 0) mbar_arrive_expect_tx(&S.full[st], frame_bytes * nfr);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kT / kFW; i++) {
                    if (f0 + i * kFW < f_end && !RG_SKIP_LOAD) {
                        const double* pf = p0 + (int64_t)i * kFW * n_in;
#pragma unroll
                        for (int j = 0; j < RPL; j++)
                            if (len[j]) bulk_load(sbase + i * kFW * (kCP * 8) + dst[j], pf + src[j], len[j], &S.full[st]);
                    }
                }
            } else {
                // odd n_in: odd frames are staged from one cell earlier; copies are clipped to the caller's buffer
                unsigned mine = 0;
#pragma unroll
                for (int i = 0; i < kT / kFW; i++) {
                    const int64_t f = f0 + i * kFW;
                    if (f >= f_end || RG_SKIP_LOAD) continue;
                    const int64_t fs = f * n_in - (f & 1);   // first double of the (shifted) frame
#pragma unroll
                    for (int j = 0; j < RPL; j++) {
                        const int64_t room = cells_left - (fs + src[j]);
                        mine += (unsigned)max((int64_t)0, min((int64_t)len[j], room * 8));
                    }
                }
                const unsigned total = __reduce_add_sync(0xffffffffu, mine);
                if (lane == 0) mbar_arrive_expect_tx(&S.full[st], total);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kT / kFW; i++) {
                    const int64_t f = f0 + i * kFW;
                    if (f >= f_end || RG_SKIP_LOAD) continue;
                    const int64_t fs = f * n_in - (f & 1);
#pragma unroll
                    for (int j = 0; j < RPL; j++) {
                        const int64_t room = cells_left - (fs + src[j]);
                        const unsigned n = (unsigned)max((int64_t)0, min((int64_t)len[j], room * 8));
                        if (n) bulk_load(sbase + i * kFW * (kCP * 8) + dst[j], vin + fs + src[j], n, &S.full[st]);
                    }
                }
            }
        }
        return;
    }

    // =============================== consumers ===============================
    // warp = tile row; quarter-warp q = one output cell of a quad; lane & 7 = frame
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(RG_REGS_CONSUMER));
    const int q = lane >> 3, t = lane & 7;
    mbar_wait(&S.slots_ready, 0);
    // Per CTA, not per sub-block: this quarter-warp's output cell (column inside the tile row) in each of the row's
    // quads, and the quads' slot ranges (first entry | slot pairs << 16)
    // Staging slice: frame row (t & 3), column (cell ^ 2 (t & 3)): the 16 lanes of a half-warp STS (2 cells of a pair x
    // up to 8 frames) then fall into 16 different banks, and a pair of cells stays a contiguous 16 bytes.
    int o_col[kQuadsRow];
    unsigned q_rng[kQuadsRow];
#pragma unroll
    for (int k = 0; k < kQuadsRow; k++) {
        o_col[k] = (2 * (int)S.pair_of[warp * kPairsRow + 2 * k + (q >> 1)] + (q & 1)) ^ (2 * (t & 3));
        const unsigned beg = S.quad_beg[warp * kQuadsRow + k], end = S.quad_beg[warp * kQuadsRow + k + 1];
        q_rng[k] = beg | (((end - beg) >> 3) << 16);
    }
    const char* val_q = reinterpret_cast<const char*>(S.val) + q * 16;    // entry(w, c): see the slot layout above
    const char* lidx_q = reinterpret_cast<const char*>(S.lidx) + q * 4;
    double* outw = S.out_s[warp];
    // frames of this lane that start 8 bytes off are staged one cell later (f_begin is even)
    const unsigned shift = (unsigned)(odd_in & t & 1) * 8u;
    const bool row_live = warp < th;
    // write-out: half-warp -> frame, lane -> 2 cells
    const int hf = lane >> 4, c2 = (lane & 15) * 2;
    const int64_t row_off = out_base + (int64_t)warp * w_out;
    const bool row_vec = tw == kTW && out_aligned;   // 16-byte stores where the frame's row starts on an even double
    double* op = vout + (f_begin + hf) * n_out + row_off + c2;   // advanced by kT frames per sub-block
    const int64_t out_step = (int64_t)kT * n_out;
    for (int s = 0; s < nsub; s++, op += out_step) {
        const int st = s % kNST;
        const int64_t f0 = f_begin + (int64_t)s * kT;
        // A waiter on `full` is woken by every completing bulk copy (hundreds per stage): sixteen warps polling it
        // burn more issue slots than the whole computation.  So ONE warp watches `full` and relays the phase change
        // through `ready`, whose only event is that one arrival.
        if (warp == 0) {
            mbar_wait(&S.full[st], (unsigned)((s / kNST) & 1));
            if (lane == 0) mbar_arrive(&S.ready[st]);
        } else {
            mbar_wait(&S.ready[st], (unsigned)((s / kNST) & 1));
        }
        const char* in0 = reinterpret_cast<const char*>(S.in_s[st] + t * kCP) + shift;
        double acc[kQuadsRow];
#pragma unroll
        for (int k = 0; k < kQuadsRow; k++) acc[k] = 0.0;
        if (row_live && !RG_SKIP_COMPUTE) {
            // the row's quads are walked two at a time (neighbours in the length-sorted order, the second never
            // shorter than the first): two independent chains per lane
            auto two_slots = [&](const char* lp, const char* vp, double& c0) {
                const uint32_t l2 = *reinterpret_cast<const uint32_t*>(lp);   // offsets of slots 2j, 2j+1: one 4-byte load
                const double2 v2 = *reinterpret_cast<const double2*>(vp);     // weights of slots 2j, 2j+1: one 16-byte load
                const double x0 = *reinterpret_cast<const double*>(in0 + (l2 & 0xffffu));
                const double x1 = *reinterpret_cast<const double*>(in0 + (l2 >> 16));
                c0 = dadd(c0, dmul(v2.x, x0));
                c0 = dadd(c0, dmul(v2.y, x1));
            };
#if RG_CHAINS == 2
#pragma unroll
            for (int k = 0; k < kQuadsRow; k += 2) {
                const unsigned begA = q_rng[k] & 0xffffu, begB = q_rng[k + 1] & 0xffffu;
                const int nA = (int)(q_rng[k] >> 16), nB = (int)(q_rng[k + 1] >> 16);
                const char* lA = lidx_q + begA * 2;
                const char* vA = val_q + begA * 8;
                const char* lB = lidx_q + begB * 2;
                const char* vB = val_q + begB * 8;
                double a0 = 0.0, b0 = 0.0;
                int j = 0;
#pragma unroll 1
                for (; j < nA; j++, lA += 16, vA += 64, lB += 16, vB += 64) {
                    two_slots(lA, vA, a0);
                    two_slots(lB, vB, b0);
                }
#pragma unroll 1
                for (; j < nB; j++, lB += 16, vB += 64) two_slots(lB, vB, b0);
                acc[k] = a0;
                acc[k + 1] = b0;
            }
#else
            // four quads at a time (ascending slot counts n0 <= n1 <= n2 <= n3): four independent chains per lane
#pragma unroll
            for (int k = 0; k < kQuadsRow; k += 4) {
                const int n0 = (int)(q_rng[k] >> 16), n1 = (int)(q_rng[k + 1] >> 16), n2 = (int)(q_rng[k + 2] >> 16),
                          n3 = (int)(q_rng[k + 3] >> 16);
                const char* l0 = lidx_q + (q_rng[k] & 0xffffu) * 2;
                const char* v0 = val_q + (q_rng[k] & 0xffffu) * 8;
                const char* l1 = lidx_q + (q_rng[k + 1] & 0xffffu) * 2;
                const char* v1 = val_q + (q_rng[k + 1] & 0xffffu) * 8;
                const char* l2 = lidx_q + (q_rng[k + 2] & 0xffffu) * 2;
                const char* v2 = val_q + (q_rng[k + 2] & 0xffffu) * 8;
                const char* l3 = lidx_q + (q_rng[k + 3] & 0xffffu) * 2;
                const char* v3 = val_q + (q_rng[k + 3] & 0xffffu) * 8;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int j = 0;
#pragma unroll 1
                for (; j < n0; j++, l0 += 16, v0 += 64, l1 += 16, v1 += 64, l2 += 16, v2 += 64, l3 += 16, v3 += 64) {
                    two_slots(l0, v0, a0);
                    two_slots(l1, v1, a1);
                    two_slots(l2, v2, a2);
                    two_slots(l3, v3, a3);
                }
#pragma unroll 1
                for (; j < n1; j++, l1 += 16, v1 += 64, l2 += 16, v2 += 64, l3 += 16, v3 += 64) {
                    two_slots(l1, v1, a1);
                    two_slots(l2, v2, a2);
                    two_slots(l3, v3, a3);
                }
#pragma unroll 1
                for (; j < n2; j++, l2 += 16, v2 += 64, l3 += 16, v3 += 64) {
                    two_slots(l2, v2, a2);
                    two_slots(l3, v3, a3);
                }
#pragma unroll 1
                for (; j < n3; j++, l3 += 16, v3 += 64) two_slots(l3, v3, a3);
                acc[k] = a0;
                acc[k + 1] = a1;
                acc[k + 2] = a2;
                acc[k + 3] = a3;
            }
#endif
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[st]);  // this warp is done with the stage: the producers may refill it
        // ---- write-out of the warp's tile row: 8 frames x 256 B, staged four frames at a time in the warp's
        // private slice; 16 B per lane, two frames per store instruction ----
        if (RG_FAKE_STORE) {
            if (row_live) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < kQuadsRow; k += 2) { s0 += acc[k]; s1 += acc[k + 1]; }
#pragma unroll
                for (int i = 0; i < kT; i += 2) {
                    const int64_t f = f0 + i + hf;
                    if (f < f_end) *reinterpret_cast<double2*>(op + (int64_t)i * n_out) = make_double2(s0, s1);
                }
            }
        } else if (RG_DIRECT_STORE) {
            if (row_live && !RG_SKIP_STORE && f0 + t < f_end) {
                double* o = vout + (f0 + t) * n_out + row_off;
#pragma unroll
                for (int k = 0; k < kQuadsRow; k++) {
                    const int c = o_col[k] ^ (2 * (t & 3));
                    if (c < tw) o[c] = acc[k];
                }
            }
        } else if (row_live && !RG_SKIP_STORE) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if ((t >> 2) == h) {
#pragma unroll
                    for (int k = 0; k < kQuadsRow; k++) outw[(t & 3) * kOS + o_col[k]] = acc[k];
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < kT / 2; i += 2) {
                    const int64_t f = f0 + 4 * h + i + hf;
                    if (f < f_end) {
                        double* o = op + (int64_t)(4 * h + i) * n_out;
                        const double* si = outw + (i + hf) * kOS + (c2 ^ (2 * (i + hf)));
                        if (row_vec && (((f * n_out + row_off) & 1) == 0)) {
                            if (RG_STORE_CS) __stcs(reinterpret_cast<double2*>(o), *reinterpret_cast<const double2*>(si));
                            else *reinterpret_cast<double2*>(o) = *reinterpret_cast<const double2*>(si);
                        } else {
                            if (c2 < tw) o[0] = si[0];   // (a pair of cells is never split by the swizzle)
                            if (c2 + 1 < tw) o[1] = si[1];
                        }
                    }
                }
                __syncwarp();  // the staging slice is free again
            }
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
    if (tile_info[tile * kTileInfoInts] >= 0) return;
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
    *tile_rows_ints_host = n * kRowInts;
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
    const int spare = ((h_in * w_in) % 2 != 0) ? 2 : 0;   // odd frames are staged one cell later
    k_plan_tiles<<<(unsigned)n_tiles, kPairs, 0, st>>>(h_in, w_in, h_out, w_out, tiles_x, spare, row_ptr, col,
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
    const int tiles_y = (int)(n_tiles / tiles_x);
    if ((uintptr_t)values_in % 16 != 0) {
        // bulk copies need 16-byte aligned sources; a misaligned values pointer takes the generic kernel
        return rg_apply_csr(device, stream, n_frames, n_in, n_out, row_ptr, col, val, values_in, values_out);
    }
    // An odd number of doubles in values_in leaves its last cell without a 16-byte partner inside the buffer:
    // the last frame then goes through the generic kernel (n_in odd and n_frames odd only).
    int64_t staged_frames = n_frames;
    if ((n_in & 1) && (n_frames & 1)) staged_frames = n_frames - 1;
    RG_CUDA(cudaFuncSetAttribute(k_apply_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BulkSmem)));
    if (n_generic_tiles < n_tiles && staged_frames > 0) {
        const int64_t chunk = 65535LL * kFB;
        for (int64_t f = 0; f < staged_frames; f += chunk) {
            const int64_t nf = staged_frames - f < chunk ? staged_frames - f : chunk;
            dim3 grid((unsigned)n_tiles, (unsigned)ceil_div(nf, kFB));
            k_apply_bulk<<<grid, kBulkThreads, sizeof(BulkSmem), st>>>(
                nf, n_in, (staged_frames - f) * n_in, h_out, w_out, tiles_x, tiles_y, tile_info, tile_rows, slot_val,
                slot_lidx, values_in + f * n_in, values_out + f * n_out);
            RG_LAUNCH_CHECK("k_apply_bulk");
        }
    }
    if (n_generic_tiles > 0 && staged_frames > 0) {
        constexpr int FT = 8;
        const int64_t chunk = 65535LL * FT;
        for (int64_t f = 0; f < staged_frames; f += chunk) {
            const int64_t nf = staged_frames - f < chunk ? staged_frames - f : chunk;
            dim3 grid((unsigned)ceil_div(n_out, 256), (unsigned)ceil_div(nf, FT));
            k_apply_generic_tiles<FT><<<grid, 256, 0, st>>>(nf, n_in, h_out, w_out, tiles_x, row_ptr, col, val,
                                                            tile_info, values_in + f * n_in, values_out + f * n_out);
            RG_LAUNCH_CHECK("k_apply_generic_tiles");
        }
    }
    if (staged_frames < n_frames)
        return rg_apply_csr(device, stream, 1, n_in, n_out, row_ptr, col, val, values_in + staged_frames * n_in,
                            values_out + staged_frames * n_out);
    return RG_OK;
}
