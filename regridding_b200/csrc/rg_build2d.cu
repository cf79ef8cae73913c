// rg_build2d.cu -- 2D first-order conservative weights build (Ramshaw 1985 edge sweep).
//
// Replaces weights_conservative_2d + _coalesce of the reference
// (regridding/_weights/_weights_conservative_2d/_weights_conservative_2d.py:80-830,
//  regridding/_weights/_weights_arrays.py:44-73).
//
// The reference walks every sweep line sequentially.  At every sweep VERTEX its state
// collapses to (inside?, static cell) and the next segment restarts from the exact
// vertex coordinates (c2d.py:326-327, 386-387), so segments are independent given the
// state at their first vertex.  This file therefore runs ONE THREAD PER SEGMENT:
//
//   K1 k_cell_area        signed input-cell areas (grid_volume, _grids.py:50-140)
//   K2 k_boundary_*       boundary edges of each static grid in the reference's scan
//                         order + two levels of bounding boxes (exactly conservative
//                         w.r.t. the reference's own per-edge bbox pre-check)
//   K2 k_line_starts      exact start state of every sweep line (bbox test, winding
//                         number over the whole boundary, cell location; c2d.py:308-322)
//   K2 k_vertex_guess     Newton cell location of every other sweep vertex: a GUESS of
//                         the walk state there
//   K3 k_walk<Count>      walks each segment from its guessed start, records the end
//                         state and histograms the fragments per input cell
//   K3 k_repair           warp per line: wherever a segment's end state differs from the
//                         start state its successor used, the successor is re-walked
//                         from the true state (sequential propagation), so the chain of
//                         states is EXACTLY the reference's sequential walk
//   K3 k_walk<Emit>       re-walks with the verified starts and scatters
//                         (output cell, emission rank, weight) into per-input-cell buckets
//   K4 k_bucket_sort      sorts each bucket by (output cell, emission rank): the order the
//                         reference's stable argsort produces; no atomic decides any order
//   K4 k_bucket_emit      merges equal pairs in NumPy's reduceat association and writes the
//                         public (input, output)-sorted arrays
#include "rg_common.cuh"
#include "rg_geom.cuh"
#include "rg_boundary.cuh"
#include <mutex>
#include <vector>

// resident CTAs of 128 threads per SM the walk kernels are compiled for (8 -> 64 registers per thread)
#ifndef RG_COUNT_MINB
#define RG_COUNT_MINB 8
#endif
#ifndef RG_EMIT_MINB
#define RG_EMIT_MINB 8
#endif

namespace rg {

// ---------------------------------------------------------------------------
// data structures
// ---------------------------------------------------------------------------

struct PassParams {
    GridView sweep, stat;
    Boundary bnd;  // of the static grid
    int pass;      // 0..3 = (sweep OUTPUT axis 0, axis 1, sweep INPUT axis 0, axis 1), c2d.py:116,184
    int axis;
    int sweep_input;
    int nlines, nseg;
    int ncy_in, ncy_out;
    int ncx_st, ncy_st;
    int32_t* guess;             // [sweep vertices] guessed state at each sweep vertex
    int32_t* line_start;        // [nlines] exact state at the first vertex of each line
    int32_t* line_bad;          // [nlines] does the chain of segment states of the line need a repair?
    int32_t* seg_start;         // [sweep vertices] state each segment starts from
    int32_t* seg_end;           // [sweep vertices] state each segment ends in
    int max_iter;
    int64_t cell_lo, cell_hi;   // input-cell band
    uint8_t* seg_hit;           // [sweep vertices] banded builds: does the segment emit anything inside the band?
    int banded;
    // piece cache: what the count walk found, so that the emit walk does not repeat the geometry
    uint8_t* pc_n;              // [sweep vertices] cached entries of the segment, kPieceNone: walk again
    int32_t* pc_cell;           // [kPieceCache][sweep vertices] static cell of the piece; -1: "move to" (re-entry point)
    double* pc_x;               // [kPieceCache][sweep vertices] end point of the piece / the point moved to
    double* pc_y;
    int64_t pc_stride;          // sweep vertices
    // line-sharded builds: this rank walks the sweep lines of every part_world-th block of 32 lines
    int part_rank, part_world;
    int nl_slots;               // line slots of this rank (== nlines when part_world == 1)
};

// The four sweep passes are independent of each other: every phase launches them together.
struct Pass4 {
    PassParams p[4];
    int64_t tstart[5];  // first thread of every pass in the per-segment launches (multiples of 256)
    int lstart[5];      // first line slot of every pass in the per-line launches
};

__device__ __forceinline__ int pass_of_thread(const Pass4& Q, int64_t tid)
{
    return (tid >= Q.tstart[1]) + (tid >= Q.tstart[2]) + (tid >= Q.tstart[3]);
}
__device__ __forceinline__ int pass_of_slot(const Pass4& Q, int slot)
{
    return (slot >= Q.lstart[1]) + (slot >= Q.lstart[2]) + (slot >= Q.lstart[3]);
}

// One raw fragment: key = output cell << 32 | emission rank, val = weight.  16-byte records: one store per
// fragment in the emit walk, one all-to-all (or one peer read) per band in sharded builds.
struct __align__(16) Frag {
    uint64_t key;
    double val;
};

constexpr int kPieceCache = 4;      // cached entries per segment (a segment makes 1.8 pieces on average)
constexpr int kPieceNone = 255;

constexpr int kStateOutside = -1;
constexpr int kStateUnknown = -2;
constexpr int kStateInvalid = -3;

// flags[] slots
constexpr int kFlagOverflow = 0;   // a verified walk exceeded max_iter
constexpr int kFlagRepairs = 1;    // number of re-walked segments
constexpr int kFlagUnknown = 2;    // number of vertices whose guess was "unknown"
constexpr int kFlagSeqOverflow = 3;  // piece index overflowed the 29-bit rank field
constexpr int kFlagBandMismatch = 4; // band build: a walked segment did not start in the state its predecessor ended in
constexpr int kFlagCapacity = 5;     // band build: the caller's fragment / triplet buffers are too small
constexpr int kFlagMaxBucket = 6;    // standard build: fragments of the fullest input cell (sizes one-walk rebuilds)

__device__ __forceinline__ int64_t vertex_of(const PassParams& P, int L, int k)
{
    return P.axis ? (int64_t)L * P.sweep.ny + k : (int64_t)k * P.sweep.ny + L;
}
__device__ __forceinline__ int64_t vertex_step(const PassParams& P) { return P.axis ? 1 : P.sweep.ny; }

// Line-sharded builds deal the sweep lines out block-cyclically in blocks of 32 (a warp's worth of
// consecutive lines keeps the axis-0 loads coalesced; the cyclic deal balances lines that cross
// little of the static grid against lines that cross all of it).
__device__ __forceinline__ int line_of_slot(const PassParams& P, int s)
{
    if (P.part_world == 1) return s;
    return ((((s >> 5) * P.part_world) + P.part_rank) << 5) | (s & 31);
}

// ---------------------------------------------------------------------------
// K1: cell areas.  grid_volume accumulates, per cell (a, b), first the axis-0 pass
// then the axis-1 pass, each as "-= area(edge a); += area(edge a+1)":
//     (((0 - A_a) + A_a+1) - B_b) + B_b+1                      (_grids.py:68-73, 134-140)
// area = 0.5 * fma(x1, y2, -RN(x2*y1)); the JIT peels the first loop iteration (edge
// index 0) and there fuses the other product: 0.5 * fma(-x2, y1, RN(x1*y2)).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double tri_area(double x1, double y1, double x2, double y2, bool first)
{
    if (first) return dmul(0.5, dfma(-x2, y1, dmul(x1, y2)));
    return dmul(0.5, dfma(x1, y2, -dmul(x2, y1)));
}

__global__ void k_cell_area(GridView g, double* __restrict__ area)
{
    const int ncx = g.nx - 1, ncy = g.ny - 1;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)ncx * ncy) return;
    const int a = (int)(c / ncy), b = (int)(c % ncy);
    const int64_t v = (int64_t)a * g.ny + b;
    const double x00 = g.x[v], y00 = g.y[v];
    const double x01 = g.x[v + 1], y01 = g.y[v + 1];
    const double x10 = g.x[v + g.ny], y10 = g.y[v + g.ny];
    const double x11 = g.x[v + g.ny + 1], y11 = g.y[v + g.ny + 1];
    // axis 0: transposed view with x and y swapped; lines i' = b, b+1 ; edge from (a, i') to (a+1, i')
    const double A0 = tri_area(y00, x00, y10, x10, b == 0);
    const double A1 = tri_area(y01, x01, y11, x11, false);
    // axis 1: lines i = a, a+1 ; edge from (i, b) to (i, b+1)
    const double B0 = tri_area(x00, y00, x01, y01, a == 0);
    const double B1 = tri_area(x10, y10, x11, y11, false);
    double r = dsub(0.0, A0);
    r = dadd(r, A1);
    r = dsub(r, B0);
    r = dadd(r, B1);
    area[c] = r;
}

// the same for the cells [c_lo, c_hi) only (band builds); area[] is indexed by the global cell id
__global__ void k_cell_area_range(GridView g, int64_t c_lo, int64_t c_hi, double* __restrict__ area)
{
    const int ncy = g.ny - 1;
    const int64_t c = c_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    const int a = (int)(c / ncy), b = (int)(c % ncy);
    const int64_t v = (int64_t)a * g.ny + b;
    const double x00 = g.x[v], y00 = g.y[v];
    const double x01 = g.x[v + 1], y01 = g.y[v + 1];
    const double x10 = g.x[v + g.ny], y10 = g.y[v + g.ny];
    const double x11 = g.x[v + g.ny + 1], y11 = g.y[v + g.ny + 1];
    const double A0 = tri_area(y00, x00, y10, x10, b == 0);
    const double A1 = tri_area(y01, x01, y11, x11, false);
    const double B0 = tri_area(x00, y00, x01, y01, a == 0);
    const double B1 = tri_area(x10, y10, x11, y11, false);
    double r = dsub(0.0, A0);
    r = dadd(r, A1);
    r = dsub(r, B0);
    r = dadd(r, B1);
    area[c] = r;
}

// ---------------------------------------------------------------------------
// the segment walk: state machine of c2d.py:324-387 restricted to one segment
// ---------------------------------------------------------------------------
template <class Sink>
__device__ inline int walk_segment(const PassParams& P, double x1, double y1, double xv, double yv,
                                   int state, Sink& sink, bool& overflow)
{
    const GridView& g = P.stat;
    int last_edge = -1;   // local edge id just crossed (inside state)
    int last_cell = -2;   // flat id of the cell just left (outside state)
    int piece = 0;
    for (int it = 0; it < P.max_iter; it++) {
        if (state < 0) {
            double t;
            const int s = boundary_entry(P.bnd, x1, y1, xv, yv, last_cell, t);
            if (s < 0) return kStateOutside;  // c2d.py:539-540: advance to the next sweep vertex
            seg_point(x1, y1, xv, yv, t, x1, y1);  // c2d.py:526-530, then x1 = x2 at :386
            state = P.bnd.cell[s];
            last_edge = edge_local_id(P.bnd, s);
            last_cell = -2;
            continue;
        }
        // _step_inside_static, c2d.py:561-749
        const int ci = state / P.ncy_st, cj = state - ci * P.ncy_st;
        const int64_t a = (int64_t)ci * g.ny + cj;
        // cell vertices in the order of indices_cell_vertex (_grids.py:153-158)
        double vx[4], vy[4];
        vx[0] = g.x[a];            vy[0] = g.y[a];
        vx[1] = g.x[a + g.ny];     vy[1] = g.y[a + g.ny];
        vx[2] = g.x[a + g.ny + 1]; vy[2] = g.y[a + g.ny + 1];
        vx[3] = g.x[a + 1];        vy[3] = g.y[a + 1];
        int hit_v = -1;
        double t = 0.0;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            if (hit_v >= 0 || v == last_edge) continue;
            const int pv = (v + 3) & 3;
            double tv;
            if (seg_hit(x1, y1, xv, yv, vx[pv], vy[pv], vx[v], vy[v], tv)) {
                hit_v = v;
                t = tv;
            }
        }
        double x2 = xv, y2 = yv;
        if (hit_v >= 0) seg_point(x1, y1, xv, yv, t, x2, y2);
        sink.piece(P, x1, y1, x2, y2, ci, cj, piece);  // always (c2d.py:712-725)
        piece++;
        if (hit_v < 0) return state;  // reached the sweep vertex inside `state`
        // cell_normals (_grids.py:143-148)
        const int ni = ci + (hit_v == 2) - (hit_v == 0);
        const int nj = cj + (hit_v == 3) - (hit_v == 1);
        if (ni < 0 || nj < 0 || ni >= P.ncx_st || nj >= P.ncy_st) {
            last_cell = state;  // c2d.py:727-737
            state = kStateOutside;
            last_edge = -1;
        } else {
            state = ni * P.ncy_st + nj;
            last_edge = (hit_v + 2) & 3;
        }
        x1 = x2;
        y1 = y2;
    }
    overflow = true;
    return kStateInvalid;
}

// _calc_and_save_weights + _index_input_output (c2d.py:757-830, 876-912): which
// (input cell, output cell) pairs a piece on sweep line L, segment k feeds.
struct PieceCells {
    int n;             // number of valid sides
    int64_t in[2], out[2];
    int side[2];       // 0 = left (L-1), 1 = right (L)
};

__device__ __forceinline__ PieceCells piece_cells(const PassParams& P, int L, int k, int ci, int cj)
{
    PieceCells r;
    r.n = 0;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int i = side ? L : L - 1;
        if (side == 0 ? (i < 0) : (i >= P.nlines - 1)) continue;
        const int si = P.axis ? i : k, sj = P.axis ? k : i;  // axis 0: (line, segment) -> (segment, line)
        int64_t fin, fout;
        if (P.sweep_input) { fin = (int64_t)si * P.ncy_in + sj; fout = (int64_t)ci * P.ncy_out + cj; }
        else               { fin = (int64_t)ci * P.ncy_in + cj; fout = (int64_t)si * P.ncy_out + sj; }
        if (fin < P.cell_lo || fin >= P.cell_hi) continue;
        r.in[r.n] = fin; r.out[r.n] = fout; r.side[r.n] = side;
        r.n++;
    }
    return r;
}

struct CountSink {
    int32_t* hist;
    int L, k, delta;
    int total;  // fragments of this segment that fall inside the band
    // piece cache (v < 0: off): entries written so far and the point the last one ended at
    int64_t v;
    int n_cached;
    double cx, cy;
    __device__ __forceinline__ void push(const PassParams& P, int cell, double x, double y)
    {
        if (n_cached < kPieceCache) {
            const int64_t at = (int64_t)n_cached * P.pc_stride + v;
            P.pc_cell[at] = cell;
            P.pc_x[at] = x;
            P.pc_y[at] = y;
        }
        n_cached++;
    }
    __device__ __forceinline__ void piece(const PassParams& P, double x1, double y1, double x2, double y2,
                                          int ci, int cj, int)
    {
        const PieceCells pc = piece_cells(P, L, k, ci, cj);
        for (int q = 0; q < pc.n; q++) atomicAdd(&hist[pc.in[q]], delta);
        total += pc.n;
        if (v >= 0) {
            if (x1 != cx || y1 != cy) push(P, -1, x1, y1);  // the piece starts at a re-entry point
            push(P, ci * P.ncy_st + cj, x2, y2);
            cx = x2;
            cy = y2;
        }
    }
};

struct EmitSink {
    const int64_t* boff;
    int32_t* cursor;
    Frag* frag;
    const double* area_in;
    const double* w_in;
    int32_t* flags;
    int L, k;
    int64_t cap = INT64_MAX;  // records the fragment buffer holds (band builds size it by estimate)
    // one-walk band builds: no offsets yet -- every input cell of the band owns `bucket_cap` slots of `frag`
    int bucket_cap = 0;
    int64_t cell_base = 0;
    __device__ __forceinline__ void piece(const PassParams& P, double x1, double y1, double x2, double y2,
                                          int ci, int cj, int piece_idx)
    {
        const PieceCells pc = piece_cells(P, L, k, ci, cj);
        if (pc.n == 0) return;
        // area_triangle (geometry.py:993) with the axis sign (c2d.py:783-784); the JIT evaluates the
        // negated form RN(x2*y1) - x1*y2 (fused) times +-0.5.
        const double neg2 = dfma(-x1, y2, dmul(x2, y1));
        const double area = dmul(neg2, P.axis == 0 ? 0.5 : -0.5);
        if (piece_idx >= (1 << 29)) atomicOr(&flags[kFlagSeqOverflow], 1);
        // both sides of the line: all loads and the two slot reservations are issued before anything is used
        // (this code is bound by the latency of these accesses, not by arithmetic)
        const bool two = pc.n > 1;
        const int64_t in0 = pc.in[0], in1 = pc.in[two ? 1 : 0];
        const double vol0 = area_in[in0];
        const double vol1 = area_in[in1];
        const int64_t base0 = bucket_cap ? (in0 - cell_base) * bucket_cap : boff[in0];
        const int64_t base1 = bucket_cap ? (in1 - cell_base) * bucket_cap : boff[in1];
        double wi0 = 1.0, wi1 = 1.0;
        if (w_in) { wi0 = w_in[in0]; wi1 = w_in[in1]; }
        const int old0 = atomicAdd(&cursor[in0], 1);
        const int old1 = two ? atomicAdd(&cursor[in1], 1) : 0;
        if (bucket_cap ? (old0 >= bucket_cap || old1 >= bucket_cap) : (base0 + old0 >= cap || base1 + old1 >= cap)) {
            atomicOr(&flags[kFlagCapacity], 1);
            return;
        }
        // emission rank inside one (input, output) pair: pass, then line (the cell right of line L
        // comes before the cell left of line L+1), then piece order inside the segment.
        const uint32_t seq = ((uint32_t)P.pass << 30) | (uint32_t)piece_idx;
        {
            const double a = pc.side[0] ? -area : area;
            // with weights_input: (area * w) / vol, the fastmath reassociation seen in the JIT
            const double w = w_in ? ddiv(dmul(a, wi0), vol0) : ddiv(a, vol0);
            frag[base0 + old0] = Frag{ ((uint64_t)pc.out[0] << 32) | seq | (pc.side[0] == 0 ? (1u << 29) : 0u), w };
        }
        if (two) {
            const double a = pc.side[1] ? -area : area;
            const double w = w_in ? ddiv(dmul(a, wi1), vol1) : ddiv(a, vol1);
            frag[base1 + old1] = Frag{ ((uint64_t)pc.out[1] << 32) | seq | (pc.side[1] == 0 ? (1u << 29) : 0u), w };
        }
    }
};

// ---------------------------------------------------------------------------
// K2: exact line-start states (c2d.py:293-322).  One CTA per line: the boundary winding number is
// latency bound (the boundary has 4 (n - 1) edges), so 256 threads share it.
// ---------------------------------------------------------------------------
__device__ inline void line_start_body(const PassParams& P, int L, const double* __restrict__ bbox2, double* s_w)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* bbox_static = bbox2 + 4 * (P.sweep_input ? 1 : 0);
    const int64_t v0 = vertex_of(P, L, 0);
    const double px = P.sweep.x[v0], py = P.sweep.y[v0];
    // point_is_inside_box_2d (geometry.py:64-105)
    const bool in_box = bbox_static[0] <= px && px <= bbox_static[2] && bbox_static[1] <= py && py <= bbox_static[3];
    int state = kStateOutside;
    if (in_box) {
        // point_is_inside_polygon over grid_boundary (_grids.py:167-215): counter-clockwise in index space;
        // the scan-order edges of the i = 0 face and of the j = ny-1 face run the other way round.  Only edges whose
        // y-range contains py contribute (every other edge adds exactly 0), so the two bounding-box levels cull whole
        // groups; the threads of the CTA take the level-1 groups (32 edges each) in turn.
        const Boundary& b = P.bnd;
        double w = 0.0;
        for (int g1 = threadIdx.x; g1 < b.n_g1; g1 += 256) {
            const BBox B1 = b.bb1[g1];
            if (!(B1.ylo <= py && py <= B1.yhi)) continue;
            const int se = min(b.n_edges, (g1 + 1) * 32);
            for (int s = g1 * 32; s < se; s++) {
                const double x0 = dsub(b.x3[s], px), y0 = dsub(b.y3[s], py);
                const double x1 = dsub(b.x4[s], px), y1 = dsub(b.y4[s], py);
                const bool reversed = (s < b.ne_a0) || (s >= 2 * b.ne_a0 + b.ne_a1);
                w += reversed ? winding_edge(x1, y1, x0, y0) : winding_edge(x0, y0, x1, y1);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);  // halves: exact in any order
        if (lane == 0) s_w[warp] = w;
        __syncthreads();
        if (warp != 0) return;
        w = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) w += s_w[q];
        if (w != 0.0) {
            int found = kLocUnknown;
            // (the result does not depend on the seed -- see locate_newton; the affine map through three corners of
            // the grid saves most of the iterations the grid centre needs)
            if (lane == 0) {
                double si, sj;
                affine_seed(P.stat, px, py, si, sj);
                found = locate_newton(P.stat, px, py, si, sj);
            }
            found = __shfl_sync(0xffffffffu, found, 0);
            if (found < 0) {
                // index_of_point_brute (_grids.py:223-279): lowest row-major containing cell
                const int64_t nc = (int64_t)P.ncx_st * P.ncy_st;
                int64_t best = INT64_MAX;
                for (int64_t base = 0; base < nc && best == INT64_MAX; base += 32) {
                    const int64_t c = base + lane;
                    bool hit = false;
                    if (c < nc) hit = cell_contains(P.stat, (int)(c / P.ncy_st), (int)(c % P.ncy_st), px, py);
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) best = base + (__ffs(m) - 1);
                }
                found = best == INT64_MAX ? kStateOutside : (int)best;
            }
            state = found;
        }
    }
    if (threadIdx.x == 0) {
        P.line_start[L] = state;
        P.line_bad[L] = 0;
    }
}

__global__ void __launch_bounds__(256) k_line_starts(const __grid_constant__ Pass4 Q, const double* __restrict__ bbox2)
{
    __shared__ double s_w[8];
    const int p = pass_of_slot(Q, (int)blockIdx.x);
    const PassParams& P = Q.p[p];
    const int L = line_of_slot(P, (int)blockIdx.x - Q.lstart[p]);
    if (L >= P.nlines) return;
    line_start_body(P, L, bbox2, s_w);
}

// K2: guessed state of every sweep vertex
__global__ void k_vertex_guess(GridView sweep, GridView stat, int32_t* __restrict__ guess, int32_t* flags)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (int64_t)sweep.nx * sweep.ny) return;
    const int r = locate_guess(stat, sweep.x[v], sweep.y[v]);
    guess[v] = r;
    if (r == kLocUnknown) atomicAdd(&flags[kFlagUnknown], 1);
}

// ---------------------------------------------------------------------------
// K3: walks
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool segment_of_thread(const PassParams& P, int64_t tid, int& L, int& k)
{
    if (tid >= (int64_t)P.nl_slots * P.nseg) return false;
    // consecutive lanes take consecutive memory: along the line for axis 1, across lines for axis 0
    int slot;
    if (P.axis) { slot = (int)(tid / P.nseg); k = (int)(tid % P.nseg); }
    else        { k = (int)(tid / P.nl_slots); slot = (int)(tid % P.nl_slots); }
    L = line_of_slot(P, slot);
    return L < P.nlines;
}

// K2 (line-sharded builds): guessed state at the first vertex of every segment this rank walks
__global__ void k_vertex_guess_part(const __grid_constant__ Pass4 Q, int32_t* flags)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid >= Q.tstart[4]) return;
    const int p = pass_of_thread(Q, gtid);
    const PassParams& P = Q.p[p];
    int32_t* guess = P.guess;
    int L, k;
    if (!segment_of_thread(P, gtid - Q.tstart[p], L, k)) return;
    if (k == 0) return;  // the line start is located exactly by k_line_starts
    const int64_t v = vertex_of(P, L, k);
    const int r = locate_guess(P.stat, P.sweep.x[v], P.sweep.y[v]);
    guess[v] = r;
    if (r == kLocUnknown) atomicAdd(&flags[kFlagUnknown], 1);
}

__global__ void __launch_bounds__(128, RG_COUNT_MINB) k_walk_count(const __grid_constant__ Pass4 Q, int32_t* __restrict__ hist)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid >= Q.tstart[4]) return;
    const int p = pass_of_thread(Q, gtid);
    const PassParams& P = Q.p[p];
    int L, k;
    if (!segment_of_thread(P, gtid - Q.tstart[p], L, k)) return;
    const int64_t v = vertex_of(P, L, k), v2 = v + vertex_step(P);
    const int start = (k == 0) ? P.line_start[L] : P.guess[v];
    P.seg_start[v] = start;
    if (start == kStateUnknown) {
        P.seg_end[v] = kStateInvalid;
        if (P.banded) P.seg_hit[v] = 1;  // decided by the repair pass
        P.pc_n[v] = kPieceNone;
        return;
    }
    const double x1 = P.sweep.x[v], y1 = P.sweep.y[v];
    CountSink sink{ hist, L, k, 1, 0, v, 0, x1, y1 };
    bool overflow = false;
    P.seg_end[v] = walk_segment(P, x1, y1, P.sweep.x[v2], P.sweep.y[v2], start, sink, overflow);
    if (P.banded) P.seg_hit[v] = sink.total > 0;
    P.pc_n[v] = (uint8_t)((sink.n_cached <= kPieceCache && !overflow) ? sink.n_cached : kPieceNone);
}

// Does every segment start from the state its predecessor ended in?  (Thread per segment; the common answer
// is yes for every line, and then k_repair has nothing to do.)
__global__ void k_chain_check(const __grid_constant__ Pass4 Q)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid >= Q.tstart[4]) return;
    const int p = pass_of_thread(Q, gtid);
    const PassParams& P = Q.p[p];
    int L, k;
    if (!segment_of_thread(P, gtid - Q.tstart[p], L, k)) return;
    if (k == 0) return;
    const int64_t v = vertex_of(P, L, k);
    if (P.seg_end[v - vertex_step(P)] != P.seg_start[v]) P.line_bad[L] = 1;
}

// One warp per flagged line: make the chain of states equal to the sequential walk.
__global__ void k_repair(const __grid_constant__ Pass4 Q, int32_t* __restrict__ hist, int32_t* __restrict__ flags)
{
    const int lane = threadIdx.x & 31;
    const int gslot = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (gslot >= Q.lstart[4]) return;
    const int p = pass_of_slot(Q, gslot);
    const PassParams& P = Q.p[p];
    const int L = line_of_slot(P, gslot - Q.lstart[p]);
    if (L >= P.nlines) return;
    if (!P.line_bad[L]) return;
    const int64_t step = vertex_step(P);
    for (int base = 1; base < P.nseg; base += 32) {
        const int k = base + lane;
        const int64_t v = vertex_of(P, L, min(k, P.nseg - 1));
        while (true) {
            bool bad = false;
            if (k < P.nseg) bad = ((volatile int32_t*)P.seg_end)[v - step] != ((volatile int32_t*)P.seg_start)[v];
            const unsigned m = __ballot_sync(0xffffffffu, bad);
            if (!m) break;
            if (lane == __ffs(m) - 1) {
                const int old_start = P.seg_start[v];
                const int new_start = P.seg_end[v - step];
                const int64_t v2 = v + step;
                const double x1 = P.sweep.x[v], y1 = P.sweep.y[v], x2 = P.sweep.x[v2], y2 = P.sweep.y[v2];
                bool overflow = false;
                if (old_start != kStateUnknown) {
                    CountSink undo{ hist, L, k, -1, 0, -1, 0, 0.0, 0.0 };
                    walk_segment(P, x1, y1, x2, y2, old_start, undo, overflow);
                }
                CountSink redo{ hist, L, k, 1, 0, -1, 0, 0.0, 0.0 };
                const int e = walk_segment(P, x1, y1, x2, y2, new_start, redo, overflow);
                P.seg_start[v] = new_start;
                P.seg_end[v] = e;
                P.pc_n[v] = kPieceNone;  // the cached pieces belong to the wrong start: the emit walk redoes them
                if (P.banded) P.seg_hit[v] = redo.total > 0;
                atomicAdd(&flags[kFlagRepairs], 1);
                __threadfence();
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128, RG_EMIT_MINB)
k_walk_emit(const __grid_constant__ Pass4 Q, const int64_t* __restrict__ boff, int32_t* __restrict__ cursor,
            Frag* __restrict__ frag,
            const double* __restrict__ area_in, const double* __restrict__ w_in, int32_t* __restrict__ flags)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid >= Q.tstart[4]) return;
    const int p = pass_of_thread(Q, gtid);
    const PassParams& P = Q.p[p];
    int L, k;
    if (!segment_of_thread(P, gtid - Q.tstart[p], L, k)) return;
    const int64_t v = vertex_of(P, L, k), v2 = v + vertex_step(P);
    if (P.banded && !P.seg_hit[v]) return;  // the count walk saw nothing of this segment inside the band
    EmitSink sink{ boff, cursor, frag, area_in, w_in, flags, L, k };
    const int nc = P.pc_n[v];
    if (nc != kPieceNone) {
        // replay the pieces the count walk recorded: same points, same cells, same order -- no geometry
        double x1 = P.sweep.x[v], y1 = P.sweep.y[v];
        int piece = 0;
        // all entries are loaded before the first one is used (independent loads in flight together)
        int cells[kPieceCache];
        double xs[kPieceCache], ys[kPieceCache];
#pragma unroll
        for (int e = 0; e < kPieceCache; e++) {
            const int64_t at = (int64_t)e * P.pc_stride + v;
            cells[e] = -1; xs[e] = 0.0; ys[e] = 0.0;
            if (e < nc) { cells[e] = P.pc_cell[at]; xs[e] = P.pc_x[at]; ys[e] = P.pc_y[at]; }
        }
#pragma unroll
        for (int e = 0; e < kPieceCache; e++) {
            if (e >= nc) break;
            const int cell = cells[e];
            const double x = xs[e], y = ys[e];
            if (cell >= 0) {
                const int ci = cell / P.ncy_st;
                sink.piece(P, x1, y1, x, y, ci, cell - ci * P.ncy_st, piece);
                piece++;
            }
            x1 = x;
            y1 = y;
        }
        return;
    }
    bool overflow = false;
    walk_segment(P, P.sweep.x[v], P.sweep.y[v], P.sweep.x[v2], P.sweep.y[v2], P.seg_start[v], sink, overflow);
    if (overflow) atomicOr(&flags[kFlagOverflow], 1);
}

// ---------------------------------------------------------------------------
// K4: per-input-cell buckets -> public layout
// ---------------------------------------------------------------------------

// Sort each bucket by key = (output cell << 32 | emission rank) and count its distinct output cells.
// A CTA owns kSortCells consecutive buckets = one contiguous range of fragments: the range is staged in shared
// memory with coalesced loads, every thread insertion-sorts its own bucket there (buckets hold ~14 fragments),
// and the range is written back coalesced.  CTAs whose range does not fit (or that hold a bucket longer than 64)
// sort in global memory: one lane per bucket, the whole warp (odd-even transposition) for the long ones.
#ifndef RG_SORT_CELLS
#define RG_SORT_CELLS 64
#endif
#ifndef RG_SORT_THREADS
#define RG_SORT_THREADS 128
#endif
#ifndef RG_SORT_CAP
#define RG_SORT_CAP 1500
#endif
constexpr int kSortCells = RG_SORT_CELLS;     // buckets per CTA
constexpr int kSortThreads = RG_SORT_THREADS; // threads per CTA: one per bucket for the bookkeeping, all of them for the
                                              // per-fragment phases
constexpr int kSortCap = RG_SORT_CAP;  // fragments staged per CTA (~31 KB of shared memory with the maps: 7 CTAs per SM;
                                       // smaller tiles hide the NVLink latency of the sharded gather better: -9 % merge time)
static_assert(kSortThreads >= kSortCells && kSortThreads % 32 == 0 && kSortCells % 32 == 0, "sort CTA shape");
constexpr int kMaxSrc = 16;     // source ranks of a sharded build

// Line-sharded builds: the fragments of a band arrive as one chunk per source rank, each chunk bucketed by
// input cell.  cntT[c * W + s] = fragments of band cell c in the chunk of source s; src_off = exclusive scan of
// the counts in [s][c] order, dst_off = exclusive scan in [c][s] order (position in the merged bucket array).
// src[s] points at the chunk of source s MINUS src_off[s * Cb] records (so that src[s][src_off[s * Cb + c]] is
// the first fragment of cell c from source s); the chunks may live in peer memory (NVLink loads).
struct GatherSrc {
    const int32_t* cntT;
    const int64_t* src_off;
    const int64_t* dst_off;
    const Frag* src[kMaxSrc];
    int W;
    int64_t Cb;
};

__device__ __forceinline__ Frag load_frag_nc(const Frag* p)
{
    const ulonglong2 q = __ldg(reinterpret_cast<const ulonglong2*>(p));
    return Frag{ q.x, __longlong_as_double((long long)q.y) };
}

// 16-byte asynchronous global -> shared copy (LDGSTS): all the copies of a staging loop are in flight together
__device__ __forceinline__ void cp_async_frag(Frag* smem_dst, const Frag* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// global-memory fallback: buckets already in place in `frag`
__device__ inline void sort_buckets_global(Frag* __restrict__ frag, int64_t beg, int64_t end, bool valid,
                                           int32_t* __restrict__ nuniq, int64_t c)
{
    const int lane = threadIdx.x & 31;
    const int64_t len = end - beg;
    if (len <= 64) {
        for (int64_t e = beg + 1; e < end; e++) {
            const Frag x = frag[e];
            int64_t f = e - 1;
            while (f >= beg && frag[f].key > x.key) {
                frag[f + 1] = frag[f];
                f--;
            }
            frag[f + 1] = x;
        }
    }
    unsigned longmask = __ballot_sync(0xffffffffu, len > 64);
    while (longmask) {
        const int src = __ffs(longmask) - 1;
        longmask &= longmask - 1;
        const int64_t b = __shfl_sync(0xffffffffu, beg, src);
        const int64_t n = __shfl_sync(0xffffffffu, end, src) - b;
        for (int64_t phase = 0; phase < n; phase++) {
            for (int64_t p = (phase & 1) + 2 * lane; p + 1 < n; p += 64) {
                const Frag f0 = frag[b + p], f1 = frag[b + p + 1];
                if (f0.key > f1.key) {
                    frag[b + p] = f1;
                    frag[b + p + 1] = f0;
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    if (valid) {
        int32_t u = 0;
        uint32_t prev = 0xffffffffu;
        for (int64_t e = beg; e < end; e++) {
            const uint32_t o = (uint32_t)(frag[e].key >> 32);
            u += (e == beg) || (o != prev);
            prev = o;
        }
        nuniq[c] = u;
    }
}

// Shared-memory layout of the two sort kernels (dynamic): the staged fragments, the bucket (local cell) of
// every staged fragment, and per local cell: first slot, length, distinct output cells.
struct SortSmem {
    Frag frag[kSortCap];
    uint16_t cell[kSortCap];
    int beg[kSortCells];
    int len[kSortCells];
    int uniq[kSortCells];
    // gather variant only: stage position of the k-th fragment of every bucket (bucket b owns pos[beg .. beg + len))
    uint16_t pos[kSortCap];
    int base[kMaxSrc + 1];
    int is_long;
};

// Rank sort, one thread per FRAGMENT: keys are unique inside a bucket, so the final slot of a fragment is the
// number of smaller keys in its bucket.  No dependent stores, no divergence between the lanes of a bucket (they
// read the same keys: shared-memory broadcasts), and the sorted records go straight to global memory.
__global__ void __launch_bounds__(kSortThreads)
k_bucket_sort(const int64_t* __restrict__ boff, int64_t n_cells, Frag* __restrict__ frag, int32_t* __restrict__ nuniq,
              const int32_t* __restrict__ abort_flag = nullptr, const Frag* __restrict__ strided = nullptr, int bucket_cap = 0,
              int32_t* __restrict__ max_bucket = nullptr)
{
    // `strided` (one-walk band builds): bucket c was written to strided[c * bucket_cap ...]; it is gathered from there
    // and leaves, sorted, at the dense offsets `boff` like in the two-walk pipeline.  `max_bucket`: longest bucket seen.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(smem_raw);
    if (abort_flag && *abort_flag) return;  // band builds: the fragment buffer was too small, nothing valid to sort
    const int64_t c0 = (int64_t)blockIdx.x * kSortCells;
    const bool owner = threadIdx.x < kSortCells;  // this thread keeps the books of bucket c
    const int64_t c = c0 + threadIdx.x;
    int64_t beg = 0, end = 0;
    if (owner && c < n_cells) {
        beg = boff[c];
        end = boff[c + 1];
    }
    const int64_t len = end - beg;
    const int64_t lo = boff[c0], hi = boff[min(c0 + (int64_t)kSortCells, n_cells)];
    if (threadIdx.x == 0) S.is_long = 0;
    __syncthreads();
    if (len > 64) S.is_long = 1;
    if (max_bucket && owner) {
        const int m = __reduce_max_sync(0xffffffffu, (int)min(len, (int64_t)INT32_MAX));   // owners fill whole warps
        if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_bucket, m);
    }
    __syncthreads();
    if (hi - lo <= kSortCap && !S.is_long) {
        const int n = (int)(hi - lo);
        if (!strided)
            for (int e = threadIdx.x; e < n; e += kSortThreads) cp_async_frag(&S.frag[e], frag + lo + e);
        if (owner) {
            const int b = (int)(beg - lo), n_mine = (int)len;
            S.beg[threadIdx.x] = b;
            S.len[threadIdx.x] = n_mine;
            S.uniq[threadIdx.x] = 0;
            for (int e = b; e < b + n_mine; e++) S.cell[e] = (uint16_t)threadIdx.x;
        }
        if (strided) {
            __syncthreads();
            for (int e = threadIdx.x; e < n; e += kSortThreads) {
                const int cl = S.cell[e];
                cp_async_frag(&S.frag[e], strided + (c0 + cl) * bucket_cap + (e - S.beg[cl]));
            }
        }
        cp_async_wait_all();
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += kSortThreads) {
            const int cl = S.cell[e];
            const int bb = S.beg[cl], m = S.len[cl];
            const Frag x = S.frag[e];
            const uint32_t o = (uint32_t)(x.key >> 32);
            int rank = 0, dup = 0;
            for (int f = 0; f < m; f++) {
                const uint64_t k2 = S.frag[bb + f].key;
                rank += k2 < x.key;
                dup += (k2 < x.key) && ((uint32_t)(k2 >> 32) == o);
            }
            frag[lo + bb + rank] = x;
            if (!dup) atomicAdd(&S.uniq[cl], 1);  // the first fragment of its (input, output) pair
        }
        __syncthreads();
        if (owner && c < n_cells) nuniq[c] = S.uniq[threadIdx.x];
        return;
    }
    if (owner) {
        if (strided) {   // the global-memory sort works in place on the dense layout
            for (int64_t e = 0; e < len; e++) frag[beg + e] = strided[c * bucket_cap + e];
            __syncwarp();
        }
        sort_buckets_global(frag, beg, end, c < n_cells, nuniq, c);  // warps 0..3 are complete
    }
}

// Sharded builds: gather the band's buckets from the W source chunks and rank-sort them.  Every source's share
// of the CTA's kSortCells cells is ONE contiguous range of its chunk: the CTA copies the W ranges into the stage with
// 16-byte asynchronous copies (all in flight together -- they may cross NVLink); a bucket is then the union of
// its W pieces in the stage.
__global__ void __launch_bounds__(kSortThreads)
k_bucket_gather_sort(const __grid_constant__ GatherSrc G, int64_t n_cells, Frag* __restrict__ frag,
                     int32_t* __restrict__ nuniq)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SortSmem& S = *reinterpret_cast<SortSmem*>(smem_raw);
    const int W = G.W;
    const int64_t c0 = (int64_t)blockIdx.x * kSortCells;
    const int64_t c1 = min(c0 + (int64_t)kSortCells, n_cells);
    const bool owner = threadIdx.x < kSortCells;  // this thread keeps the books of bucket c
    const int64_t c = c0 + threadIdx.x;
    int64_t beg = 0, end = 0;
    if (owner && c < n_cells) {
        beg = G.dst_off[c * W];
        end = G.dst_off[(c + 1) * W];
    }
    const int64_t len = end - beg;
    const int64_t lo = G.dst_off[c0 * W], hi = G.dst_off[c1 * W];
    if (threadIdx.x == 0) S.is_long = 0;
    if (threadIdx.x < W)  // fragments of source s inside this CTA's cells
        S.base[threadIdx.x + 1] = (int)min(G.src_off[(int64_t)threadIdx.x * G.Cb + c1] -
                                           G.src_off[(int64_t)threadIdx.x * G.Cb + c0], (int64_t)kSortCap + 1);
    // this cell's pieces: count and first fragment (relative to the CTA's first cell) per source
    int p_cnt[kMaxSrc], p_off[kMaxSrc];
#pragma unroll
    for (int s = 0; s < kMaxSrc; s++) {
        p_cnt[s] = 0;
        p_off[s] = 0;
        if (s < W && owner && c < n_cells) {
            p_cnt[s] = G.cntT[c * W + s];
            p_off[s] = (int)min(G.src_off[(int64_t)s * G.Cb + c] - G.src_off[(int64_t)s * G.Cb + c0], (int64_t)kSortCap + 1);
        }
    }
    __syncthreads();
    if (len > 64) S.is_long = 1;
    if (threadIdx.x == 0) {  // base[s] = fragments of the sources before s
        int acc = 0;
        S.base[0] = 0;
        for (int s = 1; s <= W; s++) {
            acc = min(acc + S.base[s], kSortCap + 1);
            S.base[s] = acc;
        }
    }
    __syncthreads();
    if (hi - lo <= kSortCap && !S.is_long) {
        const int n = (int)(hi - lo);
        for (int s = 0; s < W; s++) {
            const Frag* src = G.src[s] + G.src_off[(int64_t)s * G.Cb + c0];
            const int ns = S.base[s + 1] - S.base[s];
            for (int e = threadIdx.x; e < ns; e += kSortThreads) cp_async_frag(&S.frag[S.base[s] + e], src + e);
        }
        // while the copies fly: where the pieces of this thread's cell land in the stage
        if (owner) {
            const int b = (int)(beg - lo);
            S.beg[threadIdx.x] = b;
            S.len[threadIdx.x] = (int)len;
            S.uniq[threadIdx.x] = 0;
            int k = b;
#pragma unroll
            for (int s = 0; s < kMaxSrc; s++) {
                const int at = S.base[s < W ? s : 0] + p_off[s];
                for (int e = 0; e < p_cnt[s]; e++) {
                    S.cell[at + e] = (uint16_t)threadIdx.x;
                    S.pos[k++] = (uint16_t)(at + e);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += kSortThreads) {
            const int cl = S.cell[e];
            const int bb = S.beg[cl], m = S.len[cl];
            const Frag x = S.frag[e];
            const uint32_t o = (uint32_t)(x.key >> 32);
            int rank = 0, dup = 0;
            for (int f = 0; f < m; f++) {
                const uint64_t k2 = S.frag[S.pos[bb + f]].key;
                rank += k2 < x.key;
                dup += (k2 < x.key) && ((uint32_t)(k2 >> 32) == o);
            }
            frag[lo + bb + rank] = x;
            if (!dup) atomicAdd(&S.uniq[cl], 1);
        }
        __syncthreads();
        if (owner && c < n_cells) nuniq[c] = S.uniq[threadIdx.x];
        return;
    }
    // ---- global-memory path ----
    if (!owner) return;  // warps 0..3 are complete
    if (c < n_cells) {
        int64_t w = beg;
        for (int s = 0; s < W; s++) {
            const int ns = G.cntT[c * W + s];
            const Frag* src = G.src[s] + G.src_off[(int64_t)s * G.Cb + c];
            for (int e = 0; e < ns; e++) frag[w++] = load_frag_nc(src + e);
        }
    }
    __syncwarp();
    sort_buckets_global(frag, beg, end, c < n_cells, nuniq, c);
}

// Merge the runs of equal (input, output) pair: np.add.reduceat association over the
// emission order (warr.py:59-72) and write the public arrays.
__global__ void k_bucket_emit(const int64_t* __restrict__ boff, int64_t bstride, int64_t cell_offset,
                              const int64_t* __restrict__ colptr, int64_t n_cells,
                              const Frag* __restrict__ frag,
                              int64_t* __restrict__ ii, int64_t* __restrict__ io, double* __restrict__ vv,
                              const int32_t* __restrict__ abort_flag = nullptr)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    if (abort_flag && *abort_flag) return;
    const int64_t beg = boff[c * bstride], end = boff[(c + 1) * bstride];
    if (beg >= end) return;
    int64_t w = colptr[c];
    // One 16-byte load per fragment (the kernel is bound by the L1 wavefronts of these uncoalesced loads): the run
    // sum is accumulated while scanning.  np.add.reduceat of a run a[0..n) is a[0] + pairwise_sum(a[1..n)), and the
    // pairwise sum of fewer than 8 values is the sequential sum from -0.0; longer runs (rare) are summed again
    // from memory in NumPy's blocked order.
    Frag x = frag[beg];
    uint32_t o = (uint32_t)(x.key >> 32);
    double a0 = x.val, rest = -0.0;
    int64_t run = beg;
    for (int64_t e = beg + 1; e <= end; e++) {
        const bool more = e < end;
        Frag y = x;
        if (more) y = frag[e];
        const uint32_t oy = (uint32_t)(y.key >> 32);
        if (more && oy == o) {
            rest = dadd(rest, y.val);
            continue;
        }
        const int64_t n = e - run;
        ii[w] = c + cell_offset;
        io[w] = (int64_t)o;
        vv[w] = n == 1 ? a0 : (n - 1 < 8 ? dadd(a0, rest) : np_reduceat_segment<2>(&frag[run].val, n));
        w++;
        o = oy;
        a0 = y.val;
        rest = -0.0;
        run = e;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct Layout {
    // sizes
    int64_t nxi, nyi, nxo, nyo, Vi, Vo, Ci, Co;
    // carved pointers
    double* area_in;
    double* bbox;  // [2][4]: input grid, output grid
    Boundary bnd[2];
    int32_t* guess[2];       // [0]: output vertices located in the input grid, [1]: input vertices in the output grid
    int32_t* line_start[4];
    int32_t* line_bad[4];
    int32_t* seg_start[4];
    int32_t* seg_end[4];
    uint8_t* seg_hit[4];
    uint8_t* pc_n[4];
    int32_t* pc_cell[4];
    double* pc_x[4];
    double* pc_y[4];
    int32_t* hist;
    int64_t* boff;
    int32_t* cursor;
    int32_t* nuniq;
    int64_t* colptr;
    int64_t* scan_scratch;
    unsigned long long* scan_status[2];   // band build: status words of its two single-launch scans
    int32_t* flags;
    // band builds (rg_build2d_band)
    uint8_t* raster8;        // [kRasterN^2] byte marks (plain stores), packed into `raster`
    uint32_t* raster;        // [kRasterN^2 / 8] 4 bits per raster cell (k_band_raster_pack): does a cell of the band touch it / its 2x1, 1x2, 2x2 block?
    uint8_t* rel[2];         // [output vertices] can the pass-0 / pass-1 segment starting here have a piece in the band?
    struct BandInfo* info;
    unsigned long long* vq;  // [vq_cap] pass << 40 | sweep vertex of the run starts located "outside"
    int64_t vq_cap;
    size_t bytes;
};

constexpr int kRasterN = 512;
struct BandInfo {            // device memory, written by k_band_finalize
    int32_t rect[4][4];      // per pass: first line, lines, first segment, segments of the rectangle that is walked
    int64_t tstart[5];       // first thread of every pass in the per-segment launches (multiples of 256)
    int32_t ext[2][4];       // scratch: per OUTPUT pass the extent (Lmin, Lmax, kmin, kmax) of its relevant segments
    float sx, sy;            // raster cells per unit length over the input grid's bbox
    unsigned long long work[2];   // dynamic work distribution of the count / emit walks: next unclaimed rectangle position
    unsigned int vq_n;            // run starts located "outside" that wait for their exact check (k_band_verify_outside)
};

static Layout make_layout(void* ws, int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo)
{
    Layout l;
    l.nxi = nxi; l.nyi = nyi; l.nxo = nxo; l.nyo = nyo;
    l.Vi = nxi * nyi; l.Vo = nxo * nyo;
    l.Ci = (nxi - 1) * (nyi - 1); l.Co = (nxo - 1) * (nyo - 1);
    Carver c(ws);
    l.area_in = c.take<double>(l.Ci);
    l.bbox = c.take<double>(8);
    carve_boundary(c, l.bnd[0], nxi, nyi);
    carve_boundary(c, l.bnd[1], nxo, nyo);
    l.guess[0] = c.take<int32_t>(l.Vo);
    l.guess[1] = c.take<int32_t>(l.Vi);
    for (int p = 0; p < 4; p++) {
        const bool sweep_in = p >= 2;
        const int64_t nx = sweep_in ? nxi : nxo, ny = sweep_in ? nyi : nyo;
        const int axis = p & 1;
        l.line_start[p] = c.take<int32_t>(axis ? nx : ny);
        l.line_bad[p] = c.take<int32_t>(axis ? nx : ny);
        l.seg_start[p] = c.take<int32_t>(nx * ny);
        l.seg_end[p] = c.take<int32_t>(nx * ny);
        l.seg_hit[p] = c.take<uint8_t>(nx * ny);
        l.pc_n[p] = c.take<uint8_t>(nx * ny);
        l.pc_cell[p] = c.take<int32_t>(kPieceCache * nx * ny);
        l.pc_x[p] = c.take<double>(kPieceCache * nx * ny);
        l.pc_y[p] = c.take<double>(kPieceCache * nx * ny);
    }
    l.hist = c.take<int32_t>(l.Ci + 1);
    l.boff = c.take<int64_t>(l.Ci + 1);
    l.cursor = c.take<int32_t>(l.Ci + 1);
    l.nuniq = c.take<int32_t>(l.Ci + 1);
    l.colptr = c.take<int64_t>(l.Ci + 1);
    l.scan_scratch = c.take<int64_t>(scan_scratch_elems(l.Ci));
    l.scan_status[0] = c.take<unsigned long long>(2 * scan_status_elems(l.Ci));
    l.scan_status[1] = l.scan_status[0] + scan_status_elems(l.Ci);
    l.flags = c.take<int32_t>(8);
    l.raster8 = c.take<uint8_t>((size_t)kRasterN * kRasterN);
    l.raster = c.take<uint32_t>((size_t)kRasterN * kRasterN / 8);
    l.rel[0] = c.take<uint8_t>(l.Vo);
    l.rel[1] = c.take<uint8_t>(l.Vo);
    l.info = c.take<BandInfo>(1);
    l.vq_cap = 8 * (nxo + nyo) + 1024;
    l.vq = c.take<unsigned long long>((size_t)l.vq_cap);
    l.bytes = c.total();
    return l;
}

static int check_sizes(int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo)
{
    if (nxi < 2 || nyi < 2 || nxo < 2 || nyo < 2) return fail(RG_E_ARG, "rg_build2d: grids need at least 2x2 vertices");
    if (nxi * nyi >= INT32_MAX || nxo * nyo >= INT32_MAX)
        return fail(RG_E_TOO_LARGE, "rg_build2d: grid exceeds the int32 cell index range");
    return RG_OK;
}

static PassParams make_pass(const Layout& l, int p, const double* xin, const double* yin,
                            const double* xout, const double* yout, int64_t cell_lo, int64_t cell_hi,
                            int part_rank = 0, int part_world = 1)
{
    PassParams P;
    const GridView gin{ xin, yin, (int)l.nxi, (int)l.nyi };
    const GridView gout{ xout, yout, (int)l.nxo, (int)l.nyo };
    P.pass = p;
    P.sweep_input = p >= 2;
    P.axis = p & 1;
    P.sweep = P.sweep_input ? gin : gout;
    P.stat = P.sweep_input ? gout : gin;
    P.bnd = l.bnd[P.sweep_input ? 1 : 0];
    P.nlines = P.axis ? P.sweep.nx : P.sweep.ny;
    P.nseg = P.axis ? P.sweep.ny - 1 : P.sweep.nx - 1;
    P.ncy_in = (int)l.nyi - 1;
    P.ncy_out = (int)l.nyo - 1;
    P.ncx_st = P.stat.nx - 1;
    P.ncy_st = P.stat.ny - 1;
    P.guess = l.guess[P.sweep_input ? 1 : 0];
    P.line_start = l.line_start[p];
    P.line_bad = l.line_bad[p];
    P.seg_start = l.seg_start[p];
    P.seg_end = l.seg_end[p];
    P.max_iter = 4 * (P.ncx_st + P.ncy_st) + 64;
    P.cell_lo = cell_lo;
    P.cell_hi = cell_hi;
    P.seg_hit = l.seg_hit[p];
    P.pc_n = l.pc_n[p];
    P.pc_cell = l.pc_cell[p];
    P.pc_x = l.pc_x[p];
    P.pc_y = l.pc_y[p];
    P.pc_stride = (int64_t)P.sweep.nx * P.sweep.ny;
    P.banded = (cell_lo > 0 || cell_hi < l.Ci) ? 1 : 0;
    P.part_rank = part_rank;
    P.part_world = part_world;
    if (part_world == 1) {
        P.nl_slots = P.nlines;
    } else {
        const int nblocks = (P.nlines + 31) / 32;
        const int mine = nblocks > part_rank ? (nblocks - part_rank + part_world - 1) / part_world : 0;
        P.nl_slots = mine * 32;
    }
    return P;
}

// the sort kernels use more than the 48 KB of shared memory a kernel gets by default (per device, idempotent)
static int sort_smem_opt_in(int device)
{
    static bool done[64] = {};  // per device; the attribute is idempotent, so a race only repeats the calls
    if (device >= 0 && device < 64 && done[device]) return RG_OK;
    RG_CUDA(cudaFuncSetAttribute(k_bucket_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem)));
    RG_CUDA(cudaFuncSetAttribute(k_bucket_gather_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem)));
    if (device >= 0 && device < 64) done[device] = true;
    return RG_OK;
}

static Pass4 make_pass4(const Layout& l, const double* xin, const double* yin, const double* xout, const double* yout,
                        int64_t cell_lo, int64_t cell_hi, int part_rank, int part_world)
{
    Pass4 Q;
    Q.tstart[0] = 0;
    Q.lstart[0] = 0;
    for (int p = 0; p < 4; p++) {
        Q.p[p] = make_pass(l, p, xin, yin, xout, yout, cell_lo, cell_hi, part_rank, part_world);
        Q.tstart[p + 1] = Q.tstart[p] + ceil_div((int64_t)Q.p[p].nl_slots * Q.p[p].nseg, 256) * 256;
        Q.lstart[p + 1] = Q.lstart[p] + Q.p[p].nl_slots;
    }
    return Q;
}

}  // namespace rg

using namespace rg;

extern "C" int rg_build2d_workspace_bytes(int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo, size_t* bytes_host)
{
    if (!bytes_host) return fail(RG_E_ARG, "rg_build2d_workspace_bytes: null output");
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    *bytes_host = make_layout(nullptr, nxi, nyi, nxo, nyo).bytes;
    return RG_OK;
}

extern "C" int rg_grid_area(int device, void* stream, int64_t nx, int64_t ny,
                            const double* x, const double* y, double* area)
{
    if (nx < 2 || ny < 2 || !x || !y || !area) return fail(RG_E_ARG, "rg_grid_area: bad argument");
    if (nx * ny >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_grid_area: grid too large");
    RG_CUDA(cudaSetDevice(device));
    const GridView g{ x, y, (int)nx, (int)ny };
    const int64_t nc = (nx - 1) * (ny - 1);
    k_cell_area<<<(unsigned)ceil_div(nc, 256), 256, 0, (cudaStream_t)stream>>>(g, area);
    RG_LAUNCH_CHECK("k_cell_area");
    return RG_OK;
}

static int count_impl(int device, void* stream,
                      int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                      const double* xin, const double* yin, const double* xout, const double* yout,
                      int64_t cell_lo, int64_t cell_hi, int part_rank, int part_world,
                      void* workspace, size_t workspace_bytes, int64_t* n_fragments_host,
                      int n_bounds, const int64_t* cell_bounds_host, int64_t* frag_offsets_host)
{
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    if (!xin || !yin || !xout || !yout || !workspace || !n_fragments_host)
        return fail(RG_E_ARG, "rg_build2d_count: null pointer");
    Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_count: workspace too small");
    if (cell_lo < 0 || cell_hi > l.Ci || cell_lo > cell_hi) return fail(RG_E_ARG, "rg_build2d_count: bad cell band");
    if (part_world < 1 || part_rank < 0 || part_rank >= part_world)
        return fail(RG_E_ARG, "rg_build2d_count: bad line partition");
    if (n_bounds < 0 || n_bounds > 1024 || (n_bounds > 0 && (!cell_bounds_host || !frag_offsets_host)))
        return fail(RG_E_ARG, "rg_build2d_count: bad band bounds");
    for (int b = 0; b < n_bounds; b++)
        if (cell_bounds_host[b] < 0 || cell_bounds_host[b] > l.Ci)
            return fail(RG_E_ARG, "rg_build2d_count: band bound outside the input cells");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const GridView gin{ xin, yin, (int)nxi, (int)nyi };
    const GridView gout{ xout, yout, (int)nxo, (int)nyo };
    const int T = 256;

    RG_CUDA(cudaMemsetAsync(l.flags, 0, sizeof(int32_t) * 8, st));
    RG_CUDA(cudaMemsetAsync(l.hist, 0, sizeof(int32_t) * (size_t)(l.Ci + 1), st));
    // K1
    k_cell_area<<<(unsigned)ceil_div(l.Ci, T), T, 0, st>>>(gin, l.area_in);
    RG_LAUNCH_CHECK("k_cell_area");
    // bounding boxes + boundaries of both grids
    {
        const GridView gv[2] = { gin, gout };
        double* const bb[2] = { l.bbox, l.bbox + 4 };
        rc = build_boundaries(st, 2, gv, l.bnd, bb);
        if (rc) return rc;
    }
    const Pass4 Q = make_pass4(l, xin, yin, xout, yout, cell_lo, cell_hi, part_rank, part_world);
    if (Q.lstart[4] > 0 && Q.tstart[4] > 0) {
        if (part_world == 1) {
            // vertex guesses: output vertices in the input grid, input vertices in the output grid
            k_vertex_guess<<<(unsigned)ceil_div(l.Vo, T), T, 0, st>>>(gout, gin, l.guess[0], l.flags);
            k_vertex_guess<<<(unsigned)ceil_div(l.Vi, T), T, 0, st>>>(gin, gout, l.guess[1], l.flags);
        } else {
            // only the vertices on this rank's lines (the two axis passes of a grid use different lines)
            k_vertex_guess_part<<<(unsigned)ceil_div(Q.tstart[4], T), T, 0, st>>>(Q, l.flags);
        }
        RG_LAUNCH_CHECK("k_vertex_guess");
        k_line_starts<<<(unsigned)Q.lstart[4], 256, 0, st>>>(Q, l.bbox);
        RG_LAUNCH_CHECK("k_line_starts");
        k_walk_count<<<(unsigned)ceil_div(Q.tstart[4], 128), 128, 0, st>>>(Q, l.hist);
        RG_LAUNCH_CHECK("k_walk_count");
        k_chain_check<<<(unsigned)ceil_div(Q.tstart[4], T), T, 0, st>>>(Q);
        RG_LAUNCH_CHECK("k_chain_check");
        k_repair<<<(unsigned)ceil_div((int64_t)Q.lstart[4] * 32, 128), 128, 0, st>>>(Q, l.hist, l.flags);
        RG_LAUNCH_CHECK("k_repair");
    }
    rc = exclusive_scan_i32_i64(st, l.hist, l.boff, l.Ci, l.scan_scratch);
    if (rc) return rc;
    int64_t total = 0;
    RG_CUDA(cudaMemcpyAsync(&total, l.boff + l.Ci, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    for (int b = 0; b < n_bounds; b++)
        RG_CUDA(cudaMemcpyAsync(frag_offsets_host + b, l.boff + cell_bounds_host[b], sizeof(int64_t),
                                cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    if (total >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_build2d_count: more than 2^31 fragments");
    *n_fragments_host = total;
    return RG_OK;
}

extern "C" int rg_build2d_count(int device, void* stream,
                                int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                                const double* xin, const double* yin, const double* xout, const double* yout,
                                int64_t cell_lo, int64_t cell_hi,
                                void* workspace, size_t workspace_bytes, int64_t* n_fragments_host)
{
    return count_impl(device, stream, nxi, nyi, nxo, nyo, xin, yin, xout, yout, cell_lo, cell_hi, 0, 1,
                      workspace, workspace_bytes, n_fragments_host, 0, nullptr, nullptr);
}

extern "C" int rg_build2d_fill(int device, void* stream,
                               int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                               const double* xin, const double* yin, const double* xout, const double* yout,
                               const double* w_in, int64_t cell_lo, int64_t cell_hi,
                               void* workspace, size_t workspace_bytes,
                               void* frags, int64_t n_fragments, int64_t* nnz_host)
{
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    if (!workspace || !nnz_host || (n_fragments > 0 && !frags))
        return fail(RG_E_ARG, "rg_build2d_fill: null pointer");
    Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_fill: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaMemsetAsync(l.cursor, 0, sizeof(int32_t) * (size_t)(l.Ci + 1), st));
    {
        const Pass4 Q = make_pass4(l, xin, yin, xout, yout, cell_lo, cell_hi, 0, 1);
        k_walk_emit<<<(unsigned)ceil_div(Q.tstart[4], 128), 128, 0, st>>>(Q, l.boff, l.cursor, (Frag*)frags,
                                                                         l.area_in, w_in, l.flags);
        RG_LAUNCH_CHECK("k_walk_emit");
    }
    rc = sort_smem_opt_in(device);
    if (rc) return rc;
    k_bucket_sort<<<(unsigned)ceil_div(l.Ci, kSortCells), kSortThreads, sizeof(SortSmem), st>>>(l.boff, l.Ci, (Frag*)frags,
                                                                                           l.nuniq, nullptr, nullptr, 0,
                                                                                           l.flags + kFlagMaxBucket);
    RG_LAUNCH_CHECK("k_bucket_sort");
    rc = exclusive_scan_i32_i64(st, l.nuniq, l.colptr, l.Ci, l.scan_scratch);
    if (rc) return rc;
    int64_t nnz = 0;
    int32_t flags[8];
    RG_CUDA(cudaMemcpyAsync(&nnz, l.colptr + l.Ci, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(flags, l.flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    if (flags[kFlagOverflow] || flags[kFlagSeqOverflow])
        return fail(RG_E_WALK, "rg_build2d_fill: a sweep walk did not terminate (degenerate or folded grid)");
    *nnz_host = nnz;
    return RG_OK;
}

extern "C" int rg_build2d_emit(int device, void* stream,
                               int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                               int64_t cell_lo, int64_t cell_hi,
                               void* workspace, size_t workspace_bytes,
                               const void* frags, int64_t n_fragments,
                               int64_t* ii, int64_t* io, double* v, int64_t nnz)
{
    (void)cell_lo; (void)cell_hi; (void)n_fragments;
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    if (!workspace) return fail(RG_E_ARG, "rg_build2d_emit: null workspace");
    if (nnz > 0 && (!ii || !io || !v || !frags)) return fail(RG_E_ARG, "rg_build2d_emit: null pointer");
    Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_emit: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz > 0) {
        k_bucket_emit<<<(unsigned)ceil_div(l.Ci, 256), 256, 0, st>>>(l.boff, 1, 0, l.colptr, l.Ci, (const Frag*)frags, ii, io, v);
        RG_LAUNCH_CHECK("k_bucket_emit");
    }
    return RG_OK;
}

// Debug/diagnostic: copy the 8 status counters of the last build out of the workspace.
extern "C" int rg_build2d_stats(int device, void* stream, int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                                void* workspace, int32_t* stats_host /* 8 */)
{
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
    RG_CUDA(cudaSetDevice(device));
    RG_CUDA(cudaMemcpyAsync(stats_host, l.flags, sizeof(int32_t) * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return RG_OK;
}

// ---------------------------------------------------------------------------
// line-sharded build: every rank walks 1/W of the sweep lines of all four passes, fragments travel
// to the rank that owns their input-row band, which sorts and merges them.  The emission rank inside a
// fragment key does not depend on who walked the segment, so the merged band is bit-identical to the
// same rows of a single-GPU build.
// ---------------------------------------------------------------------------
extern "C" int rg_build2d_part_count(int device, void* stream,
                                     int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                                     const double* xin, const double* yin, const double* xout, const double* yout,
                                     int part_rank, int part_world,
                                     void* workspace, size_t workspace_bytes, int64_t* n_fragments_host,
                                     int n_bounds, const int64_t* cell_bounds_host, int64_t* frag_offsets_host,
                                     size_t* counts_offset_host, int64_t* frag_offsets_dev_or_null)
{
    int rc = count_impl(device, stream, nxi, nyi, nxo, nyo, xin, yin, xout, yout, 0, (nxi - 1) * (nyi - 1),
                        part_rank, part_world, workspace, workspace_bytes, n_fragments_host,
                        n_bounds, cell_bounds_host, frag_offsets_host);
    if (rc) return rc;
    if (counts_offset_host) {
        Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
        *counts_offset_host = (size_t)((char*)l.hist - (char*)workspace);
    }
    if (frag_offsets_dev_or_null && n_bounds > 0)  // for the peers of a P2P exchange (they read it after a barrier)
        RG_CUDA(cudaMemcpyAsync(frag_offsets_dev_or_null, frag_offsets_host, sizeof(int64_t) * n_bounds,
                                cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return RG_OK;
}

extern "C" int rg_build2d_part_fill(int device, void* stream,
                                    int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                                    const double* xin, const double* yin, const double* xout, const double* yout,
                                    const double* w_in, int part_rank, int part_world,
                                    void* workspace, size_t workspace_bytes,
                                    void* frags, int64_t n_fragments)
{
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    if (!workspace || (n_fragments > 0 && !frags))
        return fail(RG_E_ARG, "rg_build2d_part_fill: null pointer");
    if (part_world < 1 || part_rank < 0 || part_rank >= part_world)
        return fail(RG_E_ARG, "rg_build2d_part_fill: bad line partition");
    Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_part_fill: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    RG_CUDA(cudaMemsetAsync(l.cursor, 0, sizeof(int32_t) * (size_t)(l.Ci + 1), st));
    {
        const Pass4 Q = make_pass4(l, xin, yin, xout, yout, 0, l.Ci, part_rank, part_world);
        if (Q.tstart[4] > 0) {
            k_walk_emit<<<(unsigned)ceil_div(Q.tstart[4], 128), 128, 0, st>>>(Q, l.boff, l.cursor, (Frag*)frags,
                                                                             l.area_in, w_in, l.flags);
            RG_LAUNCH_CHECK("k_walk_emit");
        }
    }
    return RG_OK;
}

namespace rg {

struct MergeLayout {
    int32_t* cntT;
    int64_t* src_off;
    int64_t* dst_off;
    int32_t* nuniq;
    int64_t* colptr;
    int64_t* scan_scratch;
    size_t bytes;
};

static MergeLayout make_merge_layout(void* ws, int64_t Cb, int W)
{
    MergeLayout m;
    Carver c(ws);
    const int64_t n = Cb * W;
    m.cntT = c.take<int32_t>(n + 1);
    m.src_off = c.take<int64_t>(n + 1);
    m.dst_off = c.take<int64_t>(n + 1);
    m.nuniq = c.take<int32_t>(Cb + 1);
    m.colptr = c.take<int64_t>(Cb + 1);
    m.scan_scratch = c.take<int64_t>(scan_scratch_elems(n));
    m.bytes = c.total();
    return m;
}

__global__ void k_merge_transpose(const int32_t* __restrict__ cnt, int32_t* __restrict__ cntT, int64_t Cb, int W)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index into cntT: c * W + s
    if (i >= Cb * W) return;
    const int64_t c = i / W;
    const int s = (int)(i - c * W);
    cntT[i] = cnt[(int64_t)s * Cb + c];
}

struct CountSrc {
    const int32_t* src[kMaxSrc];
};

// counts[s][c] = src[s][c]: the band slices of the sources' per-cell counts (possibly peer memory)
__global__ void k_gather_counts(const __grid_constant__ CountSrc S, int64_t Cb, int32_t* __restrict__ counts)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cb) return;
    counts[(int64_t)blockIdx.y * Cb + c] = __ldg(S.src[blockIdx.y] + c);
}

}  // namespace rg

extern "C" int rg_build2d_merge_workspace_bytes(int64_t n_cells, int n_src, size_t* bytes_host)
{
    if (!bytes_host || n_cells < 0 || n_src < 1 || n_src > kMaxSrc)
        return fail(RG_E_ARG, "rg_build2d_merge_workspace_bytes: bad argument");
    if (n_cells * n_src >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_build2d_merge_workspace_bytes: too large");
    *bytes_host = make_merge_layout(nullptr, n_cells, n_src).bytes;
    return RG_OK;
}

extern "C" int rg_build2d_gather_counts(int device, void* stream, int64_t n_cells, int n_src,
                                        const int32_t* const* src_counts_host, int32_t* counts)
{
    if (n_cells < 0 || n_src < 1 || n_src > kMaxSrc || !src_counts_host || (n_cells > 0 && !counts))
        return fail(RG_E_ARG, "rg_build2d_gather_counts: bad argument");
    RG_CUDA(cudaSetDevice(device));
    if (n_cells == 0) return RG_OK;
    CountSrc S;
    memset(&S, 0, sizeof(S));
    for (int s = 0; s < n_src; s++) {
        if (!src_counts_host[s]) return fail(RG_E_ARG, "rg_build2d_gather_counts: null source");
        S.src[s] = src_counts_host[s];
    }
    k_gather_counts<<<dim3((unsigned)ceil_div(n_cells, 256), n_src), 256, 0, (cudaStream_t)stream>>>(S, n_cells, counts);
    RG_LAUNCH_CHECK("k_gather_counts");
    return RG_OK;
}

extern "C" int rg_build2d_merge(int device, void* stream, int64_t n_cells, int n_src,
                                const int32_t* counts /* [n_src][n_cells] */,
                                const void* const* src_chunks_host, const int64_t* src_sizes_host,
                                void* workspace, size_t workspace_bytes,
                                void* frags, int64_t* nnz_host)
{
    if (n_cells < 0 || n_src < 1 || n_src > kMaxSrc || !workspace || !nnz_host || (n_cells > 0 && !counts) ||
        !src_chunks_host || !src_sizes_host)
        return fail(RG_E_ARG, "rg_build2d_merge: bad argument");
    int64_t n_recv = 0;
    for (int s = 0; s < n_src; s++) {
        if (src_sizes_host[s] < 0 || (src_sizes_host[s] > 0 && !src_chunks_host[s]))
            return fail(RG_E_ARG, "rg_build2d_merge: bad source chunk");
        n_recv += src_sizes_host[s];
    }
    if (n_recv > 0 && !frags) return fail(RG_E_ARG, "rg_build2d_merge: null fragment buffer");
    if (n_cells * n_src >= INT32_MAX || n_recv >= INT32_MAX) return fail(RG_E_TOO_LARGE, "rg_build2d_merge: too large");
    MergeLayout m = make_merge_layout(workspace, n_cells, n_src);
    if (workspace_bytes < m.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_merge: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    *nnz_host = 0;
    if (n_cells == 0) return RG_OK;
    int rc0 = sort_smem_opt_in(device);
    if (rc0) return rc0;
    const int64_t n = n_cells * n_src;
    k_merge_transpose<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(counts, m.cntT, n_cells, n_src);
    RG_LAUNCH_CHECK("k_merge_transpose");
    int rc = exclusive_scan_i32_i64(st, counts, m.src_off, n, m.scan_scratch);
    if (rc) return rc;
    rc = exclusive_scan_i32_i64(st, m.cntT, m.dst_off, n, m.scan_scratch);
    if (rc) return rc;
    GatherSrc G;
    memset(&G, 0, sizeof(G));
    G.cntT = m.cntT; G.src_off = m.src_off; G.dst_off = m.dst_off; G.W = n_src; G.Cb = n_cells;
    int64_t before = 0;
    for (int s = 0; s < n_src; s++) {
        G.src[s] = (const Frag*)src_chunks_host[s] - before;  // src_off[s * Cb] == fragments of the sources before s
        before += src_sizes_host[s];
    }
    k_bucket_gather_sort<<<(unsigned)ceil_div(n_cells, kSortCells), kSortThreads, sizeof(SortSmem), st>>>(
        G, n_cells, (Frag*)frags, m.nuniq);
    RG_LAUNCH_CHECK("k_bucket_gather_sort");
    rc = exclusive_scan_i32_i64(st, m.nuniq, m.colptr, n_cells, m.scan_scratch);
    if (rc) return rc;
    int64_t nnz = 0, total = 0;
    RG_CUDA(cudaMemcpyAsync(&nnz, m.colptr + n_cells, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaMemcpyAsync(&total, m.src_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RG_CUDA(cudaStreamSynchronize(st));
    if (total != n_recv) return fail(RG_E_ARG, "rg_build2d_merge: the counts do not add up to the chunk sizes");
    *nnz_host = nnz;
    return RG_OK;
}

extern "C" int rg_build2d_merge_emit(int device, void* stream, int64_t n_cells, int n_src, int64_t cell_offset,
                                     void* workspace, size_t workspace_bytes,
                                     const void* frags,
                                     int64_t* ii, int64_t* io, double* v, int64_t nnz)
{
    if (n_cells < 0 || n_src < 1 || n_src > kMaxSrc || !workspace) return fail(RG_E_ARG, "rg_build2d_merge_emit: bad argument");
    if (nnz > 0 && (!ii || !io || !v || !frags)) return fail(RG_E_ARG, "rg_build2d_merge_emit: null pointer");
    MergeLayout m = make_merge_layout(workspace, n_cells, n_src);
    if (workspace_bytes < m.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_merge_emit: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    if (nnz > 0 && n_cells > 0) {
        k_bucket_emit<<<(unsigned)ceil_div(n_cells, 256), 256, 0, (cudaStream_t)stream>>>(
            m.dst_off, n_src, cell_offset, m.colptr, n_cells, (const Frag*)frags, ii, io, v);
        RG_LAUNCH_CHECK("k_bucket_emit");
    }
    return RG_OK;
}

// ===========================================================================
// Band build: ONE large grid pair built by W ranks with no exchange of fragments (rg_build2d_band).
//
// Rank r owns a band of input rows [row_lo, row_hi).  The public layout is sorted by input cell first
// (_weights_arrays.py:54-59), so the bands concatenate in rank order.  A rank only walks the sweep segments that
// can produce a piece inside its band:
//   * INPUT-line passes: the lines / segments bounding the band's cells (closed form in index space);
//   * OUTPUT-line passes: a piece inside band cell C lies inside C, so a segment can only contribute if its
//     bounding box meets the bounding box of some cell of the band.  The band's cells are rasterised into a
//     kRasterN^2 bitmap over the input grid's bbox (every cell marks every raster cell its bbox touches) and a
//     segment is RELEVANT iff the raster cells under its own bbox contain a mark: exact and conservative.
// Exactness of the walk states: the reference walks a line sequentially (c2d.py:324-387); here every walked
// segment starts from the located cell of its first vertex, and the chain "end state of k-1 == start state of k"
// is verified for every walked pair -- the predecessor of a relevant segment is always walked too (halo).  The FIRST
// segment of a run of walked segments has no walked predecessor: its start state is verified EXACTLY instead
// (band_start_exact: strictly inside the located cell beyond rounding error, or boundary winding number 0; the first
// segment of a LINE starts from the exact line-start state).  Every state a rank uses is thus verified by the rank
// itself -- a band needs no agreement with the other ranks, hence no collective.  Any failure (never observed: it
// needs a vertex within rounding error of a cell edge) raises kFlagBandMismatch and the caller rebuilds that band
// with the sequentially verified banded build (rg_build2d_count/_fill), on its own.
//
// No host synchronisation inside: buffers are sized by the caller's estimate, overflow raises kFlagCapacity
// and the counts come back in `counts_dev` (the caller reads them once, after everything was enqueued).
// ===========================================================================
namespace rg {

struct BandParams {
    int row_lo, row_hi;        // band of input rows
    const uint8_t* rel[2];     // relevance of the pass-0 / pass-1 segment starting at each output vertex
    const BandInfo* info;
    unsigned long long* vq;    // run starts located "outside" (filled by the emit walk, checked by k_band_verify_outside)
    int64_t vq_cap;
};

// raster cells per unit length along x (axis 0) / y (axis 1) over the input grid's bbox: ONE expression for the kernel
// that marks and the kernel that tests
__device__ __forceinline__ float band_raster_scale(const double* __restrict__ bbox, int axis)
{
    return (float)(kRasterN / (bbox[2 + axis] - bbox[axis]));
}

__device__ __forceinline__ int raster_index(double x, double lo, float scale)
{
    // any NON-DECREASING function of x does (overlapping intervals then give overlapping index ranges), as long as
    // marking and testing use the same one: the difference is rounded to fp32 (monotone), scaled and floored
    const int t = __float2int_rd(__fmul_rn((float)(x - lo), scale));
    return min(max(t, 0), kRasterN - 1);
}

// every cell of the band marks the raster cells its bounding box touches
__global__ void k_band_raster(GridView g, int row_lo, int row_hi, const double* __restrict__ bbox,
                              const BandInfo* __restrict__ info, uint8_t* __restrict__ raster8)
{
    const int ncy = g.ny - 1;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)(row_hi - row_lo) * ncy) return;
    const int a = row_lo + (int)(c / ncy), b = (int)(c % ncy);
    const int64_t v = (int64_t)a * g.ny + b;
    const double x0 = g.x[v], x1 = g.x[v + 1], x2 = g.x[v + g.ny], x3 = g.x[v + g.ny + 1];
    const double y0 = g.y[v], y1 = g.y[v + 1], y2 = g.y[v + g.ny], y3 = g.y[v + g.ny + 1];
    const float sx = band_raster_scale(bbox, 0), sy = band_raster_scale(bbox, 1);
    const int ix0 = raster_index(fmin(fmin(x0, x1), fmin(x2, x3)), bbox[0], sx);
    const int ix1 = raster_index(fmax(fmax(x0, x1), fmax(x2, x3)), bbox[0], sx);
    const int iy0 = raster_index(fmin(fmin(y0, y1), fmin(y2, y3)), bbox[1], sy);
    const int iy1 = raster_index(fmax(fmax(y0, y1), fmax(y2, y3)), bbox[1], sy);
    for (int ix = ix0; ix <= ix1; ix++)
        for (int iy = iy0; iy <= iy1; iy++) raster8[ix * kRasterN + iy] = 1;   // plain stores: no ordering needed
}

// byte marks -> 4 bits per raster cell (x, y), 8 cells per 32-bit word: bit 0 = the cell itself, bit 1 = the cell or
// (x, y+1), bit 2 = the cell or (x+1, y), bit 3 = any cell of the 2x2 block at (x, y).  A segment whose ends lie in
// the same or in neighbouring raster cells (every segment of a grid finer than the raster) is then tested EXACTLY
// with one load: bit (dx << 1 | dy) of the nibble at the lower corner of its raster bounding box.
__global__ void k_band_raster_pack(const uint8_t* __restrict__ raster8, uint32_t* __restrict__ raster)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;   // word q holds cells 8q .. 8q+7 of the flattened raster
    if (q >= kRasterN * kRasterN / 8) return;
    const int x = (q * 8) / kRasterN, y0 = (q * 8) % kRasterN;
    const bool has_x1 = x + 1 < kRasterN;
    uint32_t word = 0;
    bool r_prev = raster8[x * kRasterN + y0] != 0;
    bool d_prev = has_x1 && raster8[(x + 1) * kRasterN + y0] != 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int y1 = y0 + t + 1;
        const bool r_next = y1 < kRasterN && raster8[x * kRasterN + y1] != 0;
        const bool d_next = has_x1 && y1 < kRasterN && raster8[(x + 1) * kRasterN + y1] != 0;
        const uint32_t nib = (uint32_t)r_prev | ((uint32_t)(r_prev || r_next) << 1) | ((uint32_t)(r_prev || d_prev) << 2) |
                             ((uint32_t)(r_prev || r_next || d_prev || d_next) << 3);
        word |= nib << (4 * t);
        r_prev = r_next;
        d_prev = d_next;
    }
    raster[q] = word;
}

// Relevance of the two sweep segments that start at every OUTPUT vertex (i, j) (pass 1: to (i, j+1); pass 0: to
// (i+1, j)), the extent of the relevant segments in (line, segment) space and the bounding box of the output grid:
// ONE streaming pass over the output grid.  A segment is relevant iff a raster cell under its bounding box is marked
// (indices are clamped to the raster: segments beyond the input grid's bbox can only over-report).  The packed
// raster is served by L1; a warp takes 32 consecutive vertices of a row: the raster cell of the right neighbour
// comes from the next lane, the one of the lower neighbour from its coordinates (that row is in L2 by the time it is
// reached again).  Extents and bbox are reduced in registers over the grid-stride loop, then per CTA (a few atomics
// per CTA).  The kernel is instruction bound (ncu): fp32 raster indices, 32-bit index arithmetic, one-word test when
// both ends of a segment fall into one raster cell.
constexpr int kRelThreads = 256;
__global__ void __launch_bounds__(kRelThreads, 6) k_band_relevance(GridView gout, const double* __restrict__ bbox_in,
                                                                   const uint32_t* __restrict__ raster,
                                                                   uint8_t* __restrict__ rel0, uint8_t* __restrict__ rel1,
                                                                   BandInfo* __restrict__ info)
{
    __shared__ int s_ext[kRelThreads / 32][8];
    const uint32_t* __restrict__ s_raster = raster;   // 128 KB, a warp touches one or two sectors of it: served by L1
    const int nx = gout.nx, ny = gout.ny;
    const double x_lo = bbox_in[0], y_lo = bbox_in[1];
    const float sx = band_raster_scale(bbox_in, 0), sy = band_raster_scale(bbox_in, 1);
    const int lane = threadIdx.x & 31;
    int ext[8] = { INT32_MAX, -1, INT32_MAX, -1, INT32_MAX, -1, INT32_MAX, -1 };  // pass 0: Lmin Lmax kmin kmax; pass 1
    auto cell_of = [&](double x, double y) -> uint32_t {
        return (unsigned)raster_index(x, x_lo, sx) | ((unsigned)raster_index(y, y_lo, sy) << 16);
    };
    auto marked = [&](uint32_t ca, uint32_t cb) -> bool {
        const int ixa = (int)(ca & 0xffffu), iya = (int)(ca >> 16), ixb = (int)(cb & 0xffffu), iyb = (int)(cb >> 16);
        const int xa = min(ixa, ixb), xb = max(ixa, ixb), ya = min(iya, iyb), yb = max(iya, iyb);
        const int dx = xb - xa, dy = yb - ya;
        if ((dx | dy) <= 1) {   // one load: nibble of the lower corner, bit by the shape of the bounding box
            const int cell = xa * kRasterN + ya;
            return (s_raster[cell >> 3] >> (4 * (cell & 7) + ((dx << 1) | dy))) & 1u;
        }
        for (int a = xa; a <= xb; a++)
            for (int b = ya; b <= yb; b++) {
                const int cell = a * kRasterN + b;
                if ((s_raster[cell >> 3] >> (4 * (cell & 7))) & 1u) return true;
            }
        return false;
    };
    // a chunk = 32 consecutive vertices of ONE row (the last chunk of a row is partial)
    const unsigned cpr = (unsigned)(ny + 31) / 32u;
    const unsigned nchunk = cpr * (unsigned)nx;
    const unsigned wstride = gridDim.x * (blockDim.x >> 5);
    unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (; w < nchunk; w += wstride) {
        const int i = (int)(w / cpr), j = (int)(w - (unsigned)i * cpr) * 32 + lane;
        const bool in = j < ny;
        const bool has_r = in && j + 1 < ny, has_d = in && i + 1 < nx;
        const int64_t v = (int64_t)i * ny + j;
        double x = 0.0, y = 0.0, xd = 0.0, yd = 0.0;
        if (in) { x = gout.x[v]; y = gout.y[v]; }
        if (has_d) { xd = gout.x[v + ny]; yd = gout.y[v + ny]; }
        const uint32_t c = cell_of(x, y);
        uint32_t cr = __shfl_down_sync(0xffffffffu, c, 1);
        if (lane == 31 && has_r) cr = cell_of(gout.x[v + 1], gout.y[v + 1]);
        const uint32_t cd = cell_of(xd, yd);
        const bool r1 = has_r && marked(c, cr);
        const bool r0 = has_d && marked(c, cd);
        if (in) {
            rel0[v] = r0;
            rel1[v] = r1;
        }
        if (r0) { ext[0] = min(ext[0], j); ext[1] = max(ext[1], j); ext[2] = min(ext[2], i); ext[3] = max(ext[3], i); }  // pass 0: line = j, segment = i
        if (r1) { ext[4] = min(ext[4], i); ext[5] = max(ext[5], i); ext[6] = min(ext[6], j); ext[7] = max(ext[7], j); }  // pass 1: line = i, segment = j
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int other = __shfl_xor_sync(0xffffffffu, ext[q], o);
            ext[q] = (q & 1) ? max(ext[q], other) : min(ext[q], other);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) s_ext[threadIdx.x >> 5][q] = ext[q];
    }
    __syncthreads();
    const int nw = (int)(blockDim.x >> 5);
    if (threadIdx.x < 8) {
        const int q = threadIdx.x;
        int e = s_ext[0][q];
        for (int u = 1; u < nw; u++) e = (q & 1) ? max(e, s_ext[u][q]) : min(e, s_ext[u][q]);
        if (e != ((q & 1) ? -1 : INT32_MAX)) {
            if (q & 1) atomicMax(&info->ext[q >> 2][q & 3], e);
            else atomicMin(&info->ext[q >> 2][q & 3], e);
        }
    }
}

// First launch of a band build: status flags and counters cleared, bounding boxes initialised for their reductions,
// the extents of the relevant output segments reset (or set to everything when the band is the whole grid), and the
// rectangles of the INPUT-line passes, which are known in closed form: the lines / segments bounding the band's cells.
__global__ void k_band_begin(const __grid_constant__ Pass4 Q, int row_lo, int row_hi, int full, int nx_out, int ny_out,
                             int32_t* __restrict__ flags, int64_t* __restrict__ counts, double* __restrict__ bbox2,
                             BandInfo* __restrict__ info)
{
    const int t = threadIdx.x;
    if (t < 8) {
        flags[t] = 0;
        counts[t] = 0;
        bbox2[t] = ((t & 3) < 2) ? INFINITY : -INFINITY;
        info->ext[t >> 2][t & 3] = (t & 1) ? -1 : INT32_MAX;
    }
    __syncthreads();
    if (t != 0) return;
    info->work[0] = info->work[1] = 0ull;
    info->sx = info->sy = 0.0f;
    info->vq_n = 0u;
    if (full) {
        // pass 0 (axis 0): line = j, segment = i; pass 1 (axis 1): line = i, segment = j
        info->ext[0][0] = 0; info->ext[0][1] = ny_out - 1; info->ext[0][2] = 0; info->ext[0][3] = nx_out - 2;
        info->ext[1][0] = 0; info->ext[1][1] = nx_out - 1; info->ext[1][2] = 0; info->ext[1][3] = ny_out - 2;
    }
    for (int p = 0; p < 4; p++) {
        const PassParams& P = Q.p[p];
        if (!P.sweep_input) continue;
        int L0, nL, k0, nK;
        if (P.axis) {   // line = input vertex row; lines row_lo .. row_hi bound the band's cells
            L0 = row_lo; nL = min(row_hi, P.nlines - 1) - row_lo + 1; k0 = 0; nK = P.nseg;
        } else {        // segment = input row
            L0 = 0; nL = P.nlines; k0 = max(row_lo - 1, 0); nK = row_hi - k0;
        }
        if (nL <= 0 || nK <= 0) { nL = 0; nK = 0; }
        info->rect[p][0] = L0; info->rect[p][1] = nL; info->rect[p][2] = k0; info->rect[p][3] = nK;
    }
}

// Located state of every output vertex that starts a walked segment (relevant, or the halo before a relevant one) or
// ends a relevant one, of either output pass: thread per vertex of the bounding rectangle of the two passes.
__global__ void k_band_guess_out(GridView gout, GridView gin, const uint8_t* __restrict__ rel0,
                                 const uint8_t* __restrict__ rel1, const BandInfo* __restrict__ info,
                                 int32_t* __restrict__ guess, int32_t* __restrict__ flags)
{
    // pass 0 (axis 0): line = j, segment = i; pass 1 (axis 1): line = i, segment = j
    const int* r0 = info->rect[0];
    const int* r1 = info->rect[1];
    int i_lo = INT32_MAX, i_hi = -1, j_lo = INT32_MAX, j_hi = -1;
    if (r0[1] > 0) { j_lo = min(j_lo, r0[0]); j_hi = max(j_hi, r0[0] + r0[1] - 1); i_lo = min(i_lo, r0[2]); i_hi = max(i_hi, r0[2] + r0[3]); }
    if (r1[1] > 0) { i_lo = min(i_lo, r1[0]); i_hi = max(i_hi, r1[0] + r1[1] - 1); j_lo = min(j_lo, r1[2]); j_hi = max(j_hi, r1[2] + r1[3]); }
    if (i_hi < i_lo || j_hi < j_lo) return;
    i_hi = min(i_hi, gout.nx - 1);
    j_hi = min(j_hi, gout.ny - 1);
    const int ni = i_hi - i_lo + 1, nj = j_hi - j_lo + 1;
    const int64_t total = (int64_t)ni * nj;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = i_lo + (int)(t / nj), j = j_lo + (int)(t % nj);
        const int64_t v = (int64_t)i * gout.ny + j;
        const int ny = gout.ny;
        bool need = rel1[v] || rel0[v];                                   // starts a relevant segment
        if (j + 1 < ny) need = need || rel1[v + 1];                      // starts the halo of pass 1
        if (i + 1 < gout.nx) need = need || rel0[v + ny];                // starts the halo of pass 0
        if (j > 0) need = need || rel1[v - 1];                           // ends a relevant segment of pass 1
        if (i > 0) need = need || rel0[v - ny];                          // ends a relevant segment of pass 0
        int r = kStateInvalid;  // never read by a walked segment (an invalid start raises the mismatch flag)
        if (need) {
            r = locate_guess(gin, gout.x[v], gout.y[v]);
            if (r == kLocUnknown) atomicAdd(&flags[kFlagUnknown], 1);
        }
        guess[v] = r;
    }
}

// located states of the input vertices around the band (the sweep vertices of passes 2 and 3)
__global__ void k_band_guess_in(GridView gin, GridView gout, int64_t v_lo, int64_t v_hi, int32_t* __restrict__ guess,
                                int32_t* __restrict__ flags)
{
    const int64_t v = v_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= v_hi) return;
    const int r = locate_guess(gout, gin.x[v], gin.y[v]);
    guess[v] = r;
    if (r == kLocUnknown) atomicAdd(&flags[kFlagUnknown], 1);
}

// rectangles in (line, segment) space that the per-segment kernels cover: the OUTPUT-line passes, from the extents
// of their relevant segments (the input-line passes: k_band_begin), and the first thread of every pass
__global__ void k_band_finalize(const __grid_constant__ Pass4 Q, BandInfo* info)
{
    if (threadIdx.x != 0) return;
    for (int p = 0; p < 4; p++) {
        const PassParams& P = Q.p[p];
        if (P.sweep_input) continue;
        int L0 = 0, nL = 0, k0 = 0, nK = 0;
        const int* e = info->ext[P.axis];  // ext[0]: pass 0 (axis 0), ext[1]: pass 1 (axis 1)
        if (e[1] >= e[0] && e[3] >= e[2]) {
            L0 = e[0]; nL = e[1] - e[0] + 1;
            k0 = max(e[2] - 1, 0);            // halo: the predecessor of the first relevant segment
            nK = e[3] - k0 + 1;
        }
        info->rect[p][0] = L0; info->rect[p][1] = nL; info->rect[p][2] = k0; info->rect[p][3] = nK;
    }
    info->tstart[0] = 0;
    for (int p = 0; p < 4; p++)
        info->tstart[p + 1] = info->tstart[p] + ((int64_t)info->rect[p][1] * info->rect[p][3] + 255) / 256 * 256;
}

__device__ __forceinline__ bool band_relevant(const PassParams& P, const BandParams& B, int L, int k, int64_t v)
{
    if (P.sweep_input) return P.axis ? (L >= B.row_lo && L <= B.row_hi) : (k >= B.row_lo && k < B.row_hi);
    return B.rel[P.axis][v] != 0;
}
// a segment is walked when it is relevant or the predecessor (halo) of a relevant one
__device__ __forceinline__ bool band_walked(const PassParams& P, const BandParams& B, int L, int k, int64_t v)
{
    return band_relevant(P, B, L, k, v) || (k + 1 < P.nseg && band_relevant(P, B, L, k + 1, v + vertex_step(P)));
}

__device__ __forceinline__ bool band_segment(const Pass4& Q, const BandInfo& I, int64_t gtid, int& p, int& L, int& k)
{
    p = (gtid >= I.tstart[1]) + (gtid >= I.tstart[2]) + (gtid >= I.tstart[3]);
    const int64_t t = gtid - I.tstart[p];
    const int L0 = I.rect[p][0], nL = I.rect[p][1], k0 = I.rect[p][2], nK = I.rect[p][3];
    if (t >= (int64_t)nL * nK) return false;
    // consecutive lanes take consecutive memory: along the line for axis 1, across lines for axis 0
    int slot;
    if (Q.p[p].axis) { slot = (int)(t / nK); k = k0 + (int)(t % nK); }
    else             { k = k0 + (int)(t / nL); slot = (int)(t % nL); }
    L = L0 + slot;
    return true;
}

// The walks visit the (line, segment) rectangles of the four passes, of which only the band's stripe is walked: a
// static assignment leaves some warps with many long walks and others with none, so in the count walk every WARP
// claims kClaim consecutive positions at a time from a global counter until the rectangles are exhausted (the emit
// walk replays cached pieces: short, uniform work, for which the static grid-stride loop measured faster).
constexpr int kClaim = 128;   // positions per claim (one hot counter: ~1 ns per atomic, 4 M positions at config 3)
__device__ __forceinline__ int64_t band_claim(BandInfo* info, int which)
{
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(&info->work[which], (unsigned long long)kClaim);
    return (int64_t)__shfl_sync(0xffffffffu, base, 0);
}

__global__ void __launch_bounds__(128, RG_COUNT_MINB) k_band_walk_count(const __grid_constant__ Pass4 Q, const BandParams B,
                                                            int32_t* __restrict__ hist, int32_t* __restrict__ flags)
{
    const BandInfo& I = *B.info;
    const int64_t total = I.tstart[4];
    int64_t claim = 0;
    for (int sub = kClaim;; sub += 32) {
        if (sub >= kClaim) {
            claim = band_claim(const_cast<BandInfo*>(B.info), 0);
            sub = 0;
        }
        if (claim >= total) break;
        const int64_t gtid = claim + sub + (threadIdx.x & 31);
        int p, L, k;
        if (gtid >= total || !band_segment(Q, I, gtid, p, L, k)) continue;
        const PassParams& P = Q.p[p];
        const int64_t v = vertex_of(P, L, k), v2 = v + vertex_step(P);
        const bool relevant = band_relevant(P, B, L, k, v);
        if (!relevant && !(k + 1 < P.nseg && band_relevant(P, B, L, k + 1, v2))) continue;
        const int start = P.guess[v];   // (verified by the chain check: exactly, when no walked predecessor ends here)
        P.seg_start[v] = start;
        if (start <= kStateUnknown) {   // unknown or never located: the sequentially verified build decides
            P.seg_end[v] = kStateInvalid;
            atomicOr(&flags[kFlagBandMismatch], 1);
            continue;
        }
        const double x1 = P.sweep.x[v], y1 = P.sweep.y[v];
        CountSink sink{ hist, L, k, 1, 0, relevant ? v : (int64_t)-1, 0, x1, y1 };
        bool overflow = false;
        P.seg_end[v] = walk_segment(P, x1, y1, P.sweep.x[v2], P.sweep.y[v2], start, sink, overflow);
        P.seg_hit[v] = sink.total > 0;
        P.pc_n[v] = (uint8_t)((relevant && sink.n_cached <= kPieceCache && !overflow) ? sink.n_cached : kPieceNone);
        if (overflow) atomicOr(&flags[kFlagOverflow], 1);
    }
}

// EXACT check of the located state of a sweep vertex: the start state of the first segment of a run of walked
// segments, which no walked predecessor verifies.  "In cell c" holds when the vertex lies strictly inside the
// intersection of the four inner half-planes of c's edges, by more than the rounding error of the cross products
// (each is one fused multiply-add of two differences: |error| <= 4 eps (|ex Py| + |ey Px|); the margin asked for is
// 45 eps) -- in a mesh without overlapping cells the sequential walk of the reference can then only be in c.
// "Outside" holds when the winding number of the static grid's boundary polygon around the vertex is exactly 0: the
// reference's own line-start test (c2d.py:308-317).  Anything else (a vertex within rounding error of a cell edge,
// an unknown state) is not verified: the caller raises the mismatch flag and the band is rebuilt sequentially.
// (returns 1: verified, 0: not verified, 2: located "outside" -- the caller evaluates the winding number with its
// whole warp, band_winding_warp)
__device__ __noinline__ int band_start_exact(const PassParams& P, int64_t v)
{
    const int s = P.guess[v];
    const double px = P.sweep.x[v], py = P.sweep.y[v];
    if (s == kStateOutside) return 2;
    if (s < 0) return 0;
    const GridView& g = P.stat;
    const int i0 = s / P.ncy_st, j0 = s - i0 * P.ncy_st;
    const double* gx = g.x + ((int64_t)i0 * g.ny + j0);
    const double* gy = g.y + ((int64_t)i0 * g.ny + j0);
    const double x00 = gx[0], x01 = gx[1], x10 = gx[g.ny], x11 = gx[g.ny + 1];
    const double y00 = gy[0], y01 = gy[1], y10 = gy[g.ny], y11 = gy[g.ny + 1];
    // edges (i0,j0)->(i0+1,j0)->(i0+1,j0+1)->(i0,j0+1)->(i0,j0): cross product of the edge with (vertex -> point)
    int pos = 0, neg = 0;
    auto edge = [&](double ax, double ay, double bx, double by) {
        const double ex = bx - ax, ey = by - ay, qx = px - ax, qy = py - ay;
        const double t = ey * qx;
        const double c = dfma(ex, qy, -t);
        const double margin = 1e-14 * (fabs(ex * qy) + fabs(t));
        pos += c > margin;
        neg += c < -margin;
    };
    edge(x00, y00, x10, y10);
    edge(x10, y10, x11, y11);
    edge(x11, y11, x01, y01);
    edge(x01, y01, x00, y00);
    return (pos == 4 || neg == 4) ? 1 : 0;
}

// boundary winding number around (px, py), the 32 lanes of the warp sharing the edge groups (all lanes call with the
// same point; contributions are multiples of 1/2: the partial sums add up exactly in any order)
__device__ __forceinline__ double band_winding_warp(const Boundary& b, double px, double py)
{
    double w = 0.0;
    for (int g1 = threadIdx.x & 31; g1 < b.n_g1; g1 += 32) {
        const BBox B1 = b.bb1[g1];
        if (B1.ylo <= py && py <= B1.yhi) w += boundary_winding_group(b, g1, px, py);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    return w;
}

// Chain check of one walked segment (folded into the emit walk, which visits the same rectangle): the end state of
// a walked predecessor must equal the start state this segment used; the first segment of a run (its predecessor is
// not walked) must start from an EXACTLY verified state; and the end state of a relevant segment whose successor is
// not walked must equal the located state of the next vertex.  Every state a rank uses is therefore verified by the
// rank itself: a band needs no agreement with the other ranks.
// (`need_winding`: the run starts from a located "outside" -- the caller's warp evaluates the winding number)
__device__ __forceinline__ bool band_chain_bad(const PassParams& P, const BandParams& B, int L, int k, int64_t v,
                                               bool relevant, bool next_relevant, bool& need_winding)
{
    const int64_t step = vertex_step(P);
    bool bad = false;
    if (k >= 1 && (relevant || band_relevant(P, B, L, k - 1, v - step))) {
        bad = P.seg_end[v - step] != P.seg_start[v];
    } else {   // first segment of a run (or of the line): no walked predecessor
        const int e = band_start_exact(P, v);
        bad = e == 0;
        need_winding = e == 2;
    }
    if (relevant && k + 1 < P.nseg && !next_relevant &&
        !(k + 2 < P.nseg && band_relevant(P, B, L, k + 2, v + 2 * step)))
        bad = bad || (P.seg_end[v] != P.guess[v + step]);
    return bad;
}

__global__ void __launch_bounds__(128, RG_EMIT_MINB)
k_band_walk_emit(const __grid_constant__ Pass4 Q, const BandParams B, const int64_t* __restrict__ boff,
                 int32_t* __restrict__ cursor, Frag* __restrict__ frag, int64_t frag_capacity,
                 const double* __restrict__ area_in, const double* __restrict__ w_in, int32_t* __restrict__ flags)
{
    if (flags[kFlagCapacity]) return;
    const BandInfo& I = *B.info;
    const int64_t total = I.tstart[4];
    for (int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gtid < total; gtid += (int64_t)gridDim.x * blockDim.x) {
        int p, L = 0, k = 0;
        const bool seg = band_segment(Q, I, gtid, p, L, k);
        const PassParams& P = Q.p[p];
        int64_t v = 0, v2 = 0;
        bool relevant = false, walked = false, bad = false, need_winding = false;
        if (seg) {
            v = vertex_of(P, L, k);
            v2 = v + vertex_step(P);
            relevant = band_relevant(P, B, L, k, v);
            const bool next_relevant = k + 1 < P.nseg && band_relevant(P, B, L, k + 1, v2);
            walked = relevant || next_relevant;
            if (walked) bad = band_chain_bad(P, B, L, k, v, relevant, next_relevant, need_winding);
        }
        if (need_winding) {
            // run start located "outside": exact when the boundary winding number is 0.  The starts of neighbouring
            // lines sit in neighbouring lanes, so they are queued and shared out warp by warp (k_band_verify_outside)
            const unsigned at = atomicAdd(&const_cast<BandInfo*>(B.info)->vq_n, 1u);
            if ((int64_t)at < B.vq_cap) B.vq[at] = ((unsigned long long)p << 40) | (unsigned long long)v;
            else bad = true;
        }
        if (bad) atomicOr(&flags[kFlagBandMismatch], 1);
        if (!walked || !relevant || !P.seg_hit[v]) continue;
        EmitSink sink{ boff, cursor, frag, area_in, w_in, flags, L, k, frag_capacity };
        const int nc = P.pc_n[v];
        if (nc != kPieceNone) {
            double x1 = P.sweep.x[v], y1 = P.sweep.y[v];
            int piece = 0;
            int cells[kPieceCache];
            double xs[kPieceCache], ys[kPieceCache];
#pragma unroll
            for (int e = 0; e < kPieceCache; e++) {
                const int64_t at = (int64_t)e * P.pc_stride + v;
                cells[e] = -1; xs[e] = 0.0; ys[e] = 0.0;
                if (e < nc) { cells[e] = P.pc_cell[at]; xs[e] = P.pc_x[at]; ys[e] = P.pc_y[at]; }
            }
#pragma unroll
            for (int e = 0; e < kPieceCache; e++) {
                if (e >= nc) break;
                const int cell = cells[e];
                const double x = xs[e], y = ys[e];
                if (cell >= 0) {
                    const int ci = cell / P.ncy_st;
                    sink.piece(P, x1, y1, x, y, ci, cell - ci * P.ncy_st, piece);
                    piece++;
                }
                x1 = x;
                y1 = y;
            }
            continue;
        }
        bool overflow = false;
        walk_segment(P, P.sweep.x[v], P.sweep.y[v], P.sweep.x[v2], P.sweep.y[v2], P.seg_start[v], sink, overflow);
        if (overflow) atomicOr(&flags[kFlagOverflow], 1);
    }
}

// ONE-WALK band build: the count walk and the emit walk in one -- a segment is walked once and its fragments go
// straight into fixed-capacity buckets (`bucket_cap` slots per input cell of the band, the capacity learned from an
// earlier build of the same shape; a bucket that overflows raises kFlagCapacity and the caller falls back to the
// two-walk build).  No histogram, no piece cache, no replay; the per-cell counts are the cursors, scanned AFTER the walk
// for the dense offsets the sort writes to.  The chain check, which needs the end states of all segments, runs as its
// own kernel beside the sort (k_band_chain_check).
__global__ void __launch_bounds__(128, RG_COUNT_MINB)
k_band_walk_once(const __grid_constant__ Pass4 Q, const BandParams B, int32_t* __restrict__ cursor, Frag* __restrict__ frag_strided,
                 int bucket_cap, int64_t cell_base, const double* __restrict__ area_in, const double* __restrict__ w_in,
                 int32_t* __restrict__ flags)
{
    const BandInfo& I = *B.info;
    const int64_t total = I.tstart[4];
    int64_t claim = 0;
    for (int sub = kClaim;; sub += 32) {
        if (sub >= kClaim) {
            claim = band_claim(const_cast<BandInfo*>(B.info), 0);
            sub = 0;
        }
        if (claim >= total) break;
        const int64_t gtid = claim + sub + (threadIdx.x & 31);
        int p, L, k;
        if (gtid >= total || !band_segment(Q, I, gtid, p, L, k)) continue;
        const PassParams& P = Q.p[p];
        const int64_t v = vertex_of(P, L, k), v2 = v + vertex_step(P);
        const bool relevant = band_relevant(P, B, L, k, v);
        if (!relevant && !(k + 1 < P.nseg && band_relevant(P, B, L, k + 1, v2))) continue;
        const int start = P.guess[v];   // (verified by the chain check: exactly, when no walked predecessor ends here)
        P.seg_start[v] = start;
        if (start <= kStateUnknown) {   // unknown or never located: the sequentially verified build decides
            P.seg_end[v] = kStateInvalid;
            atomicOr(&flags[kFlagBandMismatch], 1);
            continue;
        }
        EmitSink sink{ nullptr, cursor, frag_strided, area_in, w_in, flags, L, k, INT64_MAX, bucket_cap, cell_base };
        bool overflow = false;
        P.seg_end[v] = walk_segment(P, P.sweep.x[v], P.sweep.y[v], P.sweep.x[v2], P.sweep.y[v2], start, sink, overflow);
        if (overflow) atomicOr(&flags[kFlagOverflow], 1);
    }
}

// chain check of every walked segment (one-walk builds; the two-walk build folds it into its emit walk)
__global__ void __launch_bounds__(256) k_band_chain_check(const __grid_constant__ Pass4 Q, const BandParams B,
                                                          int32_t* __restrict__ flags)
{
    const BandInfo& I = *B.info;
    const int64_t total = I.tstart[4];
    for (int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gtid < total; gtid += (int64_t)gridDim.x * blockDim.x) {
        int p, L, k;
        if (!band_segment(Q, I, gtid, p, L, k)) continue;
        const PassParams& P = Q.p[p];
        const int64_t v = vertex_of(P, L, k), v2 = v + vertex_step(P);
        const bool relevant = band_relevant(P, B, L, k, v);
        const bool next_relevant = k + 1 < P.nseg && band_relevant(P, B, L, k + 1, v2);
        if (!relevant && !next_relevant) continue;
        bool need_winding = false;
        bool bad = band_chain_bad(P, B, L, k, v, relevant, next_relevant, need_winding);
        if (need_winding) {
            const unsigned at = atomicAdd(&const_cast<BandInfo*>(B.info)->vq_n, 1u);
            if ((int64_t)at < B.vq_cap) B.vq[at] = ((unsigned long long)p << 40) | (unsigned long long)v;
            else bad = true;
        }
        if (bad) atomicOr(&flags[kFlagBandMismatch], 1);
    }
}

// exact check of the queued run starts that were located "outside": one warp per vertex, the lanes share the boundary
// edge groups; a non-zero winding number (the vertex is inside the boundary polygon, or on it) is a mismatch
__global__ void __launch_bounds__(128) k_band_verify_outside(const __grid_constant__ Pass4 Q, const BandParams B,
                                                             int32_t* __restrict__ flags)
{
    const int64_t n = min((int64_t)B.info->vq_n, B.vq_cap);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp; t < n; t += nwarps) {
        const unsigned long long e = B.vq[t];
        const PassParams& P = Q.p[(int)(e >> 40)];
        const int64_t v = (int64_t)(e & ((1ull << 40) - 1));
        const double w = band_winding_warp(P.bnd, P.sweep.x[v], P.sweep.y[v]);
        if ((threadIdx.x & 31) == 0 && w != 0.0) atomicOr(&flags[kFlagBandMismatch], 1);
    }
}

// counts[0] = fragments of the band, counts[1] = triplets; capacity flags (stage 0: after the count scan, 1: after the
// unique-pair scan, 2: final report of all flags)
__global__ void k_band_counts(int stage, const int64_t* __restrict__ total, int64_t capacity, int64_t* __restrict__ counts,
                              int32_t* __restrict__ flags)
{
    if (threadIdx.x != 0) return;
    if (stage < 2) {
        counts[stage] = *total;
        if (*total > capacity || *total >= INT32_MAX) flags[kFlagCapacity] = 1;
    } else {
        for (int q = 0; q < 6; q++) counts[2 + q] = flags[q];
        if (stage == 3) counts[2 + kFlagBandMismatch] = 1;
    }
}

}  // namespace rg

// side stream + events of the band build's two-stream preparation: one set per device, created on first use
namespace rg {
struct BandSide {
    cudaStream_t stream;
    cudaStream_t capture;   // one-walk builds are captured on this stream and replayed as CUDA graphs
    cudaEvent_t fork, join, emitted, verified;
};
static BandSide* band_side(int device)
{
    static std::mutex mu;
    static BandSide* table[64] = { nullptr };
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!table[device]) {
        BandSide* b = new BandSide();
        if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&b->capture, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&b->fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b->join, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b->emitted, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b->verified, cudaEventDisableTiming) != cudaSuccess) {
            delete b;
            return nullptr;
        }
        table[device] = b;
    }
    return table[device];
}
}  // namespace rg

namespace rg {
struct BandGraphKey {
    int device;
    int64_t dims[4];
    const void* ptrs[12];
    int64_t nums[7];
};
struct BandGraphEntry {
    BandGraphKey key;
    cudaGraphExec_t exec;
    uint64_t stamp;
};
static std::mutex& band_graph_mutex()
{
    static std::mutex m;
    return m;
}
static std::vector<BandGraphEntry>& band_graph_cache()
{
    static std::vector<BandGraphEntry> c;
    return c;
}
}  // namespace rg

static int band_impl(int device, void* stream,
                     int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                     const double* xin, const double* yin, const double* xout, const double* yout,
                     const double* w_in, int64_t row_lo, int64_t row_hi,
                     void* workspace, size_t workspace_bytes,
                     void* frags, int64_t frag_capacity,
                     int64_t* ii, int64_t* io, double* v, int64_t nnz_capacity,
                     int64_t* counts_dev /* [8] */, void* frags_strided, int64_t bucket_cap)
{
    int rc = check_sizes(nxi, nyi, nxo, nyo);
    if (rc) return rc;
    if (!xin || !yin || !xout || !yout || !workspace || !counts_dev || !frags || !ii || !io || !v)
        return fail(RG_E_ARG, "rg_build2d_band: null pointer");
    if (row_lo < 0 || row_hi > nxi - 1 || row_lo >= row_hi) return fail(RG_E_ARG, "rg_build2d_band: bad row band");
    if (frag_capacity < 1 || nnz_capacity < 1) return fail(RG_E_ARG, "rg_build2d_band: bad capacity");
    Layout l = make_layout(workspace, nxi, nyi, nxo, nyo);
    if (workspace_bytes < l.bytes) return fail(RG_E_WORKSPACE, "rg_build2d_band: workspace too small");
    RG_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const GridView gin{ xin, yin, (int)nxi, (int)nyi };
    const GridView gout{ xout, yout, (int)nxo, (int)nyo };
    const int T = 256;
    const int64_t ncy = nyi - 1;
    const int64_t cell_lo = row_lo * ncy, cell_hi = row_hi * ncy, nb = cell_hi - cell_lo;

    const bool full = row_lo == 0 && row_hi == nxi - 1;
    BandSide* side = band_side(device);
    if (!side) return fail((int)cudaErrorUnknown, "rg_build2d_band: cannot create the side stream");
    cudaStream_t ss = side->stream;
    const Pass4 Q = make_pass4(l, xin, yin, xout, yout, cell_lo, cell_hi, 0, 1);
    BandParams B;
    B.row_lo = (int)row_lo; B.row_hi = (int)row_hi;
    B.rel[0] = l.rel[0]; B.rel[1] = l.rel[1];
    B.info = l.info;
    B.vq = l.vq; B.vq_cap = l.vq_cap;
    BoundarySet S;   // both grids
    memset(&S, 0, sizeof(S));
    S.n = 2;
    S.g[0] = gin; S.g[1] = gout; S.b[0] = l.bnd[0]; S.b[1] = l.bnd[1]; S.bbox[0] = l.bbox; S.bbox[1] = l.bbox + 4;
    const int g1max = l.bnd[0].n_g1 > l.bnd[1].n_g1 ? l.bnd[0].n_g1 : l.bnd[1].n_g1;
    const int g2max = l.bnd[0].n_g2 > l.bnd[1].n_g2 ? l.bnd[0].n_g2 : l.bnd[1].n_g2;

    // The preparation runs on TWO streams.  Everything that leads to the relevant output segments is a chain of short
    // dependent launches (bbox of the input grid -> raster of the band -> relevance of the output segments -> extents ->
    // located states of the output vertices); everything else the walks need -- cleared histograms, cell areas, the
    // located states of the input vertices around the band -- does not depend on it and runs beside it on a side stream.
    // (No exact line starts: the first segment of a line starts from the located state of its vertex like the first
    // segment of any run, and the chain check verifies that state exactly.)
    k_band_begin<<<1, 32, 0, st>>>(Q, (int)row_lo, (int)row_hi, full ? 1 : 0, (int)nxo, (int)nyo, l.flags, counts_dev, l.bbox, l.info);
    RG_LAUNCH_CHECK("k_band_begin");
    k_boundary_edges_bb1<<<dim3((unsigned)ceil_div((int64_t)g1max * 32, T), 2), T, 0, st>>>(S);
    k_boundary_bb2<<<dim3((unsigned)ceil_div((int64_t)g2max * 32, T), 2), T, 0, st>>>(S);
    RG_LAUNCH_CHECK("k_boundary");
    RG_CUDA(cudaEventRecord(side->fork, st));
    // The host enqueues ~25 operations here, ~3 us each, which is what the start of the build is bound by: the head of
    // the main chain goes out first (the GPU then has ~100 us of work), the side stream's work next (it starts at once,
    // beside it), the tail of the main chain last.
    if (!full) {
        RG_CUDA(cudaMemsetAsync(l.raster8, 0, (size_t)kRasterN * kRasterN, st));
        k_bbox<<<dim3(kNumSM * 4, 1), T, 0, st>>>(S);
        RG_LAUNCH_CHECK("k_bbox");
        k_band_raster<<<(unsigned)ceil_div(nb, T), T, 0, st>>>(gin, (int)row_lo, (int)row_hi, l.bbox, l.info, l.raster8);
        RG_LAUNCH_CHECK("k_band_raster");
        k_band_raster_pack<<<kRasterN * kRasterN / 8 / 256, 256, 0, st>>>(l.raster8, l.raster);
        k_band_relevance<<<kNumSM * 8, kRelThreads, 0, st>>>(gout, l.bbox, l.raster, l.rel[0], l.rel[1], l.info);
        RG_LAUNCH_CHECK("k_band_relevance");
    } else {
        k_bbox<<<dim3(kNumSM * 4, 2), T, 0, st>>>(S);
        RG_LAUNCH_CHECK("k_bbox");
    }
    // ---- side stream
    RG_CUDA(cudaStreamWaitEvent(ss, side->fork, 0));
    if (full) {
        // the band is the whole grid (per-slice builds without host synchronisation): everything is relevant
        RG_CUDA(cudaMemsetAsync(l.rel[0], 1, (size_t)l.Vo, ss));
        RG_CUDA(cudaMemsetAsync(l.rel[1], 1, (size_t)l.Vo, ss));
    }
    {
        // input vertices of rows row_lo - 1 .. row_hi + 1: every start / end vertex of a walked segment of passes 2, 3
        const int64_t r0 = row_lo > 0 ? row_lo - 1 : 0, r1 = (row_hi + 2 < nxi ? row_hi + 2 : nxi);
        const int64_t v_lo = r0 * nyi, v_hi = r1 * nyi;
        k_band_guess_in<<<(unsigned)ceil_div(v_hi - v_lo, T), T, 0, ss>>>(gin, gout, v_lo, v_hi, l.guess[1], l.flags);
        RG_LAUNCH_CHECK("k_band_guess_in");
    }
    RG_CUDA(cudaMemsetAsync(l.scan_status[0], 0, sizeof(unsigned long long) * 2 * scan_status_elems(l.Ci), ss));
    if (bucket_cap <= 0) RG_CUDA(cudaMemsetAsync(l.hist + cell_lo, 0, sizeof(int32_t) * (size_t)(nb + 1), ss));
    RG_CUDA(cudaMemsetAsync(l.cursor + cell_lo, 0, sizeof(int32_t) * (size_t)(nb + 1), ss));
    // areas of the band's cells (k_cell_area works on the rows [row_lo, row_hi) of the grid: a view of those rows would
    // move the peeled first-edge pattern of grid_volume, so the full-grid kernel runs on the band's cell range)
    k_cell_area_range<<<(unsigned)ceil_div(nb, T), T, 0, ss>>>(gin, cell_lo, cell_hi, l.area_in);
    RG_LAUNCH_CHECK("k_cell_area_range");
    RG_CUDA(cudaEventRecord(side->join, ss));
    // ---- main stream, tail of the chain
    if (!full) {
        k_band_finalize<<<1, 32, 0, st>>>(Q, l.info);
        k_band_guess_out<<<kNumSM * 8, T, 0, st>>>(gout, gin, l.rel[0], l.rel[1], l.info, l.guess[0], l.flags);
        RG_LAUNCH_CHECK("k_band_guess_out");
    }
    // ---- join
    RG_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    if (full) {
        k_band_finalize<<<1, 32, 0, st>>>(Q, l.info);
        k_band_guess_out<<<kNumSM * 8, T, 0, st>>>(gout, gin, l.rel[0], l.rel[1], l.info, l.guess[0], l.flags);
        RG_LAUNCH_CHECK("k_band_guess_out");
    }
    // grid-stride: the amount of work of a partial band is only known on the device.  The whole-grid band (per-slice
    // builds) has one position per segment of the four passes: one thread each for the emit, whose scattered stores
    // and cursor atomics want neighbouring segments close in TIME (1.15 ms against 1.97 ms for the persistent loop).
    const unsigned walk_grid = kNumSM * 16;
    unsigned emit_grid = walk_grid;
    if (row_lo == 0 && row_hi == nxi - 1) {
        const int64_t nl[4] = { nyo, nxo, nyi, nxi }, nk[4] = { nxo - 1, nyo - 1, nxi - 1, nyi - 1 };
        int64_t total = 0;
        for (int p = 0; p < 4; p++) total += ceil_div(nl[p] * nk[p], 256) * 256;
        emit_grid = (unsigned)ceil_div(total, 128);
    }
    rc = sort_smem_opt_in(device);
    if (rc) return rc;
    if (bucket_cap > 0) {
        // ONE walk: fragments go straight into fixed-capacity buckets; the cursors are the counts
        k_band_walk_once<<<walk_grid, 128, 0, st>>>(Q, B, l.cursor, (Frag*)frags_strided, (int)bucket_cap, cell_lo, l.area_in, w_in,
                                                    l.flags);
        RG_LAUNCH_CHECK("k_band_walk_once");
        // chain check + exact check of the run starts located "outside": beside the scan / sort / merge
        RG_CUDA(cudaEventRecord(side->emitted, st));
        RG_CUDA(cudaStreamWaitEvent(ss, side->emitted, 0));
        k_band_chain_check<<<kNumSM * 8, 256, 0, ss>>>(Q, B, l.flags);
        k_band_verify_outside<<<kNumSM * 4, 128, 0, ss>>>(Q, B, l.flags);
        RG_LAUNCH_CHECK("k_band_verify_outside");
        RG_CUDA(cudaEventRecord(side->verified, ss));
        rc = exclusive_scan_i32_i64_single(st, l.cursor + cell_lo, l.boff + cell_lo, nb, l.scan_status[0], frag_capacity, counts_dev,
                                           l.flags + kFlagCapacity);
        if (rc) return rc;
        k_bucket_sort<<<(unsigned)ceil_div(nb, kSortCells), kSortThreads, sizeof(SortSmem), st>>>(
            l.boff + cell_lo, nb, (Frag*)frags, l.nuniq + cell_lo, l.flags + kFlagCapacity, (const Frag*)frags_strided, (int)bucket_cap,
            l.flags + kFlagRepairs);
        RG_LAUNCH_CHECK("k_bucket_sort");
    } else {
        k_band_walk_count<<<walk_grid, 128, 0, st>>>(Q, B, l.hist, l.flags);
        RG_LAUNCH_CHECK("k_band_walk_count");
        // (single-launch scans; the total goes to counts[0] / counts[1] with the capacity check)
        rc = exclusive_scan_i32_i64_single(st, l.hist + cell_lo, l.boff + cell_lo, nb, l.scan_status[0], frag_capacity, counts_dev,
                                           l.flags + kFlagCapacity);
        if (rc) return rc;
        // boff of the band starts at 0: cells index it globally (boff[cell]), fragments locally
        k_band_walk_emit<<<emit_grid, 128, 0, st>>>(Q, B, l.boff, l.cursor, (Frag*)frags, frag_capacity, l.area_in, w_in, l.flags);
        RG_LAUNCH_CHECK("k_band_walk_emit");
        // the exact check of the run starts located "outside" runs beside the sort / merge
        RG_CUDA(cudaEventRecord(side->emitted, st));
        RG_CUDA(cudaStreamWaitEvent(ss, side->emitted, 0));
        k_band_verify_outside<<<kNumSM * 4, 128, 0, ss>>>(Q, B, l.flags);
        RG_LAUNCH_CHECK("k_band_verify_outside");
        RG_CUDA(cudaEventRecord(side->verified, ss));
        // (the longest bucket goes to the otherwise unused "repairs" counter: the capacity a one-walk rebuild needs)
        k_bucket_sort<<<(unsigned)ceil_div(nb, kSortCells), kSortThreads, sizeof(SortSmem), st>>>(
            l.boff + cell_lo, nb, (Frag*)frags, l.nuniq + cell_lo, l.flags + kFlagCapacity, nullptr, 0, l.flags + kFlagRepairs);
        RG_LAUNCH_CHECK("k_bucket_sort");
    }
    rc = exclusive_scan_i32_i64_single(st, l.nuniq + cell_lo, l.colptr + cell_lo, nb, l.scan_status[1], nnz_capacity, counts_dev + 1,
                                       l.flags + kFlagCapacity);
    if (rc) return rc;
    k_bucket_emit<<<(unsigned)ceil_div(nb, 256), 256, 0, st>>>(l.boff + cell_lo, 1, cell_lo, l.colptr + cell_lo, nb,
                                                              (const Frag*)frags, ii, io, v, l.flags + kFlagCapacity);
    RG_LAUNCH_CHECK("k_bucket_emit");
    RG_CUDA(cudaStreamWaitEvent(st, side->verified, 0));
    // (stage 3: test hook, reports a chain mismatch whatever the walks found)
    k_band_counts<<<1, 32, 0, st>>>(getenv("RG_BAND_FORCE_MISMATCH") ? 3 : 2, nullptr, 0, counts_dev, l.flags);
    RG_LAUNCH_CHECK("k_band_counts");
    return RG_OK;
}

extern "C" int rg_build2d_band(int device, void* stream,
                               int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                               const double* xin, const double* yin, const double* xout, const double* yout,
                               const double* w_in, int64_t row_lo, int64_t row_hi,
                               void* workspace, size_t workspace_bytes,
                               void* frags, int64_t frag_capacity,
                               int64_t* ii, int64_t* io, double* v, int64_t nnz_capacity,
                               int64_t* counts_dev /* [8] */)
{
    return band_impl(device, stream, nxi, nyi, nxo, nyo, xin, yin, xout, yout, w_in, row_lo, row_hi, workspace, workspace_bytes,
                     frags, frag_capacity, ii, io, v, nnz_capacity, counts_dev, nullptr, 0);
}

extern "C" int rg_build2d_band_onewalk(int device, void* stream,
                                       int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                                       const double* xin, const double* yin, const double* xout, const double* yout,
                                       const double* w_in, int64_t row_lo, int64_t row_hi,
                                       void* workspace, size_t workspace_bytes,
                                       void* frags, int64_t frag_capacity,
                                       int64_t* ii, int64_t* io, double* v, int64_t nnz_capacity,
                                       int64_t* counts_dev /* [8] */, void* frags_strided, int64_t bucket_capacity)
{
    if (!frags_strided || bucket_capacity < 1 || bucket_capacity > 65535)
        return fail(RG_E_ARG, "rg_build2d_band_onewalk: bad bucket buffer");
    // A one-walk build is a REPEAT build of its shape (the bucket capacity comes from an earlier one), so the whole
    // enqueue -- ~30 operations on two streams, host-bound at its start -- is captured once per argument set and
    // replayed as a CUDA graph (the caller's stream may be the legacy default stream, which cannot be captured: the
    // capture runs on a library-owned stream, the graph is launched into the caller's).
    BandGraphKey key;
    memset(&key, 0, sizeof(key));
    key.device = device;
    key.dims[0] = nxi; key.dims[1] = nyi; key.dims[2] = nxo; key.dims[3] = nyo;
    key.ptrs[0] = xin; key.ptrs[1] = yin; key.ptrs[2] = xout; key.ptrs[3] = yout; key.ptrs[4] = w_in;
    key.ptrs[5] = workspace; key.ptrs[6] = frags; key.ptrs[7] = ii; key.ptrs[8] = io; key.ptrs[9] = v;
    key.ptrs[10] = counts_dev; key.ptrs[11] = frags_strided;
    key.nums[0] = row_lo; key.nums[1] = row_hi; key.nums[2] = (int64_t)workspace_bytes; key.nums[3] = frag_capacity;
    key.nums[4] = nnz_capacity; key.nums[5] = bucket_capacity; key.nums[6] = getenv("RG_BAND_FORCE_MISMATCH") ? 1 : 0;
    BandSide* side = getenv("RG_BAND_NO_GRAPH") ? nullptr : band_side(device);
    if (side && cudaSetDevice(device) == cudaSuccess && sort_smem_opt_in(device) == RG_OK) {
        std::lock_guard<std::mutex> lock(band_graph_mutex());
        std::vector<BandGraphEntry>& cache = band_graph_cache();
        static uint64_t clock = 0;
        static int misses_in_a_row = 0;   // a caller whose argument sets never recur (e.g. it keeps every result alive)
        for (BandGraphEntry& e : cache)   // must not pay for a capture (~1 ms) per build: after 4 misses, plain launches
            if (memcmp(&e.key, &key, sizeof(key)) == 0) {
                e.stamp = ++clock;
                misses_in_a_row = 0;
                RG_CUDA(cudaGraphLaunch(e.exec, (cudaStream_t)stream));
                return RG_OK;
            }
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        if (misses_in_a_row >= 4) {
            // (fall through to the plain launches below)
        } else if (++misses_in_a_row, cudaStreamBeginCapture(side->capture, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int rc = band_impl(device, side->capture, nxi, nyi, nxo, nyo, xin, yin, xout, yout, w_in, row_lo, row_hi, workspace,
                                     workspace_bytes, frags, frag_capacity, ii, io, v, nnz_capacity, counts_dev, frags_strided,
                                     bucket_capacity);
            const cudaError_t ce = cudaStreamEndCapture(side->capture, &graph);
            if (rc == RG_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                cudaGraphDestroy(graph);
                if (cache.size() >= 16) {   // drop the least recently used argument set
                    size_t old = 0;
                    for (size_t q = 1; q < cache.size(); q++)
                        if (cache[q].stamp < cache[old].stamp) old = q;
                    cudaGraphExecDestroy(cache[old].exec);
                    cache.erase(cache.begin() + (long)old);
                }
                cache.push_back(BandGraphEntry{ key, exec, ++clock });
                RG_CUDA(cudaGraphLaunch(exec, (cudaStream_t)stream));
                return RG_OK;
            }
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();   // the capture did not work out (an argument error, an unsupported call): plain launches
            if (rc != RG_OK && rc < 0) return rc;   // argument errors are the caller's
        } else {
            cudaGetLastError();
        }
    }
    return band_impl(device, stream, nxi, nyi, nxo, nyo, xin, yin, xout, yout, w_in, row_lo, row_hi, workspace, workspace_bytes,
                     frags, frag_capacity, ii, io, v, nnz_capacity, counts_dev, frags_strided, bucket_capacity);
}

// Per-slice builds (every orthogonal slice carries its own grid pair: the Python loop of
// regridding/_weights/_weights_conservative.py:110-139) enqueued back to back with NO host synchronisation:
// slice s is the whole-grid case of rg_build2d_band.  Workspace and fragment buffer are reused slice after slice
// (stream order); every slice has its own output arrays and its own 8 counters.
extern "C" int rg_build2d_batched(int device, void* stream, int64_t n_slices,
                                  int64_t nxi, int64_t nyi, int64_t nxo, int64_t nyo,
                                  const double* const* xin, const double* const* yin,
                                  const double* const* xout, const double* const* yout,
                                  const double* const* w_in_or_null,
                                  void* workspace, size_t workspace_bytes,
                                  void* frags, int64_t frag_capacity,
                                  int64_t* const* ii, int64_t* const* io, double* const* v, int64_t nnz_capacity,
                                  int64_t* counts_dev /* [n_slices][8] */)
{
    if (n_slices < 0 || !xin || !yin || !xout || !yout || !ii || !io || !v || !counts_dev)
        return fail(RG_E_ARG, "rg_build2d_batched: bad argument");
    for (int64_t s = 0; s < n_slices; s++) {
        const int rc = rg_build2d_band(device, stream, nxi, nyi, nxo, nyo, xin[s], yin[s], xout[s], yout[s],
                                       w_in_or_null ? w_in_or_null[s] : nullptr, 0, nxi - 1, workspace, workspace_bytes,
                                       frags, frag_capacity, ii[s], io[s], v[s], nnz_capacity, counts_dev + 8 * s);
        if (rc) return rc;
    }
    return RG_OK;
}
