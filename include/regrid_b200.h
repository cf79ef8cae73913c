/*
 * regrid_b200.h -- C ABI of libregrid_b200.so (CUDA, sm_100a).
 *
 * Drop-in boundary for the first-order conservative regridding path of
 * sun-data/regridding.  The reference has no FFI of its own (it is Python + Numba);
 * its boundary is the four places where Python hands arrays to a compiled kernel
 * (SURVEY.md section 8b).  Each entry point below names the reference call site it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer; the library never allocates result memory.
 *     Variable-length results use  count -> caller allocates -> emit;
 *     scratch comes from a caller-provided workspace (size-query functions);
 *   - `stream` is a cudaStream_t passed as void*; work is stream-ordered; functions
 *     that return a count through a *_host pointer synchronise that stream;
 *   - `device` is the CUDA device ordinal the pointers live on;
 *   - return value: 0 = ok, negative = argument error (RG_E_*), positive = cudaError_t.
 *     rg_last_error_string() describes the last failure on the calling thread;
 *   - coordinates/values are fp64, public indices int64, internal CSR indices int32;
 *   - no global mutable state: safe from several host threads on distinct
 *     streams/workspaces.
 */
#ifndef REGRID_B200_H
#define REGRID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_OK 0
#define RG_E_ARG (-1)        /* null pointer / non-positive size */
#define RG_E_TOO_LARGE (-2)  /* a size exceeds the int32 internal index range */
#define RG_E_WORKSPACE (-3)  /* workspace too small */
#define RG_E_WALK (-4)       /* a sweep walk did not terminate (degenerate / folded grid) */

/* ------------------------------------------------------------------------------
 * ndarray_linear_interpolation
 * replaces  _ndarray_linear_interpolation_1d / _2d and _linear_interpolation / _bilinear_interpolation,
 *           regridding/_interp_ndarray.py:192-297 (cell index clamped to [0, n - 2]: linear extrapolation outside).
 * D slices; slice d interpolates a + d * a_stride (n values, or (nx, ny) row-major) at its m (or P) fractional
 * indices x + d * x_stride (and y likewise); a stride of 0 shares the array / the indices between the slices (the
 * reference broadcasts them).  out is (D, m) / (D, P) contiguous. */
int rg_interp_linear_1d(int device, void* stream, int64_t D, int64_t n, int64_t m,
                        int64_t a_stride, int64_t x_stride, const double* a, const double* x, double* out);
int rg_interp_bilinear_2d(int device, void* stream, int64_t D, int64_t nx, int64_t ny, int64_t P,
                          int64_t a_stride, int64_t xy_stride,
                          const double* a, const double* x, const double* y, double* out);

/* Measurement helper (bench.py): fp64 FMA-chain throughput of the device in TFLOP/s (2 flops per FMA), the
 * denominator of the build's fp64 roofline; synchronises the stream. */
int rg_measure_fp64_peak(int device, void* stream, double* tflops_host);

const char* rg_last_error_string(void);
int rg_version(void);

/* ------------------------------------------------------------------------------
 * 2D conservative weights build (Ramshaw 1985 edge sweep)
 * replaces: weights_conservative_2d(grid_input, grid_output, weights_input)
 *           regridding/_weights/_weights_conservative_2d/_weights_conservative_2d.py:80-126
 *           (called from regridding/_weights/_weights_conservative.py:129-139)
 *   plus    _coalesce  regridding/_weights/_weights_arrays.py:44-73
 *
 * Grids are row-major contiguous fp64 vertex arrays of shape (nx, ny).
 * Result: the reference's public layout -- unique (input, output) pairs sorted by
 * (input, output), flat C-order CELL indices, weights summed per pair in NumPy's
 * np.add.reduceat association over the reference's emission order.
 *
 * Protocol (all three calls share one workspace, which must stay untouched between them):
 *   rg_build2d_workspace_bytes -> caller allocates `workspace`
 *   rg_build2d_count           -> *n_fragments_host  (raw fragments before merging)
 *   caller allocates frags: n_fragments records of 16 bytes, 16-byte aligned
 *                    ({uint64 key = output cell << 32 | emission rank; double weight})
 *   rg_build2d_fill            -> *nnz_host
 *   caller allocates indices_input/indices_output (int64[nnz]), values (double[nnz])
 *   rg_build2d_emit
 * cell_lo/cell_hi restrict the build to input cells with flat index in
 * [cell_lo, cell_hi) (band partition for multi-GPU builds); pass 0 and
 * (nx_in-1)*(ny_in-1) for a full build.
 * ------------------------------------------------------------------------------ */
int rg_build2d_workspace_bytes(int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                               size_t* bytes_host);

int rg_build2d_count(int device, void* stream,
                     int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                     const double* x_in, const double* y_in,
                     const double* x_out, const double* y_out,
                     int64_t cell_lo, int64_t cell_hi,
                     void* workspace, size_t workspace_bytes,
                     int64_t* n_fragments_host);

int rg_build2d_fill(int device, void* stream,
                    int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                    const double* x_in, const double* y_in,
                    const double* x_out, const double* y_out,
                    const double* weights_input_or_null,
                    int64_t cell_lo, int64_t cell_hi,
                    void* workspace, size_t workspace_bytes,
                    void* frags, int64_t n_fragments,
                    int64_t* nnz_host);

int rg_build2d_emit(int device, void* stream,
                    int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                    int64_t cell_lo, int64_t cell_hi,
                    void* workspace, size_t workspace_bytes,
                    const void* frags, int64_t n_fragments,
                    int64_t* indices_input, int64_t* indices_output, double* values, int64_t nnz);

/* Diagnostics of the last build in `workspace`: stats_host[0] = walk overflow flag,
 * [1] = segments re-walked by the chain repair, [2] = sweep vertices whose state guess
 * was unknown, [3] = emission-rank overflow flag; [4..7] reserved. */
int rg_build2d_stats(int device, void* stream, int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                     void* workspace, int32_t* stats_host);

/* Band build for multi-GPU strong scaling of ONE grid pair (the reference's parallel axis is the sweep line,
 * _weights_conservative_2d.py:286; its result layout is sorted by input cell, _weights_arrays.py:54-59, so bands
 * of input rows concatenate).  Rank r passes its band of input rows [row_lo, row_hi) and gets that band's public
 * triplets; it walks only the sweep segments whose bounding box meets a cell of the band (exact, see
 * rg_build2d.cu).  Stream-ordered, NO host synchronisation: `frags` (16-byte records) and ii / io / v are sized
 * by the caller's estimate; counts_dev[0] = fragments, [1] = triplets, [2..7] = status flags
 * (2 walk overflow, 3 longest bucket (fragments of one input cell), 4 unknown guesses, 5 rank overflow, 6 CHAIN MISMATCH -> a walk state of THIS band
 * could not be verified (every state a band uses is verified by the band itself: no agreement with other ranks is
 * needed): rebuild this band with rg_build2d_count/_fill/_emit,
 * 7 CAPACITY -> buffers too small: reallocate from counts_dev[0..1] and call again).
 * The preparation runs on the caller's stream and on a library-owned side stream (per device) that is forked from
 * and joined back into the caller's stream inside the call: the call is stream-ordered for the caller. */
int rg_build2d_band(int device, void* stream,
                    int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                    const double* x_in, const double* y_in, const double* x_out, const double* y_out,
                    const double* weights_input_or_null, int64_t row_lo, int64_t row_hi,
                    void* workspace, size_t workspace_bytes,
                    void* frags, int64_t frag_capacity,
                    int64_t* indices_input, int64_t* indices_output, double* values, int64_t nnz_capacity,
                    int64_t* counts_dev);
/* The same band in ONE walk (no count walk, no replay): every input cell of the band owns `bucket_capacity` slots of
 * `frags_strided` ((row_hi - row_lo) * (ny_in - 1) * bucket_capacity records of 16 bytes); the sort gathers the
 * buckets from there into `frags`.  rg_build2d_band reports the longest bucket of a build in counts_dev[3]: size
 * bucket_capacity from it (a little above) when the same shape is built again.  A bucket that overflows raises the
 * CAPACITY flag (counts_dev[7]): build with rg_build2d_band instead.  Results are identical to rg_build2d_band's. */
int rg_build2d_band_onewalk(int device, void* stream,
                            int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                            const double* x_in, const double* y_in, const double* x_out, const double* y_out,
                            const double* weights_input_or_null, int64_t row_lo, int64_t row_hi,
                            void* workspace, size_t workspace_bytes,
                            void* frags, int64_t frag_capacity,
                            int64_t* indices_input, int64_t* indices_output, double* values, int64_t nnz_capacity,
                            int64_t* counts_dev, void* frags_strided, int64_t bucket_capacity);


/* ------------------------------------------------------------------------------
 * Line-sharded 2D build (strong scaling of ONE large build over W ranks).
 * replaces the same reference code as rg_build2d_* (the reference parallelises the sweep lines with
 * numba.prange, _weights_conservative_2d.py:286; here the lines are dealt out across GPUs).
 *
 * Rank r of W walks the sweep lines of every W-th block of 32 lines of all four passes
 * (rg_build2d_part_count / _part_fill, same workspace as rg_build2d_*).  Its fragments come out
 * bucketed by input cell, so the fragments of every input-row band are one contiguous range:
 * frag_offsets_host[b] = first fragment of input cell cell_bounds_host[b] (n_bounds bounds, ascending;
 * also copied to frag_offsets_dev_or_null, stream-ordered, for peers that read it over NVLink).
 * The per-input-cell fragment counts (int32[(nx_in-1)*(ny_in-1)]) live in the workspace at byte
 * offset *counts_offset_host.
 * The owner of a band then needs, per source rank s, the band slice of that rank's counts and of its
 * fragments.  They either travel by all-to-all (NCCL) or stay where they are and are read in place over
 * NVLink peer mappings (regridding_b200/_parallel.py does both): rg_build2d_gather_counts copies the W
 * count slices into counts[s][c]; rg_build2d_merge takes one chunk pointer + size per source
 * (src_chunks_host[s] = first fragment of the band in source s, device or peer memory), gathers the
 * chunks cell by cell through shared memory, sorts every bucket by (output cell, emission rank) into
 * `frags` (sum of the sizes records) and counts the distinct pairs; rg_build2d_merge_emit writes the
 * band's public triplets (indices_input = cell_offset + band cell).  The emission rank does not depend
 * on which rank walked a segment: the concatenated bands equal the single-GPU result bit for bit.
 * n_src <= 16.
 * ------------------------------------------------------------------------------ */
int rg_build2d_part_count(int device, void* stream,
                          int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                          const double* x_in, const double* y_in,
                          const double* x_out, const double* y_out,
                          int part_rank, int part_world,
                          void* workspace, size_t workspace_bytes,
                          int64_t* n_fragments_host,
                          int n_bounds, const int64_t* cell_bounds_host, int64_t* frag_offsets_host,
                          size_t* counts_offset_host, int64_t* frag_offsets_dev_or_null);

int rg_build2d_part_fill(int device, void* stream,
                         int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                         const double* x_in, const double* y_in,
                         const double* x_out, const double* y_out,
                         const double* weights_input_or_null,
                         int part_rank, int part_world,
                         void* workspace, size_t workspace_bytes,
                         void* frags, int64_t n_fragments);

int rg_build2d_merge_workspace_bytes(int64_t n_cells, int n_src, size_t* bytes_host);

int rg_build2d_gather_counts(int device, void* stream, int64_t n_cells, int n_src,
                             const int32_t* const* src_counts_host /* n_src device pointers */,
                             int32_t* counts /* [n_src][n_cells] */);

int rg_build2d_merge(int device, void* stream, int64_t n_cells, int n_src,
                     const int32_t* counts /* [n_src][n_cells] */,
                     const void* const* src_chunks_host /* n_src device pointers */,
                     const int64_t* src_sizes_host /* n_src record counts */,
                     void* workspace, size_t workspace_bytes,
                     void* frags /* sum(src_sizes) records */,
                     int64_t* nnz_host);

int rg_build2d_merge_emit(int device, void* stream, int64_t n_cells, int n_src, int64_t cell_offset,
                          void* workspace, size_t workspace_bytes,
                          const void* frags,
                          int64_t* indices_input, int64_t* indices_output, double* values, int64_t nnz);

/* Signed cell areas; replaces grid_volume
 * regridding/_weights/_weights_conservative_2d/_grids.py:50-140. area: double[(nx-1)*(ny-1)]. */
int rg_grid_area(int device, void* stream, int64_t nx, int64_t ny,
                 const double* x, const double* y, double* area);

/* ------------------------------------------------------------------------------
 * 2D cell location
 * replaces: index_of_point_brute / index_of_point_secant
 *           regridding/_weights/_weights_conservative_2d/_grids.py:223-279, 356-463
 * (the reference's public find_indices is 1D only; this is the 2D extension).
 * For each of n_points query points: flat index i*(ny-1)+j of the lowest-index cell
 * whose quad contains the point under point_is_inside_polygon
 * (regridding/geometry.py:737-829), else `fill`.  Stream-ordered, no host synchronisation.  While the call
 * runs, cell_flat transiently holds up to three marker values below every cell index (INT64_MIN .. INT64_MIN + 3,
 * skipping `fill`); they are all resolved when the call's last kernel has run.  Exact for meshes without overlapping
 * cells (the assumption of the reference's own secant locator).
 * ------------------------------------------------------------------------------ */
int rg_find_indices_2d_workspace_bytes(int64_t nx, int64_t ny, int64_t n_points, size_t* bytes_host);

int rg_find_indices_2d(int device, void* stream, int64_t nx, int64_t ny,
                       const double* x, const double* y,
                       int64_t n_points, const double* px, const double* py,
                       int64_t fill, int64_t* cell_flat,
                       void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------
 * 2D multilinear (bilinear) weights on a curvilinear vertex grid + their application
 * (BASELINE config 5).  The reference stops at 1D (regridding/_weights/_weights_multilinear.py:128-131
 * raises for 2D); this extends its 1D rule (wml.py:105-119, 185-202) to two axes: values live on the
 * VERTICES, an output point takes the four vertices of its containing cell (cell_flat from
 * rg_find_indices_2d, `fill` = outside) with weights (1-u)(1-v), (1-u)v, u(1-v), uv, (u, v) solving the
 * cell's bilinear map.  bounds_mode: 0 = extrapolate (bilinear map of the nearest border cell continued),
 * 1 = nan (weights of outside points are NaN), 2 = raise (as nan; the caller raises when *n_outside_dev > 0).
 * idx4: int64[n_points][4] flat vertex indices in ascending order; w4: double[n_points][4]; both 32-byte
 * aligned.  n_outside_dev: device int32 counter (stream-ordered).
 * rg_ell4_apply: values_out[f][p] = sum_k w4[p][k] * values_in[f][idx4[p][k]] (k ascending, separately
 * rounded multiply and add from +0.0 -- the accumulation order of
 * regridding/_regrid/_regrid_from_weights.py:179-182 on the sorted triplets).
 * ------------------------------------------------------------------------------ */
int rg_multilinear2d_weights(int device, void* stream, int64_t nx, int64_t ny,
                             const double* x, const double* y,
                             int64_t n_points, const double* px, const double* py,
                             const int64_t* cell_flat, int64_t fill, int bounds_mode,
                             int64_t* idx4, double* w4, int32_t* n_outside_dev);

int rg_ell4_apply(int device, void* stream, int64_t n_frames, int64_t n_in, int64_t n_points,
                  const int64_t* idx4, const double* w4, const double* values_in, double* values_out);

/* Per-slice builds: n_slices grid pairs of one shape (BASELINE config 4: every frame carries its own grid; the
 * reference loops over the orthogonal slices in Python, _weights_conservative.py:110-139), enqueued back to back
 * without any host synchronisation.  Arrays of n_slices device pointers (host arrays of pointers); workspace and
 * `frags` are shared and reused slice after slice; counts_dev is [n_slices][8] with the layout of rg_build2d_band. */
int rg_build2d_batched(int device, void* stream, int64_t n_slices,
                       int64_t nx_in, int64_t ny_in, int64_t nx_out, int64_t ny_out,
                       const double* const* x_in, const double* const* y_in,
                       const double* const* x_out, const double* const* y_out,
                       const double* const* weights_input_or_null,
                       void* workspace, size_t workspace_bytes,
                       void* frags, int64_t frag_capacity,
                       int64_t* const* indices_input, int64_t* const* indices_output, double* const* values,
                       int64_t nnz_capacity, int64_t* counts_dev);

/* ------------------------------------------------------------------------------
 * 1D multilinear weights (weights()'s default method) and the saved-weights ordering of raw triplets
 * replaces  _weights_multilinear / _weights_from_indices_multilinear(_1d)
 *           regridding/_weights/_weights_multilinear.py:9-206 (searchsorted location, clamp / "below" fix-up,
 *           w1 = (x - x0) / (x1 - x0), w0 = 1 - w1, weights_input factors, bounds)
 *           and the ordering of _coalesce, regridding/_weights/_weights_arrays.py:44-73 (stable sort by
 *           (indices_input, indices_output); multilinear elements hold no repeated pair).
 * x_in (D, n) and x_out (D, m) row-major; the D elements come back one after the other, 2 m triplets each, sorted
 * by (input, output); bounds: 0 extrapolate, 1 nan, 2 raise (the count of outside points is returned and the caller
 * raises).  At most 65535 spectra and 2^31 triplets per call. */
int rg_multilinear1d_workspace_bytes(int64_t D, int64_t m, size_t* bytes_host);
int rg_multilinear1d_weights(int device, void* stream, int64_t D, int64_t n, int64_t m,
                             const double* x_in, const double* x_out, const double* weights_input_or_null, int bounds,
                             int64_t* indices_input, int64_t* indices_output, double* values,
                             int64_t* n_outside_host_or_null, void* workspace, size_t workspace_bytes);
/* raw triplets of ONE element (n of them, flat indices below n_in / n_out) -> sorted by (input, output), stable */
int rg_sort_triplets_workspace_bytes(int64_t n, size_t* bytes_host);
int rg_sort_triplets(int device, void* stream, int64_t n, int64_t n_in, int64_t n_out,
                     const int64_t* indices_input, const int64_t* indices_output, const double* values,
                     int64_t* out_indices_input, int64_t* out_indices_output, double* out_values,
                     void* workspace, size_t workspace_bytes);

/* ------------------------------------------------------------------------------
 * fill(method="gauss_seidel"): red-black Gauss-Seidel relaxation of missing cells
 * replaces: _fill_gauss_seidel_2d / _iteration_gauss_seidel_2d  regridding/_fill/_gauss_seidel.py:83-139
 *           (called from fill_gauss_seidel, regridding/_fill/_gauss_seidel.py:13-59)
 * a: double[num_t][num_y][num_x], updated in place (the guess already stored in the missing cells), periodic in
 * both axes.  idx_lists_host: 6 device pointers [colour][level] to int32 flat indices of the missing cells with
 * (i + j) & 1 == colour and level = (i == num_x-1 && num_x odd) + (j == num_y-1 && num_y odd) -- the levels
 * reproduce the order of the reference's sequential sweep across the periodic wrap; counts_host: their lengths.
 * One cooperative launch runs all iterations; the result equals the reference's bit for bit.
 * ------------------------------------------------------------------------------ */
int rg_fill_gauss_seidel_2d(int device, void* stream, double* a, int64_t num_t, int64_t num_y, int64_t num_x,
                            const int32_t* const* idx_lists_host, const int64_t* counts_host, int64_t num_iterations);

/* ------------------------------------------------------------------------------
 * shared-weights apply
 * replaces: _regrid_from_weights(weights, values_input, values_output)
 *           regridding/_regrid/_regrid_from_weights.py:165-182
 *           (called from regridding/_regrid/_regrid_from_weights.py:146-150)
 *
 * rg_csr_from_coo turns the public (input, output)-sorted COO into CSR by OUTPUT
 * row with ascending input index inside each row -- the accumulation order of the
 * reference's sequential loop -- so that rg_apply_csr is bit-identical to it.
 * Negative (wrap-around) indices must be normalised by the caller.
 * ------------------------------------------------------------------------------ */
int rg_csr_workspace_bytes(int64_t nnz, int64_t n_out, size_t* bytes_host);

int rg_csr_from_coo(int device, void* stream, int64_t nnz, int64_t n_in, int64_t n_out,
                    const int64_t* indices_input, const int64_t* indices_output, const double* values,
                    int32_t* row_ptr /* n_out+1 */, int32_t* col /* nnz */, double* val /* nnz */,
                    void* workspace, size_t workspace_bytes);

/* values_in: (n_frames, n_in) row-major, values_out: (n_frames, n_out) row-major; every
 * element of values_out is written (rows without weights get +0.0). */
int rg_apply_csr(int device, void* stream, int64_t n_frames, int64_t n_in, int64_t n_out,
                 const int32_t* row_ptr, const int32_t* col, const double* val,
                 const double* values_in, double* values_out);

/* Planned (shared-memory staged) apply for weights between 2D cell grids
 * (h_in, w_in) -> (h_out, w_out): the same arithmetic and bits as rg_apply_csr, organised
 * for HBM bandwidth.  The plan analyses a CSR once and rg_apply_planned then streams any
 * number of frames through it:
 *   rg_apply_plan_build  per 4x32 output tile: the footprint of input cells it references,
 *                        tile-local u16 indices, and the SLOT layout of its entries (the two
 *                        output cells that share a half-warp are aligned so that their
 *                        shared-memory gathers fall into different banks); returns the number
 *                        of slot entries in *n_slots_host;
 *   rg_apply_plan_slots  fills the caller-allocated slot arrays (weights + byte offsets in
 *                        the order the kernel consumes them).
 * Tiles whose footprint does not fit on chip are counted in *n_generic_tiles_host and
 * served by the generic per-cell kernel.
 * Buffers (caller-allocated; sizes from rg_apply_plan_sizes): tile_info int32[tile_info_ints]
 * (8-byte aligned), tile_rows int32[tile_rows_ints], lidx uint16[nnz]; slot_val
 * double[n_slots], slot_lidx uint16[n_slots] (both 16-byte aligned). */
int rg_apply_plan_sizes(int64_t h_out, int64_t w_out, int64_t* n_tiles_host,
                        int64_t* tile_info_ints_host, int64_t* tile_rows_ints_host);

int rg_apply_plan_build(int device, void* stream, int64_t nnz,
                        int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out,
                        const int32_t* row_ptr, const int32_t* col,
                        int32_t* tile_info, int32_t* tile_rows, uint16_t* lidx,
                        int64_t* n_generic_tiles_host, int64_t* n_slots_host);

int rg_apply_plan_slots(int device, void* stream, int64_t h_out, int64_t w_out,
                        const int32_t* row_ptr, const double* val,
                        const int32_t* tile_info, const uint16_t* lidx,
                        int64_t n_slots, double* slot_val, uint16_t* slot_lidx);

int rg_apply_planned(int device, void* stream, int64_t n_frames,
                     int64_t h_in, int64_t w_in, int64_t h_out, int64_t w_out,
                     const int32_t* row_ptr, const int32_t* col, const double* val,
                     const int32_t* tile_info, const int32_t* tile_rows,
                     const double* slot_val, const uint16_t* slot_lidx,
                     int64_t n_generic_tiles,
                     const double* values_in, double* values_out);

/* ------------------------------------------------------------------------------
 * transposed weights
 * replaces the per-element arithmetic of transpose_weights_conservative
 *           regridding/_weights/_weights_transposed/_weights_transposed.py:236-249
 *           and its cell volumes (_cell_volume_1d :306-322 = cell_length, c1d/_grids.py:10-35;
 *           _cell_volume_2d :325-340 = rg_grid_area)
 * values_transposed[e] = values[e] [/ w[ii[e]]^2] * volume_input[ii[e]] / volume_output[io[e]]
 * in NumPy's evaluation order; negative indices wrap.  The index arrays themselves are
 * only swapped by the caller (transpose_weights, :13-52).
 * ------------------------------------------------------------------------------ */
int rg_cell_length_1d(int device, void* stream, int64_t S, int64_t n, const double* x /* (S, n) */,
                      double* length /* (S, n-1) */);

int rg_transpose_conservative(int device, void* stream, int64_t nnz, int64_t n_in, int64_t n_out,
                              const int64_t* indices_input, const int64_t* indices_output,
                              const double* values, const double* volume_input /* n_in */,
                              const double* volume_output /* n_out */, const double* weights_input /* n_in | NULL */,
                              double* values_transposed);

/* ------------------------------------------------------------------------------
 * 1D conservative, batched over S independent spectra
 * replaces: weights_conservative_1d(x_input, x_output, weights_input, weights_output, start, stop)
 *           regridding/_weights/_weights_conservative_1d/_weights_conservative_1d.py:12-56, 60-189
 *           (called from regridding/_weights/_weights_conservative.py:89-97)
 * x_in: (S, n) edges, x_out: (S, m) edges.  Triplets of spectrum s are written at
 * offset s*(n+m) of the output arrays (capacity n+m each), counts[s] of them, in the
 * reference's emission order; indices are the reference's (possibly negative,
 * complemented) indices.
 * ------------------------------------------------------------------------------ */
int rg_cons1d_batched(int device, void* stream, int64_t S, int64_t n, int64_t m,
                      const double* x_in, const double* x_out, const double* weights_input_or_null,
                      int64_t* indices_input, int64_t* indices_output, double* values,
                      int64_t* counts);

/* Fused 1D conservative regrid (weights never materialised): values_in (S, n-1) -> values_out (S, m-1),
 * accumulating in the reference's order (rfw.py:179-182 over c1d.py emission order). */
int rg_regrid1d_conservative(int device, void* stream, int64_t S, int64_t n, int64_t m,
                             const double* x_in, const double* x_out, const double* weights_input_or_null,
                             const double* values_in, double* values_out);

/* ------------------------------------------------------------------------------
 * 1D cell location
 * replaces: _find_indices_brute_1d      regridding/_find_indices/_find_indices_brute.py:24-51
 *           _find_indices_searchsorted_1d regridding/_find_indices/_find_indices_searchsorted.py:24-62
 *           (called from regridding/_find_indices/_find_indices.py:112-123)
 * method: 0 = brute, 1 = searchsorted.  x_in (D, n), x_out (D, m) -> out (D, m).
 * ------------------------------------------------------------------------------ */
int rg_find_indices_1d(int device, void* stream, int method, int64_t D, int64_t n, int64_t m,
                       const double* x_in, const double* x_out, int64_t fill, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* REGRID_B200_H */
