/*
 * improver_b200 — C ABI of the B200 (sm_100a) neighbourhood-processing engine.
 *
 * This is the drop-in boundary for the hot path of metoppv/improver's
 * `improver.nbhood` / `improver.threshold`.  The reference is
 * pure Python: the "FFI" a maintainer would add is a ctypes binding (shown in
 * INTEGRATION.md).  Every entry point takes plain pointers and sizes.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller unless marked
 *     HOST; arrays are C-contiguous, spatial dims last: (..., y, x);
 *   - `stream` is a cudaStream_t passed as void*; work is stream-ordered, no
 *     entry point synchronises or allocates;
 *   - return 0 on success, non-zero on error (message: nbh_last_error());
 *   - float32 in / float32 out like the reference (FLOAT_DTYPE,
 *     improver/metadata/constants/__init__.py:9); sums are accumulated in
 *     float64 like improver/nbhood/nbhood.py:282-285.
 */
#ifndef IMPROVER_B200_H
#define IMPROVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBH_ABI_VERSION 2

/* comparison operators, improver/utilities/probability_manipulation.py:15-66 */
enum { NBH_OP_GT = 0, NBH_OP_GE = 1, NBH_OP_LT = 2, NBH_OP_LE = 3, NBH_OP_EQ = 4 };

/* flags */
enum {
  NBH_SUM_ONLY = 1,   /* NeighbourhoodProcessing(sum_only=True), nbhood.py:188 */
  NBH_WRAP_X   = 2,   /* extension: periodic x (global grids); no reference parity */
  NBH_WEIGHTED = 4,   /* circular weighted_mode, nbhood.py:55-92 */
  NBH_EXACT_SUMS = 8, /* single-member square launches: only kernels whose window sums do not depend on
                         where a tile / band starts (a y-band then reproduces the whole grid bit for bit) */
  NBH_LAND_SEA = 16   /* mask2d is a land mask (non-zero = land): every point is normalised over the points
                         of its OWN surface type, i.e. the land pass + sea pass + filled(0) sum of
                         improver/cli/nbhood_land_and_sea.py:148-184 in one launch (re_mask is implied) */
};

/* One threshold with its fuzzy bounds (improver/threshold.py:302-333).
 * lo == hi selects the deterministic comparison (threshold.py:438-439). */
typedef struct NbhThreshold {
  double value;
  double lo;
  double hi;
  int32_t op;       /* NBH_OP_* */
  int32_t _pad;
} NbhThreshold;

/* Description of one neighbourhood launch.
 *
 * Input is viewed as planes x members of 2-D (ny, nx) float32 slices:
 *   slice(p, m) = in + p * plane_stride + m * member_stride      (elements)
 * For NeighbourhoodProcessing.process (nbhood.py:457-463) every slice is its
 * own plane (n_members = 1).  For the fused Threshold -> neighbourhood ->
 * realization-mean chain, a plane is one lead time and the members are the
 * realizations collapsed by the mean (cube_manipulation.py:93-131).
 * Output: float32 (n_planes, max(n_thresholds,1), ny, nx).
 */
typedef struct NbhDesc {
  int64_t n_planes;
  int32_t n_members;
  int32_t ny;
  int32_t nx;
  int32_t cells;            /* radius in grid cells: int(radius / spacing), spatial.py:152-156 */
  int64_t plane_stride;
  int64_t member_stride;
  int32_t n_thresholds;     /* 0: neighbourhood of the input itself */
  int32_t flags;
  const NbhThreshold* thresholds;   /* HOST pointer, n_thresholds entries */
  const uint8_t* mask2d;    /* (ny,nx) non-zero = valid point (mask_cube, nbhood.py:453-456) or NULL */
  const uint8_t* mask_nd;   /* same shape as input, non-zero = MASKED (np.ma mask) or NULL;
                               requires n_members == 1 */
  /* --- ABI 2 --- */
  const int32_t* plane_cells;  /* HOST array of n_planes radii in grid cells (lead-time dependent radii,
                                  nbhood.py:132-149 / 438-455: one radius per slice) or NULL: `cells` for all */
  /* y-band tiling of one huge grid: this launch sees rows
   * [y_offset, y_offset + ny) of a grid of ny_global rows.  The band must include the halo rows it
   * received from its neighbours; `halo_top` / `halo_bottom` of its first / last rows are such
   * halo rows (their outputs are not produced).  ny_global == 0: not tiled. */
  int32_t y_offset;
  int32_t ny_global;
  int32_t halo_top;
  int32_t halo_bottom;
  /* per-slice statistics of the WHOLE grid (nbh_slice_stats_f32 on every band + min / max
   * all-reduce), n_planes * n_members * max(n_thresholds, 1) entries, DEVICE pointer; required
   * when ny_global != 0 and the launch is not sum_only (the reference's clip and its trimming
   * bounding boxes, nbhood.py:252-256 / 341-409, are properties of the whole slice) */
  const struct NbhSliceStats* ext_stats;
} NbhDesc;

/* Exact statistics of one 2-D slice after thresholding / masking (10 x int32).  The *min fields
 * combine across bands with MIN, the *max fields with MAX. */
typedef struct NbhSliceStats {
  int32_t y0min, y0max, x0min, x0max;   /* bounding box of values != 0 (rows global: + y_offset) */
  int32_t y1min, y1max, x1min, x1max;   /* bounding box of values != 1 */
  int32_t vmin_key, vmax_key;           /* nanmin / nanmax of the slice as ordered-int keys:
                                           key = bits >= 0 ? bits : bits ^ 0x7fffffff */
} NbhSliceStats;

int nbh_version(void);
const char* nbh_last_error(void);
/* number of SMs and opt-in shared memory of the current device */
int nbh_device_info(int32_t* sm_count, int64_t* smem_optin_bytes);

/* Per-slice statistics of a band (or of a whole grid): slice index
 * (plane * n_members + member) * max(n_thresholds, 1) + threshold.  Replaces the per-slice
 * nanmin / nanmax and np.argwhere bounding boxes of improver/nbhood/nbhood.py:252-256, 353-382. */
int nbh_slice_stats_f32(const NbhDesc* desc, const float* in, NbhSliceStats* stats, void* stream);

/* y-band tiling, the data movement of the halo exchange that does not need a communicator:
 * the rows a band sends to its neighbours / receives from them are packed into / unpacked from
 * contiguous buffers, so that the exchange itself is one ncclSend / ncclRecv pair per neighbour
 * (the caller owns the communicator: torch.distributed in improver_b200.distributed, or
 * ncclGroupStart(); ncclSend(send_up..); ncclRecv(recv_up..); ... ncclGroupEnd() from C).
 *   band      float32 (n_slices, rows, nx): own rows of every slice
 *   ext       float32 (n_slices, halo_top + rows + halo_bottom, nx): the extended band
 *   nbh_halo_pack:   send_up   <- first `cells` rows of every slice (n_slices, cells, nx) or NULL
 *                    send_down <- last  `cells` rows                                   or NULL
 *                    and copies the own rows into ext
 *   nbh_halo_unpack: ext top halo <- recv_up (n_slices, halo_top, nx), bottom halo <- recv_down */
int nbh_halo_pack(const float* band, float* ext, float* send_up, float* send_down, int64_t n_slices,
                  int32_t rows, int32_t nx, int32_t cells, int32_t halo_top, int32_t halo_bottom,
                  void* stream);
int nbh_halo_unpack(float* ext, const float* recv_up, const float* recv_down, int64_t n_slices,
                    int32_t rows, int32_t nx, int32_t halo_top, int32_t halo_bottom, void* stream);

/* The whole exchange in one stream-ordered call: pack, one ncclSend / ncclRecv pair per
 * neighbouring rank inside one NCCL group (NVLink / NVSwitch, no host staging), unpack.
 *   nccl_comm  the caller's ncclComm_t (ranks ordered top band = 0 ... bottom band = n_ranks - 1)
 *   ext        float32 (n_slices, halo_top + rows + halo_bottom, nx) with halo_top = rank > 0 ? cells : 0,
 *              halo_bottom = rank < n_ranks - 1 ? cells : 0
 * NCCL is resolved at run time from the library already loaded into the process. */
size_t nbh_halo_exchange_workspace_bytes(int64_t n_slices, int32_t nx, int32_t cells);
int nbh_halo_exchange(void* nccl_comm, int32_t rank, int32_t n_ranks, const float* band, float* ext,
                      int64_t n_slices, int32_t rows, int32_t nx, int32_t cells, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Scratch the caller must provide (zero-initialisation not required). */
size_t nbh_workspace_bytes(const NbhDesc* desc);

/* Replaces NeighbourhoodProcessing._calculate_neighbourhood with method
 * "square" (improver/nbhood/nbhood.py:244-409, boxsum at
 * improver/utilities/neighbourhood_tools.py:129-211), optionally fused with
 * Threshold._calculate_truth_value (improver/threshold.py:407-471) in front
 * and the realization mean behind.
 *   out       float32 (n_planes, T, ny, nx)
 *   out_mask  uint8 same shape as out or NULL: 1 where the reference's re_mask
 *             would mask the point (nbhood.py:314-315)
 *   status    int32[2] device: [0] = 1 if an unmasked NaN was read
 *             (nbhood.py:167-168), [1] = number of slices the trimming fix-up touched
 */
int nbh_square_f32(const NbhDesc* desc, const float* in, float* out, uint8_t* out_mask,
                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Same contract for method "circular" (nbhood.py:55-92 kernel,
 * scipy.ndimage.correlate(mode="nearest") at nbhood.py:401). */
int nbh_circular_f32(const NbhDesc* desc, const float* in, float* out, uint8_t* out_mask,
                     int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces GeneratePercentilesFromANeighbourhood.pad_and_unpad_cube
 * (nbhood.py:649-667): in float32 (n_slices, ny, nx) ->
 * out float32 (n_slices, n_pct, ny, nx); pct: HOST float32 percentiles. */
int nbh_percentiles_f32(const float* in, float* out, const float* pct, int32_t n_pct,
                        int64_t n_slices, int32_t ny, int32_t nx, int32_t cells,
                        int32_t* status, void* workspace, size_t workspace_bytes, void* stream);
size_t nbh_percentiles_workspace_bytes(int64_t n_slices, int32_t ny, int32_t nx, int32_t cells);

/* Replaces the Threshold.process hot loop (improver/threshold.py:643-713).
 *   in        float32 (n_collapse, n_points) ; mask_nd uint8 same shape or NULL
 *   out       float32 (T, n_points): mean truth value over unmasked members
 *   contrib   int32 (n_points) or NULL: number of unmasked members (threshold.py:679)
 *   n_collapse = 1 when no coordinate is collapsed.
 */
int nbh_threshold_f32(const float* in, const uint8_t* mask_nd, float* out, int32_t* contrib,
                      const NbhThreshold* thresholds, int32_t n_thresholds,
                      int64_t n_collapse, int64_t n_points, int32_t* status, void* stream);

/* Mean over the leading axis: replaces collapse_realizations (improver/utilities/cube_manipulation.py:93-131,
 * iris MEAN = numpy's float32 mean over axis 0).  in float32 (n_collapse, n_points) -> out float32 (n_points). */
int nbh_collapse_mean_f32(const float* in, float* out, int64_t n_collapse, int64_t n_points, void* stream);

/* boxsum of the utility API (improver/utilities/neighbourhood_tools.py:129-211) in float64, any
 * odd box (box_y, box_x): in float64 (n_slices, rows, cols) -> out float64 (n_slices, rows - box_y,
 * cols - box_x); out[i][j] = sum of in[i+1 .. i+box_y][j+1 .. j+box_x] (cumsum = 1), or the
 * four-corner difference of an already cumulative input (cumsum = 0). */
int nbh_boxsum_f64(const double* in, double* out, int64_t n_slices, int32_t rows, int32_t cols, int32_t box_y,
                   int32_t box_x, int32_t cumsum, void* stream);

/* out = float32(a * double(in) + b): the unit conversion of the data into the threshold units
 * (improver/threshold.py:430-431 `cube.convert_units`, cf_units on a float32 array). */
int nbh_linear_map_f32(const float* in, float* out, double a, double b, int64_t n, void* stream);

/* Zone collapse of ApplyNeighbourhoodProcessingWithAMask.collapse_mask_coord
 * (improver/nbhood/use_nbhood.py:158-192): NaN-aware weighted mean over the zone axis with renormalised
 * weights.  values float32 (n_lead, n_zones, n_points), weights float32 (n_zones, n_points) (masked
 * weights as 0) -> out float32 (n_lead, n_points), NaN where no zone is valid. */
int nbh_zone_collapse_f32(const float* values, const float* weights, float* out, int64_t n_lead, int32_t n_zones,
                          int64_t n_points, void* stream);

/* Threshold "in vicinity": replaces Threshold._vicinity_processing (improver/threshold.py:473-518)
 * with maximum_within_vicinity (improver/utilities/spatial.py:740-855) inside the accumulation
 * loop of Threshold.process (threshold.py:661-707).
 *   in        float32 (n_collapse, n_planes, ny, nx); mask_nd uint8 same shape or NULL (non-zero = masked)
 *   landmask  uint8 (ny, nx) or NULL: non-zero = land; land and sea points are spread separately
 *   vicinity_cells  HOST array of n_vicinity radii in grid cells (square of side 2 r + 1)
 *   out       float32 (n_vicinity, T, n_planes, ny, nx): mean over unmasked members of the
 *             maximum truth value within the vicinity
 *   contrib   int32 (n_planes, ny, nx) or NULL: number of unmasked members (threshold.py:679)
 */
int nbh_threshold_vicinity_f32(const float* in, const uint8_t* mask_nd, const uint8_t* landmask, float* out,
                               int32_t* contrib, const NbhThreshold* thresholds, int32_t n_thresholds,
                               const int32_t* vicinity_cells, int32_t n_vicinity, int64_t n_collapse,
                               int64_t n_planes, int32_t ny, int32_t nx, int32_t* status, void* stream);

/* Maximum / minimum within a square vicinity of every point of every 2-D slice: replaces
 * maximum_within_vicinity / minimum_within_vicinity (improver/utilities/spatial.py:805-903) as used by
 * OccurrenceWithinVicinity.process (:1176-1255).  op: 0 maximum, 1 minimum.
 *   in float32 (n_planes, ny, nx) -> out float32 (n_vicinity, n_planes, ny, nx); masked points
 *   (mask_nd) neither contribute nor are written (0 is stored there). */
int nbh_vicinity_f32(const float* in, const uint8_t* mask_nd, const uint8_t* landmask, float* out,
                     const int32_t* vicinity_cells, int32_t n_vicinity, int32_t op, int64_t n_planes,
                     int32_t ny, int32_t nx, int32_t* status, void* stream);

/* Recursive filter: replaces RecursiveFilter._run_recursion together with the symmetric halo
 * (pad_cube_with_halo), the zeroing of coefficients around masked points
 * (OrographicSmoothingCoefficients.zero_masked) and remove_halo_from_cube, i.e. the per-slice body of
 * RecursiveFilter.process (improver/nbhood/recursive_filter.py:155-202, 262-286, 403-436).
 *   in       float32 (n_slices, ny, nx)
 *   coeff_x  float32 (ny, nx - 1), coeff_y float32 (ny - 1, nx): smoothing coefficients
 *   mask     uint8 or NULL, non-zero = masked point; one (ny, nx) plane for every slice
 *            (mask_per_slice = 0) or (n_slices, ny, nx) (mask_per_slice = 1)
 *   mask_zeros  non-zero: points whose input is exactly 0 are masked as well (:395-396)
 *   pad      halo width in cells = 2 * edge_width; iterations >= 1
 *   out      float32 (n_slices, ny, nx)
 * Every float32 operation is rounded as numpy rounds it, so results are bit-identical. */
int nbh_recursive_filter_f32(const float* in, const float* coeff_x, const float* coeff_y, const uint8_t* mask,
                             int32_t mask_per_slice, int32_t mask_zeros, float* out, int64_t n_slices,
                             int32_t ny, int32_t nx, int32_t pad, int32_t iterations, void* workspace,
                             size_t workspace_bytes, void* stream);
size_t nbh_recursive_filter_workspace_bytes(int64_t n_slices, int32_t ny, int32_t nx, int32_t pad,
                                            int32_t coeff_per_slice);

/* One directional sweep in place: RecursiveFilter._recurse_forward / _recurse_backward
 * (improver/nbhood/recursive_filter.py:55-153).  grid float32 (n_slices, ny, nx);
 * axis 0 = y with coeffs (ny - 1, nx), axis 1 = x with coeffs (ny, nx - 1);
 * dirs 1 = forward, 2 = backward, 3 = forward then backward. */
int nbh_recurse_f32(float* grid, const float* coeffs, int64_t n_slices, int32_t ny, int32_t nx, int32_t axis,
                    int32_t dirs, void* stream);

/* Measurement aid (bench.py roofline): when enabled, the dominant kernel of
 * the next nbh_square_f32 / nbh_circular_f32 / nbh_percentiles_f32 call is
 * bracketed by CUDA events on the caller's stream.  nbh_profile_last_ms
 * synchronises on those events and returns the kernel's duration. */
int nbh_profile_enable(int on);
int nbh_profile_last_ms(float* ms);
/* name of the kernel that was bracketed (static string, "" if none) */
const char* nbh_profile_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* IMPROVER_B200_H */
