/* mmb200.h - C ABI of the B200-native MagellanMapper blob-detection library.
 *
 * The reference (sanderslab/magellanmapper) is pure Python and has no FFI for
 * this path; its arithmetic lives in scikit-image / scipy wheels.  Each entry
 * point below replaces one of those call sites (reference file:line cited) and
 * is what a ctypes binding on the reference side would load.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types.
 *   - every function returns 0 on success or a negative MMB_ERR_* code;
 *     mmb_last_error() returns a thread-local message for the last failure.
 *   - all volume buffers are CALLER-OWNED DEVICE pointers.  A float volume is
 *     [Z][Y][pitch] with `pitch >= X` elements per row (rows may be padded).
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it
 *     unless documented otherwise.
 *   - variable-length outputs use a caller-sized buffer + a device counter; the
 *     counter keeps counting past `capacity`, so overflow is detectable and is
 *     reported (MMB_ERR_OVERFLOW), never silently truncated.
 */
#ifndef MMB200_H
#define MMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMB_VERSION 1

enum {
  MMB_OK = 0,
  MMB_ERR_INVALID = -1,      /* bad argument */
  MMB_ERR_CUDA = -2,         /* CUDA runtime error, see mmb_last_error() */
  MMB_ERR_OVERFLOW = -3,     /* candidate / edge buffer too small */
  MMB_ERR_UNSUPPORTED = -4   /* shape or option outside the implemented range */
};

/* input voxel types accepted by the converting entry points */
enum { MMB_U8 = 0, MMB_U16 = 1, MMB_F32 = 2, MMB_F64 = 3 };

/* One local-maximum candidate of the (z, y, x, scale) LoG cube. 20 bytes. */
typedef struct mmb_cand {
  int32_t z, y, x;   /* voxel coordinate inside the volume passed in */
  int32_t s;         /* index into the sigma ladder */
  float resp;        /* -sigma^2 * LoG response at the peak */
} mmb_cand;

/* ROI-profile preprocessing keys (magmap/settings/roi_prof.py:72-84). */
typedef struct mmb_preproc_params {
  double clip_vmin;          /* lower percentile, 0-100                         */
  double clip_vmax;          /* upper percentile, 0-100                         */
  double max_thresh;         /* config.near_max[chl] * max_thresh_factor        */
  double clip_min, clip_max; /* clip after stretching                           */
  double unsharp_strength;   /* 0 disables the sigma=8 unsharp mask             */
  double erosion_threshold;  /* 0 disables; erode when block mean exceeds this  */
} mmb_preproc_params;

int mmb_version(void);
const char* mmb_last_error(void);

/* ---- host -> device feed of a sub-box -----------------------------------------
 * Replaces the numpy slicing `roi = image5d[0, z0:z1, y0:y1, x0:x1]` /
 * `cls.img[sub_roi_slices]` that hands a sub-ROI to a worker
 * (magmap/cv/stack_detect.py:362, :105-107): copies `planes` pieces of
 * `piece_bytes` contiguous bytes each, `src_pitch_bytes` apart in HOST memory
 * (pinned for a true asynchronous DMA), to a dense device buffer - e.g. the rows
 * y0..y1 of every z-plane of a (Z, Y, X) volume.  Asynchronous on `stream`.      */
int mmb_upload_pieces(void* dst_device, const void* src_host, int64_t planes,
                      int64_t piece_bytes, int64_t src_pitch_bytes, void* stream);

/* ---- import metadata: intensity bounds ------------------------------------------
 * Replaces the np.percentile calls behind the image metadata `near_min` /
 * `near_max` that saturate_roi consumes (magmap/io/importer.py:1368-1377 and
 * :571-583 per z-plane, calc_intensity_bounds :1415-1444 over a whole array;
 * numpy 'linear' method, float64): out[g * nq + i] = percentile q_percent[i] of
 * group g, where a group is one z-plane (`per_plane` != 0, Z groups) or the whole
 * (Z, Y, X) view (one group).  uint8 / uint16 input with element strides, so one
 * channel of a channel-last array is read in place; other dtypes return
 * MMB_ERR_UNSUPPORTED.  `out_device` holds groups * nq doubles, `work`
 * mmb_percentiles_work_bytes(groups) bytes.  Asynchronous on `stream`.             */
int64_t mmb_percentiles_work_bytes(int n_groups);
int mmb_percentiles(const void* in, int dtype, const int64_t in_strides[3],
                    int Z, int Y, int X, int per_plane, const double* q_percent,
                    int nq, double* out_device, void* work, void* stream);

/* ---- img_as_float ---------------------------------------------------------
 * Replaces skimage.util.img_as_float inside blob_log (reference call site
 * magmap/cv/detector.py:931): out = in * scale, element strides given per axis
 * so a channel of a channel-last (z,y,x,c) array can be read in place.        */
int mmb_to_float(const void* in, int dtype, const int64_t in_strides[3],
                 int Z, int Y, int X, float* out, int64_t pitch, double scale,
                 void* stream);

/* ---- isotropic resize -------------------------------------------------------
 * Replaces cv_nd.make_isotropic (magmap/cv/cv_nd.py:1071-1106, called from
 * magmap/cv/detector.py:893-897), i.e. skimage.transform.resize(roi, out_shape,
 * mode='reflect', preserve_range=True).astype(roi.dtype): Gaussian anti-aliasing on
 * the axes that shrink (sigma = (in/out - 1)/2, truncate 4), then order-1
 * scipy.ndimage.zoom(grid_mode=True), all in float64; integer dtypes are truncated
 * back to integers.  `edge_mode`: 0 = 'reflect' (ndimage 'mirror'), 1 = 'edge'
 * (ndimage 'nearest'; the reference's choice for ROIs one voxel thick).  The result
 * is a pitched float32 volume [Zo][Yo][pitch_out].                                  */
int mmb_resize_linear(const void* in, int dtype, const int64_t in_strides[3], int Z, int Y,
                      int X, float* out, int Zo, int Yo, int Xo, int64_t pitch_out,
                      int edge_mode, void* stream);

/* ---- spectral unmixing ------------------------------------------------------
 * One step of the loop at magmap/cv/detector.py:910-921:
 * target = max(target - factor * other, 0) on two pitched float32 volumes.         */
int mmb_unmix_subtract(float* target, const float* other, int Z, int Y, int X, int64_t pitch,
                       double factor, void* stream);

/* ---- saturate_roi + denoise_roi per preprocessing block --------------------
 * Replaces the loop at magmap/cv/stack_detect.py:122-150 that calls
 * plot_3d.saturate_roi (plot_3d.py:55-111) and plot_3d.denoise_roi
 * (plot_3d.py:114-172) on every (bz,by,bx) block anchored at the chunk origin:
 * exact np.percentile (linear) -> stretch -> clip -> sigma=8 'nearest' unsharp
 * mask -> octahedron(1) erosion when the stretched block mean > threshold.
 * Blocks up to 32 voxels a side: one CTA per block, block resident in shared memory.
 * Larger blocks - the whole ROI of the GUI path (block shape = volume shape,
 * magmap/gui/visualizer.py:2742-2743), the `lowres` profile - run as whole-volume
 * kernels (radix-select percentiles, 65-tap 'nearest' blur per block); any size.
 * Equal clip_vmin and clip_vmax skip the stretch (denoise_roi on its own).      */
int mmb_preprocess_blocks(const void* in, int dtype, const int64_t in_strides[3],
                          int Z, int Y, int X, int bz, int by, int bx,
                          const mmb_preproc_params* p, float* out, int64_t pitch,
                          void* stream);

/* ---- one scale of the LoG cube ---------------------------------------------
 * Replaces `-scipy.ndimage.gaussian_laplace(img, sigma) * sigma**2`
 * (skimage blob_log, reached from magmap/cv/detector.py:931): truncate=4,
 * radius=int(4*sigma+0.5), mode='reflect' on all faces of the volume given.
 * `work` must hold mmb_log_work_bytes() bytes; `out` may not alias `in`.      */
int64_t mmb_log_work_bytes(int Z, int Y, int64_t pitch);
int mmb_log_scale(const float* in, float* out, void* work, int Z, int Y, int X,
                  int64_t pitch, double sigma, void* stream);

/* Separable Gaussian-derivative passes exposed one by one for parity tests and
 * per-pass roofline measurement.  axis: 0=z 1=y 2=x.  mode: 0 = one input ->
 * (g*in0, h*in0); 1 = (in0,in1) -> (g*in0, h*in0 + g*in1); 2 = (in0,in1) ->
 * scale*(h*in0 + g*in1) in out0.  g = sampled Gaussian, h = its second
 * derivative, as scipy.ndimage._filters._gaussian_kernel1d builds them.       */
int mmb_log_pass(const float* in0, const float* in1, float* out0, float* out1,
                 int Z, int Y, int X, int64_t pitch, int axis, int mode,
                 double sigma, double scale, void* stream);

/* ---- 4-D local maxima + threshold ------------------------------------------
 * Replaces skimage.feature.peak_local_max(cube, threshold_abs=thr,
 * footprint=ones((3,3,3,3)), exclude_border=False) for scale `s` given the
 * neighbouring scales (NULL where absent = 'nearest' at the ends of the scale
 * axis).  Emits voxels with z in [z_lo, z_hi).  *counter is a device int that
 * the caller zeroes; it counts every peak even beyond `capacity`.             */
int mmb_localmax_compact(const float* prev, const float* cur, const float* next,
                         int Z, int Y, int X, int64_t pitch, int s, float thr,
                         int z_lo, int z_hi, mmb_cand* out, int capacity,
                         int* counter, void* stream);

/* ---- within-volume overlap pruning ------------------------------------------
 * Replaces skimage.feature.blob._prune_blobs (sphere-overlap test on
 * [z,y,x,sigma] rows) with the order-independent resolution described in
 * DESIGN.md.  keep[i] = 1 if candidate i survives.  Synchronous on `stream`.
 * Candidate order matters only for ties: on equal sigma the candidate with the
 * higher response (then lower C-order index) is the one removed, as in
 * scikit-image.  `dims` = {Y, X} of the volume the coordinates refer to (for
 * the tie order), `num_sigma` = ladder length.                                */
int mmb_prune_within(const mmb_cand* cand, int n, const double* sigmas,
                     int num_sigma, double overlap, int Y, int X,
                     uint8_t* keep, void* stream);
/* Same result for candidates listed by ascending z: the pair search then stops
 * at the cut-off distance 2*sigma_max*sqrt(3)+1 in z instead of visiting all
 * pairs.  Used for the single prune over the gathered candidates of every
 * z-slab in the seamless multi-GPU mode (no reference counterpart: one
 * _prune_blobs call over the whole volume, detector.py:931).                 */
int mmb_prune_within_zsorted(const mmb_cand* cand, int n, const double* sigmas,
                             int num_sigma, double overlap, int Y, int X,
                             uint8_t* keep, void* stream);

/* ---- seam matching ------------------------------------------------------------
 * Replaces detector._find_close_blobs inside remove_close_blobs
 * (magmap/cv/detector.py:1000-1085): box test |d| <= tol per axis between
 * master and check rows (int32 z,y,x triples).  check_hit[j] = 1 if any master
 * matches; master_last[i] = largest matching check index or -1 (the match that
 * wins the reference's repeated fancy-index assignment).                     */
int mmb_prune_seams(const int32_t* master_zyx, int n_master,
                    const int32_t* check_zyx, int n_check, const int32_t tol[3],
                    int32_t* master_last, uint8_t* check_hit, void* stream);

/* ---- fused per-chunk driver -----------------------------------------------------
 * Replaces StackDetector.detect_sub_roi's arithmetic
 * (magmap/cv/stack_detect.py:122-158): optional block preprocessing (pre !=
 * NULL, else plain img_as_float with `scale`), the num_sigma LoG scales kept in
 * a 3-deep ring, local maxima, overlap pruning.  Survivors are compacted to
 * the front of `cand`; *n_out (host) receives their count and *n_peaks (host,
 * may be NULL) the number of local maxima before pruning.  Synchronous.
 * `work` must hold mmb_detect_work_bytes() bytes.                             */
int64_t mmb_detect_work_bytes(int Z, int Y, int64_t pitch, int capacity);
int mmb_detect_chunk(const void* in, int dtype, const int64_t in_strides[3],
                     int Z, int Y, int X, int64_t pitch, double scale,
                     const mmb_preproc_params* pre, int bz, int by, int bx,
                     const double* sigmas, int num_sigma, double threshold,
                     double overlap, int z_lo, int z_hi, void* work,
                     mmb_cand* cand, int capacity, int* n_out, int* n_peaks,
                     void* stream);

/* Asynchronous form: enqueues every kernel of the chunk on `stream` and returns
 * without synchronising; nothing is read back to the host.  `status` is a DEVICE
 * int32[4] written at the end of the chunk: [0] local maxima found (may exceed
 * `capacity`), [1] survivors compacted to the front of `cand`, [2] kill edges
 * found by the overlap pruning (may exceed mmb_detect_edge_capacity(capacity)),
 * [3] number of candidates whose survival depends on the iteration order of
 * scikit-image's pair set (at least one killer, none of them a root).
 * The caller copies `status` and `cand[0 .. status[1])` back after the stream
 * reaches this point; if [0] > capacity or [2] > edge capacity the results are
 * incomplete and the chunk must be redone with a larger capacity.  `work` and
 * `cand` may be reused by the next enqueue on the same stream.  num_sigma <= 64. */
int mmb_detect_chunk_enqueue(const void* in, int dtype, const int64_t in_strides[3],
                             int Z, int Y, int X, int64_t pitch, double scale,
                             const mmb_preproc_params* pre, int bz, int by, int bx,
                             const double* sigmas, int num_sigma, double threshold,
                             double overlap, int z_lo, int z_hi, void* work,
                             mmb_cand* cand, int capacity, int32_t* status,
                             void* stream);
int mmb_detect_edge_capacity(int capacity);

/* ---- blob tables of a chunked stack -------------------------------------------------
 * One detected blob as it leaves a chunk: chunk-local voxel, index into the channel's
 * sigma ladder, LoG response, linear index of the chunk in the C-ordered chunk grid,
 * position of the channel in the channel list.  32 bytes.                             */
typedef struct mmb_row {
  int32_t z, y, x;
  int32_t s;
  float resp;
  int32_t chunk;
  int32_t channel;
  int32_t reserved;
} mmb_row;

/* Tags the survivors of one chunk (`cand[0 .. n)` of mmb_detect_chunk_enqueue) with
 * their chunk and channel: the device-side form of Blobs.format_blobs + the chunk
 * columns chunking.merge_blobs appends (magmap/cv/detector.py:325-364,
 * magmap/cv/chunking.py:410-445).                                                     */
int mmb_rows_from_cands(const mmb_cand* cand, int n, int chunk, int channel, mmb_row* out,
                        void* stream);

/* Appends the candidates `cand[i]` with keep[i] != 0 (all when `keep` is NULL) that lie,
 * after adding `shift` (z, y, x), inside the box [lo, hi) to `dst` at *counter (a DEVICE
 * int32 that keeps counting past `capacity`, so overflow is detectable).  The seamless
 * multi-GPU mode uses it to turn tile-local local maxima into the owned part of the
 * global candidate list, and to list the survivors of the single global _prune_blobs
 * (no reference counterpart: the reference has one process-local list, detector.py:931). */
int mmb_cands_append(const mmb_cand* cand, int n, const uint8_t* keep, const int32_t shift[3],
                     const int32_t lo[3], const int32_t hi[3], mmb_cand* dst, int capacity,
                     int32_t* counter, void* stream);

/* Chunk-grid geometry (chunking.stack_splitter, stack_detect.setup_blocks): HOST
 * pointers, read before the call returns.                                            */
typedef struct mmb_stack_geom {
  int32_t grid[3];           /* chunks per axis (z, y, x), each <= 128                  */
  int32_t overlap[3];        /* Blocks.overlap                                          */
  int32_t tol[3];            /* Blocks.tol (inclusive box tolerance, < 64)              */
  int32_t pad[3];            /* Blocks.overlap_padding                                  */
  const int32_t* start[3];   /* start[a][j]: first voxel of chunk section j on axis a  */
  const int32_t* size[3];    /* size[a][j]: its extent                                  */
  int32_t n_channels;        /* channels detected in this pass                          */
  int32_t num_sigma;         /* row pitch of `sigmas`                                   */
  const double* sigmas;      /* [n_channels][num_sigma] sigma ladders                   */
  const int32_t* channel_ids;/* [n_channels] value of the table's channel column       */
} mmb_stack_geom;

/* Replaces chunking.merge_blobs -> StackPruner.prune_blobs_mp (prune_overlap,
 * detector.remove_close_blobs) -> the final column layout of detect_blobs_blocks
 * (magmap/cv/stack_detect.py:680-861, :644-677, :455-467; detector.py:1009-1085) on
 * the rows of every chunk, in any order: rows are put into merge order (chunks in grid
 * order, channels in request order, peak_local_max order inside a detection), seams
 * are pruned axis by axis exactly as the reference walks them, and the float64 table is
 * written to `out_table` (capacity n rows): 8 columns z, y, x, radius, confirmed,
 * truth, channel, region when `final_layout`, else the 11 Blobs.Cols columns.
 * *n_out (DEVICE int32) = rows written.  seam_counts (DEVICE, may be NULL):
 * [n_channels][3][128][4] int32 = per channel, axis and seam {rows in the slab, rows
 * after pruning, rows in the ratio region, matched checks} for
 * detector.meas_pruning_ratio.  Asynchronous on `stream`; `work` holds
 * mmb_stack_tables_work_bytes(n) bytes.  A grid of one chunk just orders the rows
 * (peak_local_max order) and formats them.                                           */
int64_t mmb_stack_tables_work_bytes(int n_rows);
int mmb_stack_tables(const mmb_row* rows, int n, const mmb_stack_geom* geom, int final_layout,
                     double* out_table, int32_t* n_out, int32_t* seam_counts, void* work,
                     void* stream);

/* ---- intensity co-localisation ----------------------------------------------
 * The voxel work of colocalizer.colocalize_blobs (magmap/cv/colocalizer.py:340-441):
 * for every channel, every blob of that channel owns the voxels of its
 * skimage.morphology.ball(2) neighbourhood that no blob of the same channel with a
 * larger index reaches (the grey dilation of the index-labelled mask); the ROI
 * intensities of the owned voxels are summed for EVERY channel.  `roi`: (z, y, x, c)
 * with element strides; `blobs`: (n, 4) int32 rows z, y, x, channel (all inside the
 * ROI); outputs `sums` (n, C) float64 and `counts` (n) int32, so that the reference's
 * np.mean(roi[mask == b, c]) is sums[b][c] / counts[b].  Integer ROIs are summed
 * exactly.  `work`: mmb_coloc_work_bytes bytes.                                    */
int64_t mmb_coloc_work_bytes(int Z, int Y, int X);
int mmb_coloc_sums(const void* roi, int dtype, const int64_t strides[4], int Z, int Y, int X,
                   int C, const int32_t* blobs, int n, double* sums, int32_t* counts,
                   void* work, void* stream);

/* number of kernels this library has launched in this process (bench.py's
 * gpu_launches).                                                              */
int64_t mmb_launch_count(void);

/* Optional per-kernel timing with CUDA events recorded on the launching stream
 * around every kernel (bench.py's roofline).  mmb_profile_enable(1) clears and
 * starts recording, (0) stops.  mmb_profile_collect synchronises the recorded
 * events and sums, per kernel kind, elapsed milliseconds, launch count and work
 * units (voxels; candidates for the prune kernels; pairs for the seam match).
 * Kinds: 0 to_float, 1 preprocess, 2 log_x, 3 log_y, 4 log_z, 5 localmax,
 * 6 prune_edges, 7 prune_resolve, 8 compact, 9 seam_match, 10 log_xy (the fused
 * x -> y sweep; log_x / log_y then only count the volumes it does not serve).     */
#define MMB_PROF_NKINDS 11
int mmb_profile_enable(int on);
int mmb_profile_collect(double ms[MMB_PROF_NKINDS], int64_t launches[MMB_PROF_NKINDS],
                        double units[MMB_PROF_NKINDS]);

#ifdef __cplusplus
}
#endif
#endif /* MMB200_H */
