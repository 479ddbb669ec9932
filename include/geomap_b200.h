/*
 * geomap_b200 - C ABI of the B200-native tiled-detection hot path.
 *
 * Drop-in boundary for the data-parallel path of Abolfazlmsl/Oriented-Object-Detection
 * (large-map tiled OBB detection).  The reference has no FFI of its own: its boundary is
 * the set of module-level Python functions of Detect_OBB.py / Train_OBB.py.  Each entry
 * point below names the reference function (file:line) whose arithmetic it replaces; the
 * Python mirror in oriented_object_detection_b200/detect.py binds them with ctypes and
 * keeps the reference's function names, argument meaning and error behaviour
 * (INTEGRATION.md shows the binding a maintainer of the reference would add).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types in any signature.
 *   - "dev" pointers are CUDA device pointers owned by the caller; "host" pointers are
 *     ordinary host memory.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  Device entry points are asynchronous on `stream` unless stated.
 *   - every function returns an int status: GM_OK (0), a negative GM_E* for an invalid
 *     argument / too-small buffer, or a positive cudaError_t value.  Nothing throws.
 *   - nothing persistent is allocated by the device entry points: scratch memory is a
 *     caller-provided workspace sized by the matching *_workspace_bytes query.
 *   - thread-safe per stream (no global mutable state except the *_host helpers, which
 *     keep a per-device scratch arena guarded by a mutex).
 *
 * Detection record (reference docstring Detect_OBB.py:207-208, built at :256-262):
 *   (x1,y1,x2,y2,x3,y3,x4,y4, cls, conf, angle) in map pixels.  On device it is SoA:
 *   boxes double[n][8] (map coordinates: an fp32 network output plus an integer tile offset is exact in
 *   float64 and not in fp32), cls int32[n], conf float[n], angle double[n].
 */
#ifndef GEOMAP_B200_H
#define GEOMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_OK 0
#define GM_EINVAL (-1)     /* bad argument */
#define GM_ENOSPC (-2)     /* caller buffer / workspace too small */
#define GM_ERANGE (-3)     /* size outside the supported range (e.g. tile wider than GM_MAX_TILE) */
#define GM_ENODEV (-4)     /* no CUDA device / not an sm_100 device */

#define GM_MAX_SCALES 8    /* Gaussian scales in the DT-Edge gradient stack */
#define GM_MAX_RADIUS 15   /* ksize <= 31  (sigma <= 5.0) */
#define GM_MAX_TILE 1024   /* tile side supported by the per-warp chamfer scan */

/* One tile of the overlapped plan: rows [y0,y0+h) x cols [x0,x0+w) of the map; its pixels
 * start at element px_off of the packed tile batch (tiles are stored back to back, each
 * one contiguous, ragged edge tiles keep their true h x w). */
typedef struct gm_tile {
    int32_t y0, x0, h, w;
    int64_t px_off;
} gm_tile;

/* DT-Edge parameters = the config globals of Detect_OBB.py:29-32. */
#define GM_DTEDGE_GENERIC_GRAD 1   /* flags: use the generic gradient kernel even for the configured
                                     (0, 0.6, 1.2, 2.4) stack (parity tests of both kernels) */
#define GM_DTEDGE_OTSU 2           /* flags: DT_BIN_METHOD = "otsu" (Detect_OBB.py:109-111, Train_OBB.py:633-635):
                                     edges = Otsu threshold of the min-max normalised 8-bit gradient; p_hi unused */
typedef struct gm_dtedge_params {
    double  sigmas[GM_MAX_SCALES];    /* MS_SIGMAS; 0 = no blur */
    double  p_hi;                     /* DT_P_HI (DT_BIN_METHOD = "percentile", the default; ignored with GM_DTEDGE_OTSU) */
    int32_t n_sigmas;                 /* len(MS_SIGMAS), 1..GM_MAX_SCALES */
    int32_t morph_open;               /* DT_MORPH_OPEN iterations, 0..8 (n erosions then n dilations, like cv2) */
    int32_t layout;                   /* 0 = HWC [h][w][4] (Detect), 1 = CHW [4][h][w] (Train) */
    int32_t flags;                    /* GM_DTEDGE_* bits; 0 = the reference's default configuration */
} gm_dtedge_params;

/* ---- library ------------------------------------------------------------------------ */
int         gm_version(void);                 /* 100*major + minor */
const char* gm_status_string(int status);
int         gm_device_check(void);            /* GM_OK iff the current device is sm_100 */
int64_t     gm_launch_count(void);            /* kernels launched by this library so far (process-wide) */

/* ---- a1: tile plan  (Detect_OBB.py:210-223; ragged tiles kept; row-major) -------------- */
/* Number of tiles; optional outputs: grid rows/cols and the total pixel count of all tiles. */
int64_t gm_tile_plan_count(int32_t H, int32_t W, int32_t tile_size, int32_t overlap,
                           int32_t* rows, int32_t* cols, int64_t* total_px);
/* Fill `tiles_host[cap]`; tile rows [row_begin,row_end) only (row band of one rank; pass
 * 0,-1 for all).  px_off restarts at 0 for the first tile written.  Returns the count. */
int64_t gm_tile_plan_fill(int32_t H, int32_t W, int32_t tile_size, int32_t overlap,
                          int32_t row_begin, int32_t row_end,
                          gm_tile* tiles_host, int64_t cap, int64_t* total_px);

/* ---- a2: 3-channel tile gather  (build_multich 3-ch branch, Detect_OBB.py:92-93) ------- */
/* map_dev: uint8 [H][W][3] BGR.  out_dev: packed tiles, tile t at byte 3*px_off, [h][w][3]. */
int gm_tile_gather_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                      const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                      uint8_t* out_dev, void* stream);

/* ---- a3: 4-channel [R,G,B,DT-Edge] build  (Detect_OBB.py:95-133, Train_OBB.py:615-664) -- */
size_t gm_dtedge_workspace_bytes(int64_t total_px, int32_t n_tiles);
/* out_dev: packed tiles, tile t at byte 4*px_off, HWC or CHW per params->layout. */
int gm_dtedge_build_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                       const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                       int64_t total_px, const gm_dtedge_params* params,
                       uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                       void* stream);
/* Tiles [tile_begin, tile_begin + tile_count) of the same plan only (the other arguments describe the
 * WHOLE plan; every tile owns disjoint slices of the workspace and of out_dev).  This is what lets the
 * host upload a map in row chunks and build the tile rows whose pixels have arrived while the next
 * chunk is still in flight (ranges may be issued on different streams). */
int gm_dtedge_build_range_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                             const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                             int64_t total_px, int32_t tile_begin, int32_t tile_count,
                             const gm_dtedge_params* params,
                             uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                             void* stream);
/* Same build, with a CUDA event between the six kernels (grad, select_grad, edge_open, chamfer,
 * select_dist, tail); synchronises and returns the stage durations in ms (bench.py roofline). */
#define GM_DTEDGE_STAGES 6
int gm_dtedge_build_timed(const uint8_t* map_dev, int32_t H, int32_t W,
                          const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                          int64_t total_px, const gm_dtedge_params* params,
                          uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                          void* stream, float* stage_ms_host /* [GM_DTEDGE_STAGES] */);
/* Selection statistics since the last reset (device-global counters of the current device): how many
 * per-tile order-statistic selections were decided by the sampled bracket and how many fell back to the
 * two-pass method: {k_select_grad proven, fell back, k_select_dist proven, fell back, list overflow, rank
 * outside bracket 0, outside bracket 1, bracket too wide}.  Both paths return the exact order statistics
 * (np.percentile inputs, Detect_OBB.py:113-114, :128); tuning / tests only. */
int gm_dtedge_select_stats(uint32_t* counts8_host, int32_t reset);
/* Debug/parity taps into the workspace of the last build on it (device pointers, valid
 * until the workspace is reused): S = max_s(gx^2+gy^2) uint32[total_px]; chamfer field
 * uint32[total_px] (16.16 fixed point); zero mask (opened edges), bit-packed: row y of tile
 * t (index ti) occupies ceil(w/32) uint32 words starting at word
 * (px_off >> 5) + ti * (GM_MAX_TILE + 1) + y * ceil(w/32), bit x&31 of word x>>5. */
int gm_dtedge_workspace_views(void* workspace_dev, int64_t total_px, int32_t n_tiles,
                              uint32_t** S_dev, uint32_t** zero_bits_dev, uint32_t** chamfer_dev);

/* ---- a10: rotated IoU  (compute_polygon_iou, Detect_OBB.py:144-154) -------------------- */
/* boxes: double[n][8] corner lists (the reference's coordinates are Python floats; a tile
 * offset added to an fp32 network output is exact in float64 but not in fp32).  Arithmetic is
 * pair-local fp32 for convex quads; a concave SIMPLE quad (valid for shapely) sends the pair through the float64
 * piecewise clip.  Invalid quads (zero area, self-intersecting ring) give 0. */
int gm_rotated_iou_pairs(const double* boxes_a_dev, const double* boxes_b_dev,
                         const int32_t* idx_a_dev, const int32_t* idx_b_dev, int64_t n_pairs,
                         float* iou_dev, void* stream);
/* The same pair list in float64 - the value the reference itself computes (shapely on Python floats,
 * Detect_OBB.py:148-154), concave simple quads included: iou_dev double[n_pairs]. */
int gm_rotated_iou_pairs_f64(const double* boxes_a_dev, const double* boxes_b_dev,
                             const int32_t* idx_a_dev, const int32_t* idx_b_dev, int64_t n_pairs,
                             double* iou_dev, void* stream);
/* Dense n x m matrix, no early-out (the roofline kernel): iou_dev float[n][m].  Convex quads only: a pair with a
 * concave quad reads 0 here (use gm_rotated_iou_pairs or gm_polygon_iou_host for such boxes).
 * The boxes are prepared once per call into stream-ordered scratch (cudaMallocAsync / cudaFreeAsync on `stream`,
 * 64 B per row box + 160 B per column box, from the current device's default memory pool): the stream must belong
 * to the current device. */
int gm_rotated_iou_matrix(const double* boxes_a_dev, int32_t n, const double* boxes_b_dev, int32_t m,
                          float* iou_dev, void* stream);
/* Same arithmetic, each COLUMN (box b_j against every a_i) reduced to a checksum instead of stored
 * (pure-FP32 throughput measurement: no n*m store traffic): col_sum[j] = sum_i IoU(a_i, b_j). */
int gm_rotated_iou_matrix_sum(const double* boxes_a_dev, int32_t n, const double* boxes_b_dev, int32_t m,
                              double* col_sum_dev /* [m] */, void* stream);

/* ---- a5: post-network decode of one batch of tiles (Ultralytics OBB predictor tail) ------ */
/* head_dev: float [n_tiles][4+nc+1][A] = (cx,cy,w,h, cls probs..., theta) in network-input
 * pixels.  Per tile: conf filter (best class prob > conf_thr), conf-desc
 * order, probiou fast-NMS(iou_probiou), first max_det, regularise, undo the letterbox of an
 * (h x w) tile into the net_h x net_w network input (scale_boxes: gain = min(net_h/h, net_w/w)), corners.  Output slots: tile t owns [t*max_det, t*max_det+count[t]);
 * tile-local corners float[.][8] exactly as `results[0].obb.xyxyxyxy` would hold them. */
size_t gm_decode_workspace_bytes(int32_t n_tiles, int32_t n_anchors);
int gm_decode_tiles(const float* head_dev, int32_t n_tiles, int32_t n_classes, int32_t n_anchors,
                    const gm_tile* tiles_dev, int32_t net_h, int32_t net_w,
                    float conf_thr, float iou_probiou, int32_t max_det,
                    float* boxes_local_dev, int32_t* cls_dev, float* conf_dev, int32_t* count_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- f1: predictor pre-processing of a batch of equally sized tiles (Ultralytics LetterBox + preprocess,
 *          reached through model(net_input, conf=...), Detect_OBB.py:76-85) ---------------------- */
/* LetterBox(net_size, auto=auto_rect, stride, scaleup=True) geometry of a (tile_h x tile_w) tile. */
int gm_letterbox_shape(int32_t tile_h, int32_t tile_w, int32_t net_size, int32_t stride, int32_t auto_rect,
                       int32_t* new_h, int32_t* new_w, int32_t* top, int32_t* left, int32_t* out_h, int32_t* out_w);
/* tiles_dev: n_tiles records of the packed batch `packed_dev` (channels = 3 BGR or 4 RGB+DT, tile t at byte
 * channels*px_off), all of size tile_h x tile_w.  out_dev: float32 [n_tiles][channels][out_h][out_w]:
 * cv2 INTER_LINEAR resize (bit-exact) if the letterbox changes the size, 114 padding, BGR->RGB for 3
 * channels, / 255. */
int gm_letterbox_tiles(const uint8_t* packed_dev, int32_t channels, const gm_tile* tiles_dev, int32_t n_tiles,
                       int32_t tile_h, int32_t tile_w, int32_t net_size, int32_t stride, int32_t auto_rect,
                       float* out_dev, void* stream);

/* ---- a6-a9 + per-tile a11  (detect_symbols body, Detect_OBB.py:228-264) ----------------- */
/* In: tile-local corners float[n][8], cls, conf, tile_id[n] (non-decreasing; e.g. the slots of
 * gm_decode_tiles compacted, or any per-tile detector output).  Steps: + (x0,y0) of the tile
 * (exact, float64), border filter on the box centre (margin_px <= c <= dim - margin_px on both
 * axes of the ragged tile; skipped if margin_px <= 0), strike angle (degrees, float64) for
 * class `angle_class` else 0, exact greedy rotated NMS(iou_merge) inside each tile.
 * Out (compacted, tiles in order, survivors of a tile in stable confidence-descending order):
 * boxes double[.][8] map coordinates, cls, conf, angle double, src = input index,
 * count int64[1] (negative: -(pairs needed) when edge_capacity was too small; rerun). */
size_t gm_tile_postprocess_workspace_bytes(int64_t n, int64_t edge_capacity);
/* The same call with the caller's bound on the detections of ONE tile (a detector's max_det; Ultralytics: 300).  With
 * 0 < max_per_tile <= 320 the per-tile NMS runs as one CTA per tile - boxes ranked by confidence, one `IoU >= thr` bit per
 * same-class pair, a score-ordered sweep of the bit masks (the greedy rule of merge_detections, Detect_OBB.py:183-198) -
 * instead of the general grid / sort / edge-list engine; same outputs, same float64 threshold decisions.  The bound is a
 * promise about speed, not correctness: a tile that exceeds it is still resolved exactly, only slowly.  max_per_tile = 0
 * is gm_tile_postprocess.  Same workspace. */
int gm_tile_postprocess_bounded(const float* boxes_local_dev, const int32_t* cls_dev, const float* conf_dev,
                                const int32_t* tile_id_dev, int64_t n,
                                const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_class,
                                int32_t margin_px, int32_t angle_class, double iou_merge, int64_t edge_capacity,
                                int32_t max_per_tile,
                                double* out_boxes_dev, int32_t* out_cls_dev, float* out_conf_dev,
                                double* out_angle_dev, int32_t* out_src_dev, int64_t* out_count_dev,
                                void* workspace_dev, size_t workspace_bytes, void* stream);
int gm_tile_postprocess(const float* boxes_local_dev, const int32_t* cls_dev, const float* conf_dev,
                        const int32_t* tile_id_dev, int64_t n,
                        const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_class,
                        int32_t margin_px, int32_t angle_class, double iou_merge, int64_t edge_capacity,
                        double* out_boxes_dev, int32_t* out_cls_dev, float* out_conf_dev,
                        double* out_angle_dev, int32_t* out_src_dev, int64_t* out_count_dev,
                        void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- a11: class-wise exact greedy rotated NMS  (merge_detections, Detect_OBB.py:176-200) */
/* Keep-set identical to the sequential reference: stable confidence-descending order, a box
 * survives iff no already-kept same-class box has IoU >= iou_thr (float64 decision).
 * cls in [0, max_class].  Outputs: order_dev int32[n] = stable conf-desc permutation of the
 * input (what the reference's in-place sort leaves in the caller's list); keep_dev uint8[n]
 * by INPUT index; kept_idx_dev int32[<=n] = kept input indices in output order;
 * n_kept_dev int64[1] (negative: -(pairs needed) when edge_capacity was too small; rerun).
 * edge_capacity = max overlapping same-class pairs with IoU >= thr held (0 -> 16 n + 1024). */
size_t gm_nms_workspace_bytes(int64_t n, int64_t edge_capacity);
int gm_nms_global(const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev, int64_t n,
                  int32_t max_class, double iou_thr, int64_t edge_capacity,
                  int32_t* order_dev, uint8_t* keep_dev, int32_t* kept_idx_dev, int64_t* n_kept_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);

/* Threshold-adjacent pairs since the last reset (device-global counters of the current device; synchronises).  Every
 * `IoU >= thr` decision of gm_nms_global / gm_tile_postprocess / gm_fuse_scales is taken the reference's way, on a
 * float64 IoU (Detect_OBB.py:193, :394): the fp32 IoU decides unless it lies within 1e-4 of the threshold, in which
 * case the pair is recomputed in float64.  counts2_host[0] = pairs that took the float64 path, [1] = those of them whose
 * float64 IoU lies within 1e-5 of the threshold - the pairs a different polygon library could decide differently.
 * Counted per pair-discovery launch (a call repeated after an edge-capacity overflow counts its pairs again). */
int gm_threshold_adjacent_stats(uint64_t* counts2_host, int32_t reset);

/* ---- a12: dual-scale late fusion  (cross_scale_consensus_filter, Detect_OBB.py:347-423) - */
/* Detections of all scales concatenated in ascending-scale order, list order inside a scale;
 * scale_id int32[n] non-decreasing.  n_scales == 1 is the reference's passthrough.  Output:
 * kept input indices in the reference's output order; n_kept as above. */
size_t gm_fuse_workspace_bytes(int64_t n, int64_t edge_capacity);
int gm_fuse_scales(const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev,
                   const int32_t* scale_id_dev, int64_t n, int32_t n_scales, int32_t max_class,
                   double iou_partner, double conf_low, double conf_high, int64_t edge_capacity,
                   int32_t* kept_idx_dev, int64_t* n_kept_dev,
                   void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- f2: evaluation matching  (Detect_OBB.py:456-480, :512-565, :574-607, :609-648) -------- */
/* Detections and ground truths are grouped into independent segments (one image, or one image and
 * class): segment s holds detections [det_off[s], det_off[s+1]) in PROCESSING ORDER (list order for
 * _match_dets_to_gts_pixel, stable score-descending for compute_pr_for_class) and GTs
 * [gt_off[s], gt_off[s+1]); *_off are int64[n_segments + 1] on the device.  Boxes are double[.][8].
 * mat_off[s] = sum over earlier segments of n_det_s * n_gt_s; n_pairs = mat_off[n_segments]. */
/* iou[mat_off[s] + i * n_gt_s + j] = float64 IoU(det i, gt j) of segment s; -1 when both class arrays are
 * given and the classes differ (such a pair is never a candidate, Detect_OBB.py:468). */
int gm_eval_iou_segments(const double* det_dev, const int32_t* det_cls_dev, const double* gt_dev,
                         const int32_t* gt_cls_dev, const int64_t* det_off_dev, const int64_t* gt_off_dev,
                         const int64_t* mat_off_dev, int32_t n_segments, int64_t n_pairs,
                         double* iou_dev, void* stream);
/* The reference's sequential rule, replayed for n_thr IoU thresholds on the same IoU values: a detection
 * takes the unused GT of strictly largest IoU > 0 (first on ties) and is matched iff that IoU >= thr.
 * match[t * n_det + i] = segment-local GT index or -1.  used_scratch: uint8[n_thr * n_gt]. */
int gm_eval_match_greedy(const double* iou_dev, const int64_t* det_off_dev, const int64_t* gt_off_dev,
                         const int64_t* mat_off_dev, int32_t n_segments, int64_t n_det, int64_t n_gt,
                         const double* thr_dev, int32_t n_thr, uint8_t* used_scratch_dev,
                         int32_t* match_dev, void* stream);
/* evaluate_center_hit: match[i] = first unused same-class VALID GT polygon (convex, non-zero area) that
 * strictly contains the centre (mean of the 4 corners) of detection i, or -1.  used_scratch: uint8[n_gt]. */
int gm_eval_center_hit(const double* det_dev, const int32_t* det_cls_dev, const double* gt_dev,
                       const int32_t* gt_cls_dev, const int64_t* det_off_dev, const int64_t* gt_off_dev,
                       int32_t n_segments, uint8_t* used_scratch_dev, int32_t* match_dev, void* stream);

/* ---- (e) cross-band exchange records (multi-GPU merge; no counterpart in the single-process reference) ----
 * The global merge_detections (Detect_OBB.py:291) runs over the survivors of ALL row bands.  A rank ships its
 * survivors as `capacity` fixed-size 80-byte records {corners double[8], class int32, confidence float32, strike
 * angle double}; rows at or beyond *count_dev are blank (NaN corners, class -1, confidence -inf), so no rank ever
 * needs another rank's count on the host.  angle_dev may be NULL (angles read as 0). */
#define GM_BAND_RECORD_BYTES 80
int gm_band_pack(const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev, const double* angle_dev,
                 int64_t n_rows /* length of the input arrays */, const int64_t* count_dev, int64_t capacity,
                 uint8_t* records_dev /* [capacity][80] */, void* stream);
/* Records of all ranks (the all_gather result, total = world * capacity rows) -> SoA arrays.  cls_owned is the
 * class for rows this rank resolves (class % world == rank) and -1 for all others (classes are independent in
 * merge_detections, Detect_OBB.py:193); *n_valid_dev = number of non-blank rows. */
int gm_band_unpack(const uint8_t* records_dev, int64_t total, int32_t world, int32_t rank, double* boxes_dev,
                   int32_t* cls_dev, int32_t* cls_owned_dev, float* conf_dev, double* angle_dev, int64_t* n_valid_dev,
                   void* stream);
/* keep[i] = 0 where cls_owned[i] < 0 (before the keep flags of all ranks are OR-ed together). */
int gm_band_mask_keep(uint8_t* keep_dev, const int32_t* cls_owned_dev, int64_t total, void* stream);
/* Kept rows, in the stable confidence order `order_dev` (gm_nms_global), compacted into the output arrays (each
 * `total` rows long; the first *n_out_dev are written).  out_index = the row's position in the gathered list. */
size_t gm_band_extract_workspace_bytes(int64_t total);
int gm_band_extract(const int32_t* order_dev, const uint8_t* keep_dev, int64_t total, const double* boxes_dev,
                    const int32_t* cls_dev, const float* conf_dev, const double* angle_dev, double* out_boxes_dev,
                    int32_t* out_cls_dev, float* out_conf_dev, double* out_angle_dev, int64_t* out_index_dev,
                    int64_t* n_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- (e) seam-band exchange: the cross-band merge with per-rank work independent of the number of ranks ------------
 * The global merge_detections (Detect_OBB.py:291, :176-200) is class-wise greedy NMS over the concatenated list of all
 * bands.  Its overlap graph (same class, IoU >= thr) splits into components; a component whose boxes all lie in one band
 * is resolved by that band's rank alone.  gm_band_merge_local runs the exact NMS of the rank's OWN survivors with one
 * extension: a box that may overlap a box of another band (its AABB grown by extent_bound reaches a rectangle of
 * foreign box centres), or that has a higher-priority undecided-here neighbour, is DEFERRED unless a locally kept
 * higher-priority box already suppresses it.  Every box it decides has the single-rank verdict.  The deferred boxes -
 * the seam band - are written as fixed-size records for ONE all_gather; gm_band_merge_finish resolves the gathered seam
 * boxes of all ranks (the same engine; identical on every rank, list order = rank order), applies the verdicts to the
 * rank's own rows and compacts the rank's kept records in stable confidence order.  The union of the ranks' outputs,
 * merged by (confidence desc, rank, local order), is the single-rank result - members and order.
 *   boxes/cls/conf: n_rows rows of which the first *count_dev are valid (gm_tile_postprocess output); rows beyond are
 *     blanked IN PLACE.  cls in [0, max_class] (batched maps: map * n_classes + class).
 *   foreign_rects_host: n_rects (<= 8) closed rectangles {x0, y0, x1, y1} covering every position a box centre of another
 *     rank can take (safe regions of the foreign tiles, Detect_OBB.py:167-174).  extent_bound: an upper bound, over the
 *     boxes of ALL ranks, of the Chebyshev distance of a corner from its box's centre (the mean of the four corners,
 *     the point the border filter tests); each rank checks its own boxes (status bit below; header word 2 = its max).
 *   seam_records_dev: uint8 [(seam_capacity + 1)][GM_BAND_RECORD_BYTES]; row 0 = header, identical layout on all ranks.
 *   meta_dev int64[4] = {kept rows of this rank, status bits OR-ed over all ranks, seam rows of all ranks, survivors of
 *     all ranks}.  A non-zero status means the result must not be used: rerun with the named capacity / bound raised. */
#define GM_SEAM_EDGE_OVERFLOW 1ULL       /* more overlapping pairs than edge_capacity (local or seam phase) */
#define GM_SEAM_CAPACITY_OVERFLOW 4ULL   /* a rank deferred more boxes than seam_capacity */
#define GM_SEAM_EXTENT_EXCEEDED 8ULL     /* a box reaches farther than extent_bound from its centre: candidates may have been missed */
#define GM_SEAM_INPUT_OVERFLOW 16ULL     /* *count_dev was negative (the per-tile stage overflowed its pair buffer) */
#define GM_SEAM_CHAIN_ESCAPES 32ULL      /* gm_band_merge_finish on a RESTRICTED rank range: a chain of overlaps reaches a rank left
                                            out, so a box of this rank has no verdict - call it again on all ranks (block_count 0);
                                            local to the rank: the gathered records are all it needs, no collective */
size_t gm_band_merge_workspace_bytes(int64_t n_rows, int32_t world, int64_t seam_capacity, int64_t edge_capacity);
int gm_band_merge_local(double* boxes_dev, int32_t* cls_dev, float* conf_dev, int64_t n_rows,
                        const int64_t* count_dev, int32_t max_class, double iou_thr, int64_t edge_capacity,
                        const float* foreign_rects_host, int32_t n_rects, float extent_bound,
                        int32_t world, int64_t seam_capacity, uint8_t* seam_records_dev,
                        void* workspace_dev, size_t workspace_bytes, void* stream);
/* workspace_dev must be the one gm_band_merge_local used (it holds the local verdicts; the call may be repeated).
 * block_begin / block_count: resolve only the seam boxes of ranks [block_begin, block_begin + block_count) (must contain
 * `rank`; block_count <= 0 = all ranks).  With a restricted range, outside_rects_host are the rectangles of box centres of
 * the ranks LEFT OUT (same form as foreign_rects_host): a box of the resolved set that may overlap such a box is deferred
 * again, deferral propagates down its chains, and GM_SEAM_CHAIN_ESCAPES reports when it reaches a box of this rank. */
int gm_band_merge_finish(const uint8_t* gathered_dev /* [world][(seam_capacity + 1)][80] */, int32_t world, int32_t rank,
                         int64_t seam_capacity, int32_t block_begin, int32_t block_count,
                         const float* outside_rects_host, int32_t n_rects, float extent_bound, const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev,
                         const double* angle_dev /* may be NULL */, int64_t n_rows, int32_t max_class, double iou_thr,
                         int64_t edge_capacity, double* out_boxes_dev, int32_t* out_cls_dev, float* out_conf_dev,
                         double* out_angle_dev /* may be NULL */, int32_t* out_src_dev, int64_t* meta_dev,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- f3: label side of the training tilers  (Train_OBB.py:44-146, :290-428) ------------------ */
/* Full tiles only (the training tilers skip ragged edge tiles, Train_OBB.py:90-91): rows x cols tiles at
 * multiples of stride = tile_size - overlap; tile_id = row * cols + col is the reference's running counter.
 * span = ceil(tile_size / stride): a label's anchor falls into at most span^2 tiles.  Returns rows * cols. */
int64_t gm_train_tile_grid(int32_t H, int32_t W, int32_t tile_size, int32_t overlap, int32_t* rows, int32_t* cols,
                           int32_t* span);
/* labels_dev: double[n][8] corner coordinates in pixels.  For label i and candidate a in [0, span^2):
 * flag[i*span^2 + a] = 1 iff the label belongs to that tile (anchor = midpoint of corners 1 and 4 inside the
 * tile, bounding-box coverage >= cov_threshold); then tile_id[.] is the tile and coords[.][8] the label shifted
 * to the tile, clipped to [0, tile_size] and divided by tile_size.  Sorting the flagged entries by
 * (tile_id, i) gives every tile's label table in the reference's row order. */
int gm_train_label_tiles(const double* labels_dev, int64_t n, int32_t H, int32_t W, int32_t tile_size,
                         int32_t overlap, double cov_threshold, uint8_t* flag_dev, int32_t* tile_id_dev,
                         double* coords_dev, void* stream);

/* ---- host-buffer conveniences (synchronous; copies inside) ------------------------------ */
/* build_multich(bgr, out_channels) on one host crop (Detect_OBB.py:87-133). */
int gm_build_multich_host(const uint8_t* bgr_host, int32_t h, int32_t w, int32_t out_channels,
                          const gm_dtedge_params* params, uint8_t* out_host);
/* compute_polygon_iou(box1, box2) (Detect_OBB.py:144-154), float64 arithmetic on the device. */
int gm_polygon_iou_host(const double* box1_host, const double* box2_host, double* iou_host);

/* FP32 FFMA peak micro-benchmark (roofline denominator for the IoU kernel): runs
 * `iters` dependent-chain-free FFMA per thread on a full grid, returns TFLOP/s. */
int gm_ffma_peak(int32_t iters, double* tflops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOMAP_B200_H */
