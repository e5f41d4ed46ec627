/* cm3d_b200 - C ABI of the B200-native 2D-mask -> 3D lifting path.
 *
 * The reference (meharkhurana03/cm3d) has no FFI: its hot path is the body of the
 * frame loop / mask loop in src/{nuscenes,kitti,waymo}/2d_to_3d.py.  These entry
 * points replace the regions the reference brackets with its own stopwatches
 * ("io" tail, "points in mask", "medoid"), batched over many frames:
 *
 *   cm3d_masks_*            src/nuscenes/2d_to_3d.py:425-428 (pycocotools RLE decode), :526-527,
 *                           :543-544 (3x3 erosion, bool (W,H) view)
 *   cm3d_aggregate_sweeps   src/nuscenes/2d_to_3d.py:437-465 (close-point removal,
 *                           rotate/translate per sweep, hstack);
 *                           src/kitti/2d_to_3d.py:1066-1083; src/waymo/2d_to_3d.py:472-486
 *   cm3d_project_membership src/nuscenes/2d_to_3d.py:553-617 (clone, global->camera chain,
 *                           view_points, bounds/depth test, floor, mask lookup);
 *                           src/kitti/2d_to_3d.py:1238-1351; src/waymo/2d_to_3d.py:557-616
 *   cm3d_scan_segments      (no reference counterpart: sizes of the boolean-index results)
 *   cm3d_compact_segments   src/nuscenes/2d_to_3d.py:617-620 (track_points, gather)
 *   cm3d_medoid             src/nuscenes/2d_to_3d.py:116-119,641-663 (cdist medoid, centroid)
 *   cm3d_hull_obb           src/kitti/2d_to_3d.py:855-876,1481-1484,1524 (open3d OBB of the hull vertices -> yaw)
 *   cm3d_nearest_lane       src/nuscenes/2d_to_3d.py:277-302 (closest lane point per centroid)
 *
 * Conventions: every pointer is a DEVICE pointer; the caller (PyTorch) owns and
 * sizes every buffer; nothing is allocated, nothing throws; all launches are
 * asynchronous on `stream` (a cudaStream_t passed as void*); return 0 on success,
 * a negative CM3D_E* code on bad arguments, or -(1000 + cudaError_t) when a launch
 * fails.  One process per GPU.
 *
 * Descriptor tables are plain int32 arrays (floats stored by bit pattern) built on
 * the host by cm3d_b200/batch.py; their layouts are fixed by the enums below.
 *
 * Geometry of a batch:
 *   raw points   one float buffer; sweep s starts at a 4-float-aligned offset, is
 *                npts*stride floats long and padded to a multiple of 4 floats.
 *   tiles        every sweep is cut into tiles of CM3D_TILE raw points; tiles of a
 *                frame are contiguous and in sweep order.  Tile t owns slots
 *                [t*CM3D_TILE, t*CM3D_TILE + tile_cnt[t]) of the aggregated cloud
 *                (survivors of the close-point filter, original order).  The index
 *                of a point in the reference's `aggr_pc_points` is
 *                tile_prefix[t] + (slot - t*CM3D_TILE).
 *   instances    numbered frame-major over the batch; instance i owns
 *                [seg_off[i], seg_off[i+1]) of seg_point_idx / seg_xyzw.
 */
#ifndef CM3D_B200_H
#define CM3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM3D_ABI_VERSION 13
#define CM3D_TILE 1024          /* points per tile: compaction / count granule */
#define CM3D_MAX_INST 254       /* instances per frame (hit ids are one byte, 0 = none, 255 = overflow) */
#define CM3D_MAX_VCAMS 16       /* (camera, mask size) combinations per frame */
#define CM3D_MEDOID_COLS 256    /* columns of the distance matrix per full medoid work item */

enum {
    CM3D_OK = 0,
    CM3D_EINVAL = -1,           /* bad argument */
    CM3D_ELIMIT = -2            /* a per-frame limit above is exceeded */
};

/* transform chain: CM3D_MAX_CHAIN ops of CM3D_OP_WORDS words: kind, 12 floats, 3 pad */
enum { CM3D_OP_END = 0, CM3D_OP_T = 1, CM3D_OP_R = 2, CM3D_OP_A = 3 };
#define CM3D_MAX_CHAIN 4
#define CM3D_OP_WORDS 16
#define CM3D_CHAIN_WORDS (CM3D_MAX_CHAIN * CM3D_OP_WORDS)

/* sweep_desc[s][CM3D_SW_WORDS] */
enum { CM3D_SW_RAW_LO = 0, CM3D_SW_RAW_HI, CM3D_SW_NPTS, CM3D_SW_STRIDE, CM3D_SW_FRAME,
       CM3D_SW_TILE_BASE, CM3D_SW_CHAIN, CM3D_SW_FOURTH, CM3D_SW_WORDS };
/* frame_desc[f][CM3D_FR_WORDS] */
enum { CM3D_FR_TILE_BEGIN = 0, CM3D_FR_TILE_END, CM3D_FR_VCAM_BEGIN, CM3D_FR_NVCAMS,
       CM3D_FR_INST_BEGIN, CM3D_FR_NINST, CM3D_FR_CLOSE_BITS, CM3D_FR_USE_CLOSE,
       CM3D_FR_MIN_DEPTH_BITS, CM3D_FR_CNT_OFF, CM3D_FR_MIN_MEDOID_PTS, CM3D_FR_LIST_BEGIN,
       CM3D_FR_TREF = 12 /* 3 floats: q = p + tref is the point the cull planes are evaluated on */,
       CM3D_FR_CHAIN_SIG = 15 /* sum kind_k * 4^k when all vcams of the frame share it, else -1 */,
       CM3D_FR_FLOOR_BITS = 16, CM3D_FR_USE_FLOOR = 17 /* default-off ground threshold: keep z > floor */,
       CM3D_FR_WORDS = 20 };
/* vcam_desc[v][CM3D_VC_WORDS]: one per (camera, mask size) of a frame */
enum { CM3D_VC_CHAIN = 0, CM3D_VC_VIEWPAD = 1 /* 12 floats */, CM3D_VC_W = 13, CM3D_VC_H,
       CM3D_VC_LIST_BEGIN /* into cam_inst_list, relative to the frame's LIST_BEGIN */,
       CM3D_VC_LIST_COUNT, CM3D_VC_FRAME, CM3D_VC_GRID_OFF /* word offset into vcam_grid */,
       CM3D_VC_GRID_NX /* cells per row, ceil(W/CM3D_CELL) */,
       CM3D_VC_PLANES = 20 /* 5 x (nx,ny,nz,d) floats: conservative frustum planes, see below */,
       CM3D_VC_FLAGS = 40 /* bit 0: viewpad is [[fx,0,cx,0],[0,fy,cy,0],[0,0,1,0]] */,
       CM3D_VC_WORDS = 44 };
/* Cull planes (host-built in fp64 by cm3d_b200/batch.py, optional: (0,0,0,1) x 5 disables them).
 * For q = fl(p + tref) and S = |qx|+|qy|+|qz|, a point whose plane value n.q + d is below
 * -S * 2^-17 for ANY of the five planes (depth, left, right, top, bottom) provably fails the
 * reference's depth / image-bounds test in that vcam, whatever the rounding of the exact chain;
 * only the survivors run the exact fp32 chain.  The planes never change a result. */
#define CM3D_CELL 32            /* pixels per side of a vcam_grid cell */
/* inst_desc[i][CM3D_IN_WORDS] */
enum { CM3D_IN_BITS_LO = 0, CM3D_IN_BITS_HI, CM3D_IN_W, CM3D_IN_H, CM3D_IN_PITCH, CM3D_IN_VCAM,
       CM3D_IN_FRAME, CM3D_IN_LOCAL, CM3D_IN_WORDS };
/* error flag words written by the kernels (errflags[CM3D_ERR_WORDS]) */
enum { CM3D_ERR_SEG_OVERFLOW = 0 /* total members needed when > seg_cap */,
       CM3D_ERR_RLE_SIZE = 1     /* 1 + first instance whose runs do not cover W*H */,
       CM3D_ERR_WORDS = 4 };

int cm3d_abi_version(void);
const char *cm3d_error_string(int code);

/* ---- masks: COCO runs or dense uint8 -> bit planes -> 3x3-eroded bit planes + bbox ------- */

/* Dense (H,W) uint8 masks -> bit planes (bit x&31 of word y*pitch + x/32 = mask[y][x] != 0).
 * src_off[i] = byte offset of instance i in `masks`. */
int cm3d_masks_pack_dense(const uint8_t *masks, const int64_t *src_off, const int32_t *inst_desc,
                          int n_inst, int max_words, uint32_t *bits, void *stream);

/* pycocotools compressed `counts` strings -> run lengths, on the device.  counts = the strings of
 * all instances back to back, byte_off[n_inst+1] their byte offsets.  Instance i's runs are
 * written to runs[byte_off[i] ..) and zero-padded to byte_off[i+1] (runs has as many uint32 as
 * counts has bytes), so byte_off doubles as the run_off of cm3d_masks_fill_rle. */
int cm3d_masks_decode_counts(const uint8_t *counts, const int64_t *byte_off, int n_inst,
                             uint32_t *runs, void *stream);

/* COCO run lengths (alternating 0-run,1-run; row-major over the (H,W) image) -> bit planes.
 * `bits` must be zero on entry.  run_start is scratch of the same length as runs.
 * row_range[2*i], [2*i+1] = first / last image row of instance i that holds a set pixel
 * ({INT_MAX,-1} when the mask is empty): lets the erosion skip the empty rows. */
int cm3d_masks_fill_rle(const uint32_t *runs, const int64_t *run_off, uint32_t *run_start,
                        const int32_t *inst_desc, int n_inst, int max_runs, uint32_t *bits,
                        int32_t *row_range, int32_t *errflags, void *stream);

/* cv2.erode(mask, ones(3,3)) on bit planes; bbox[i] = {xmin,ymin,xmax,ymax} of the eroded
 * set bits ({INT_MAX,INT_MAX,-1,-1} if none).  row_range (optional, from cm3d_masks_fill_rle; NULL =
 * look at every row): rows outside it are written as zeros without reading the input. */
int cm3d_masks_erode3x3(const uint32_t *bits_in, const int32_t *inst_desc, const int32_t *row_range,
                        int n_inst, int max_words, uint32_t *bits_out, int32_t *bbox, void *stream);

/* Instance lookup grid: for every vcam, cell (cx,cy) of CM3D_CELL^2 pixels holds
 * ceil(list_count/32) words whose bit k says "the eroded bbox of the vcam's k-th instance touches
 * this cell" (vcam_grid[GRID_OFF + (cy*NX + cx)*nwords + k/32]).  max_cells = largest NX*NY. */
int cm3d_build_vcam_grid(const int32_t *vcam_desc, int n_vcams, int max_cells, const int32_t *frame_desc,
                         const int32_t *cam_inst_list, const int32_t *inst_bbox, uint32_t *vcam_grid,
                         void *stream);

/* ---- sweeps -> aggregated cloud (tile-compacted SoA) ---------------------------------------- */

/* xyzw: 4 arrays of n_slots floats (x | y | z | 4th row), n_slots = n_tiles*CM3D_TILE.
 * tile_sweep[t] = sweep the tile belongs to. */
int cm3d_aggregate_sweeps(const float *raw, const int32_t *tile_sweep, int n_tiles,
                          const int32_t *sweep_desc, const int32_t *frame_desc,
                          const uint32_t *chains, float *xyzw, int32_t *tile_cnt, void *stream);

/* ---- projection + mask membership ------------------------------------------------------------- */

/* hits[slot]: up to four frame-local instance ids (+1) in ascending order, one per byte,
 * 0 = none; byte 3 == 255 means "more than four, recompute".
 * tile_inst_cnt[cnt_off(f) + j*ntiles(f) + (t - tile_begin(f))] = members of instance j in tile t.
 * pix (optional, may be NULL): pix[v*n_slots + slot] = fx | fy<<16 of the point in the
 * frame's v-th vcam, or -1 when it fails the depth / image-bounds test. */
int cm3d_project_membership(const float *xyzw, const int32_t *tile_cnt, const int32_t *tile_sweep,
                            int n_tiles, const int32_t *sweep_desc, const int32_t *frame_desc,
                            const int32_t *vcam_desc, const int32_t *cam_inst_list,
                            const int32_t *inst_desc, const int32_t *inst_bbox,
                            const uint32_t *chains, const uint32_t *bits, const uint32_t *vcam_grid,
                            uint32_t *hits, uint16_t *tile_inst_cnt, int32_t *pix, void *stream);

/* Scans.  tile_prefix[t] = index in the frame's aggr_pc_points of tile t's first point;
 * frame_n[f] = N of frame f; tile_inst_base = exclusive prefix of tile_inst_cnt over the
 * frame's tiles (same layout, int32); seg_off[n_inst_total+1] = exclusive prefix of the
 * per-instance member counts.  Medoid schedule: item_inst[n_inst_total] = the instances ordered
 * by work-item count, largest first; item_off[n_inst_total+1] = exclusive prefix of
 * cm3d_medoid_items(M, frame minimum) in that order.  medoid_best is reset.
 * Sets errflags[CM3D_ERR_SEG_OVERFLOW] when seg_off[end] > seg_cap. */
int cm3d_scan_segments(const int32_t *tile_cnt, const uint16_t *tile_inst_cnt,
                       const int32_t *frame_desc, int n_frames, int max_inst_per_frame,
                       int n_inst_total, const int32_t *inst_desc, int64_t seg_cap,
                       int32_t *tile_prefix, int32_t *frame_n, int32_t *tile_inst_base,
                       int32_t *seg_off, int32_t *item_off, int32_t *item_inst,
                       unsigned long long *medoid_best, int32_t *errflags, void *stream);

/* Ordered (ascending point index) per-instance index lists + gathered points.
 * seg_xyzw: 4 arrays of seg_cap floats.  Does nothing when the overflow flag is set. */
int cm3d_compact_segments(const float *xyzw, const int32_t *tile_cnt, const int32_t *tile_prefix,
                          const int32_t *tile_sweep, int n_tiles, const int32_t *sweep_desc,
                          const int32_t *frame_desc, const int32_t *vcam_desc,
                          const int32_t *cam_inst_list, const int32_t *inst_desc,
                          const int32_t *inst_bbox, const uint32_t *chains, const uint32_t *bits,
                          const uint32_t *vcam_grid, const uint32_t *hits, const int32_t *tile_inst_base,
                          const int32_t *seg_off, int32_t *seg_point_idx, float *seg_xyzw,
                          int64_t seg_cap, int max_inst_per_frame, const int32_t *errflags,
                          void *stream);

/* ---- medoid ---------------------------------------------------------------------------------- */

/* medoid_local[i] = argmin_j sum_k ||p_k - p_j|| with torch.cdist / sum(axis=0) arithmetic
 * (-1 if the instance has fewer than its frame's minimum points); medoid_point_idx = that
 * point's index in aggr_pc_points; centroid[4*i..] = its x,y,z,4th row (NaN when absent).
 * col_sums (optional, seg_cap floats) receives every column sum.  max_items bounds the grid:
 * it must be >= item_off[n_inst_total] (seg_cap/CM3D_MEDOID_COLS + 2*n_inst_total always is).
 * item_off / item_inst: the schedule written by cm3d_scan_segments.
 *
 * Screen + verify (same result, ~3x less work; csrc/medoid.cu has the error bounds): with
 * screen_sums (seg_cap floats), screen_min (5 * n_inst_total words) and screen_min_pts > 0 given and
 * col_sums NULL, instances with screen_min_pts <= M <= 2^19 points get approximate column sums
 * first (bit-identical squared distances, MUFU square root, flat accumulation) and only the
 * columns within the proven error bound of the approximate minimum are summed exactly, in the
 * reference's order; the argmin runs over those.  Instances whose squared norms share one binade
 * (nuScenes' global frame) have an exactly symmetric squared-distance matrix and are screened over
 * the pairs i <= j only (screen_flags bit 0 turns that off); with sym_ws (5 * seg_cap words of
 * scratch, optional) instances that straddle two binades are screened the same way on a copy
 * permuted by binade, pairs across the binades in both orders (bit 1 turns that off).  Smaller
 * instances, instances with
 * coordinates outside the fast square root's range and every call with col_sums take the all-exact
 * path.  screen_stats (optional, 1 word, zeroed by the caller): receives the number of verified
 * columns.  item_info (optional scratch, 4 * max_items words, 16-byte aligned): item -> {instance, index
 * inside it, segment offset, segment length}, so that the blocks of the item-grid launches find their work
 * with one 16-byte load instead of a search and three dependent loads. */
int cm3d_medoid(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off,
                const int32_t *seg_point_idx, const int32_t *item_off, const int32_t *item_inst,
                int n_inst_total,
                int max_items, unsigned long long *medoid_best, float *col_sums,
                float *screen_sums, uint32_t *screen_min, int screen_min_pts, int screen_flags,
                float *sym_ws, int32_t *screen_stats, int32_t *item_info, int32_t *medoid_local, int32_t *medoid_point_idx, float *centroid,
                const int32_t *errflags, void *stream);
#define CM3D_SCREEN_MIN_PTS 512   /* default screen_min_pts */

/* Work items of an instance with m member points (host helper; 0 below min_pts, 1 below 32 points,
 * else ceil(floor32(m)/CM3D_MEDOID_COLS) full items + one tail item when m % 32 != 0). */
int cm3d_medoid_items(int m, int min_pts);

/* ---- KITTI orientation (open3d is not in the reference tree: graded against the reference's get_depth_bbox run over a
 * Qhull-based stand-in) ------------------------------------------------------------------------- */

/* open3d's oriented bounding box of every instance with at least min_pts points and the reference's yaw
 * (src/kitti/2d_to_3d.py:855-876,1524): obb[16*i..] = yaw, centre xyz, wlh (after the reference's
 * axis shuffle), R' row-major (9 floats); NaN for skipped instances.  mode 0 = open3d 0.15's published
 * algorithm: PCA of the CONVEX-HULL VERTICES (found on the GPU by gift wrapping, csrc/obb.cu); a flat
 * cloud gets the reference's fallback box (first point, extent 1, identity -> yaw 0, kitti:1483-1484).
 * mode 1 = PCA of all member points (round 1's estimator; diagnostics only).  hull_ws: scratch of
 * cm3d_hull_obb_ws_words(seg_cap) int32 words (mode 0).  hull_info (optional, n_inst_total ints):
 * hull vertex count; -1 = flat / fallback; -(count+1) = wrapping closed with an inconsistent face
 * count (exactly degenerate input); 0 = skipped.  order (optional, n_inst_total ints): instance handled by
 * block b - cm3d_scan_segments' item_inst lists the instances largest first, which keeps the launch's tail short. */
int64_t cm3d_hull_obb_ws_words(int64_t seg_cap);
int cm3d_hull_obb(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off, int n_inst_total,
                  int min_pts, int mode, const int32_t *order, int32_t *hull_ws, int64_t hull_ws_words,
                  float *obb, int32_t *hull_info, const int32_t *errflags, void *stream);

/* ---- default-off extensions (north star (3)/(4); never executed by the reference: PARITY UNPINNED) ---- */

/* Per-instance outlier filter by neighbour counting: keep[p] = 1 when at least min_neighbors members
 * of the same instance (the point itself included) lie within `radius` of member p
 * (d2 = (dx*dx + dy*dy) + dz*dz in fp32 <= fl(radius*radius)).  kept_count[i] = survivors of
 * instance i; item_first[n_inst_total+1] is scratch; max_items >= sum ceil(M_i/256)
 * (seg_cap/256 + n_inst_total always is).  Stands where the reference has its dead
 * `clusters_hdbscan` (src/kitti/2d_to_3d.py:159-174). */
int cm3d_neighbor_filter(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off, int n_inst_total,
                         int max_items, float radius, int min_neighbors, int32_t *item_first,
                         uint8_t *keep, int32_t *kept_count, void *stream);

/* Segment offsets + medoid schedule from per-instance counts staged in seg_off[1..n_inst_total]
 * (pass kept_count = seg_off + 1 above): the batch-level half of cm3d_scan_segments. */
int cm3d_schedule_segments(const int32_t *frame_desc, const int32_t *inst_desc, int n_inst_total,
                           int64_t seg_cap, int32_t *seg_off, int32_t *item_off, int32_t *item_inst,
                           unsigned long long *medoid_best, int32_t *errflags, void *stream);

/* Ordered compaction of the kept points of every instance into new segments (offsets seg_off2). */
int cm3d_filter_segments(const float *seg_xyzw, const int32_t *seg_point_idx, int64_t seg_cap,
                         const int32_t *seg_off, const uint8_t *keep, const int32_t *seg_off2,
                         int n_inst_total, float *seg_xyzw2, int32_t *seg_point_idx2,
                         const int32_t *errflags, void *stream);

/* Block-level orientation / extent search: box[8*i..] = centre xyz, extent along the heading, across
 * it, up, heading in [0, pi/2) about `up_axis` (0/1/2), footprint area - the heading among
 * k*pi/(2*n_angles), k < n_angles, with the smallest axis-aligned footprint of the rotated points
 * (first minimum); NaN for instances with fewer than min_pts points. */
int cm3d_box_search(const float *seg_xyzw, int64_t seg_cap, const int32_t *seg_off, int n_inst_total,
                    int up_axis, int n_angles, int min_pts, float *box, const int32_t *errflags,
                    void *stream);

/* ---- pass 2: closest lane point --------------------------------------------------------------- */

/* For every centroid (n x 2 doubles) the nearest of m lane points (m x 2 doubles, 16-byte aligned):
 * idx_out[c] = argmin_k sqrt(dx*dx + dy*dy) in binary64 (first minimum, like numpy.argmin over
 * scipy's cdist row; -1 when m == 0), dist_out[c] = that distance.  Replaces
 * lane_yaws_distances_and_coords' cdist + argmin + min (src/nuscenes/2d_to_3d.py:277-302). */
int cm3d_nearest_lane(const double *centroids_xy, int n, const double *lane_xy, int m,
                      int32_t *idx_out, double *dist_out, void *stream);

/* Self-test: counts (adds to *mismatches) the floats in [2^-101, FLT_MAX] U {0} on which the
 * medoid kernel's branch-free square root differs from IEEE sqrt.rn.f32.  Must stay 0. */
int cm3d_selftest_sqrt(unsigned long long *mismatches, void *stream);

/* Self-test: *max_ulp (zeroed by the caller) = largest distance in ulps between the screen kernel's
 * approximate square root (MUFU.SQRT) and IEEE sqrt.rn.f32 over 0 and [2^-101, FLT_MAX].  The
 * screen error bound assumes <= 4; the GPU suite pins <= 2. */
int cm3d_selftest_sqrt_approx(unsigned *max_ulp, void *stream);

/* ---- the whole launch sequence of one batch in ONE call ------------------------------------------------------ */

/* What `Lifter.run` does call by call (masks -> grid -> aggregate -> project -> scan -> gather -> [hull boxes] -> medoid),
 * for callers whose batches are short enough that ~15 foreign-function calls and ~30 allocations per batch show
 * (one frame per call, Waymo's 16-frame batches).  Every pointer is device memory owned by the caller; sizes are
 * those the individual entry points document.  The call zeroes `out` (out_words words: the label block that holds
 * frame_n .. errflags, in any layout - only the field pointers are used) and bits_raw, then launches.
 * stream_medoid: NULL or == stream -> everything on `stream`; otherwise the medoid runs on stream_medoid behind an
 * event recorded after the gather (the hull boxes stay on `stream`, next to it) and stream_medoid finally waits for
 * the boxes, so that "stream_medoid is done" means "the batch is done".  Optional (may be NULL): row_range with
 * masks_kind 0, obb / hull_info / hull_ws when want_obb == 0, sym_ws, screen_stats (zeroed by the call), item_info,
 * screen_sums / screen_min when screen_min_pts == 0.  *launches (optional) receives the number of kernels launched. */
typedef struct cm3d_batch_args {
    int32_t n_frames, n_inst, n_tiles, n_vcams, max_cells, max_words, max_runs, max_inst_per_frame;
    int32_t masks_kind;                 /* 0 dense uint8 masks, 1 COCO run lengths, 2 COCO `counts` strings */
    int32_t want_obb, obb_mode, obb_min_pts;
    int32_t screen_min_pts, screen_flags, max_items, reserved;
    int64_t bits_words, seg_cap, hull_ws_words, out_words, mask_bytes;
    const float *raw;
    const int32_t *tile_sweep, *sweep_desc, *frame_desc, *vcam_desc, *cam_inst_list, *inst_desc;
    const uint32_t *chains;
    const uint8_t *mask;
    const int64_t *mask_off;
    int32_t *out, *frame_n, *seg_off, *item_off, *medoid_local, *medoid_point_idx;
    float *centroid;
    int32_t *errflags;
    uint32_t *runs, *run_start;
    int32_t *row_range;
    uint32_t *bits_raw, *bits;
    int32_t *bbox;
    uint32_t *vcam_grid;
    float *xyzw;
    int32_t *tile_cnt, *tile_prefix;
    uint32_t *hits;
    uint16_t *tile_inst_cnt;
    int32_t *tile_inst_base;
    unsigned long long *medoid_best;
    int32_t *item_inst, *seg_point_idx;
    float *seg_xyzw, *screen_sums;
    uint32_t *screen_min;
    float *sym_ws;
    int32_t *screen_stats, *item_info;
    float *obb;
    int32_t *hull_info, *hull_ws;
    void *stream, *stream_medoid;
    int32_t *launches;                  /* host pointer */
} cm3d_batch_args;

int cm3d_lift_batch(const cm3d_batch_args *a);
/* sizeof(cm3d_batch_args) as this library was compiled: a binding checks its own mirror of the struct against it. */
int cm3d_batch_args_size(void);

/* ---- host-side packer (no CUDA call inside; releases nothing, allocates nothing the caller sees) ---------- */

/* The batch as flat arrays (cm3d_b200/batch.py: pack_frames_native builds it from FrameSpecs).  Ops of all
 * chains back to back: the sweeps' chains in sweep order, then the cameras' chains in camera order
 * (frame-major); op_begin has n_sweeps + n_cams + 1 entries; pointers travel as 64-bit integers. */
typedef struct cm3d_pack_input {
    int32_t n_frames, n_sweeps, n_cams, n_inst;
    const int32_t *fr_n_sweeps, *fr_n_cams, *fr_n_inst, *fr_fourth, *fr_min_pts, *fr_use_close, *fr_use_floor;
    const float *fr_close, *fr_min_dist, *fr_floor;
    const uint64_t *sw_ptr;
    const int32_t *sw_npts, *sw_stride;
    const int32_t *op_begin, *op_kind;
    const uint64_t *op_ptr;
    const uint64_t *cam_K;
    const int32_t *in_cam, *in_W, *in_H;
    const int64_t *in_counts_off;
    const uint8_t *counts;
} cm3d_pack_input;

/* plan[16]: 0 n_tiles, 1 n_vcams, 2 n_chains, 3 raw floats, 4 meta words, 5 mask bytes, 6..12 word offsets of
 * tile_sweep, sweep_desc, frame_desc, vcam_desc, cam_inst_list, inst_desc, chains in meta, 13 max instances per
 * frame, 14 raw points.  CM3D_ELIMIT when a frame exceeds CM3D_MAX_INST / CM3D_MAX_VCAMS. */
int cm3d_pack_plan(const cm3d_pack_input *in, int64_t *plan);
/* sizeof(cm3d_pack_input) as compiled (for a binding's own mirror of the struct). */
int cm3d_pack_input_size(void);

/* Fills raw (plan[3] floats), meta (plan[4] words), mask (plan[5] bytes: the counts strings), mask_off
 * (n_inst + 1), vcam_keys (plan[1] x {frame, camera, W, H}) and out[8] = cnt_total, bits_words, max_words,
 * grid_words, max_cells, max_runs.  Same bytes as the Python packer (cull planes to fp64 rounding). */
int cm3d_pack_fill(const cm3d_pack_input *in, const int64_t *plan, float *raw, int32_t *meta, uint8_t *mask,
                   int64_t *mask_off, int32_t *vcam_keys, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* CM3D_B200_H */
