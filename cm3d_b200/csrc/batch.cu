// cm3d_lift_batch: the launch sequence of one batch behind ONE foreign-function call.
//
// `Lifter.run` (cm3d_b200/lifter.py) makes ~15 C-ABI calls and ~30 workspace allocations per batch; for a
// 64-frame nuScenes batch (15 ms of kernels) that is noise, for one frame per call or Waymo's 16-frame batches
// (0.5 - 1.3 ms of kernels) it is the bound.  This is the same sequence - the same entry points, in the same
// order, with the same arguments - driven from C over ONE caller-owned workspace.
//
// Replaces the per-frame body of the reference scripts (src/nuscenes/2d_to_3d.py:425-663,
// src/kitti/2d_to_3d.py:1001-1524, src/waymo/2d_to_3d.py:472-653) like the individual calls do.
#include "common.cuh"

#define CM3D_TRY(expr)                  \
    do {                                \
        const int rc__ = (expr);        \
        if (rc__ != CM3D_OK) return rc__; \
    } while (0)
#define CM3D_CUDA(expr)                                          \
    do {                                                         \
        const cudaError_t e__ = (expr);                          \
        if (e__ != cudaSuccess) return -(1000 + (int)e__);       \
    } while (0)

extern "C" int cm3d_batch_args_size(void) { return (int)sizeof(cm3d_batch_args); }

extern "C" int cm3d_lift_batch(const cm3d_batch_args *a)
{
    if (!a) return CM3D_EINVAL;
    if (a->n_frames < 0 || a->n_inst < 0 || a->n_tiles < 0 || a->seg_cap < 0 || a->out_words < 0) return CM3D_EINVAL;
    if (!a->out || !a->frame_n || !a->seg_off || !a->item_off || !a->medoid_local || !a->medoid_point_idx ||
        !a->centroid || !a->errflags)
        return CM3D_EINVAL;
    cudaStream_t st = (cudaStream_t)a->stream;
    cudaStream_t st2 = a->stream_medoid ? (cudaStream_t)a->stream_medoid : st;
    const int I = a->n_inst, T = a->n_tiles;
    int launches = 0;

    CM3D_CUDA(cudaMemsetAsync(a->out, 0, (size_t)a->out_words * 4, st));

    // ---- masks -> eroded bit planes (+ bbox) -> per-vcam instance lookup grid
    if (I) {
        const int32_t *row_range = nullptr;
        if (a->masks_kind == 0) {
            CM3D_TRY(cm3d_masks_pack_dense(a->mask, a->mask_off, a->inst_desc, I, a->max_words, a->bits_raw, st));
            launches += 1;
        } else {
            CM3D_CUDA(cudaMemsetAsync(a->bits_raw, 0, (size_t)a->bits_words * 4, st));
            const uint32_t *runs = reinterpret_cast<const uint32_t *>(a->mask);
            if (a->masks_kind == 2) {
                CM3D_TRY(cm3d_masks_decode_counts(a->mask, a->mask_off, I, a->runs, st));
                runs = a->runs;
                launches += 1;
            }
            CM3D_TRY(cm3d_masks_fill_rle(runs, a->mask_off, a->run_start, a->inst_desc, I, a->max_runs, a->bits_raw,
                                         a->row_range, a->errflags, st));
            row_range = a->row_range;
            launches += 2;
        }
        CM3D_TRY(cm3d_masks_erode3x3(a->bits_raw, a->inst_desc, row_range, I, a->max_words, a->bits, a->bbox, st));
        launches += 2;
        CM3D_TRY(cm3d_build_vcam_grid(a->vcam_desc, a->n_vcams, a->max_cells, a->frame_desc, a->cam_inst_list, a->bbox,
                                      a->vcam_grid, st));
        launches += 1;
    }

    // ---- sweeps -> aggregated cloud -> projection + membership -> scans -> ordered gather
    CM3D_TRY(cm3d_aggregate_sweeps(a->raw, a->tile_sweep, T, a->sweep_desc, a->frame_desc, a->chains, a->xyzw, a->tile_cnt, st));
    CM3D_TRY(cm3d_project_membership(a->xyzw, a->tile_cnt, a->tile_sweep, T, a->sweep_desc, a->frame_desc, a->vcam_desc,
                                     a->cam_inst_list, a->inst_desc, a->bbox, a->chains, a->bits, a->vcam_grid, a->hits,
                                     a->tile_inst_cnt, nullptr, st));
    launches += T ? 2 : 0;
    CM3D_TRY(cm3d_scan_segments(a->tile_cnt, a->tile_inst_cnt, a->frame_desc, a->n_frames, a->max_inst_per_frame, I,
                                a->inst_desc, a->seg_cap, a->tile_prefix, a->frame_n, a->tile_inst_base, a->seg_off,
                                a->item_off, a->item_inst, a->medoid_best, a->errflags, st));
    launches += 2;
    CM3D_TRY(cm3d_compact_segments(a->xyzw, a->tile_cnt, a->tile_prefix, a->tile_sweep, T, a->sweep_desc, a->frame_desc,
                                   a->vcam_desc, a->cam_inst_list, a->inst_desc, a->bbox, a->chains, a->bits, a->vcam_grid,
                                   a->hits, a->tile_inst_base, a->seg_off, a->seg_point_idx, a->seg_xyzw, a->seg_cap,
                                   a->max_inst_per_frame, a->errflags, st));
    launches += T ? 1 : 0;

    // ---- second phase on its own stream, behind the gather
    cudaEvent_t ev = nullptr;
    if (st2 != st) {
        CM3D_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st2, ev, 0);
        cudaEventDestroy(ev);                     // released once the recorded work has completed
        if (e != cudaSuccess) return -(1000 + (int)e);
    }

    // ---- KITTI: hull-vertex box + yaw; stays on the front stream, next to the medoid
    bool hull = false;
    if (a->want_obb && I) {
        if (!a->obb) return CM3D_EINVAL;
        CM3D_TRY(cm3d_hull_obb(a->seg_xyzw, a->seg_cap, a->seg_off, I, a->obb_min_pts, a->obb_mode, a->item_inst, a->hull_ws,
                               a->hull_ws_words, a->obb, a->hull_info, a->errflags, st));
        launches += 1;
        hull = true;
    }

    // ---- medoid (screen + verify where screen_min_pts > 0)
    if (I) {
        const bool screen = a->screen_min_pts > 0 && a->screen_sums && a->screen_min;
        if (screen && a->screen_stats) CM3D_CUDA(cudaMemsetAsync(a->screen_stats, 0, 4, st2));
        CM3D_TRY(cm3d_medoid(a->seg_xyzw, a->seg_cap, a->seg_off, a->seg_point_idx, a->item_off, a->item_inst, I, a->max_items,
                             a->medoid_best, nullptr, screen ? a->screen_sums : nullptr, screen ? a->screen_min : nullptr,
                             screen ? a->screen_min_pts : 0, a->screen_flags, screen ? a->sym_ws : nullptr,
                             screen ? a->screen_stats : nullptr, a->item_info, a->medoid_local, a->medoid_point_idx,
                             a->centroid, a->errflags, st2));
        // expand_items, k_medoid, finalize; + classify, screen, verify; + screen_sym, screen_min; + permute; + prune
        const bool ws = screen && a->sym_ws;
        launches += screen ? (((a->screen_flags & 1) ? 6 : ((a->screen_flags & 2) || !ws ? 8 : 9)) + (ws && !(a->screen_flags & 4) ? 1 : 0)) : 3;
    }
    if (hull && st2 != st) {                      // "stream_medoid is done" must cover the boxes too
        CM3D_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st2, ev, 0);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) return -(1000 + (int)e);
    }
    if (a->launches) *a->launches = launches;
    return CM3D_OK;
}
