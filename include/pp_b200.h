/*
 * pp_b200.h -- C ABI of libpp_b200.so: the B200 (sm_100a) implementation of the
 * PointPillars pre/post-processing hot path of michalp0lak/ObjectDetection_3D.
 *
 * Every entry point is what a binding of the reference's ops layer would call; the
 * reference interface each one replaces is cited as file:line (paths inside the
 * reference checkout).  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers unless the name ends in _host.
 *  - The caller owns every buffer (outputs and workspace); the library never allocates
 *    or retains device memory in the device-pointer entry points.  Workspace sizes come
 *    from the matching *_workspace_bytes function; workspaces need no initialisation.
 *  - Work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises
 *    unless stated.  Data-dependent sizes (pillar count, keep count) are returned through
 *    device scalars so that the caller decides when to synchronise.
 *  - Return value: 0 on success, negative PP_ERR_* otherwise; pp_last_error() returns a
 *    thread-local message.  No exception crosses the boundary.  The Python mirror turns
 *    PP_ERR_INVALID into ValueError/AssertionError like the reference's own checks
 *    (ops/ops_torch.py:555-562, 643-646).
 */
#ifndef PP_B200_H
#define PP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_OK 0
#define PP_ERR_INVALID (-1)   /* bad argument */
#define PP_ERR_CUDA (-2)      /* launch / runtime failure, see pp_last_error() */
#define PP_ERR_WORKSPACE (-3) /* workspace too small */

typedef void *pp_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

PP_API int pp_version(void);
PP_API const char *pp_last_error(void);
/* number of kernels this library has launched in this process (for gpu_launches accounting) */
PP_API int64_t pp_launch_count(void);
/* Opt-in per-kernel timing with CUDA events on the launching stream (used by bench.py for the roofline
 * numbers).  pp_profile_enable(1) clears and starts collecting, (0) stops; pp_profile_report waits for
 * the last event and writes one line per kernel: "<name> <launches> <total_ms>\n". */
PP_API int pp_profile_enable(int on);
PP_API int pp_profile_report(char *buf, size_t buf_bytes);

/* ------------------------------------------------------------------------------------------
 * Stage 1 -- hard voxelization.
 * Replaces ops/ops_numba.py:109-168 points_to_voxel and its two kernels (:171-240 given /
 * shuffled order, :242-308 reflectance pre-order), called from VoxelGenerator.generate (:56-60),
 * CustomVoxelGenerator.generate (:95-99) and PointPillarsVoxelization.forward
 * (model/PointPillars.py:330-354).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double range[6];     /* coors_range xyzxyz, exactly converted to double                 */
    double vsize[3];     /* voxel_size, exactly converted to double                         */
    int32_t range_is_f64; /* numba promotion: (p - range) is done in f64 iff range is f64   */
    int32_t vsize_is_f64; /* the division is f64 iff range or voxel_size is f64             */
    int32_t grid[3];     /* np.round((range[3:]-range[:3])/vsize) as in ops_numba.py:144-145 */
    int32_t max_points;  /* P, points kept per pillar                                        */
    int32_t max_voxels;  /* pillar cap; the (cap+1)-th new pillar BREAKS the pass (:223,:291) */
    int32_t num_feats;   /* C, floats per point (>= 3; >= 4 for PP_ORDER_REFLECTANCE_DESC)    */
} pp_voxel_cfg;

enum {
    PP_ORDER_GIVEN = 0,            /* process points in array order (replay of the shuffled order) */
    PP_ORDER_REFLECTANCE_DESC = 1, /* points[:,3] descending; ties: lower original index first, -0.0 == +0.0 (the reference's
                                      numba quicksort leaves the order of ties unspecified: replay it with PP_ORDER_PERM) */
    PP_ORDER_PERM = 2              /* caller-supplied permutation: position p reads points[perm[p]] */
};

/* rows the caller must allocate for voxels/coors/num_points: min(max_voxels, N, cells).  N < 2^27, cells < 2^31. */
PP_API int64_t pp_voxelize_max_rows(int64_t n_points, const pp_voxel_cfg *cfg);
PP_API size_t pp_voxelize_workspace_bytes(int64_t n_points, const pp_voxel_cfg *cfg, int order);

/*
 * points     (N, C) f32
 * perm       (N) int32, only for PP_ORDER_PERM (else NULL)
 * voxels     (rows, P, C) f32   rows < *voxel_num are fully written (zero padded)
 * coors      (rows, 3) int32    xyz cell index, like the numpy return (:164)
 * num_points (rows) int32
 * voxel_num  device int32 scalar = number of pillars (the reference's slice bound, :164-166)
 * pillar_map optional (may be NULL): (gx*gy*gz) int32 written with the pillar id of every
 *            occupied cell and -1 elsewhere (cell = (z*gy + y)*gx + x, i.e. the (D,H,W) order of
 *            the BEV canvas), for the fused scatter.
 */
PP_API int pp_voxelize(const float *points, int64_t n_points, const pp_voxel_cfg *cfg, int order,
                const int32_t *perm, float *voxels, int32_t *coors, int32_t *num_points,
                int32_t *voxel_num, int32_t *pillar_map, void *workspace, size_t workspace_bytes,
                pp_stream_t stream);

/*
 * pp_voxelize fused with the single-layer PillarFeatureNet (pp_pillar_features below): the warp that gathers a pillar
 * also decorates it and runs it through Linear + BN + ReLU + max while its points are still in registers, so the frame
 * has one launch less and the voxels are not read back.  Same outputs as pp_voxelize followed by pp_pillar_features
 * (bit-identical features); feat (rows, units + 1).  Needs num_feats == 4, max_points <= 32, units <= 64.
 */
typedef struct {
    const float *weight;  /* (units, 9) */
    const float *scale;   /* (units) BatchNorm folded: gamma / sqrt(var + eps) */
    const float *shift;   /* (units) beta - mean * scale */
    int32_t units;
    float vx, vy, x_off, y_off; /* pillar centre = cell * v + off, model/PointPillars.py:500-508 */
    float *feat;
} pp_pfn_fused;
PP_API int pp_voxelize_features(const float *points, int64_t n_points, const pp_voxel_cfg *cfg, int order,
                         const int32_t *perm, float *voxels, int32_t *coors, int32_t *num_points,
                         int32_t *voxel_num, int32_t *pillar_map, const pp_pfn_fused *pfn, void *workspace,
                         size_t workspace_bytes, pp_stream_t stream);

/*
 * The whole of stages 1 + 2 of one frame in one call: pp_voxelize_features AND the dense scatter of the pillar
 * features into the BEV pseudo-image.  Replaces PointPillarsVoxelization.forward (model/PointPillars.py:330-354) +
 * PillarFeatureNet.forward (:480-526) + SparseConvTensor(...).dense().view (:565-571) for batch size 1.
 * canvas: (1, (units+1)*D, H, W) float32 of THIS frame (for a batch pass each frame's slice), written completely:
 * the gather kernel stores each pillar's features straight from its registers (channel c of cell (z,y,x) at
 * canvas[(c*D+z)*H*W + y*W + x]); the zeros of all other cells are written by the per-point kernels before their
 * dependency waits, i.e. off the critical path.  canvas must be 32-byte aligned, (units+1)*D*H*W a multiple of 8.
 * pfn->feat (rows, units+1) and pillar_map may be NULL here (they are by-products, not inputs of the scatter).
 * Same restrictions as pp_voxelize_features.
 */
PP_API int pp_voxelize_scatter(const float *points, int64_t n_points, const pp_voxel_cfg *cfg, int order,
                        const int32_t *perm, float *voxels, int32_t *coors, int32_t *num_points,
                        int32_t *voxel_num, int32_t *pillar_map, const pp_pfn_fused *pfn, float *canvas,
                        void *workspace, size_t workspace_bytes, pp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2 -- pillar decoration, PFN and dense scatter.
 * ---------------------------------------------------------------------------------------- */
enum {
    PP_COORS_XYZ_I32 = 0,  /* (M,3) int32 x,y,z  (pp_voxelize output) + a single batch index */
    PP_COORS_BZYX_I32 = 1, /* (M,4) int32 b,z,y,x (SparseMiddleExtractor after .int())        */
    PP_COORS_BZYX_I64 = 2  /* (M,4) int64 b,z,y,x (PointPillars.voxelize output)              */
};
enum { PP_NUM_I32 = 0, PP_NUM_I64 = 1 };

/*
 * Decoration only.  Replaces PillarFeatureNet.forward up to the padding mask,
 * model/PointPillars.py:490-521 (+ get_paddings_indicator, model/utils.py:442-458).
 * out (M, P, C+5) = [feat(C) | xyz - mean | x - (cx*vx + x_off), y - (cy*vy + y_off)] * (slot < n).
 * m_dev: optional device int32 scalar overriding M (rows >= *m_dev are skipped), may be NULL.
 */
PP_API int pp_decorate(const float *voxels, const void *num_points, int num_kind, const void *coors,
                int coors_kind, int64_t M, const int32_t *m_dev, int P, int C, float vx, float vy,
                float x_off, float y_off, float *out, pp_stream_t stream);

/*
 * One PFNLayer in eval mode.  Replaces PFNLayer.forward, model/PointPillars.py:388-423:
 * Linear(bias=False) -> BatchNorm1d (running stats, folded to scale/shift by the caller)
 * -> ReLU -> max over all P slots.  last_layer: out (M, U); else out (M, P, 2U).
 */
PP_API int pp_pfn_layer(const float *in, int64_t M, int P, int Cin, const float *weight /* (U,Cin) */,
                 const float *scale, const float *shift, int U, int last_layer, float *out,
                 pp_stream_t stream);

/*
 * Fused single-layer PillarFeatureNet (decorate + Linear + BN + ReLU + max + num_points channel).
 * Replaces PillarFeatureNet.forward, model/PointPillars.py:480-526, for len(feat_channels) == 1.
 * feat (M, U+1); the last channel is float(num_points) (:526).
 */
PP_API int pp_pillar_features(const float *voxels, const void *num_points, int num_kind, const void *coors,
                       int coors_kind, int64_t M, const int32_t *m_dev, int P, int C, float vx,
                       float vy, float x_off, float y_off, const float *weight /* (U, C+5) */,
                       const float *scale, const float *shift, int U, float *feat,
                       pp_stream_t stream);

/*
 * Dense scatter.  Replaces SparseMiddleExtractor.forward's
 * SparseConvTensor(feat, coors, (D,H,W), B).dense().view(N, C*D, H, W), model/PointPillars.py:565-571.
 * canvas (B, C*D, H, W) f32 is written exactly once (zeros included): no memset needed.
 * Duplicate coordinates: the highest row index wins (index_put order).
 * map_ws: workspace of B*D*H*W int32.  batch_index is used only with PP_COORS_XYZ_I32.
 */
PP_API size_t pp_scatter_workspace_bytes(int B, int D, int H, int W);
PP_API int pp_scatter_dense(const float *feat, const void *coors, int coors_kind, int64_t M,
                     const int32_t *m_dev, int C, int batch_index, int B, int D, int H, int W,
                     float *canvas, void *map_ws, size_t map_ws_bytes, pp_stream_t stream);

/* Same scatter, fed by the cell -> pillar map that pp_voxelize already produced (pillar_map argument):
 * no map build, one kernel.  feat (M, C); pillar_map (B*D*H*W) int32 in (b, z, y, x) order, -1 = empty. */
PP_API int pp_scatter_mapped(const float *feat, const int32_t *pillar_map, int C, int B, int D, int H, int W,
                             float *canvas, pp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage 3 -- boxes: codec, corners, IoU, NMS.  Boxes are 9-parameter
 * [x, y, z_bottom, dx, dy, dz, rx, ry, rz] (config.yaml:5).
 * ---------------------------------------------------------------------------------------- */
/* BBoxCoder.encode / decode, model/utils.py:276-306 / :309-337.  (K,9) each. */
PP_API int pp_box_encode(const float *src, const float *dst, int64_t K, float *out, pp_stream_t stream);
PP_API int pp_box_decode(const float *anchors, const float *deltas, int64_t K, float *out, pp_stream_t stream);
/* limit_period, model/utils.py:339-350 */
PP_API int pp_limit_period(const float *val, int64_t n, float offset, float period, float *out,
                    pp_stream_t stream);
/* Anchor3DRangeGenerator.grid_anchors for one range, model/utils.py:168-264.
 * out (D,H,W,S,R,9); sizes (S,3) and rots (R,3) are HOST arrays (tiny, passed by value). */
PP_API int pp_grid_anchors(const float *range6_host, const float *sizes_host, int S, const float *rots_host,
                    int R, int D, int H, int W, float *out, pp_stream_t stream);
/* bbox2corners3D, ops/ops_torch.py:160-256: (N,9) -> (N,8,3) */
PP_API int pp_box_corners3d(const float *boxes, int64_t N, float *corners, pp_stream_t stream);
/* bbox2rotated_corners2D, ops/ops_torch.py:13-114: (N,9) -> (N,4) xy bounding rectangle */
PP_API int pp_box_aabb2d(const float *boxes, int64_t N, float *rect, pp_stream_t stream);

enum { PP_IOU = 0, PP_IOF = 1, PP_GIOU = 2 };
/* bbox_iou2D, ops/ops_torch.py:538-607: (m,4),(n,4) -> (m,n); same op order, no FMA contraction,
 * so results are bit-identical to the eager torch ops on identical rectangles. */
PP_API int pp_bbox_iou2d(const float *b1, int64_t m, const float *b2, int64_t n, int mode, float eps,
                  float *out, pp_stream_t stream);
/* Rotated BEV IoU (extension named by the north star; the reference has no rotated-rectangle IoU, its nms_dim == 2
 * form is the AABB above).  Footprint of a 9-parameter box = centre (x,y), size (dx,dy), yaw rz; (m,9),(n,9) -> (m,n).
 * The intersection area is evaluated in float64 (a in b's frame: the unit square cut by b's two axes, Green's integral
 * edge by edge, no polygon built); the float32 result is within rounding of the float64 oracle
 * (tests/test_rotated_iou.py).  iou(a, b) == iou(b, a) bit for bit. */
PP_API int pp_iou_rotated_bev(const float *b1, int64_t m, const float *b2, int64_t n, float *out, pp_stream_t stream);
/* box3d_overlap, ops/ops_torch.py:692-755 (pytorch3d _C.iou_box3d, a third-party kernel that is not part of the
 * reference checkout: parity unpinned; pinned against an independent float64 computation instead).
 * corners (N,8,3) in the reference's corner order; vol (may be NULL) and iou (N,M).  Exact convex intersection of the
 * two parallelepipeds (v0; v1-v0, v3-v0, v4-v0), iou = vol / (vol1 + vol2 - vol), evaluated in float64 (divergence theorem
 * face by face, each face area a branch-free line integral): the float32 results are within rounding of
 * the float64 oracle, symmetric bit for bit, and independent of the argument order also for faces that are coplanar only
 * up to the rounding of the corners.  0 when the xy bounding rectangles of the corners do not overlap. */
PP_API int pp_box3d_overlap(const float *corners1, int64_t n, const float *corners2, int64_t m, float *vol, float *iou,
                     pp_stream_t stream);
/* check_coplanar + check_nonzero, ops/ops_torch.py:610-690: flags[i] bit 0 = not coplanar, bit 1 = zero-area face */
PP_API int pp_box3d_check(const float *corners, int64_t n, float eps, int32_t *flags, pp_stream_t stream);
/* iou_jit, ops/ops_numba.py:7-36 (eps added to widths, evaluated in f64 like numba) */
PP_API int pp_iou_jit(const float *boxes, int64_t N, const float *query, int64_t K, double eps, float *out,
               pp_stream_t stream);

/*
 * One class of multiclass_nms, model/utils.py:376-424 (nms_dim == 2 form):
 * candidates = score > score_thr (strict), sorted by descending score (ties: lower index first),
 * AABB of the rotated box (bbox2rotated_corners2D), greedy suppression with iou > iou_thr (strict).
 * scores: element i at scores[i * score_stride].
 * keep (N) int64: kept ORIGINAL indices in descending-score order; keep_count device int32 scalar.
 */
enum { PP_NMS_AABB2D = 0, PP_NMS_ROT_BEV = 1, PP_NMS_BOX3D = 2 };
PP_API size_t pp_nms_workspace_bytes(int64_t N);
PP_API int pp_nms(const float *boxes9, const float *scores, int64_t score_stride, int64_t N, float score_thr,
           float iou_thr, int64_t *keep, int32_t *keep_count, void *workspace, size_t workspace_bytes,
           pp_stream_t stream);

/* Same greedy NMS with a selectable pair test.  PP_NMS_AABB2D = pp_nms (the reference's nms_dim == 2 form);
 * PP_NMS_ROT_BEV = rotated BEV footprints (x, y, dx, dy, rz) with the pair IoU of pp_iou_rotated_bev (extension named
 * by the north star, no counterpart in the reference): the footprints' bounding rectangles drive the tile prefilter,
 * the exact pair test runs only for pairs whose rectangles overlap (queued per tile, evaluated with full warps).
 * PP_NMS_BOX3D = the reference's nms_dim == 3 form: oriented 3-D IoU of pp_box3d_overlap on bbox2corners3D(boxes). */
PP_API size_t pp_nms_workspace_bytes_mode(int64_t N, int iou_mode);
PP_API int pp_nms_mode(const float *boxes9, const float *scores, int64_t score_stride, int64_t N, float score_thr,
                float iou_thr, int iou_mode, int64_t *keep, int32_t *keep_count, void *workspace,
                size_t workspace_bytes, pp_stream_t stream);

/*
 * Fused head post-processing, Anchor3DHead.get_bboxes_single, model/PointPillars.py:1040-1092.
 * Head tensors stay in the conv layout (channels, H, W); row r = (y*W + x)*A + a of the reference's
 * permute(1,2,0).reshape(-1, k) views reads channels a*k .. a*k+k-1 at (y, x); A = S*R anchors per position.
 *  pp_head_max_scores     max_scores[r] = max_c sigmoid(cls[a*ncls + c, y, x])                         (:1050-1058)
 *  pp_head_select_decode  for the K selected rows (rows == NULL: rows 0..K-1): anchor generated on the fly
 *                         (model/utils.py:168-264, never materialised), boxes (K,9) = BBoxCoder.decode, scores (K,ncls)
 *                         = sigmoid(logits), dir_bits (K,3) = argmax of the three direction logit pairs  (:1045-1065)
 *  pp_head_direction_fixup  boxes[:, 6:9] = limit_period(rot - off, 1, pi) + off + pi * bit, in place    (:1085-1092)
 * range6 / sizes (S,3) / rots (R,3) are HOST arrays like pp_grid_anchors'.
 */
PP_API int pp_head_max_scores(const float *cls, int anchors_per_pos, int ncls, int H, int W, float *max_scores,
                       pp_stream_t stream);
PP_API int pp_head_select_decode(const float *cls, const float *reg, const float *dirs, const int64_t *rows, int64_t K,
                          const float *range6_host, const float *sizes_host, int S, const float *rots_host, int R,
                          int ncls, int H, int W, float *boxes, float *scores, int32_t *dir_bits, pp_stream_t stream);
PP_API int pp_head_direction_fixup(float *boxes, const int32_t *dir_bits, int64_t K, float dir_offset, pp_stream_t stream);

/*
 * `_, topk_inds = max_scores.topk(nms_pre)` of Anchor3DHead.get_bboxes_single, model/PointPillars.py:1056-1065:
 * rows (k) int64 = the indices of the k largest of scores (n), in descending score order; equal scores: lower index
 * first (torch.topk leaves that order unspecified).  Radix select + ordered compaction + stable sort of the survivors;
 * k is clamped to n.  Workspace: pp_head_topk_workspace_bytes(n, k).
 */
PP_API size_t pp_head_topk_workspace_bytes(int64_t n, int64_t k);
PP_API int pp_head_topk(const float *scores, int64_t n, int64_t k, int64_t *rows, void *workspace, size_t workspace_bytes,
                 pp_stream_t stream);

/*
 * Target assignment reductions of Anchor3DHead.assign_bboxes, model/PointPillars.py:964-978, without the (G, A)
 * IoU matrix: for every anchor the best IoU over the ground truths and the FIRST ground truth reaching it (:968),
 * for every ground truth its best IoU over the anchors (:971), and the low-quality-match flag (:976-978):
 * lowq[a] = any g with gt_max[g] >= lo_thr and iou(g, a) == gt_max[g].
 * iou_mode PP_NMS_AABB2D: gt (G,4), anchors (A,4) rectangles (bbox2rotated_corners2D outputs, bbox_iou2D pair test);
 * PP_NMS_BOX3D: gt (G,8,3), anchors (A,8,3) corners (bbox2corners3D outputs, box3d_overlap pair test).
 */
PP_API int pp_assign_overlaps(const float *gt, int64_t G, const float *anchors, int64_t A, int iou_mode, float lo_thr,
                       float *max_ov, int32_t *argmax, float *gt_max, uint8_t *lowq, pp_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Either side of the path.  Workspace for the three compacting calls:
 * pp_compact_workspace_bytes(number of points / of canvas cells).
 * ---------------------------------------------------------------------------------------- */
PP_API size_t pp_compact_workspace_bytes(int64_t n);
/* PointPillars.preprocess, model/PointPillars.py:241-266: optional global_outlier_check (ops/ops_numpy.py:111-115:
 * keep |p - mean| < mean(norm) + 5 std(norm)), range filter lo <= xyz < hi (:251-252), feature selection (:266).
 * out (n, n_features) holds *out_count rows in the input order.  range6 and features are HOST arrays.
 * outlier_check: PP_OUTLIER_OFF; PP_OUTLIER_EXACT = the statistics in numpy's own float32 order of operations (the
 * column means as the sequential sums numpy makes of a reduction over the non-contiguous axis, the 1-D reductions
 * pairwise in blocks of 128): the kept rows are bit-exact with the reference's on float32 input; needs
 * pp_preprocess_workspace_bytes(n).  PP_OUTLIER_FAST = float64 statistics, fully parallel (pp_compact_workspace_bytes(n)
 * suffices): rows within rounding of the 5-sigma threshold can differ. */
enum { PP_OUTLIER_OFF = 0, PP_OUTLIER_EXACT = 1, PP_OUTLIER_FAST = 2 };
PP_API size_t pp_preprocess_workspace_bytes(int64_t n);
PP_API int pp_preprocess_points(const float *points, int64_t n, int C, int outlier_check, const float *range6_host,
                         const int32_t *features_host, int n_features, float *out, int32_t *out_count,
                         void *workspace, size_t workspace_bytes, pp_stream_t stream);
/* min / max of xyz -> out6 (device): the point_cloud_range of CustomVoxelizer.voxelize, model/utils.py:17-18 */
PP_API int pp_points_minmax(const float *points, int64_t n, int C, float *out6, void *workspace, size_t workspace_bytes,
                     pp_stream_t stream);
/* np.sum(vox, axis=1) / vp with vp appended, model/utils.py:34-43: (M,P,C),(M) -> (M, C+1) */
PP_API int pp_voxel_centroids(const float *voxels, const int32_t *num_points, int64_t M, int P, int C, float *out,
                       pp_stream_t stream);
/* SubmanifoldSparseRPN.forward's dense -> sparse step, model/PointPillars.py:766-789: cells of x (B,C,H,W) with any
 * non-zero channel in (b, y, x) row-major order -> coords (nnz,3) int32, values (nnz,C); *nnz device scalar. */
PP_API int pp_dense_to_sparse(const float *x, int B, int C, int H, int W, int32_t *coords, float *values, int32_t *nnz,
                       void *workspace, size_t workspace_bytes, pp_stream_t stream);

/* Stable radix sort of (u32 key, u32 value) pairs, ascending; building block exposed for tests. */
PP_API size_t pp_sort_workspace_bytes(int64_t n);
PP_API int pp_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out,
                      uint32_t *vals_out, int64_t n, void *workspace, size_t workspace_bytes,
                      pp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H */
