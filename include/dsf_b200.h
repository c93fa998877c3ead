/* dsf_b200 - C ABI of the B200-native DSF model-fitting hot path (libdsf_b200.so).
 *
 * Every entry point: plain pointers and sizes, no torch types, returns 0 on success or a
 * negative DsfStatus; never throws, never allocates per call (except dsf_mano_create), never
 * synchronises the device; work is enqueued on the cudaStream_t passed last (a CUstream /
 * torch.cuda.current_stream().cuda_stream handle).  All tensors are fp32 row-major and owned
 * by the caller; indices are int32.  Re-entrant; the only library-owned state is the constant
 * set behind a DsfMano handle.  CUDA-graph capturable.
 *
 * Each declaration names the reference interface it replaces (paths relative to the DSF repo).
 */
#ifndef DSF_B200_H
#define DSF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dsfStream_t; /* cudaStream_t */

enum DsfStatus {
    DSF_OK = 0,
    DSF_ERR_BAD_ARG = -1,
    DSF_ERR_CUDA = -2,
    DSF_ERR_UNSUPPORTED = -3,
    DSF_ERR_NO_DEVICE = -4
};

#define DSF_NV 778        /* MANO vertices */
#define DSF_NVW 779       /* + wrist-cap centre (mano_layer.py:636-637) */
#define DSF_NJ 16         /* kinematic joints */
#define DSF_NJOUT 21      /* + 5 fingertip vertices (mano_layer.py:124-131) */
#define DSF_NBETA 10
#define DSF_NPOSE 135
#define DSF_NSPHERE 66
#define DSF_VIEW_STRIDE 20 /* floats per hand in a view record, see dsf_view_setup */

/* Rasteriser flags (the `flags` argument of the raster / render / fit entry points).
 * DSF_RASTER_PERSPECTIVE_CORRECT = pytorch3d's RasterizationSettings.perspective_correct.  The reference
 * builds RasterizationSettings(image_size, blur_radius=0, faces_per_pixel=1) without it
 * (render_model/mano_layer.py:946-950); in the pytorch3d 0.4.0 it pins (README.md:41) that argument defaults to
 * False and MeshRasterizer.forward passes it on unchanged, so flags = 0 - z interpolated with the screen-space
 * barycentrics - is the reference's behaviour.  Later pytorch3d releases infer True for perspective cameras;
 * the flag reproduces that. */
#define DSF_RASTER_PERSPECTIVE_CORRECT 1
/* DSF_RASTER_SEPARATE_BACKWARD: the fused steps run the raster backward as its own kernel even when the
 * rasteriser's epilogue could emit the vertex gradient (flags without PERSPECTIVE_CORRECT).  The epilogue works
 * with integer pixel moments and px = 1 - (2 q + 1) / S, which is exact when S is a power of two (the direct
 * R x R raster); in the literal 640-pixel raster the float32 sample coordinates deviate from it by up to one
 * ulp, which the per-pixel kernel sees and the moments do not - callers in literal mode set this flag. */
#define DSF_RASTER_SEPARATE_BACKWARD 2

const char* dsf_last_error_string(void);
int dsf_version(void);

/* ------------------------------------------------------------------------------------------
 * M0  MANO constants  - replaces MANO_SMPL.__init__ buffer set-up, render_model/mano_layer.py:98-269
 * Host arrays in the layouts the reference builds; copied to the current device once. */
typedef struct DsfManoHost {
    const float* v_template;  /* (778,3)                       :112-113 */
    const float* shapedirs;   /* (10,2334)  reshape(-1,10).T   :116-120 */
    const float* posedirs;    /* (135,2334)                    :142-145 */
    const float* j_regressor; /* (778,16)   regressed joints   :123 (tip columns are implied) */
    const float* hands_comp;  /* (45,45)                       :135-136 */
    const float* hands_mean;  /* (45)                          :138-139 */
    const float* weights;     /* (778,16)                      :149-154 */
    const int* parents;       /* (16) parents[0] ignored       :147 */
    const int* faces;         /* (n_faces,3) incl. the 16 wrist-fan faces :102-106 */
    int n_faces;
} DsfManoHost;

typedef struct DsfMano DsfMano;
int dsf_mano_create(const DsfManoHost* host, DsfMano** out);
int dsf_mano_free(DsfMano* h);
/* floats of caller-provided scratch per call of dsf_mano_forward (kept for dsf_mano_backward); the scratch (and the
 * workspace of the fused steps, which starts with it) must be 16-byte aligned: its rows travel by TMA / cp.async */
long dsf_mano_workspace_floats(int batch);

/* A (B, ...) parameter block addressed with row strides so the slices of a (B,62) tensor
 * (mano_layer.py:1073-1076) can be passed without copies. cam may be NULL (raw MANO metres). */
typedef struct DsfManoParams {
    const float* quat;  int ld_quat;  int quat_dim; /* 3 axis-angle or 4 quaternion (w,x,y,z) */
    const float* theta; int ld_theta; int ncomp;    /* PCA coefficients, ncomp <= 45 */
    const float* beta;  int ld_beta;                /* 10 */
    const float* cam;   int ld_cam;                 /* (scale, tx, ty, tz) or NULL */
} DsfManoParams;

typedef struct DsfManoGrads {
    float* quat;  int ld_quat;
    float* theta; int ld_theta;
    float* beta;  int ld_beta;
    float* cam;   int ld_cam;   /* may be NULL when params.cam is NULL */
} DsfManoGrads;

/* M1-M4  replaces MANO_SMPL.forward (mano_layer.py:573-641) and get_mano_vertices (:643-678).
 * verts (B,779,3), joints (B,21,3), Rs (B,15,3,3) or NULL.
 * With cam: out = (mano * unit_scale) * cam.scale + cam.trans, unit_scale = 1000 [* global_scale].
 * Without cam: unit_scale is applied alone (pass 1 for MANO_SMPL.forward). */
int dsf_mano_forward(const DsfMano* h, int batch, const DsfManoParams* p, float unit_scale,
                     float* verts, float* joints, float* Rs, float* workspace, dsfStream_t stream);

/* autograd of the above: cotangents g_verts (B,779,3) / g_joints (B,21,3) (either may be NULL)
 * -> parameter gradients (overwritten).  workspace must be the one forward filled;
 * verts/joints are forward's outputs. */
int dsf_mano_backward(const DsfMano* h, int batch, const DsfManoParams* p, float unit_scale,
                      const float* verts, const float* joints, const float* g_verts,
                      const float* g_joints, const DsfManoGrads* g, float* workspace,
                      dsfStream_t stream);

/* ------------------------------------------------------------------------------------------
 * R0/R3  per-hand camera + crop set-up - replaces Render.points3DToImg / comToBounds /
 * Offset2Trans / resize / affine_grid (mano_layer.py:1318-1324, :1133-1169, :1233-1260) and
 * the pytorch3d PerspectiveCameras / RasterizationSettings built at :939-952.
 * mode 0 "direct": R x R raster with crop-space intrinsics, samples at crop pixel centres (i + 1/2; what a
 *         pytorch3d MeshRasterizer of image_size R with those intrinsics does - BASELINE's benchmark configs).
 * mode 2 "direct, index-aligned": the same raster with the principal point moved by half a crop pixel, so that
 *         sample i sits at crop coordinate i - the convention of M, JointTrans and of the literal chain, whose
 *         crop index c reads sensor coordinate (c - t) / s.  Use it to fit crops the data loader cut with M.
 * mode 1 "literal": the S x S raster -> (H,W) resize -> crop chain, evaluated only at the
 *         raster pixel each crop pixel reads (S = max(W,H)); M_in (B,3,3) optional (M_render /
 *         getDepth pass their own, must be axis-aligned), else recomputed like render() does.
 * Outputs: view (B,20) = [fxn,fyn,pxn,pyn, zc,zhalf,bg, ax,bx,ay,by, x_lo,x_hi,y_lo,y_hi(as float), affine,
 *          S, 0,0,0]
 *          (affine = 1 when sample index = a * ndc + b holds exactly, i.e. mode 0; the rasteriser then
 *          converts run ends analytically instead of searching xs; S = side of the square raster whose
 *          pixels the samples are: R in mode 0, max(W,H) in mode 1),
 *          xs (B,R), ys (B,R) NDC sample coordinates (NaN = reads zero padding), M_out (B,3,3) or NULL. */
int dsf_view_setup(int mode, int batch, const float* center3d, const float* cube,
                   const float* intr4, int W, int H, int R, const float* M_in, float* view,
                   float* xs, float* ys, float* M_out, dsfStream_t stream);

/* R1/R4  replaces self.rasterizer(meshes) (mano_layer.py:1083, pytorch3d 0.4.0
 * _C.rasterize_meshes, faces_per_pixel=1, blur_radius=0, perspective_correct = flags & 1) fused with the
 * background fill (:1084-1085) and normalize_img (:1289-1299).
 * verts_cam (NM,779,3) camera-space mm; faces come from the handle.
 * img (NM,R,R) normalised depth; pix_to_face (NM,R,R) int32 (-1 bg, index local to the mesh);
 * optional Fragments outputs zbuf (NM,R,R), bary (NM,R,R,3), dists (NM,R,R) with -1 background.
 * optional fused m2d loss: with target (NM,R,R) the epilogue also writes, per mesh and tile,
 * [sum |target - img| * mask, mask count] (union mask at thr) to loss_parts_tile (NM, tiles, 2). */
int dsf_raster_forward(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                       const float* xs, const float* ys, int R, float* img, int* pix_to_face,
                       float* zbuf, float* bary, float* dists, const float* target, float thr,
                       float* loss_parts_tile, int flags, dsfStream_t stream);
/* number of tiles per mesh = length of the per-mesh partial-sum records of loss_parts_tile */
int dsf_raster_tiles(int R);

/* R2  replaces _C.rasterize_meshes_backward + the index_put to verts + the camera chain, for
 * the zbuf-only gradient DSF uses.  g_img (NM,R,R) is the cotangent of img.
 * g_verts_cam (NM,779,3) overwritten. */
int dsf_raster_backward(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                        const float* xs, const float* ys, int R, const int* pix_to_face,
                        const float* g_img, float* g_verts_cam, int flags, dsfStream_t stream);

/* R1 + R4 + L2 + R2 in one rasteriser launch, for camera-space vertices from any source (flags must not
 * contain DSF_RASTER_PERSPECTIVE_CORRECT): rasterise + normalise, the inline m2d loss against target (NM,R,R)
 * (train_render.py:728-732: union mask at thr, per-mesh normalised, x loss_weight / norm_batch; norm_batch <= 0
 * means n_mesh) and d loss / d verts_cam.  Without perspective correction the depth of a face is affine in the
 * sample position, so the zbuf cotangent of a face reduces to three integer moments of sign(img - target) over
 * its pixels; the epilogue accumulates them, evaluates one closed-form gradient per touched face and gathers per
 * vertex - no pix_to_face plane, no second pass.  img (NM,R,R); pix_to_face (NM,R,R) or NULL (not written);
 * parts (NM,2), totals (4) as dsf_depth_loss mode 0; g_verts_cam (NM,779,3) or NULL (then the per-tile shares
 * stay in the workspace); workspace: dsf_raster_loss_workspace_floats(n_mesh, R) floats.  3 launches (raster,
 * loss fold, gradient-share reduction). */
long dsf_raster_loss_workspace_floats(int n_mesh, int R);
int dsf_raster_loss_grad(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                         const float* xs, const float* ys, int R, const float* target, float thr,
                         float loss_weight, int norm_batch, float* img, int* pix_to_face, float* parts,
                         float* totals, float* g_verts_cam, float* workspace, int flags, dsfStream_t stream);

/* ------------------------------------------------------------------------------------------
 * L1/L2  render losses.
 * mode 0: inline m2d loss, train_render.py:728-732  (union mask, per-hand normalised, x weight/B)
 * mode 1: depth_loss.forward, render_model/render_loss.py:15-21 (both < thr, global mean)
 * real/synth (B,R,R).  parts (B,2) = per-hand [sum |d|*mask, count]; totals (4) =
 * [loss, sum, count, loss * B (un-normalised, for summing over ranks)];
 * g_synth (B,R,R) or NULL = d loss / d synth. */
int dsf_depth_loss(int mode, int batch, int R, const float* real, const float* synth, float thr,
                   float weight, float* parts, float* totals, float* g_synth, dsfStream_t stream);

/* ------------------------------------------------------------------------------------------
 * C1  replaces MANO_SMPL.calculate_coll + get_sphere_radius (mano_layer.py:271-317, :373-385).
 * joints (B,21,3), mesh (B,779,3) (treated as constant, every caller detaches it).
 * out_loss (1) = mean over (B,66) of gated row sums; per_hand (B,2) = [ungated total, gated total];
 * g_joints (B,21,3) or NULL = d out_loss / d joints.  NB the reference's 0.1 gate acts per sphere
 * row (mano_layer.py:383 sums a size-1 axis), reproduced here. */
int dsf_coll_forward_backward(const DsfMano* h, int batch, const float* joints, const float* mesh,
                              float* out_loss, float* per_hand, float* g_joints,
                              dsfStream_t stream);

/* ------------------------------------------------------------------------------------------
 * P1  replaces pytorch3d _C.point_face_dist_forward / _backward as wrapped by
 * metric/meshLoss.py:21-70 and used by ICPLoss (:347-353) / JointICPLoss (:377-394), with the
 * batch-shared face list DSF always passes.  points (B,P,3), verts (B,V,3),
 * faces (F,3) int32 device pointer -> dists (B,P) squared, idxs (B,P) face index in its mesh. */
/* order_ws: (B, P + F) int32 scratch or NULL.  With it the points of each hand, and its faces, are first ordered by
 * the Morton cell of a 16^3 grid (two small extra kernels) so that a warp's points are neighbours, consecutive
 * faces share a small bounding sphere, and the scan rejects whole groups of 8 faces with one test before the
 * per-face bounding-sphere test; results are identical either way (the brute-force minimum and arg-min, lowest
 * face index on ties). */
int dsf_point_face_forward(int batch, int P, int V, int F, const float* points, const float* verts,
                           const int* faces, float* dists, int* idxs, int* order_ws, dsfStream_t stream);
/* dsf_point_face_forward plus work counters for the FP32 roofline of this (compute-bound) row: stats (device,
 * 4 x uint64, overwritten) = pairs rejected by their own bounding-sphere test | evaluated, interior branch |
 * evaluated, edge branch | group-sphere tests (all other pairs fell with their group of 8 faces). */
int dsf_point_face_stats(int batch, int P, int V, int F, const float* points, const float* verts,
                         const int* faces, float* dists, int* idxs, int* order_ws, unsigned long long* stats,
                         dsfStream_t stream);
int dsf_point_face_backward(int batch, int P, int V, int F, const float* points, const float* verts,
                            const int* faces, const int* idxs, const float* g_dists,
                            float* g_points, float* g_verts, dsfStream_t stream);

/* ------------------------------------------------------------------------------------------
 * "next" row f2 - seg_pcl + JointICPLoss.
 * dsf_sphere_set replaces MANO_SMPL.get_sphere_radius (mano_layer.py:271-317): 66 sphere centres
 * (B,66,3) from joints_centres and radii (B,66) from joints_radii + mesh (seg_pcl passes two
 * different joint sets, :407-408).
 * dsf_seg_pcl replaces MANO_SMPL.seg_pcl (:404-426): points (B,P,3) -> label 0 (palm) or the
 * finger bone 1..15 whose sphere surface is nearest, seg (B,P) int32.
 * dsf_joint_icp_forward/backward replace JointICPLoss / FingerICPLoss' distance stage
 * (metric/meshLoss.py:356-394): each point is tested only against the face subset of its own
 * label (subset k = label k+1; CSR subset_ptr (n+1), subset_faces (sum,3) device pointers);
 * other points get distance 0 / index -1, which is what the reference's where() leaves of them. */
int dsf_sphere_set(const DsfMano* h, int batch, const float* joints_centres, const float* joints_radii,
                   const float* mesh, float* centres, float* radii, dsfStream_t stream);
int dsf_seg_pcl(int batch, int P, const float* points, const float* centres, const float* radii, int* seg,
                dsfStream_t stream);
int dsf_joint_icp_forward(int batch, int P, int V, int n_subsets, const float* points, const float* verts,
                          const int* seg, const int* subset_ptr, const int* subset_faces, float* dists,
                          int* idxs, dsfStream_t stream);
int dsf_joint_icp_backward(int batch, int P, int V, const float* points, const float* verts, const int* seg,
                           const int* subset_ptr, const int* subset_faces, const int* idxs,
                           const float* g_dists, float* g_points, float* g_verts, dsfStream_t stream);

/* device pointer to the handle's face list (n_faces,3) int32, for dsf_point_face_* */
const int* dsf_mano_faces_device(const DsfMano* h, int* n_faces);

/* ------------------------------------------------------------------------------------------
 * Fused fitting step: M1+M4 -> R1/R4 -> L2 -> R2 -> MANO backward, one call, fixed launch
 * sequence (graph-capturable).  params (B,62) = [quat3|theta45|beta10|scale|trans3]; target
 * (B,R,R) normalised depth; view/xs/ys from dsf_view_setup; cube (B,3), center3d (B,3).
 * norm_batch: the number of hands the batch mean of the loss runs over (<= 0: batch); a caller that
 * splits one batch into several calls (e.g. on parallel streams) passes the full size so the
 * gradients are scaled alike, and combines totals[3] (un-normalised loss) itself.
 * crop_joints (B,n,3) or NULL: teacher joints for crop_hand of the rendered image before the loss.
 * Outputs: img (B,R,R), pix_to_face (B,R,R), verts (B,779,3), joints (B,21,3) (normalised cube
 * units, global_scale 1/125), g_params (B,62) = d loss/d params, parts (B,2), totals (4).
 * flags = 0 (the reference's rasteriser settings): depth is affine over a face, so the rasteriser's epilogue
 * reduces the loss gradient to three integer moments per face and emits the vertex gradient itself - the step
 * is 7 launches (pose, blend GEMM, skin | raster+loss+raster-backward | skin bwd, GEMM bwd, pose bwd; the loss
 * records parts / totals ride along with the last three), the pix_to_face plane is not needed and pix_to_face
 * may be NULL (not written).
 * flags & DSF_RASTER_PERSPECTIVE_CORRECT: separate raster-backward kernel, pix_to_face required. */
long dsf_fit_workspace_floats(int batch, int R);
int dsf_fit_step(const DsfMano* h, int batch, int R, const float* params, const float* center3d,
                 const float* cube, const float* view, const float* xs, const float* ys,
                 const float* target, float loss_weight, int norm_batch, const float* crop_joints,
                 int n_crop_joints, const float* crop_M, const float* intr4, float* img, int* pix_to_face,
                 float* verts, float* joints, float* g_params, float* parts, float* totals,
                 float* workspace, int flags, dsfStream_t stream);

/* The same step fed straight from the loader's row-run transport format (dsf_pack_u16_rows below; replaces
 * data/render_loader.py:738-745 + the fp32 upload + train_render.py:728-732): the rasteriser's epilogue decodes and
 * normalises the sensor pixels where it compares them, so there is no unpack launch and no fp32 target plane in
 * HBM.  rows (batch,R,2) / hand_offset (batch+1) point at the first hand of this call; hand_offset holds absolute
 * pixel offsets into payload, so the slices of a batch share one payload.  Needs flags without
 * DSF_RASTER_PERSPECTIVE_CORRECT / DSF_RASTER_SEPARATE_BACKWARD.  Results are bit-identical to dsf_fit_step on the
 * target dsf_target_from_u16_rows rebuilds. */
int dsf_fit_step_rows(const DsfMano* h, int batch, int R, const float* params, const float* center3d,
                      const float* cube, const float* view, const float* xs, const float* ys,
                      const unsigned short* rows, const unsigned int* hand_offset, const unsigned short* payload,
                      int invalid_value, float loss_weight, int norm_batch, const float* crop_joints,
                      int n_crop_joints, const float* crop_M, const float* intr4, float* img, int* pix_to_face,
                      float* verts, float* joints, float* g_params, float* parts, float* totals,
                      float* workspace, int flags, dsfStream_t stream);

/* totals (4) of one batch that ran as n_slices dsf_fit_step calls (slice_totals (n_slices,4), each with
 * norm_batch = the full batch): sums [1],[2],[3] and sets [0] = sum [3] / norm_batch.  One tiny launch, so a
 * sliced step stays free of framework ops inside a captured graph. */
int dsf_sum_totals(int n_slices, const float* slice_totals, int norm_batch, float* totals, dsfStream_t stream);

/* Multi-view fused step (BASELINE config "high-res multi-view"): MANO once per hand; per view v the posed
 * hand is rotated about center3d[b] by rot[b, v] (3x3 row-major; RotationPoints :874-885 as getDepth uses it,
 * :1204-1209), rasterised with that view's record (view / xs / ys for batch * views meshes, view-minor) and
 * compared with target (batch * views, R, R) by the union-mask m2d loss (depth + silhouette), mean over all
 * batch * views images times loss_weight; the adjoints of all views are summed into g_params (B,62).
 * img / pix_to_face (batch * views, R, R), parts (batch * views, 2), totals (4) as in dsf_fit_step. */
long dsf_fit_views_workspace_floats(int batch, int views, int R);
int dsf_fit_step_views(const DsfMano* h, int batch, int views, int R, const float* params,
                       const float* center3d, const float* cube, const float* rot, const float* view,
                       const float* xs, const float* ys, const float* target, float loss_weight,
                       float* img, int* pix_to_face, float* verts, float* joints, float* g_params,
                       float* parts, float* totals, float* workspace, int flags, dsfStream_t stream);

/* R5 - Render.render (render_model/mano_layer.py:1071-1097) as one forward and one backward call for the
 * autograd drop-in: params (B, ld_params >= 62 | 63) = [quat(3|4) | theta45 | beta10 | scale, trans3],
 * view / xs / ys / M from dsf_view_setup.  Outputs: img (B,R,R) normalised depth, pix_to_face (B,R,R),
 * verts (B,779,3) / joints (B,21,3) normalised (saved for backward), and the reference's other three
 * return values joint_uvd (B,21,3) (JointTrans :1301-1309), joint_xyz (B,21,3), mesh_xyz (B,779,3)
 * (:1093-1094; any of the three may be NULL).  workspace: dsf_render_workspace_floats(batch) floats, kept
 * untouched between forward and backward.  Backward: cotangents (any may be NULL) -> g_params, same
 * layout / leading dimension as params (only the 62 | 63 parameter columns are written). */
long dsf_render_workspace_floats(int batch);
int dsf_render_forward(const DsfMano* h, int batch, int R, const float* params, int ld_params, int quat_dim,
                       const float* center3d, const float* cube, const float* view, const float* xs,
                       const float* ys, const float* M, const float* intr4, float* img, int* pix_to_face,
                       float* verts, float* joints, float* joint_uvd, float* joint_xyz, float* mesh_xyz,
                       float* workspace, int flags, dsfStream_t stream);
int dsf_render_backward(const DsfMano* h, int batch, int R, const float* params, int ld_params, int quat_dim,
                        const float* center3d, const float* cube, const float* view, const float* xs,
                        const float* ys, const float* M, const float* intr4, const float* verts,
                        const float* joints, const int* pix_to_face, const float* g_img,
                        const float* g_joint_uvd, const float* g_joint_xyz, const float* g_mesh_xyz,
                        float* g_params, float* workspace, int flags, dsfStream_t stream);

/* "next" row f1 - replaces loader.crop_hand (data/render_loader.py:1209-1227, with uvdImg2xyzImg
 * :1190-1200): pixels whose back-projected point falls outside the box around the teacher skeleton
 * (joints (B,nj,3) normalised; offsets in mm, reference defaults 25/20/20) become background 1.0.
 * keep (B,R,R) uint8 mask optional.  The same crop can be applied inside dsf_fit_step to the
 * rendered image before the loss (crop_joints != NULL), as train_render.py:727 does. */
int dsf_crop_hand(int batch, int R, const float* img, const float* joints, int n_joints,
                  const float* center3d, const float* cube, const float* M, const float* intr4,
                  float offset_xy, float offset_z, float thickness, float* out, unsigned char* keep,
                  dsfStream_t stream);

/* "next" row f1 - replaces loader.Img2pcl (data/render_loader.py:1121-1156): nearest resize of the
 * (B,R_in,R_in) normalised depth crop to feature_size^2, foreground = value <= 0.99, each foreground
 * cell back-projected to cube-normalised xyz (uvd_nl2xyznl_tensor :1059-1073; M is inverted in the
 * kernel, img_size / flip are loader.img_size / loader.flip), then a fixed-size cloud per hand:
 * pcl (B, sample_num, 3) = [foreground list in pixel order, repeated floor(sample_num / n) times |
 * sample_num mod n of them drawn without replacement, pixel order]; n = 0 gives zeros.  The draw is
 * a counter-based hash of (seed, hand, pixel): uniform, reproducible, not torch.multinomial's stream.
 * sample_num = 0: every foreground point in pixel order, pcl has feature_size^2 rows per hand of
 * which the first count[b] are written.  count (B) int32 optional. */
int dsf_img2pcl(int batch, int R_in, int feature_size, const float* img, const float* center3d,
                const float* cube, const float* M, const float* intr4, float img_size, float flip,
                int sample_num, unsigned long long seed, float* pcl, int* count, dsfStream_t stream);

/* replaces loader.uvdImg2xyzImg (data/render_loader.py:1190-1200): per pixel camera-space xyz in mm
 * and its cube-normalised copy, both (B,3,R,R); either output may be NULL. */
int dsf_uvd_img_to_xyz(int batch, int R, const float* img, const float* center3d, const float* cube,
                       const float* M, const float* intr4, float img_size, float flip, float* xyz_img,
                       float* xyz_normal, dsfStream_t stream);

/* data format either side of the path: the cropped sensor depth as uint16 millimetres (B,R,R)
 * (0 = no reading; invalid_value, if non-zero, is the loader's `premax` marker) -> the normalised
 * fp32 target (B,R,R) the m2d loss reads.  Replaces loader.normalize_img (data/render_loader.py:
 * 738-745), which the reference runs on the CPU before a 4-byte-per-pixel upload. */
int dsf_target_from_u16(int batch, int R, const unsigned short* depth_mm, const float* center3d,
                        const float* cube, int invalid_value, float* target, dsfStream_t stream);

/* Row-run transport of the same crop (about 80 % of a hand crop is background): per row only the span from the
 * first to the last non-background pixel travels.  rows (B,R,2) uint16 = (first column, length), hand_offset (B+1)
 * uint32 = start of each hand's pixels in the packed uint16 payload.  dsf_target_from_u16_rows rebuilds the
 * normalised fp32 target on the device, bit-identical to dsf_target_from_u16 of the unpacked crop (R <= 512).
 * dsf_pack_u16_rows is the loader's side (plain host code, no device): it fills rows / hand_offset / payload from
 * (B,R,R) uint16 crops and returns the number of payload pixels (-1: capacity too small).  Background = depth 0,
 * the invalid marker, or a value at / beyond the far plane center_z + cube_z / 2. */
int dsf_target_from_u16_rows(int batch, int R, const unsigned short* rows, const unsigned int* hand_offset,
                             const unsigned short* payload, const float* center3d, const float* cube,
                             int invalid_value, float* target, dsfStream_t stream);
long dsf_pack_u16_rows(int batch, int R, const unsigned short* depth_mm, const float* center3d, const float* cube,
                       int invalid_value, unsigned short* rows, unsigned int* hand_offset, unsigned short* payload,
                       long payload_cap);

/* I1 - replaces eval_coll.py:611-626 self_intersection (with get_part_mesh :348-373) and
 * util/intersect.py:102-107 intersect_vox, i.e. trimesh's voxelized(pitch).points +
 * mesh.contains(points) on the CPU.  verts (B,n_verts,3) fp32 (mm); cap centre c = mean of the
 * vertices cap_idx[cap_ptr[c] .. cap_ptr[c+1]) and becomes water vertex n_verts + c; part p owns the
 * triangles part_faces[part_ptr[p] .. part_ptr[p+1]) (int32 triples into the water mesh, watertight);
 * pair_mask[s*n_parts+t] != 0: count the surface voxels of t whose centre lies inside s.
 * Outputs: pair_counts (B,n_parts,n_parts) int64, voxel_counts (B,n_parts) int64 or NULL,
 * volume (B) float64 = sum(pair_counts) * pitch^3 (-1 if status != 0), status (B) int32 bit 0 = a
 * one z-layer of a part's box exceeds the voxel bitmap (196 k voxels; boxes are processed in z-slabs), bit 1 = more than 10 subdivision levels
 * (trimesh raises there).  All topology arrays and outputs are device pointers; workspace of
 * dsf_intersect_workspace_bytes() bytes, 8-byte aligned.  Geometry is evaluated in float64. */
long dsf_intersect_workspace_bytes(int batch, int n_verts, int n_caps, int n_parts);
int dsf_intersect_vox(int batch, int n_verts, const float* verts, int n_caps, const int* cap_ptr,
                      const int* cap_idx, int n_parts, const int* part_ptr, const int* part_faces,
                      const unsigned char* pair_mask, double pitch, long long* pair_counts,
                      long long* voxel_counts, double* volume, int* status, void* workspace,
                      dsfStream_t stream);

/* R5 helper - replaces RotationPoints / RotationNormalPoints (render_model/mano_layer.py:874-895), the
 * rigid rotation that makes the extra camera views: out = R (p - c) + c with R (B,3,3) row-major,
 * pts / out (B,n,3), center3d (B,3) or NULL (pure rotation).  Backward: g_pts = R^T g and, when the
 * pointers are given, g_R (B,3,3) = sum_n g_n (p_n - c)^T and g_center (B,3) = sum_n (g_n - R^T g_n). */
int dsf_rotate_points(int batch, int n, const float* pts, const float* Rm, const float* center3d,
                      float* out, dsfStream_t stream);
int dsf_rotate_points_backward(int batch, int n, const float* pts, const float* Rm, const float* center3d,
                               const float* g_out, float* g_pts, float* g_R, float* g_center,
                               dsfStream_t stream);

/* ------------------------------------------------------------------------------------------
 * "next" row f4 - the synthetic-data generator extras of Render.forward / M_render and the chamfer loss.
 *
 * dsf_mask_img replaces Render.mask_img (render_model/mano_layer.py:1326-1340): pixel (row, col) of the (B,R,R)
 * normalised crop, seen as the point (2 (col + .5) / R - 1, 2 (row + .5) / R - 1, depth), becomes background 1.0
 * when it lies strictly inside one of the hand's n_mask <= 32 spheres (centres (B,n_mask,3), radii (B,n_mask));
 * the random choice of joints / offsets / radii stays with the host wrapper, which draws them with the
 * reference's own generator calls.  out may alias img.
 *
 * dsf_synth2real replaces Render.synth2real (:1222-1231): img + upsample_nearest(noise (B,R/patch,R/patch)) on
 * the foreground (img < bk_value), then - sigma != 0 - reflect padding by 2 and the normalised 5x5 Gaussian of
 * GaussianSmoothing (:808-868) in one pass.  noise may be NULL.  out must not alias img.
 *
 * dsf_chamfer_forward / _backward: nearest neighbour (squared distance, index; lowest index on ties) of every
 * point of x (B,P1,3) in y (B,P2,3) and vice versa - the K = 1 knn_points core of pytorch3d's chamfer_distance as
 * surface_loss.forward uses it (render_model/render_loss.py:44-52); backward takes d loss / d dist_x, d dist_y and
 * writes d loss / d x, d y (overwritten). */
int dsf_mask_img(int batch, int R, const float* img, int n_mask, const float* centres, const float* radii,
                 float* out, dsfStream_t stream);
int dsf_synth2real(int batch, int R, const float* img, const float* noise, int patch, float bk_value,
                   float sigma, float* out, dsfStream_t stream);
int dsf_chamfer_forward(int batch, int P1, int P2, const float* x, const float* y, float* dist_x, int* idx_x,
                        float* dist_y, int* idx_y, dsfStream_t stream);
int dsf_chamfer_backward(int batch, int P1, int P2, const float* x, const float* y, const int* idx_x,
                         const int* idx_y, const float* g_dist_x, const float* g_dist_y, float* g_x, float* g_y,
                         dsfStream_t stream);

/* number of kernel launches the last call on this thread enqueued (bench.py's gpu_launches) */
int dsf_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DSF_B200_H */
