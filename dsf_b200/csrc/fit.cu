// dsf_b200 - fused model-fitting step: MANO forward -> rasterise -> m2d depth loss -> raster backward
// -> MANO backward, as one fixed launch sequence on one stream (CUDA-graph capturable).
// This is the chain Render.render (mano_layer.py:1071-1097) + train_render.py:728-732 + loss.backward()
// runs through ~700 ATen/pytorch3d launches in the reference.
#include "raster.cuh"

#define AUX_THREADS_RED 256

// workspace of one fused step (floats): MANO scratch | per-tile loss sums | per-tile gradient flags |
// per-tile vertex-gradient shares (the first NVW*3 floats per hand double as g_verts of the unfused path) |
// the rasteriser's done-counter
static inline size_t fit_ws_parts(int n_mesh_mano) { return (size_t)WS_HANDS(n_mesh_mano) * WS_PER_HAND; }

extern "C" long dsf_fit_workspace_floats(int batch, int R) {
    const long nt = dsf_raster_tiles(R);
    return WS_HANDS(batch) * WS_PER_HAND + (long)batch * (2L * nt + nt + nt * NVW * 3) + 4;
}

static int fit_step_impl(const DsfMano* h, int batch, int R, const float* params, const float* center3d,
                         const float* cube, const float* view, const float* xs, const float* ys,
                         const float* target, const TargetRows* trows, float loss_weight, int norm_batch,
                         const float* crop_joints, int n_crop_joints,
                         const float* crop_M, const float* intr4, float* img, int* pix_to_face,
                         float* verts, float* joints, float* g_params, float* parts, float* totals,
                         float* workspace, int flags, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && params && center3d && cube && view && xs && ys && (target || trows), "null input");
    DSF_REQUIRE(img && verts && joints && g_params && parts && totals && workspace, "null output");
    DSF_REQUIRE(batch > 0 && batch <= 65535, "batch must be in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= 512, "crop size R must be in [8,512]");
    const bool fused = dsf_raster_fused_grad_ok(h, flags);
    DSF_REQUIRE(fused || pix_to_face, "pix_to_face may only be NULL when the rasteriser produces the gradient itself "
                                      "(no perspective correction)");
    DSF_REQUIRE(fused || target, "the row-run target needs the step whose rasteriser produces the gradient itself "
                                 "(flags without PERSPECTIVE_CORRECT / SEPARATE_BACKWARD)");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws_mano = workspace;
    const int n_tiles = dsf_raster_tiles(R);
    float* parts_tile = workspace + fit_ws_parts(batch);
    int* gv_flag = reinterpret_cast<int*>(parts_tile + (size_t)batch * n_tiles * 2);
    float* gv_tile = reinterpret_cast<float*>(gv_flag + (size_t)batch * n_tiles);
    float* g_verts = gv_tile;
    const float thr = 0.99f;
    DSF_REQUIRE(!crop_joints || (crop_M && intr4 && n_crop_joints > 0), "crop_joints needs crop_M, intr4 and a joint count");
    // crop_hand defaults of data/render_loader.py:1209 (offsetxy=25, offsetz=20, hand_thickness=20)
    CropParams crop = {crop_joints, crop_M, n_crop_joints, 0.f, 0.f, 0.f, 0.f, 25.f, 20.f, 20.f};
    if (crop_joints) { crop.fx = intr4[0]; crop.fy = intr4[1]; crop.px = intr4[2]; crop.py = intr4[3]; }
    const CropParams* cropp = crop_joints ? &crop : nullptr;

    // params (B,62) = [quat3 | theta45 | beta10 | scale, trans3]   (mano_layer.py:1073-1076)
    DsfManoParams p;
    p.quat = params; p.ld_quat = 62; p.quat_dim = 3;
    p.theta = params + 3; p.ld_theta = 62; p.ncomp = 45;
    p.beta = params + 48; p.ld_beta = 62;
    p.cam = params + 58; p.ld_cam = 62;
    DsfManoGrads g;
    g.quat = g_params; g.ld_quat = 62;
    g.theta = g_params + 3; g.ld_theta = 62;
    g.beta = g_params + 48; g.ld_beta = 62;
    g.cam = g_params + 58; g.ld_cam = 62;
    const float unit_scale = 1000.f * (1.f / 125.f);   // get_mano_vertices(..., global_scale=1/125), :1077
    const float gscale = loss_weight / (float)(norm_batch > 0 ? norm_batch : batch);

    int rc = dsf_mano_forward_impl(h, batch, &p, unit_scale, verts, joints, nullptr, ws_mano, st);
    if (rc) return rc;
    // rasterise + normalise; the per-tile m2d loss sums and - without perspective correction - the vertex
    // gradient itself fall out of the epilogue.  The loss records ride along with the MANO backward kernels
    // (parts by the skinning backward, totals by block 0 of the pose backward): no fold / totals launches.
    RasterFused rf = {fused ? gv_tile : nullptr, gv_flag};
    rc = dsf_raster_forward_impl(h, batch, verts, cube, center3d, view, xs, ys, R, img, pix_to_face, nullptr,
                                 nullptr, nullptr, target, thr, parts_tile, cropp, flags, &rf, st, target ? nullptr : trows);
    if (rc) return rc;
    LossFold lf = {parts_tile, n_tiles, batch, parts, totals, loss_weight};
    if (fused) {
        GradTiles gt = {gv_tile, gv_flag, parts_tile, n_tiles, gscale};
        return dsf_mano_backward_impl(h, batch, &p, unit_scale, verts, joints, nullptr, nullptr, &g, ws_mano, &gt,
                                      cube, &lf, st);
    }
    // perspective-correct depth is not affine in the sample position: separate backward kernel, which
    // recomputes d loss / d img per pixel from (target, img, N_b) - no gradient image in HBM
    rc = dsf_raster_backward_impl(h, batch, verts, cube, center3d, view, xs, ys, R, pix_to_face, nullptr,
                                  g_verts, target, img, parts_tile, gscale, thr, cropp, flags, st);
    if (rc) return rc;
    return dsf_mano_backward_impl(h, batch, &p, unit_scale, verts, joints, g_verts, nullptr, &g, ws_mano, nullptr,
                                  nullptr, &lf, st);
}

extern "C" int dsf_fit_step(const DsfMano* h, int batch, int R, const float* params, const float* center3d,
                            const float* cube, const float* view, const float* xs, const float* ys,
                            const float* target, float loss_weight, int norm_batch, const float* crop_joints,
                            int n_crop_joints,
                            const float* crop_M, const float* intr4, float* img, int* pix_to_face,
                            float* verts, float* joints, float* g_params, float* parts, float* totals,
                            float* workspace, int flags, dsfStream_t stream) {
    DSF_REQUIRE(target, "null input");
    return fit_step_impl(h, batch, R, params, center3d, cube, view, xs, ys, target, nullptr, loss_weight, norm_batch,
                         crop_joints, n_crop_joints, crop_M, intr4, img, pix_to_face, verts, joints, g_params, parts,
                         totals, workspace, flags, stream);
}

// The same step with the target still in the loader's row-run transport format (dsf_pack_u16_rows): the rasteriser's
// epilogue decodes + normalises the sensor pixels where it compares them, so neither the unpack launch nor the fp32
// target plane (64 KB per hand written and read back) exists.  rows / hand_offset point at the first hand of this
// call (hand_offset holds absolute pixel offsets into payload, so slices of a batch share one payload).
extern "C" int dsf_fit_step_rows(const DsfMano* h, int batch, int R, const float* params, const float* center3d,
                                 const float* cube, const float* view, const float* xs, const float* ys,
                                 const unsigned short* rows, const unsigned int* hand_offset,
                                 const unsigned short* payload, int invalid_value, float loss_weight, int norm_batch,
                                 const float* crop_joints, int n_crop_joints, const float* crop_M, const float* intr4,
                                 float* img, int* pix_to_face, float* verts, float* joints, float* g_params,
                                 float* parts, float* totals, float* workspace, int flags, dsfStream_t stream) {
    DSF_REQUIRE(rows && hand_offset && payload, "null row-run target");
    DSF_REQUIRE((((size_t)rows) & 3) == 0, "rows must be 4-byte aligned");
    DSF_REQUIRE(invalid_value >= 0 && invalid_value <= 65535, "invalid_value must fit uint16 (0 = none)");
    TargetRows tr = {rows, hand_offset, payload, (unsigned)invalid_value};
    return fit_step_impl(h, batch, R, params, center3d, cube, view, xs, ys, nullptr, &tr, loss_weight, norm_batch,
                         crop_joints, n_crop_joints, crop_M, intr4, img, pix_to_face, verts, joints, g_params, parts,
                         totals, workspace, flags, stream);
}

// ------------------------------------------------------------------------------------------------
// The rasteriser's fused launch on its own (R1 + R4 + L2 + R2 for camera-space vertices from any source):
// rasterise, normalise, m2d loss against the target and the per-tile shares of d loss / d verts_cam, then
// one small kernel that sums the shares and applies each mesh's loss normalisation.
// ------------------------------------------------------------------------------------------------
extern "C" long dsf_raster_loss_workspace_floats(int n_mesh, int R) {
    const long nt = dsf_raster_tiles(R);
    return (long)n_mesh * (2L * nt + nt + nt * NVW * 3) + 4;
}

__global__ void __launch_bounds__(AUX_THREADS_RED)
grad_tiles_reduce_kernel(GradTiles gt, const float* __restrict__ view, float* __restrict__ g_verts) {
    const int mesh = blockIdx.x;
    __shared__ float s_scale;
    if (threadIdx.x == 0) s_scale = grad_tiles_scale(gt, mesh, view[(size_t)mesh * VIEW + 5]);
    __syncthreads();
    const float sc = s_scale;
    for (int i = threadIdx.x; i < NVW * 3; i += AUX_THREADS_RED)
        g_verts[(size_t)mesh * NVW * 3 + i] = sc * grad_tiles_load(gt, mesh, i);
}

extern "C" int dsf_raster_loss_grad(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                                    const float* xs, const float* ys, int R, const float* target, float thr,
                                    float loss_weight, int norm_batch, float* img, int* pix_to_face, float* parts,
                                    float* totals, float* g_verts_cam, float* workspace, int flags,
                                    dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts_cam && view && xs && ys && target && img && parts && totals && workspace, "null argument");
    DSF_REQUIRE(n_mesh > 0 && n_mesh <= 65535, "n_mesh must be in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= 512, "crop size R must be in [8,512]");
    DSF_REQUIRE(dsf_raster_fused_grad_ok(h, flags & 3), "fused loss gradient needs flags without PERSPECTIVE_CORRECT "
                                                    "(use dsf_raster_forward + dsf_depth_loss + dsf_raster_backward)");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = dsf_raster_tiles(R);
    float* parts_tile = workspace;
    int* gv_flag = reinterpret_cast<int*>(parts_tile + (size_t)n_mesh * n_tiles * 2);
    float* gv_tile = reinterpret_cast<float*>(gv_flag + (size_t)n_mesh * n_tiles);
    RasterFused rf = {gv_tile, gv_flag};
    if (flags & 4096) rf.gv_tile = nullptr;         // tuning aid: forward + loss sums only
    int rc = dsf_raster_forward_impl(h, n_mesh, verts_cam, nullptr, nullptr, view, xs, ys, R, img, pix_to_face, nullptr,
                                     nullptr, nullptr, target, thr, parts_tile, nullptr, flags & 3, &rf, st);
    if (rc) return rc;
    rc = dsf_fold_totals_impl(n_mesh, n_tiles, loss_weight, parts_tile, parts, totals, st);
    if (rc) return rc;
    if (g_verts_cam) {
        GradTiles gt = {gv_tile, gv_flag, parts_tile, n_tiles, loss_weight / (float)(norm_batch > 0 ? norm_batch : n_mesh)};
        grad_tiles_reduce_kernel<<<n_mesh, AUX_THREADS_RED, 0, st>>>(gt, view, g_verts_cam);
        DSF_CHECK_LAUNCH();
    }
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// Render.render as two calls (forward, backward) for the autograd drop-in: the reference's
// Render.render (mano_layer.py:1071-1097) returns (img, joint_uvd, joint_xyz, mesh_xyz); left as
// separate torch ops around the kernels it costs ~110 launches and 1.6 ms of host time per step at
// any batch size, here it is 5 + 4 launches.
// ------------------------------------------------------------------------------------------------
#define AUX_THREADS 128

// joints / verts normalised (get_mano_vertices with global_scale 1/125) -> the three auxiliary outputs,
// same operation order as the reference: hand = x * cube / 2 + center (:1078-1079), JointTrans (:1301-1309)
// with points3DToImg (:1318-1324), joint_xyz / mesh_xyz = (hand - center) / cube * 2 (:1093-1094)
__global__ void __launch_bounds__(AUX_THREADS)
render_aux_kernel(const float* __restrict__ verts, const float* __restrict__ joints, const float* __restrict__ center,
                  const float* __restrict__ cube, const float* __restrict__ M, float fx, float fy, float px, float py,
                  float crop, float* __restrict__ joint_uvd, float* __restrict__ joint_xyz, float* __restrict__ mesh_xyz) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const float c[3] = {center[3 * b], center[3 * b + 1], center[3 * b + 2]};
    const float q[3] = {cube[3 * b], cube[3 * b + 1], cube[3 * b + 2]};
    if (mesh_xyz)
        for (int i = tid; i < NVW * 3; i += AUX_THREADS) {
            const int k = i % 3;
            const float hv = __fadd_rn(__fdiv_rn(__fmul_rn(verts[(size_t)b * NVW * 3 + i], q[k]), 2.f), c[k]);
            mesh_xyz[(size_t)b * NVW * 3 + i] = __fmul_rn(__fdiv_rn(__fsub_rn(hv, c[k]), q[k]), 2.f);
        }
    if (tid < NJOUT) {
        const float* j = joints + ((size_t)b * NJOUT + tid) * 3;
        float h[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) h[k] = __fadd_rn(__fdiv_rn(__fmul_rn(j[k], q[k]), 2.f), c[k]);
        if (joint_xyz)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                joint_xyz[((size_t)b * NJOUT + tid) * 3 + k] = __fmul_rn(__fdiv_rn(__fsub_rn(h[k], c[k]), q[k]), 2.f);
        if (joint_uvd) {
            const float* m = M + 9 * (size_t)b;
            const float u = __fadd_rn(__fdiv_rn(__fmul_rn(h[0], fx), __fadd_rn(h[2], 1e-8f)), px);
            const float v = __fadd_rn(__fdiv_rn(__fmul_rn(h[1], fy), h[2]), py);
            const float ut = __fadd_rn(__fadd_rn(__fmul_rn(m[0], u), __fmul_rn(m[1], v)), m[2]);
            const float vt = __fadd_rn(__fadd_rn(__fmul_rn(m[3], u), __fmul_rn(m[4], v)), m[5]);
            float* o = joint_uvd + ((size_t)b * NJOUT + tid) * 3;
            o[0] = __fsub_rn(__fmul_rn(__fdiv_rn(ut, crop), 2.f), 1.f);
            o[1] = __fsub_rn(__fmul_rn(__fdiv_rn(vt, crop), 2.f), 1.f);
            o[2] = __fdiv_rn(__fsub_rn(h[2], c[2]), __fdiv_rn(q[2], 2.f));
        }
    }
}

// cotangents of the auxiliary outputs -> g_verts += g_mesh_xyz, g_joints = chain of g_joint_uvd + g_joint_xyz
__global__ void __launch_bounds__(AUX_THREADS)
render_aux_bwd_kernel(const float* __restrict__ joints, const float* __restrict__ center, const float* __restrict__ cube,
                      const float* __restrict__ M, float fx, float fy, float crop, const float* __restrict__ g_uvd,
                      const float* __restrict__ g_jxyz, const float* __restrict__ g_mxyz, int have_raster,
                      float* __restrict__ g_verts, float* __restrict__ g_joints) {
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < NVW * 3; i += AUX_THREADS) {
        const size_t o = (size_t)b * NVW * 3 + i;
        const float a = have_raster ? g_verts[o] : 0.f;
        g_verts[o] = g_mxyz ? a + g_mxyz[o] : a;          // d mesh_xyz / d verts = 1
    }
    if (tid < NJOUT) {
        const size_t o = ((size_t)b * NJOUT + tid) * 3;
        float g[3] = {0.f, 0.f, 0.f};
        if (g_jxyz) { g[0] = g_jxyz[o]; g[1] = g_jxyz[o + 1]; g[2] = g_jxyz[o + 2]; }
        if (g_uvd) {
            const float q[3] = {cube[3 * b], cube[3 * b + 1], cube[3 * b + 2]};
            const float* j = joints + o;
            const float hx = j[0] * q[0] * 0.5f + center[3 * b], hy = j[1] * q[1] * 0.5f + center[3 * b + 1],
                        hz = j[2] * q[2] * 0.5f + center[3 * b + 2];
            const float* m = M + 9 * (size_t)b;
            const float s = 2.f / crop;
            const float gu = (g_uvd[o] * m[0] + g_uvd[o + 1] * m[3]) * s, gv = (g_uvd[o] * m[1] + g_uvd[o + 1] * m[4]) * s;
            const float ize = 1.f / (hz + 1e-8f), iz = 1.f / hz;
            const float ghx = gu * fx * ize, ghy = gv * fy * iz;
            const float ghz = -gu * hx * fx * ize * ize - gv * hy * fy * iz * iz + g_uvd[o + 2] / (q[2] * 0.5f);
            g[0] += ghx * q[0] * 0.5f; g[1] += ghy * q[1] * 0.5f; g[2] += ghz * q[2] * 0.5f;
        }
        g_joints[o] = g[0]; g_joints[o + 1] = g[1]; g_joints[o + 2] = g[2];
    }
}

static void render_params(const float* params, int ld, int quat_dim, float* g_params, DsfManoParams* p, DsfManoGrads* g) {
    p->quat = params; p->ld_quat = ld; p->quat_dim = quat_dim;
    p->theta = params + quat_dim; p->ld_theta = ld; p->ncomp = 45;
    p->beta = params + quat_dim + 45; p->ld_beta = ld;
    p->cam = params + quat_dim + 55; p->ld_cam = ld;
    if (g) {
        g->quat = g_params; g->ld_quat = ld;
        g->theta = g_params + quat_dim; g->ld_theta = ld;
        g->beta = g_params + quat_dim + 45; g->ld_beta = ld;
        g->cam = g_params + quat_dim + 55; g->ld_cam = ld;
    }
}

extern "C" long dsf_render_workspace_floats(int batch) {
    return WS_HANDS(batch) * WS_PER_HAND + (long)batch * (NVW * 3 + NJOUT * 3);
}

extern "C" int dsf_render_forward(const DsfMano* h, int batch, int R, const float* params, int ld_params, int quat_dim,
                                  const float* center3d, const float* cube, const float* view, const float* xs,
                                  const float* ys, const float* M, const float* intr4, float* img, int* pix_to_face,
                                  float* verts, float* joints, float* joint_uvd, float* joint_xyz, float* mesh_xyz,
                                  float* workspace, int flags, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && params && center3d && cube && view && xs && ys && M && intr4, "null input");
    DSF_REQUIRE(img && pix_to_face && verts && joints && workspace, "null output");
    DSF_REQUIRE(batch > 0 && batch <= 65535, "batch must be in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= 512, "crop size R must be in [8,512]");
    DSF_REQUIRE((quat_dim == 3 || quat_dim == 4) && ld_params >= quat_dim + 59, "params must be (B, 62 | 63)");
    cudaStream_t st = (cudaStream_t)stream;
    DsfManoParams p;
    render_params(params, ld_params, quat_dim, nullptr, &p, nullptr);
    int rc = dsf_mano_forward_impl(h, batch, &p, 1000.f * (1.f / 125.f), verts, joints, nullptr, workspace, st);
    if (rc) return rc;
    rc = dsf_raster_forward_impl(h, batch, verts, cube, center3d, view, xs, ys, R, img, pix_to_face, nullptr, nullptr,
                                 nullptr, nullptr, 0.99f, nullptr, nullptr, flags, nullptr, st);
    if (rc) return rc;
    if (joint_uvd || joint_xyz || mesh_xyz) {
        render_aux_kernel<<<batch, AUX_THREADS, 0, st>>>(verts, joints, center3d, cube, M, intr4[0], intr4[1], intr4[2],
                                                         intr4[3], (float)R, joint_uvd, joint_xyz, mesh_xyz);
        DSF_CHECK_LAUNCH();
    }
    return DSF_OK;
}

extern "C" int dsf_render_backward(const DsfMano* h, int batch, int R, const float* params, int ld_params, int quat_dim,
                                   const float* center3d, const float* cube, const float* view, const float* xs,
                                   const float* ys, const float* M, const float* intr4, const float* verts,
                                   const float* joints, const int* pix_to_face, const float* g_img,
                                   const float* g_joint_uvd, const float* g_joint_xyz, const float* g_mesh_xyz,
                                   float* g_params, float* workspace, int flags, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && params && center3d && cube && view && xs && ys && M && intr4 && verts && joints && pix_to_face,
                "null input");
    DSF_REQUIRE(g_params && workspace, "null output");
    DSF_REQUIRE(batch > 0 && batch <= 65535, "batch must be in [1,65535] per call");
    DSF_REQUIRE((quat_dim == 3 || quat_dim == 4) && ld_params >= quat_dim + 59, "params must be (B, 62 | 63)");
    cudaStream_t st = (cudaStream_t)stream;
    float* g_verts = workspace + (size_t)WS_HANDS(batch) * WS_PER_HAND;
    float* g_joints = g_verts + (size_t)batch * NVW * 3;
    DsfManoParams p;
    DsfManoGrads g;
    render_params(params, ld_params, quat_dim, g_params, &p, &g);
    int rc;
    if (g_img) {
        rc = dsf_raster_backward_impl(h, batch, verts, cube, center3d, view, xs, ys, R, pix_to_face, g_img, g_verts,
                                      nullptr, nullptr, nullptr, 0.f, 0.99f, nullptr, flags, st);
        if (rc) return rc;
    }
    render_aux_bwd_kernel<<<batch, AUX_THREADS, 0, st>>>(joints, center3d, cube, M, intr4[0], intr4[1], (float)R,
                                                         g_joint_uvd, g_joint_xyz, g_mesh_xyz, g_img ? 1 : 0, g_verts,
                                                         g_joints);
    DSF_CHECK_LAUNCH();
    return dsf_mano_backward_impl(h, batch, &p, 1000.f * (1.f / 125.f), verts, joints, g_verts, g_joints, &g, workspace,
                                  nullptr, nullptr, nullptr, st);
}

// ------------------------------------------------------------------------------------------------
// Multi-view fitting step (BASELINE config "high-res multi-view"): MANO once per hand, then per view a
// rigid rotation about the hand's centre (RotationPoints, mano_layer.py:874-885, as getDepth does with
// `rot`, :1204-1209), rasterise + fused m2d loss against that view's target, and the adjoint chain back to
// the hand's 62 parameters (the views' vertex cotangents are rotated back and summed).
// ------------------------------------------------------------------------------------------------
// verts_cam[b, v] = R[b, v] (verts[b] * cube[b] / 2) + center[b]     (rotation about center[b])
__global__ void __launch_bounds__(AUX_THREADS)
views_place_kernel(int views, const float* __restrict__ verts, const float* __restrict__ center,
                   const float* __restrict__ cube, const float* __restrict__ rot, float* __restrict__ verts_cam) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const float hx = cube[3 * b] * 0.5f, hy = cube[3 * b + 1] * 0.5f, hz = cube[3 * b + 2] * 0.5f;
    const float cx = center[3 * b], cy = center[3 * b + 1], cz = center[3 * b + 2];
    for (int i = tid; i < NVW; i += AUX_THREADS) {
        const float* p = verts + ((size_t)b * NVW + i) * 3;
        const float x = p[0] * hx, y = p[1] * hy, z = p[2] * hz;
        for (int v = 0; v < views; ++v) {
            const float* R = rot + ((size_t)b * views + v) * 9;
            float* o = verts_cam + (((size_t)b * views + v) * NVW + i) * 3;
            o[0] = (R[0] * x + R[1] * y + R[2] * z) + cx;
            o[1] = (R[3] * x + R[4] * y + R[5] * z) + cy;
            o[2] = (R[6] * x + R[7] * y + R[8] * z) + cz;
        }
    }
}

// g_verts[b] = sum_v R[b, v]^T g_cam[b, v] * cube[b] / 2; g_cam either dense (B*V,779,3) or the rasteriser's
// per-tile shares (gt.gv_tile != null: sum the flagged tiles, times the mesh's loss normalisation)
__global__ void __launch_bounds__(AUX_THREADS)
views_place_bwd_kernel(int views, const float* __restrict__ g_cam, GradTiles gt, const float* __restrict__ cube,
                       const float* __restrict__ rot, float* __restrict__ g_verts) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const float hx = cube[3 * b] * 0.5f, hy = cube[3 * b + 1] * 0.5f, hz = cube[3 * b + 2] * 0.5f;
    __shared__ float s_scale[32];
    if (gt.gv_tile && tid < views && tid < 32) s_scale[tid] = grad_tiles_scale(gt, b * views + tid, hz);
    __syncthreads();
    for (int i = tid; i < NVW; i += AUX_THREADS) {
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int v = 0; v < views; ++v) {                       // fixed order: deterministic sum
            const float* R = rot + ((size_t)b * views + v) * 9;
            float g0, g1, g2;
            if (gt.gv_tile) {
                const int mesh = b * views + v;
                const float sc = s_scale[v];
                g0 = sc * grad_tiles_load(gt, mesh, 3 * i);
                g1 = sc * grad_tiles_load(gt, mesh, 3 * i + 1);
                g2 = sc * grad_tiles_load(gt, mesh, 3 * i + 2);
            } else {
                const float* g = g_cam + (((size_t)b * views + v) * NVW + i) * 3;
                g0 = g[0]; g1 = g[1]; g2 = g[2];
            }
            ax += R[0] * g0 + R[3] * g1 + R[6] * g2;
            ay += R[1] * g0 + R[4] * g1 + R[7] * g2;
            az += R[2] * g0 + R[5] * g1 + R[8] * g2;
        }
        float* o = g_verts + ((size_t)b * NVW + i) * 3;
        o[0] = ax * hx; o[1] = ay * hy; o[2] = az * hz;
    }
}

extern "C" long dsf_fit_views_workspace_floats(int batch, int views, int R) {
    const long nm = (long)batch * views, nt = dsf_raster_tiles(R);
    const long grad = nt * NVW * 3 > NVW * 3 ? nt * NVW * 3 : NVW * 3;
    return WS_HANDS(batch) * WS_PER_HAND + (long)batch * NVW * 3 + nm * ((long)NVW * 3 + grad + 3L * nt) + 4;
}

extern "C" int dsf_fit_step_views(const DsfMano* h, int batch, int views, int R, const float* params,
                                  const float* center3d, const float* cube, const float* rot, const float* view,
                                  const float* xs, const float* ys, const float* target, float loss_weight,
                                  float* img, int* pix_to_face, float* verts, float* joints, float* g_params,
                                  float* parts, float* totals, float* workspace, int flags, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && params && center3d && cube && rot && view && xs && ys && target, "null input");
    DSF_REQUIRE(img && verts && joints && g_params && parts && totals && workspace, "null output");
    DSF_REQUIRE(batch > 0 && views > 0 && views <= 32 && (long)batch * views <= 65535,
                "views must be in [1,32], batch * views in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= 512, "crop size R must be in [8,512]");
    const bool fused = dsf_raster_fused_grad_ok(h, flags);
    DSF_REQUIRE(fused || pix_to_face, "pix_to_face may only be NULL without perspective correction");
    cudaStream_t st = (cudaStream_t)stream;
    const int nm = batch * views, n_tiles = dsf_raster_tiles(R);
    const size_t grad = (size_t)(n_tiles > 1 ? n_tiles : 1) * NVW * 3;
    float* ws_mano = workspace;
    float* g_verts = workspace + (size_t)WS_HANDS(batch) * WS_PER_HAND;
    float* verts_cam = g_verts + (size_t)batch * NVW * 3;
    float* g_cam = verts_cam + (size_t)nm * NVW * 3;              // dense cotangent, or the per-tile shares
    float* parts_tile = g_cam + (size_t)nm * grad;
    int* gv_flag = reinterpret_cast<int*>(parts_tile + (size_t)nm * n_tiles * 2);
    const float thr = 0.99f;
    DsfManoParams p;
    DsfManoGrads g;
    render_params(params, 62, 3, g_params, &p, &g);
    const float unit_scale = 1000.f * (1.f / 125.f);
    int rc = dsf_mano_forward_impl(h, batch, &p, unit_scale, verts, joints, nullptr, ws_mano, st);
    if (rc) return rc;
    views_place_kernel<<<batch, AUX_THREADS, 0, st>>>(views, verts, center3d, cube, rot, verts_cam);
    DSF_CHECK_LAUNCH();
    RasterFused rf = {fused ? g_cam : nullptr, gv_flag};
    rc = dsf_raster_forward_impl(h, nm, verts_cam, nullptr, nullptr, view, xs, ys, R, img, pix_to_face, nullptr, nullptr,
                                 nullptr, target, thr, parts_tile, nullptr, flags, &rf, st);
    if (rc) return rc;
    GradTiles gt = {};
    if (fused) {
        gt = GradTiles{g_cam, gv_flag, parts_tile, n_tiles, loss_weight / (float)nm};
    } else {
        rc = dsf_raster_backward_impl(h, nm, verts_cam, nullptr, nullptr, view, xs, ys, R, pix_to_face, nullptr, g_cam,
                                      target, img, parts_tile, loss_weight / (float)nm, thr, nullptr, flags, st);
        if (rc) return rc;
    }
    views_place_bwd_kernel<<<batch, AUX_THREADS, 0, st>>>(views, g_cam, gt, cube, rot, g_verts);
    DSF_CHECK_LAUNCH();
    LossFold lf = {parts_tile, n_tiles, nm, parts, totals, loss_weight};     // parts per mesh + totals: pose backward, block 0
    return dsf_mano_backward_impl(h, batch, &p, unit_scale, verts, joints, g_verts, nullptr, &g, ws_mano, nullptr,
                                  nullptr, &lf, st);
}
