// dsf_b200 - fused model-fitting step: MANO forward -> rasterise -> m2d depth loss -> raster backward
// -> MANO backward, as one fixed launch sequence on one stream (CUDA-graph capturable).
// This is the chain Render.render (mano_layer.py:1071-1097) + train_render.py:728-732 + loss.backward()
// runs through ~700 ATen/pytorch3d launches in the reference.
#include "common.cuh"

int dsf_mano_forward_impl(const DsfMano* h, int B, const DsfManoParams* p, float unit_scale, float* verts,
                          float* joints, float* Rs, float* ws, cudaStream_t st);
int dsf_mano_backward_impl(const DsfMano* h, int B, const DsfManoParams* p, float unit_scale,
                           const float* verts, const float* joints, const float* g_verts,
                           const float* g_joints, const DsfManoGrads* g, float* ws, cudaStream_t st);
struct CropParams {
    const float* joints;
    const float* M;
    int nj;
    float fx, fy, px, py;
    float off_xy, off_z, thick;
};
extern "C" int dsf_raster_tiles(int R);
int dsf_raster_forward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                            const float* place_off, const float* view, const float* xs, const float* ys,
                            int R, float* img, int* p2f, float* zbuf, float* bary, float* dists,
                            const float* target, float thr, float* parts_tile, const CropParams* crop,
                            cudaStream_t st);
int dsf_raster_backward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                             const float* place_off, const float* view, const float* xs, const float* ys,
                             int R, const int* p2f, const float* g_img, float* g_verts, const float* target,
                             const float* img, const float* parts, float gscale, float thr, const CropParams* crop,
                             cudaStream_t st);
int dsf_fold_loss_impl(int B, int n_tiles, float weight, const float* parts_tile, float* parts, float* totals,
                       cudaStream_t st);

extern "C" long dsf_fit_workspace_floats(int batch, int R) {
    return (long)batch * (WS_PER_HAND + 2L * dsf_raster_tiles(R) + NVW * 3);
}

extern "C" int dsf_fit_step(const DsfMano* h, int batch, int R, const float* params, const float* center3d,
                            const float* cube, const float* view, const float* xs, const float* ys,
                            const float* target, float loss_weight, int norm_batch, const float* crop_joints,
                            int n_crop_joints,
                            const float* crop_M, const float* intr4, float* img, int* pix_to_face,
                            float* verts, float* joints, float* g_params, float* parts, float* totals,
                            float* workspace, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && params && center3d && cube && view && xs && ys && target, "null input");
    DSF_REQUIRE(img && pix_to_face && verts && joints && g_params && parts && totals && workspace, "null output");
    DSF_REQUIRE(batch > 0 && batch <= 65535, "batch must be in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= 512, "crop size R must be in [8,512]");
    cudaStream_t st = (cudaStream_t)stream;
    float* ws_mano = workspace;
    const int n_tiles = dsf_raster_tiles(R);
    float* parts_tile = workspace + (size_t)batch * WS_PER_HAND;
    float* g_verts = parts_tile + (size_t)batch * n_tiles * 2;
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

    int rc = dsf_mano_forward_impl(h, batch, &p, unit_scale, verts, joints, nullptr, ws_mano, st);
    if (rc) return rc;
    // rasterise + normalise; the m2d loss partial sums fall out of the epilogue (no extra pass)
    rc = dsf_raster_forward_impl(h, batch, verts, cube, center3d, view, xs, ys, R, img, pix_to_face, nullptr,
                                 nullptr, nullptr, target, thr, parts_tile, cropp, st);
    if (rc) return rc;
    rc = dsf_fold_loss_impl(batch, n_tiles, loss_weight, parts_tile, parts, totals, st);
    if (rc) return rc;
    // backward recomputes d loss / d img per pixel from (target, img, N_b): no gradient image in HBM
    rc = dsf_raster_backward_impl(h, batch, verts, cube, center3d, view, xs, ys, R, pix_to_face, nullptr,
                                  g_verts, target, img, parts, loss_weight / (float)(norm_batch > 0 ? norm_batch : batch), thr,
                                  cropp, st);
    if (rc) return rc;
    rc = dsf_mano_backward_impl(h, batch, &p, unit_scale, verts, joints, g_verts, nullptr, &g, ws_mano, st);
    return rc;
}
