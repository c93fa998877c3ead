// dsf_b200 - internal interface between the rasteriser (raster.cu), the MANO backward (mano.cu) and the
// fused steps (fit.cu).  Not part of the C ABI.
#pragma once
#include "common.cuh"

// crop_hand parameters (data/render_loader.py:1209-1227) for the fused loss
struct CropParams {
    const float* joints;   // (B, nj, 3) normalised teacher joints, or nullptr = no crop
    const float* M;        // (B,3,3) axis-aligned crop transform
    int nj;
    float fx, fy, px, py;
    float off_xy, off_z, thick;
};

// The target of the m2d loss in the loader's row-run transport format (dsf_pack_u16_rows): the rasteriser's
// epilogue decodes and normalises it on the fly, so the sensor crop never exists as an fp32 plane in HBM.
struct TargetRows {
    const unsigned short* rows;        // (n_mesh, R, 2) uint16 (first column, length) of every row's span, or null
    const unsigned int* hand_offset;   // (n_mesh + 1) start of each hand's pixels in the payload
    const unsigned short* payload;     // packed uint16 millimetres
    unsigned invalid;                  // sensor's "no measurement" marker besides 0 (0 = none)
};

// loader.normalize_img of one sensor pixel (data/render_loader.py:738-745), same operation order:
// background / invalid -> far plane, clamp to the cube, (v - centre_z) / half_depth
__device__ __forceinline__ float target_norm(unsigned u, unsigned invalid, float cz, float hz, float far_, float near_) {
    float v = (float)u;
    if (u == 0u || (invalid && u == invalid)) v = far_;
    if (v >= far_) v = far_;
    if (v <= near_) v = near_;
    return __fdiv_rn(__fsub_rn(v, cz), hz);
}

// optional fused tail of the raster forward launch (see raster_fwd_kernel)
struct RasterFused {
    float* gv_tile;               // (n_mesh, tiles, NVW*3) per-tile vertex-gradient shares, or null
    int* gv_flag;                 // (n_mesh, tiles)
};

// loss bookkeeping done by the kernels after the rasteriser (no extra launch, no atomics): the per-mesh consumer
// of the gradient shares writes parts (n_mesh,2); the last kernel of the step (MANO pose backward, block 0)
// reduces them to totals (4) - see dsf_depth_loss for the layout.
struct LossFold {
    const float* parts_tile;      // (n_mesh, tiles, 2)
    int n_tiles;
    int n_mesh;
    float* parts;                 // (n_mesh, 2)
    float* totals;                // (4)
    float weight;
};

// what a consumer of the per-tile gradient shares needs: g_verts[mesh] = scale(mesh) * sum_t gv_tile[mesh][t]
// over the flagged tiles, scale = gscale / (N_mesh + 1e-8) / zhalf, N_mesh = mask count of the mesh (all tiles).
struct GradTiles {
    const float* gv_tile;
    const int* gv_flag;
    const float* parts_tile;      // (n_mesh, tiles, 2) [sum, count]
    int n_tiles;
    float gscale;                 // loss_weight / batch of the mean
};

__device__ __forceinline__ float grad_tiles_scale(const GradTiles& gt, int mesh, float zhalf) {
    float n = 0.f;
    for (int t = 0; t < gt.n_tiles; ++t) n += gt.parts_tile[((size_t)mesh * gt.n_tiles + t) * 2 + 1];
    return gt.gscale / (n + 1e-8f) / zhalf;
}

// sum of the flagged tile shares of element i (0 .. NVW*3) of one mesh, fixed order
__device__ __forceinline__ float grad_tiles_load(const GradTiles& gt, int mesh, int i) {
    float a = 0.f;
    for (int t = 0; t < gt.n_tiles; ++t)
        if (gt.gv_flag[(size_t)mesh * gt.n_tiles + t]) a += gt.gv_tile[((size_t)mesh * gt.n_tiles + t) * NVW * 3 + i];
    return a;
}

extern "C" int dsf_raster_tiles(int R);
bool dsf_raster_fused_grad_ok(const DsfMano* h, int flags);
int dsf_raster_forward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                            const float* place_off, const float* view, const float* xs, const float* ys,
                            int R, float* img, int* p2f, float* zbuf, float* bary, float* dists,
                            const float* target, float thr, float* parts_tile, const CropParams* crop,
                            int flags, const RasterFused* fused, cudaStream_t st, const TargetRows* trows = nullptr);
int dsf_raster_backward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                             const float* place_off, const float* view, const float* xs, const float* ys,
                             int R, const int* p2f, const float* g_img, float* g_verts, const float* target,
                             const float* img, const float* parts_tile, float gscale, float thr, const CropParams* crop,
                             int flags, cudaStream_t st);
int dsf_fold_totals_impl(int B, int n_tiles, float weight, const float* parts_tile, float* parts, float* totals,
                         cudaStream_t st);

int dsf_mano_forward_impl(const DsfMano* h, int B, const DsfManoParams* p, float unit_scale, float* verts,
                          float* joints, float* Rs, float* ws, cudaStream_t st);
// gt != null: the vertex cotangent is read from the rasteriser's per-tile shares (cube (B,3) gives zhalf).
// lf != null: loss bookkeeping rides along (parts by the skinning backward when lf->n_mesh == B, totals by block 0
// of the pose backward).
int dsf_mano_backward_impl(const DsfMano* h, int B, const DsfManoParams* p, float unit_scale,
                           const float* verts, const float* joints, const float* g_verts,
                           const float* g_joints, const DsfManoGrads* g, float* ws, const GradTiles* gt,
                           const float* cube, const LossFold* lf, cudaStream_t st);
