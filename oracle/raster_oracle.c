/* TEST INFRASTRUCTURE ONLY.  CPU oracle for the third-party (pytorch3d 0.4.0) arithmetic on the
 * DSF hot path: mesh rasterisation + backward and point-to-face distance + backward.  See the
 * headers of the two *_impl.h files for provenance and the "parity unpinned" statement.
 * Built by oracle/build.py:  gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction, no
 * fast-math: the float build is the bit-exactness contract for the CUDA rasteriser). */
#include <math.h>

#define REAL float
#define SUFFIX f32
#include "raster_oracle_impl.h"
#include "pointface_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX f64
#include "raster_oracle_impl.h"
#include "pointface_oracle_impl.h"
#undef REAL
#undef SUFFIX

/* ---- batch drivers (float build), one hand per OpenMP task: used by the CPU pipeline that
 * bench.py times as the reference arm and by the autograd wrapper in oracle/pipeline.py ---- */
void orc_batch_render_f32(const float* verts, int B, int V, const int* faces, int F, const float* view,
                          const float* xs, const float* ys, int R, int perspective_correct, float eps,
                          int zcull_mode, int* pix_to_face, float* zbuf, float* bary, float* vndc) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* vn = vndc + (long)b * V * 3;
        orc_project_f32(verts + (long)b * V * 3, V, view[8 * b], view[8 * b + 1], view[8 * b + 2],
                        view[8 * b + 3], vn);
        orc_rasterize_f32(vn, faces, F, xs + (long)b * R, R, ys + (long)b * R, R, perspective_correct, eps,
                          zcull_mode, pix_to_face + (long)b * R * R, zbuf + (long)b * R * R,
                          bary ? bary + (long)b * R * R * 3 : 0, 0);
    }
}

void orc_batch_render_backward_f32(const float* verts, int B, int V, const int* faces, const float* view,
                                   const float* xs, const float* ys, int R, const int* pix_to_face,
                                   const float* grad_zbuf, int perspective_correct, float eps,
                                   const float* vndc, float* grad_vndc_scratch, float* grad_verts) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* gvn = grad_vndc_scratch + (long)b * V * 3;
        for (int i = 0; i < V * 3; ++i) gvn[i] = 0.f;
        orc_rasterize_backward_f32(vndc + (long)b * V * 3, faces, xs + (long)b * R, R, ys + (long)b * R, R,
                                   pix_to_face + (long)b * R * R, grad_zbuf + (long)b * R * R, 0,
                                   perspective_correct, eps, gvn);
        orc_project_backward_f32(verts + (long)b * V * 3, V, view[8 * b], view[8 * b + 1], view[8 * b + 2],
                                 view[8 * b + 3], gvn, grad_verts + (long)b * V * 3);
    }
}

void orc_batch_point_face_f32(const float* points, int B, int P, const float* verts, int V, const int* faces,
                              int F, float eps, float* dists, int* idxs) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        orc_point_face_forward_f32(points + (long)b * P * 3, P, verts + (long)b * V * 3, faces, F, eps,
                                   dists + (long)b * P, idxs + (long)b * P);
}
