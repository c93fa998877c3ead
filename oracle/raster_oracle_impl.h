/* TEST INFRASTRUCTURE ONLY - included twice by raster_oracle.c (REAL=float, REAL=double).
 *
 * CPU restatement of the pytorch3d==0.4.0 rasteriser the reference calls at
 * render_model/mano_layer.py:952,1083 (MeshRasterizer -> rasterize_meshes -> _C.rasterize_meshes,
 * naive path, faces_per_pixel=1, blur_radius=0) and of its autograd backward.  pytorch3d is a
 * pinned-in-prose dependency (README.md:41) that is NOT vendored under /root/reference and not
 * installable offline, so this follows its published algorithm as recorded in SURVEY.md
 * section 8(a) rows R1/R2 and Appendix B:
 *   csrc/rasterize_meshes/rasterize_meshes_cpu.cpp  (RasterizeMeshesNaiveCpu / BackwardCpu)
 *   csrc/utils/geometry_utils.h                     (edge function, barycentrics, perspective
 *                                                    correction, point-segment distance)
 *   renderer/cameras.py::_get_sfm_calibration_matrix (screen-space intrinsics -> NDC)
 * PARITY UNPINNED for this file: the reference ships no tests or golden vectors for the
 * rasteriser, so it is cross-validated by a float64 build of itself (tie detection), float64
 * finite differences and geometric invariants only.  Three recalled semantics stay switchable:
 * perspective_correct, eps (kEpsilon) and the behind-camera rule (zcull_mode).
 */

#define FN2(a, b) a##_##b
#define FN1(a, b) FN2(a, b)
#define FN(name) FN1(name, SUFFIX)

static inline REAL FN(edge)(REAL px, REAL py, REAL ax, REAL ay, REAL bx, REAL by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* squared distance from p to segment ab (geometry_utils.h PointLineDistanceForward) */
static inline REAL FN(seg_dist)(REAL px, REAL py, REAL ax, REAL ay, REAL bx, REAL by, REAL eps) {
    REAL bax = bx - ax, bay = by - ay;
    REAL l2 = bax * bax + bay * bay;
    if (l2 <= eps) {
        REAL dx = px - bx, dy = py - by;
        return dx * dx + dy * dy;
    }
    REAL t = (bax * (px - ax) + bay * (py - ay)) / l2;
    if (t < 0) t = 0;
    if (t > 1) t = 1;
    REAL qx = ax + t * bax - px, qy = ay + t * bay - py;
    return qx * qx + qy * qy;
}

/* camera-space (x right, y down, z forward; mm) -> pytorch3d NDC with R=diag(-1,-1,1), T=0
 * (mano_layer.py:935-945).  fxn = fx/(W/2), pxn = -(px - W/2)/(W/2) (same for y with H).
 * out[v] = (x_ndc, y_ndc, z_view). */
void FN(orc_project)(const float* verts, int V, float fxn, float fyn, float pxn, float pyn, REAL* out) {
    for (int v = 0; v < V; ++v) {
        REAL x = (REAL)verts[3 * v], y = (REAL)verts[3 * v + 1], z = (REAL)verts[3 * v + 2];
        REAL xv = -x, yv = -y;
        out[3 * v] = ((REAL)fxn * xv + (REAL)pxn * z) / z;
        out[3 * v + 1] = ((REAL)fyn * yv + (REAL)pyn * z) / z;
        out[3 * v + 2] = z;
    }
}

/* Naive rasteriser evaluated at a separable grid of NDC sample points xs[nx] x ys[ny]
 * (a full S x S image is xs[i] = ys[i] = 1 - (2 i + 1)/S; a sample coordinate that is NaN
 * marks a crop pixel that reads zero padding and stays background).
 * Outputs (row-major [ny][nx]): pix_to_face (-1 bg), zbuf (-1), bary[3] (-1), dists (-1). */
void FN(orc_rasterize)(const REAL* vn, const int* faces, int F, const REAL* xs, int nx, const REAL* ys,
                       int ny, int perspective_correct, REAL eps, int zcull_mode, int* pix_to_face,
                       REAL* zbuf, REAL* bary, REAL* dists) {
    for (int yi = 0; yi < ny; ++yi) {
        for (int xi = 0; xi < nx; ++xi) {
            REAL px = xs[xi], py = ys[yi];
            int best_f = -1;
            REAL best_z = -1, bb0 = -1, bb1 = -1, bb2 = -1, best_d = -1;
            if (px == px && py == py) {
                for (int f = 0; f < F; ++f) {
                    const REAL* a = vn + 3 * faces[3 * f];
                    const REAL* b = vn + 3 * faces[3 * f + 1];
                    const REAL* c = vn + 3 * faces[3 * f + 2];
                    REAL x0 = a[0], y0 = a[1], z0 = a[2];
                    REAL x1 = b[0], y1 = b[1], z1 = b[2];
                    REAL x2 = c[0], y2 = c[1], z2 = c[2];
                    REAL zmin = z0 < z1 ? (z0 < z2 ? z0 : z2) : (z1 < z2 ? z1 : z2);
                    REAL zmax = z0 > z1 ? (z0 > z2 ? z0 : z2) : (z1 > z2 ? z1 : z2);
                    if (zcull_mode == 0 ? (zmin < eps) : (zmax < 0)) continue;
                    REAL xmin = x0 < x1 ? (x0 < x2 ? x0 : x2) : (x1 < x2 ? x1 : x2);
                    REAL xmax = x0 > x1 ? (x0 > x2 ? x0 : x2) : (x1 > x2 ? x1 : x2);
                    REAL ymin = y0 < y1 ? (y0 < y2 ? y0 : y2) : (y1 < y2 ? y1 : y2);
                    REAL ymax = y0 > y1 ? (y0 > y2 ? y0 : y2) : (y1 > y2 ? y1 : y2);
                    if (px > xmax || px < xmin || py > ymax || py < ymin) continue;
                    REAL farea = FN(edge)(x0, y0, x1, y1, x2, y2);
                    if (farea <= eps && farea >= -eps) continue;
                    REAL area = FN(edge)(x2, y2, x0, y0, x1, y1) + eps;
                    REAL w0 = FN(edge)(px, py, x1, y1, x2, y2) / area;
                    REAL w1 = FN(edge)(px, py, x2, y2, x0, y0) / area;
                    REAL w2 = FN(edge)(px, py, x0, y0, x1, y1) / area;
                    REAL b0 = w0, b1 = w1, b2 = w2;
                    if (perspective_correct) {
                        REAL t0 = w0 * z1 * z2;
                        REAL t1 = z0 * w1 * z2;
                        REAL t2 = z0 * z1 * w2;
                        REAL den = t0 + t1 + t2;
                        b0 = t0 / den;
                        b1 = t1 / den;
                        b2 = t2 / den;
                    }
                    REAL pz = b0 * z0 + b1 * z1 + b2 * z2;
                    if (pz < 0) continue;
                    int inside = b0 > 0 && b1 > 0 && b2 > 0;
                    if (!inside) continue; /* blur_radius == 0: dist >= 0 always rejects */
                    if (best_f < 0 || pz < best_z) { /* strict: lowest face index wins exact ties */
                        REAL d01 = FN(seg_dist)(px, py, x0, y0, x1, y1, eps);
                        REAL d02 = FN(seg_dist)(px, py, x0, y0, x2, y2, eps);
                        REAL d12 = FN(seg_dist)(px, py, x1, y1, x2, y2, eps);
                        REAL d = d01 < d02 ? d01 : d02;
                        d = d < d12 ? d : d12;
                        best_f = f;
                        best_z = pz;
                        bb0 = b0;
                        bb1 = b1;
                        bb2 = b2;
                        best_d = -d;
                    }
                }
            }
            int o = yi * nx + xi;
            pix_to_face[o] = best_f;
            zbuf[o] = best_z;
            if (bary) {
                bary[3 * o] = bb0;
                bary[3 * o + 1] = bb1;
                bary[3 * o + 2] = bb2;
            }
            if (dists) dists[o] = best_d;
        }
    }
}

/* RasterizeMeshesBackwardCpu for the only gradient DSF feeds (mano_layer.py:1084 keeps zbuf):
 * grad_zbuf -> grad of the NDC vertices (x_ndc, y_ndc, z), accumulated per vertex.
 * grad_bary (optional, may be NULL) is supported for completeness; grad_dists is not. */
void FN(orc_rasterize_backward)(const REAL* vn, const int* faces, const REAL* xs, int nx, const REAL* ys,
                                int ny, const int* pix_to_face, const REAL* grad_zbuf,
                                const REAL* grad_bary, int perspective_correct, REAL eps, REAL* grad_vn) {
    for (int yi = 0; yi < ny; ++yi) {
        for (int xi = 0; xi < nx; ++xi) {
            int o = yi * nx + xi;
            int f = pix_to_face[o];
            if (f < 0) continue;
            REAL px = xs[xi], py = ys[yi];
            int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
            const REAL *a = vn + 3 * i0, *b = vn + 3 * i1, *c = vn + 3 * i2;
            REAL x0 = a[0], y0 = a[1], z0 = a[2];
            REAL x1 = b[0], y1 = b[1], z1 = b[2];
            REAL x2 = c[0], y2 = c[1], z2 = c[2];
            REAL area = FN(edge)(x2, y2, x0, y0, x1, y1) + eps;
            REAL e0 = FN(edge)(px, py, x1, y1, x2, y2);
            REAL e1 = FN(edge)(px, py, x2, y2, x0, y0);
            REAL e2 = FN(edge)(px, py, x0, y0, x1, y1);
            REAL w0 = e0 / area, w1 = e1 / area, w2 = e2 / area;
            REAL b0 = w0, b1 = w1, b2 = w2;
            REAL t0 = 0, t1 = 0, t2 = 0, den = 1;
            if (perspective_correct) {
                t0 = w0 * z1 * z2;
                t1 = z0 * w1 * z2;
                t2 = z0 * z1 * w2;
                den = t0 + t1 + t2;
                b0 = t0 / den;
                b1 = t1 / den;
                b2 = t2 / den;
            }
            REAL gz = grad_zbuf ? grad_zbuf[o] : 0;
            REAL gb0 = gz * z0, gb1 = gz * z1, gb2 = gz * z2;
            if (grad_bary) {
                gb0 += grad_bary[3 * o];
                gb1 += grad_bary[3 * o + 1];
                gb2 += grad_bary[3 * o + 2];
            }
            REAL gz0 = gz * b0, gz1 = gz * b1, gz2 = gz * b2;
            REAL gw0 = gb0, gw1 = gb1, gw2 = gb2;
            if (perspective_correct) {
                REAL s = (gb0 * t0 + gb1 * t1 + gb2 * t2) / (den * den);
                REAL gt0 = gb0 / den - s, gt1 = gb1 / den - s, gt2 = gb2 / den - s;
                gw0 = gt0 * z1 * z2;
                gw1 = gt1 * z0 * z2;
                gw2 = gt2 * z0 * z1;
                gz0 += gt1 * w1 * z2 + gt2 * z1 * w2;
                gz1 += gt0 * w0 * z2 + gt2 * z0 * w2;
                gz2 += gt0 * w0 * z1 + gt1 * z0 * w1;
            }
            /* w_i = e_i / area */
            REAL ge0 = gw0 / area, ge1 = gw1 / area, ge2 = gw2 / area;
            REAL garea = -(gw0 * e0 + gw1 * e1 + gw2 * e2) / (area * area);
            REAL gx0 = 0, gy0 = 0, gx1 = 0, gy1 = 0, gx2 = 0, gy2 = 0;
            /* e0 = edge(p, v1, v2) */
            gx1 += ge0 * (py - y2); gy1 += ge0 * (x2 - px);
            gx2 += ge0 * (y1 - py); gy2 += ge0 * (px - x1);
            /* e1 = edge(p, v2, v0) */
            gx2 += ge1 * (py - y0); gy2 += ge1 * (x0 - px);
            gx0 += ge1 * (y2 - py); gy0 += ge1 * (px - x2);
            /* e2 = edge(p, v0, v1) */
            gx0 += ge2 * (py - y1); gy0 += ge2 * (x1 - px);
            gx1 += ge2 * (y0 - py); gy1 += ge2 * (px - x0);
            /* area = edge(v2, v0, v1) + eps: p := v2, a := v0, b := v1 */
            gx2 += garea * (y1 - y0); gy2 += -garea * (x1 - x0);
            gx0 += garea * (y2 - y1); gy0 += garea * (x1 - x2);
            gx1 += garea * (y0 - y2); gy1 += garea * (x2 - x0);
            grad_vn[3 * i0] += gx0; grad_vn[3 * i0 + 1] += gy0; grad_vn[3 * i0 + 2] += gz0;
            grad_vn[3 * i1] += gx1; grad_vn[3 * i1 + 1] += gy1; grad_vn[3 * i1 + 2] += gz1;
            grad_vn[3 * i2] += gx2; grad_vn[3 * i2 + 1] += gy2; grad_vn[3 * i2 + 2] += gz2;
        }
    }
}

/* chain grad of NDC vertices back to camera-space vertices through orc_project */
void FN(orc_project_backward)(const float* verts, int V, float fxn, float fyn, float pxn, float pyn,
                              const REAL* grad_vn, REAL* grad_verts) {
    (void)pxn; (void)pyn;
    for (int v = 0; v < V; ++v) {
        REAL x = (REAL)verts[3 * v], y = (REAL)verts[3 * v + 1], z = (REAL)verts[3 * v + 2];
        REAL gxn = grad_vn[3 * v], gyn = grad_vn[3 * v + 1], gzn = grad_vn[3 * v + 2];
        /* x_ndc = -fxn x / z + pxn */
        grad_verts[3 * v] = -gxn * (REAL)fxn / z;
        grad_verts[3 * v + 1] = -gyn * (REAL)fyn / z;
        grad_verts[3 * v + 2] = gzn + gxn * (REAL)fxn * x / (z * z) + gyn * (REAL)fyn * y / (z * z);
    }
}

#undef FN
#undef FN1
#undef FN2
