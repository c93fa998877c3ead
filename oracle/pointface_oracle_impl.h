/* TEST INFRASTRUCTURE ONLY - included twice by raster_oracle.c (REAL=float, REAL=double).
 *
 * CPU restatement of pytorch3d==0.4.0 `_C.point_face_dist_forward/backward`, which the reference
 * reaches from metric/meshLoss.py:52,63 (ICPLoss :347-353, JointICPLoss :377-394).  pytorch3d is
 * not vendored under /root/reference; the arithmetic follows its published sources
 * (csrc/point_mesh/point_mesh_cpu.cpp, csrc/utils/geometry_utils.h PointTriangle3Distance*,
 * PointLine3Distance*, BarycentricCoords3Forward, IsInsideTriangle) as recorded in SURVEY.md
 * section 8(a) row P1.  PARITY UNPINNED (no reference tests / golden vectors exist for it);
 * cross-validated against a float64 build, brute-force geometry and finite differences.
 * The mesh topology is shared by the whole batch (DSF always passes the same face list).
 */
#define FN2(a, b) a##_##b
#define FN1(a, b) FN2(a, b)
#define FN(name) FN1(name, SUFFIX)

typedef struct { REAL x, y, z; } FN(v3);
#define V3 FN(v3)

static inline V3 FN(sub)(V3 a, V3 b) { V3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline V3 FN(add)(V3 a, V3 b) { V3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline V3 FN(mul)(V3 a, REAL s) { V3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline REAL FN(dot)(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 FN(cross)(V3 a, V3 b) {
    V3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static inline V3 FN(ld)(const float* p) { V3 r = {(REAL)p[0], (REAL)p[1], (REAL)p[2]}; return r; }

static inline REAL FN(line3)(V3 p, V3 v0, V3 v1, REAL eps, REAL* tt_out) {
    V3 d = FN(sub)(v1, v0);
    REAL l2 = FN(dot)(d, d);
    if (l2 <= eps) {
        V3 q = FN(sub)(p, v1);
        *tt_out = -1;
        return FN(dot)(q, q);
    }
    REAL t = FN(dot)(d, FN(sub)(p, v0)) / l2;
    REAL tt = t < 0 ? 0 : (t > 1 ? 1 : t);
    V3 q = FN(sub)(p, FN(add)(v0, FN(mul)(d, tt)));
    *tt_out = tt;
    return FN(dot)(q, q);
}

/* returns squared distance; *branch = 0 plane, 1 edge v0v1, 2 edge v0v2, 3 edge v1v2 */
static inline REAL FN(point_tri)(V3 p, V3 v0, V3 v1, V3 v2, REAL eps, int* branch) {
    V3 n = FN(cross)(FN(sub)(v2, v0), FN(sub)(v1, v0));
    REAL nn = sqrt(FN(dot)(n, n));
    V3 nh = FN(mul)(n, (REAL)1 / (nn + eps));
    REAL t = FN(dot)(FN(sub)(v0, p), nh);
    V3 p0 = FN(add)(p, FN(mul)(nh, t));
    /* barycentric coordinates of the projection (BarycentricCoords3Forward) */
    V3 a = FN(sub)(v1, v0), b = FN(sub)(v2, v0), c = FN(sub)(p0, v0);
    REAL d00 = FN(dot)(a, a), d01 = FN(dot)(a, b), d11 = FN(dot)(b, b);
    REAL d20 = FN(dot)(c, a), d21 = FN(dot)(c, b);
    REAL den = d00 * d11 - d01 * d01 + eps;
    REAL w1 = (d11 * d20 - d01 * d21) / den;
    REAL w2 = (d00 * d21 - d01 * d20) / den;
    REAL w0 = (REAL)1 - w1 - w2;
    int inside = w0 >= 0 && w0 <= 1 && w1 >= 0 && w1 <= 1 && w2 >= 0 && w2 <= 1;
    if (inside && nn > eps) {
        *branch = 0;
        return t * t;
    }
    REAL tt;
    REAL e01 = FN(line3)(p, v0, v1, eps, &tt);
    REAL e02 = FN(line3)(p, v0, v2, eps, &tt);
    REAL e12 = FN(line3)(p, v1, v2, eps, &tt);
    REAL d = e01;
    *branch = 1;
    if (d > e02) { d = e02; *branch = 2; }
    if (d > e12) { d = e12; *branch = 3; }
    return d;
}

/* points (P,3), verts (V,3), faces (F,3) of ONE cloud/mesh pair -> dists (P), idxs (P).
 * Ties on equal distance: the lowest face index wins (strict < while scanning in order). */
void FN(orc_point_face_forward)(const float* points, int P, const float* verts, const int* faces, int F,
                                REAL eps, REAL* dists, int* idxs) {
    for (int i = 0; i < P; ++i) {
        V3 p = FN(ld)(points + 3 * i);
        REAL best = 0;
        int bi = -1, br;
        for (int f = 0; f < F; ++f) {
            REAL d = FN(point_tri)(p, FN(ld)(verts + 3 * faces[3 * f]), FN(ld)(verts + 3 * faces[3 * f + 1]),
                                   FN(ld)(verts + 3 * faces[3 * f + 2]), eps, &br);
            if (bi < 0 || d < best) { best = d; bi = f; }
        }
        dists[i] = best;
        idxs[i] = bi;
    }
}

static inline void FN(line3_bwd)(V3 p, V3 v0, V3 v1, REAL eps, REAL g, V3* gp, V3* g0, V3* g1) {
    REAL tt;
    (void)FN(line3)(p, v0, v1, eps, &tt);
    if (tt < 0) { /* degenerate segment: distance to v1 */
        V3 q = FN(mul)(FN(sub)(p, v1), 2 * g);
        *gp = FN(add)(*gp, q);
        *g1 = FN(sub)(*g1, q);
        return;
    }
    V3 d = FN(sub)(v1, v0);
    V3 q = FN(mul)(FN(sub)(p, FN(add)(v0, FN(mul)(d, tt))), 2 * g);   /* 2 g (p - proj) */
    *gp = FN(add)(*gp, q);
    *g0 = FN(sub)(*g0, FN(mul)(q, 1 - tt));
    *g1 = FN(sub)(*g1, FN(mul)(q, tt));
}

/* grad of dists[i] wrt its point and the three vertices of face idxs[i], accumulated. */
void FN(orc_point_face_backward)(const float* points, int P, const float* verts, const int* faces,
                                 const int* idxs, const REAL* grad_dists, REAL eps, REAL* grad_points,
                                 REAL* grad_verts) {
    for (int i = 0; i < P; ++i) {
        int f = idxs[i];
        if (f < 0) continue;
        int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        V3 p = FN(ld)(points + 3 * i), v0 = FN(ld)(verts + 3 * i0), v1 = FN(ld)(verts + 3 * i1),
           v2 = FN(ld)(verts + 3 * i2);
        REAL g = grad_dists[i];
        V3 z = {0, 0, 0}, gp = z, g0 = z, g1 = z, g2 = z;
        int br;
        (void)FN(point_tri)(p, v0, v1, v2, eps, &br);
        if (br == 0) {
            V3 e2 = FN(sub)(v2, v0), e1 = FN(sub)(v1, v0);
            V3 n = FN(cross)(e2, e1);
            REAL nn = sqrt(FN(dot)(n, n));
            V3 nh = FN(mul)(n, (REAL)1 / (nn + eps));
            V3 dv = FN(sub)(v0, p);
            REAL t = FN(dot)(dv, nh);
            REAL gt = 2 * t * g;
            gp = FN(mul)(nh, -gt);
            g0 = FN(mul)(nh, gt);
            V3 gnh = FN(mul)(dv, gt);
            /* nh = n / (|n| + eps) */
            REAL s = (REAL)1 / (nn + eps);
            REAL proj = FN(dot)(gnh, n) * s * s / (nn > 0 ? nn : 1);
            V3 gn = FN(sub)(FN(mul)(gnh, s), FN(mul)(n, proj));
            /* n = e2 x e1 : g_e2 = e1 x gn, g_e1 = gn x e2 */
            V3 ge2 = FN(cross)(e1, gn), ge1 = FN(cross)(gn, e2);
            g2 = FN(add)(g2, ge2);
            g1 = FN(add)(g1, ge1);
            g0 = FN(sub)(g0, FN(add)(ge2, ge1));
        } else if (br == 1) {
            FN(line3_bwd)(p, v0, v1, eps, g, &gp, &g0, &g1);
        } else if (br == 2) {
            FN(line3_bwd)(p, v0, v2, eps, g, &gp, &g0, &g2);
        } else {
            FN(line3_bwd)(p, v1, v2, eps, g, &gp, &g1, &g2);
        }
        if (grad_points) {
            grad_points[3 * i] += gp.x; grad_points[3 * i + 1] += gp.y; grad_points[3 * i + 2] += gp.z;
        }
        grad_verts[3 * i0] += g0.x; grad_verts[3 * i0 + 1] += g0.y; grad_verts[3 * i0 + 2] += g0.z;
        grad_verts[3 * i1] += g1.x; grad_verts[3 * i1 + 1] += g1.y; grad_verts[3 * i1 + 2] += g1.z;
        grad_verts[3 * i2] += g2.x; grad_verts[3 * i2 + 1] += g2.y; grad_verts[3 * i2 + 2] += g2.z;
    }
}

#undef V3
#undef FN
#undef FN1
#undef FN2
