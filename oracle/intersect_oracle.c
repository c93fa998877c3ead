/* TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's intersection-volume metric (row I1).
 *
 * Reference call sites: eval_coll.py:611-626 (self_intersection: per part `mesh.voxelized(pitch)`,
 * per non-adjacent pair s < t `s_mesh.contains(t_voxel.points)`, volume = count * pitch^3),
 * util/intersect.py:102-107 (intersect_vox, the same for an object / hand pair) and
 * eval_coll.py:348-373 (get_part_mesh: cap centre = mean of a vertex loop, appended after the mesh).
 *
 * The arithmetic lives in trimesh (un-vendored, unpinned; absent from this image), restated here from
 * its published algorithm:
 *   voxelized(pitch)  = voxelize_subdivide(edge_factor 2, max_iter 10): midpoint-subdivide every face
 *                       whose longest edge exceeds pitch / 2 (children decided again, level by level),
 *                       voxel index = round-half-even(vertex / pitch) over all resulting vertices,
 *                       duplicates removed, point = index * pitch.
 *   contains(points)  = only points inside the AABB; ray in the fixed direction
 *                       (0.4395064455, 0.617598629942, 0.652231566745) forwards and backwards, odd
 *                       crossing count = inside when both agree; disagreeing rays where one side hit
 *                       nothing are outside; the rest are re-cast once in a second direction (the
 *                       reference draws it at random; fixed here).
 * PARITY UNPINNED: the reference has no test or golden vector for this path and trimesh cannot be run
 * here; the restatement is checked against analytic cases in tests/ instead.
 * Everything is float64, like the reference (np.loadtxt meshes). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y, z; } d3;

static d3 mid(d3 a, d3 b) { d3 r = {(a.x + b.x) / 2, (a.y + b.y) / 2, (a.z + b.z) / 2}; return r; }
static double dist(d3 a, d3 b) {
    const double x = b.x - a.x, y = b.y - a.y, z = b.z - a.z;
    return sqrt(x * x + y * y + z * z);
}

typedef struct { long long* v; long n, cap; } KeyVec;
static void kv_push(KeyVec* k, long long a, long long b, long long c) {
    if (k->n + 3 > k->cap) {
        k->cap = k->cap ? 2 * k->cap : 4096;
        k->v = (long long*)realloc(k->v, sizeof(long long) * k->cap);
    }
    k->v[k->n++] = a; k->v[k->n++] = b; k->v[k->n++] = c;
}
static void push_vertex(KeyVec* k, d3 p, double pitch) {
    kv_push(k, (long long)nearbyint(p.x / pitch), (long long)nearbyint(p.y / pitch), (long long)nearbyint(p.z / pitch));
}

/* remesh.subdivide_to_size, one face: returns -1 when max_iter is exceeded (the reference raises) */
static int subdivide_face(KeyVec* k, d3 a, d3 b, d3 c, double max_edge, double pitch, int level, int max_iter) {
    const double e = fmax(dist(a, b), fmax(dist(b, c), dist(c, a)));
    if (!(e > max_edge)) {
        push_vertex(k, a, pitch); push_vertex(k, b, pitch); push_vertex(k, c, pitch);
        return 0;
    }
    if (level >= max_iter) return -1;
    const d3 ab = mid(a, b), bc = mid(b, c), ca = mid(c, a);
    if (subdivide_face(k, a, ab, ca, max_edge, pitch, level + 1, max_iter)) return -1;
    if (subdivide_face(k, ab, b, bc, max_edge, pitch, level + 1, max_iter)) return -1;
    if (subdivide_face(k, ca, bc, c, max_edge, pitch, level + 1, max_iter)) return -1;
    if (subdivide_face(k, ab, bc, ca, max_edge, pitch, level + 1, max_iter)) return -1;
    return 0;
}

static int cmp3(const void* pa, const void* pb) {
    const long long* a = (const long long*)pa; const long long* b = (const long long*)pb;
    for (int i = 0; i < 3; ++i) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}

/* signed crossing counts of the line p + t*dir with one triangle: adds to fwd (t > 0) or bwd (t < 0) */
static void ray_tri(d3 p, const double* dir, d3 a, d3 b, d3 c, long* fwd, long* bwd) {
    const double e1x = b.x - a.x, e1y = b.y - a.y, e1z = b.z - a.z;
    const double e2x = c.x - a.x, e2y = c.y - a.y, e2z = c.z - a.z;
    const double px = dir[1] * e2z - dir[2] * e2y, py = dir[2] * e2x - dir[0] * e2z, pz = dir[0] * e2y - dir[1] * e2x;
    const double det = e1x * px + e1y * py + e1z * pz;
    if (det == 0.0) return;
    const double tx = p.x - a.x, ty = p.y - a.y, tz = p.z - a.z;
    const double u = (tx * px + ty * py + tz * pz) / det;
    if (u < 0.0 || u > 1.0) return;
    const double qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    const double v = (dir[0] * qx + dir[1] * qy + dir[2] * qz) / det;
    if (v < 0.0 || u + v > 1.0) return;
    const double t = (e2x * qx + e2y * qy + e2z * qz) / det;
    if (t > 0.0) ++*fwd; else if (t < 0.0) ++*bwd;
}

static const double DIR0[3] = {0.4395064455, 0.617598629942, 0.652231566745};
static const double DIR1[3] = {-0.617598629942, 0.652231566745, 0.4395064455};

static int contains(d3 p, const d3* wv, const int* faces, int nf, const double* lo, const double* hi) {
    if (p.x < lo[0] || p.x > hi[0] || p.y < lo[1] || p.y > hi[1] || p.z < lo[2] || p.z > hi[2]) return 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        long fwd = 0, bwd = 0;
        const double* dir = attempt ? DIR1 : DIR0;
        for (int f = 0; f < nf; ++f) ray_tri(p, dir, wv[faces[3 * f]], wv[faces[3 * f + 1]], wv[faces[3 * f + 2]], &fwd, &bwd);
        const int cf = fwd & 1, cb = bwd & 1;
        if (cf == cb) return cf;
        if (fwd == 0 || bwd == 0) return 0;
    }
    return 0;
}

/* one hand / one scene.  verts (n_verts,3) float32; caps: centre c = mean(verts[cap_idx[cap_ptr[c]..cap_ptr[c+1])])
 * becomes water vertex n_verts + c; parts: faces part_faces[part_ptr[p]..part_ptr[p+1]) index the water mesh;
 * pair_mask[s*n_parts+t] != 0 -> count the voxel points of t inside s.  Returns the volume, or -1 on
 * max_iter overflow.  pair_counts (n_parts*n_parts) and voxel_counts (n_parts) optional. */
double orc_intersect_vox(const float* verts, int n_verts, int n_caps, const int* cap_ptr, const int* cap_idx,
                         int n_parts, const int* part_ptr, const int* part_faces, const unsigned char* pair_mask,
                         double pitch, long long* pair_counts, long long* voxel_counts) {
    const int nw = n_verts + n_caps;
    d3* wv = (d3*)malloc(sizeof(d3) * nw);
    for (int i = 0; i < n_verts; ++i) { wv[i].x = verts[3 * i]; wv[i].y = verts[3 * i + 1]; wv[i].z = verts[3 * i + 2]; }
    for (int c = 0; c < n_caps; ++c) {
        double sx = 0, sy = 0, sz = 0;
        const int n = cap_ptr[c + 1] - cap_ptr[c];
        for (int k = cap_ptr[c]; k < cap_ptr[c + 1]; ++k) { sx += wv[cap_idx[k]].x; sy += wv[cap_idx[k]].y; sz += wv[cap_idx[k]].z; }
        wv[n_verts + c].x = sx / n; wv[n_verts + c].y = sy / n; wv[n_verts + c].z = sz / n;
    }
    KeyVec* vox = (KeyVec*)calloc(n_parts, sizeof(KeyVec));
    long* nvox = (long*)calloc(n_parts, sizeof(long));
    double* lo = (double*)malloc(sizeof(double) * 3 * n_parts);
    double* hi = (double*)malloc(sizeof(double) * 3 * n_parts);
    int bad = 0;
    for (int p = 0; p < n_parts; ++p) {
        for (int c = 0; c < 3; ++c) { lo[3 * p + c] = INFINITY; hi[3 * p + c] = -INFINITY; }
        for (int f = part_ptr[p]; f < part_ptr[p + 1]; ++f) {
            for (int k = 0; k < 3; ++k) {
                const d3 q = wv[part_faces[3 * f + k]];
                lo[3 * p] = fmin(lo[3 * p], q.x); hi[3 * p] = fmax(hi[3 * p], q.x);
                lo[3 * p + 1] = fmin(lo[3 * p + 1], q.y); hi[3 * p + 1] = fmax(hi[3 * p + 1], q.y);
                lo[3 * p + 2] = fmin(lo[3 * p + 2], q.z); hi[3 * p + 2] = fmax(hi[3 * p + 2], q.z);
            }
            if (subdivide_face(&vox[p], wv[part_faces[3 * f]], wv[part_faces[3 * f + 1]], wv[part_faces[3 * f + 2]],
                               pitch / 2.0, pitch, 0, 10)) bad = 1;
        }
        qsort(vox[p].v, vox[p].n / 3, 3 * sizeof(long long), cmp3);
        long m = 0;
        for (long i = 0; i < vox[p].n / 3; ++i)
            if (i == 0 || cmp3(vox[p].v + 3 * i, vox[p].v + 3 * (i - 1)) != 0) {
                memmove(vox[p].v + 3 * m, vox[p].v + 3 * i, 3 * sizeof(long long));
                ++m;
            }
        nvox[p] = m;
        if (voxel_counts) voxel_counts[p] = m;
    }
    long long total = 0;
    for (int s = 0; s < n_parts && !bad; ++s)
        for (int t = 0; t < n_parts; ++t) {
            if (pair_counts) pair_counts[s * n_parts + t] = 0;
            if (!pair_mask[s * n_parts + t]) continue;
            long long cnt = 0;
            for (long i = 0; i < nvox[t]; ++i) {
                const d3 q = {vox[t].v[3 * i] * pitch, vox[t].v[3 * i + 1] * pitch, vox[t].v[3 * i + 2] * pitch};
                cnt += contains(q, wv, part_faces + 3 * part_ptr[s], part_ptr[s + 1] - part_ptr[s], lo + 3 * s, hi + 3 * s);
            }
            if (pair_counts) pair_counts[s * n_parts + t] = cnt;
            total += cnt;
        }
    for (int p = 0; p < n_parts; ++p) free(vox[p].v);
    free(vox); free(nvox); free(lo); free(hi); free(wv);
    return bad ? -1.0 : (double)total * pitch * pitch * pitch;
}

void orc_batch_intersect_vox(int batch, const float* verts, int n_verts, int n_caps, const int* cap_ptr,
                             const int* cap_idx, int n_parts, const int* part_ptr, const int* part_faces,
                             const unsigned char* pair_mask, double pitch, long long* pair_counts,
                             long long* voxel_counts, double* volume) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < batch; ++b)
        volume[b] = orc_intersect_vox(verts + (long)b * n_verts * 3, n_verts, n_caps, cap_ptr, cap_idx, n_parts, part_ptr,
                                      part_faces, pair_mask, pitch,
                                      pair_counts ? pair_counts + (long)b * n_parts * n_parts : 0,
                                      voxel_counts ? voxel_counts + (long)b * n_parts : 0);
}
