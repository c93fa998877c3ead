/* CPU re-statement of phase A/B of raster_fwd_kernel (dsf_b200/csrc/raster.cu): the conservative pixel run of a
 * face on a row, checked against the exact oracle-order edge-sign test on random triangles.  The claim under test:
 * the closed-form run never misses a pixel the exact test accepts (false positives are allowed and counted).
 * The GPU's approximate division (__fdividef, <= 2 ulp) is modelled by perturbing the slope by +-4 ulp.
 * Build: gcc -O2 -ffp-contract=off.  Test infrastructure only. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static float edge_rn(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);   /* each op rounds once: -ffp-contract=off */
}
static uint64_t rng_state;
static double urand(void) {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}
static float pix_to_ndc(int i, int n) { return -1.0f + (2.0f * (float)(n - 1 - i) + 1.0f) / (float)n; }

/* returns the number of missed pixels; accumulates statistics */
long check(int R, long n_tri, uint64_t seed, int ulp_shift, long* n_exact, long* n_cand, long* n_rows) {
    rng_state = seed * 0x9E3779B97F4A7C15ull + 12345;
    float* xs = (float*)malloc(sizeof(float) * R);
    for (int i = 0; i < R; ++i) xs[i] = pix_to_ndc(i, R);          /* xs == ys in direct mode */
    const float ax = -(float)R * 0.5f, bx = ((float)R - 1.f) * 0.5f;
    long missed = 0;
    for (long t = 0; t < n_tri; ++t) {
        /* triangle: centre anywhere in (and slightly around) the image, size from sub-pixel to half the image,
         * with a share of slivers and of edges that are nearly horizontal / vertical */
        const double cx = urand() * 2.6 - 1.3, cy = urand() * 2.6 - 1.3;
        const double size = pow(10.0, -3.0 + 2.9 * urand());
        float x[3], y[3];
        const int kind = (int)(urand() * 6);
        for (int k = 0; k < 3; ++k) { x[k] = (float)(cx + (urand() - 0.5) * size); y[k] = (float)(cy + (urand() - 0.5) * size); }
        if (kind == 0) y[1] = y[0] + (float)((urand() - 0.5) * 1e-6);            /* near-horizontal edge */
        if (kind == 1) x[2] = x[1] + (float)((urand() - 0.5) * 1e-6);            /* near-vertical edge */
        if (kind == 2) { x[2] = (float)(x[0] + (x[1] - x[0]) * 0.5 + (urand() - 0.5) * 1e-5 * size);   /* sliver */
                         y[2] = (float)(y[0] + (y[1] - y[0]) * 0.5 + (urand() - 0.5) * 1e-5 * size); }
        if (kind == 3) y[1] = y[0];                                                /* exactly horizontal edge */
        if (kind == 4) { x[1] = (float)(cx + (urand() - 0.5) * 2000.0 * urand());   /* a vertex far outside the image */
                         y[1] = (float)(cy + (urand() - 0.5) * 2000.0 * urand()); }
        const float x0 = x[0], y0 = y[0], x1 = x[1], y1 = y[1], x2 = x[2], y2 = y[2];
        const float farea = edge_rn(x0, y0, x1, y1, x2, y2);
        if (farea <= 1e-8f && farea >= -1e-8f) continue;                            /* culled like the kernel */
        const float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
        const float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
        int ia = 0, ib = R - 1, ja = 0, jb = R - 1;                                   /* exact bbox on the sample grid */
        while (ia < R && !(xs[ia] <= xmax)) ++ia;
        while (ib >= 0 && !(xs[ib] >= xmin)) --ib;
        while (ja < R && !(xs[ja] <= ymax)) ++ja;
        while (jb >= 0 && !(xs[jb] >= ymin)) --jb;
        if (ia > ib || ja > jb) continue;
        /* phase A: edge lines */
        float e_m[3], e_c[3];
        int use[3] = {0, 0, 0}, low[3] = {0, 0, 0};
        if (fabsf(farea) >= 1e-5f) {
            const float sg = farea > 0.f ? 1.f : -1.f;
            const float xa[3] = {x1, x2, x0}, ya[3] = {y1, y2, y0}, xb[3] = {x2, x0, x1}, yb[3] = {y2, y0, y1};
            for (int i = 0; i < 3; ++i) {
                const float dy = yb[i] - ya[i];
                float m = (xb[i] - xa[i]) / dy;
                m = m * (1.0f + (float)ulp_shift * 1.1920929e-7f);                  /* approximate division */
                const float am = fabsf(m);
                if (am < 1e30f) {
                    const float dl = 2e-6f * (1.f + fabsf(xa[i]) + (2.f + fabsf(ya[i])) * am);
                    low[i] = sg * dy > 0.f;
                    e_m[i] = m;
                    e_c[i] = fmaf(-ya[i], m, xa[i]) + (low[i] ? -dl : dl);
                    use[i] = 1;
                }
            }
        }
        for (int j = ja; j <= jb; ++j) {
            const float py = xs[j];
            float lo = -INFINITY, hi = INFINITY;
            for (int i = 0; i < 3; ++i)
                if (use[i]) {
                    const float v = fmaf(e_m[i], py, e_c[i]);
                    if (low[i]) lo = fmaxf(lo, v); else hi = fminf(hi, v);
                }
            float fa = ceilf(fmaf(ax, hi, bx) - 1e-3f), fb = floorf(fmaf(ax, lo, bx) + 1e-3f);
            int ka = fa < -2e9f ? ia : (fa > 2e9f ? R : (int)fa), kb = fb > 2e9f ? ib : (fb < -2e9f ? -1 : (int)fb);
            if (ka < ia) ka = ia;
            if (kb > ib) kb = ib;
            ++*n_rows;
            if (kb >= ka) *n_cand += kb - ka + 1;
            for (int k = ia; k <= ib; ++k) {
                const float px = xs[k];
                const float e0 = edge_rn(px, py, x1, y1, x2, y2), e1 = edge_rn(px, py, x2, y2, x0, y0),
                            e2 = edge_rn(px, py, x0, y0, x1, y1);
                const int in = (e0 > 0.f && e1 > 0.f && e2 > 0.f) || (e0 < 0.f && e1 < 0.f && e2 < 0.f);
                if (in) {
                    ++*n_exact;
                    if (k < ka || k > kb) ++missed;
                }
            }
        }
    }
    free(xs);
    return missed;
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 200000;
    long total_missed = 0;
    for (int R = 128; R <= 256; R *= 2)
        for (int shift = -4; shift <= 4; shift += 4) {
            long n_exact = 0, n_cand = 0, n_rows = 0;
            const long missed = check(R, n, 7 + R + shift, shift, &n_exact, &n_cand, &n_rows);
            printf("R=%d ulp_shift=%+d rows=%ld exact=%ld candidates=%ld missed=%ld\n", R, shift, n_rows, n_exact, n_cand, missed);
            total_missed += missed;
        }
    printf("TOTAL_MISSED %ld\n", total_missed);
    return total_missed ? 1 : 0;
}
