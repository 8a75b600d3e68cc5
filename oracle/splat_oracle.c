/*
 * splat_oracle.c -- plain-C CPU port of the GaussianRenderer.render hot path.
 * TEST INFRASTRUCTURE ONLY: used by tests/ (as a checker at sizes the torch oracle is too slow
 * for) and by bench.py's cpu_baseline / --impl reference legs.  Never linked into the product.
 *
 * Follows the reference (Loveof1ife7/mini-3d-gaussian-splatting) function by function:
 *   oracle_project      src/core/renderer.py:117-220 + src/core/gaussian_model.py:113-122,200-207
 *                       + src/utils/math_utils.py:9-26 (+ colour/opacity activations renderer.py:88-94)
 *   oracle_bin          src/core/renderer.py:222-239 (depth sort) and :263-298 (tile lists)
 *   oracle_raster_fwd   src/core/renderer.py:300-367 -- the literal scalar loop, tile by tile, pixel by pixel
 *   oracle_raster_bwd   reverse-mode derivative of that loop, textbook back-to-front walk
 *   oracle_project_bwd  reverse-mode derivative of oracle_project
 * fp32 arithmetic with the reference's operation order; built with -ffp-contract=off so the
 * compiler introduces no FMAs, and fmaf() is called explicitly where the reference's BLAS does.
 * Two documented substitutions: the 2x2 inverse and largest eigenvalue use closed
 * forms instead of LAPACK (conic within 3e-7 relative; int(radius) identical).
 * Parity pin: tests/test_oracle_c.py compares it DIRECTLY with outputs of the literal reference (tests/golden/: five
 * frames with the reference's autograd gradients up to 1 000 splats, BASELINE configs[0], a 192x128 frame where two
 * thirds of the pixels terminate early, stages 1-3 at 1080p on 200 k splats) and with oracle/splat_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference: use all host cores even when a launcher (torchrun) exported OMP_NUM_THREADS=1 */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* cam: Rv[9] Tv[3] fx fy cx cy (16 floats), as include/gsplat_b200.h */
typedef struct {
    float S[9], R[9], sig[3], q[4], qn;
    float X, Y, Z, iz, mx, my, J00, J02, J11, J12, M[9];
    float a, b01, b10, c, q00, q01, q10, q11, radius;
} Splat;

static void cov_from_params(const float* sl, const float* rot, Splat* g) {
    for (int k = 0; k < 3; ++k) g->sig[k] = expf(sl[k]);
    float n = sqrtf(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2] + rot[3] * rot[3]);
    if (n < 1e-12f) n = 1e-12f;
    g->qn = n;
    float w = rot[0] / n, x = rot[1] / n, y = rot[2] / n, z = rot[3] / n;
    float n2 = sqrtf(w * w + x * x + y * y + z * z);
    if (n2 < 1e-12f) n2 = 1e-12f;
    w /= n2; x /= n2; y /= n2; z /= n2;
    g->q[0] = w; g->q[1] = x; g->q[2] = y; g->q[3] = z;
    float* R = g->R;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 3; ++k) acc += R[r * 3 + k] * (g->sig[k] * g->sig[k]) * R[c * 3 + k];
            g->S[r * 3 + c] = acc;
        }
}

static void project_one(const float* cam, const float* p, Splat* g, float rmin, float rmax) {
    const float* Rv = cam; const float* Tv = cam + 9;
    const float fx = cam[12], fy = cam[13], cx = cam[14], cy = cam[15];
    /* MKL's K=3 accumulation: product, then two FMAs; `+ Tv` is a separate add (renderer.py:154) */
    g->X = fmaf(p[2], Rv[2], fmaf(p[1], Rv[1], p[0] * Rv[0])) + Tv[0];
    g->Y = fmaf(p[2], Rv[5], fmaf(p[1], Rv[4], p[0] * Rv[3])) + Tv[1];
    g->Z = fmaf(p[2], Rv[8], fmaf(p[1], Rv[7], p[0] * Rv[6])) + Tv[2];
    g->mx = fx * g->X / g->Z + cx;                       /* renderer.py:161 */
    g->my = -fy * g->Y / g->Z + cy;                      /* renderer.py:162 */
    float tmp[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            tmp[r * 3 + c] = Rv[r * 3] * g->S[c] + Rv[r * 3 + 1] * g->S[3 + c] + Rv[r * 3 + 2] * g->S[6 + c];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            g->M[r * 3 + c] = tmp[r * 3] * Rv[c * 3] + tmp[r * 3 + 1] * Rv[c * 3 + 1] + tmp[r * 3 + 2] * Rv[c * 3 + 2];
    g->iz = 1.0f / g->Z;
    g->J00 = fx * g->iz;
    g->J02 = -fx * g->X * g->iz * g->iz;
    g->J11 = -fy * g->iz;
    g->J12 = fy * g->Y * g->iz * g->iz;
    const float* M = g->M;
    float u0 = g->J00 * M[0] + g->J02 * M[6], u1 = g->J00 * M[1] + g->J02 * M[7], u2 = g->J00 * M[2] + g->J02 * M[8];
    float v0 = g->J11 * M[3] + g->J12 * M[6], v1 = g->J11 * M[4] + g->J12 * M[7], v2 = g->J11 * M[5] + g->J12 * M[8];
    g->a = (u0 * g->J00 + u2 * g->J02) + 1e-6f;
    g->b01 = u1 * g->J11 + u2 * g->J12;
    g->b10 = v0 * g->J00 + v2 * g->J02;
    g->c = (v1 * g->J11 + v2 * g->J12) + 1e-6f;
    float det = g->a * g->c - g->b01 * g->b10;
    g->q00 = g->c / det; g->q01 = -g->b01 / det; g->q10 = -g->b10 / det; g->q11 = g->a / det;
    float mid = 0.5f * (g->a + g->c), hd = 0.5f * (g->a - g->c);
    float r = 3.0f * sqrtf(mid + sqrtf(hd * hd + g->b10 * g->b10));
    if (r == r) { if (r < rmin) r = rmin; if (r > rmax) r = rmax; }
    g->radius = r;
}

/* outputs as gs_project_fwd; cov3d may be NULL (parameter mode) or given (covariance mode) */
void oracle_project(int64_t n, const float* xyz, const float* scaling_log, const float* rotation, const float* cov3d,
                    const float* opacity, int opacity_is_logit, const float* feat0, int64_t feat_stride,
                    const float* cam, int W, int H, float rmin, float rmax,
                    float* means2d, float* depths, float* conics, float* radii, float* colors, float* opac,
                    uint8_t* vis, int32_t* tiles_touched, int32_t* rect /* tx0,ty0,tx1,ty1 */) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        Splat g;
        if (cov3d) memcpy(g.S, cov3d + i * 9, 9 * sizeof(float));
        else cov_from_params(scaling_log + i * 3, rotation + i * 4, &g);
        project_one(cam, xyz + i * 3, &g, rmin, rmax);
        means2d[i * 2] = g.mx; means2d[i * 2 + 1] = g.my; depths[i] = g.Z;
        conics[i * 4] = g.q00; conics[i * 4 + 1] = g.q01; conics[i * 4 + 2] = g.q10; conics[i * 4 + 3] = g.q11;
        radii[i] = g.radius;
        for (int c = 0; c < 3; ++c) colors[i * 3 + c] = sigmoidf_(feat0[i * feat_stride + c]);
        opac[i] = opacity_is_logit ? sigmoidf_(opacity[i]) : opacity[i];
        const float r = g.radius;
        int v = (g.Z > 0) && (g.mx >= -r) && (g.mx < (float)W + r) && (g.my >= -r) && (g.my < (float)H + r) && (r > 0);
        vis[i] = (uint8_t)v;
        int cnt = 0, tx0 = 0, ty0 = 0, tx1 = 0, ty1 = 0;
        if (v) {                                               /* renderer.py:278-293 */
            int ir = (int)r, ix = (int)g.mx, iy = (int)g.my;
            int x0 = ix - ir < 0 ? 0 : ix - ir, x1 = ix + 1 + ir > W ? W : ix + 1 + ir;
            int y0 = iy - ir < 0 ? 0 : iy - ir, y1 = iy + 1 + ir > H ? H : iy + 1 + ir;
            if (x0 < x1 && y0 < y1) {
                tx0 = x0 / TILE; tx1 = (x1 - 1) / TILE; ty0 = y0 / TILE; ty1 = (y1 - 1) / TILE;
                cnt = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
            }
        }
        tiles_touched[i] = cnt;
        rect[i * 4] = tx0; rect[i * 4 + 1] = ty0; rect[i * 4 + 2] = tx1; rect[i * 4 + 3] = ty1;
    }
}

/* ---- depth sort + tile lists ------------------------------------------------------------- */
typedef struct { uint32_t key; int32_t id; } KeyId;

static void radix_sort_keyid(KeyId* a, KeyId* tmp, int64_t n) {   /* LSD, stable */
    for (int pass = 0; pass < 4; ++pass) {
        int64_t hist[257] = {0};
        const int sh = pass * 8;
        for (int64_t i = 0; i < n; ++i) hist[((a[i].key >> sh) & 255) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (int64_t i = 0; i < n; ++i) tmp[hist[(a[i].key >> sh) & 255]++] = a[i];
        KeyId* t = a; a = tmp; tmp = t;
    }
}

/* Returns D.  sorted_ids[V'] (splats with tiles, depth order, ties by ascending id);
 * entry_ids[D] grouped by tile; ranges[num_tiles*2]; call with entry_ids == NULL to get D only. */
int64_t oracle_bin(int64_t n, const float* depths, const int32_t* tiles_touched, const int32_t* rect, int W, int H,
                   int32_t* sorted_ids, int64_t* num_sorted_out, int32_t* entry_ids, int32_t* ranges) {
    const int tiles_x = (W + TILE - 1) / TILE, tiles_y = (H + TILE - 1) / TILE;
    const int num_tiles = tiles_x * tiles_y;
    KeyId* a = (KeyId*)malloc(sizeof(KeyId) * (size_t)(n > 0 ? n : 1));
    KeyId* t = (KeyId*)malloc(sizeof(KeyId) * (size_t)(n > 0 ? n : 1));
    int64_t m = 0, D = 0;
    for (int64_t i = 0; i < n; ++i)
        if (tiles_touched[i] > 0) {
            uint32_t bits; memcpy(&bits, depths + i, 4);
            a[m].key = bits; a[m].id = (int32_t)i; ++m;
            D += tiles_touched[i];
        }
    radix_sort_keyid(a, t, m);           /* 4 passes: result back in `a` */
    *num_sorted_out = m;
    for (int64_t j = 0; j < m; ++j) sorted_ids[j] = a[j].id;
    if (entry_ids) {
        int64_t* count = (int64_t*)calloc((size_t)num_tiles + 1, sizeof(int64_t));
        for (int64_t j = 0; j < m; ++j) {
            const int32_t* r = rect + (int64_t)a[j].id * 4;
            for (int ty = r[1]; ty <= r[3]; ++ty)
                for (int tx = r[0]; tx <= r[2]; ++tx) count[ty * tiles_x + tx + 1]++;
        }
        for (int k = 0; k < num_tiles; ++k) count[k + 1] += count[k];
        for (int k = 0; k < num_tiles; ++k) { ranges[2 * k] = (int32_t)count[k]; ranges[2 * k + 1] = (int32_t)count[k + 1]; }
        for (int64_t j = 0; j < m; ++j) {            /* append in global depth order (renderer.py:277-298) */
            const int32_t* r = rect + (int64_t)a[j].id * 4;
            for (int ty = r[1]; ty <= r[3]; ++ty)
                for (int tx = r[0]; tx <= r[2]; ++tx) entry_ids[count[ty * tiles_x + tx]++] = a[j].id;
        }
        free(count);
    }
    free(a); free(t);
    return D;
}

/* ---- compositing ---------------------------------------------------------------------------- */
typedef struct { float dx, dy, e, w, a, contrib; } Eval;

static int eval_splat(float px, float py, const float* m2, const float* q, float op, float A, Eval* ev) {
    ev->dx = px - m2[0];
    ev->dy = py - m2[1];
    float s = ev->dx * ev->dx * q[0] + (q[1] + q[2]) * ev->dx * ev->dy + ev->dy * ev->dy * q[3];   /* renderer.py:333 */
    ev->e = expf(-0.5f * s);
    ev->w = ev->e < 0.f ? 0.f : (ev->e > 1.f ? 1.f : ev->e);
    if (ev->w < 1e-5f) return 0;
    float u = op * ev->w;
    ev->a = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);
    if (ev->a <= 0.f) return 0;
    ev->contrib = (1.0f - A) * ev->a;
    if (ev->contrib <= 0.f) return 0;
    return 1;
}

/* tile_first/tile_count select a contiguous block of tiles (bounded samples for the CPU baseline);
 * pixels of other tiles are left untouched.  pix_state [H*W*4] = C(3 incl. bg), Dsum. */
void oracle_raster_fwd(int W, int H, const int32_t* entry_ids, const int32_t* ranges, const float* means2d,
                       const float* conics, const float* depths, const float* colors, const float* opac,
                       const float* bg, int any_visible, int tile_first, int tile_count,
                       float* image, float* alpha, float* depth, float* pix_state, int32_t* n_consumed) {
    const int tiles_x = (W + TILE - 1) / TILE;
    const int64_t plane = (int64_t)W * H;
#pragma omp parallel for schedule(dynamic, 4)
    for (int tile = tile_first; tile < tile_first + tile_count; ++tile) {
        const int tx = tile % tiles_x, ty = tile / tiles_x;
        const int s0 = ranges[2 * tile], s1 = ranges[2 * tile + 1];
        for (int yy = ty * TILE; yy < ty * TILE + TILE && yy < H; ++yy)
            for (int xx = tx * TILE; xx < tx * TILE + TILE && xx < W; ++xx) {
                float A = 0.f, C[3] = {bg[0], bg[1], bg[2]}, Ds = 0.f;
                int pos = s0;
                for (; pos < s1; ++pos) {
                    const int i = entry_ids[pos];
                    Eval ev;
                    if (!eval_splat((float)xx, (float)yy, means2d + 2 * (int64_t)i, conics + 4 * (int64_t)i, opac[i], A, &ev))
                        continue;
                    for (int c = 0; c < 3; ++c) C[c] += ev.contrib * colors[3 * (int64_t)i + c];
                    A = A + ev.contrib;
                    Ds += ev.contrib * depths[i];
                    if (A >= 0.995f) { ++pos; break; }
                }
                const int64_t p = (int64_t)yy * W + xx;
                if (any_visible) {
                    for (int c = 0; c < 3; ++c) {
                        float v = C[c] + (1.0f - A) * bg[c];
                        image[c * plane + p] = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
                    }
                    alpha[p] = A < 0.f ? 0.f : (A > 1.f ? 1.f : A);
                    depth[p] = Ds / (A + 1e-6f);
                } else {
                    for (int c = 0; c < 3; ++c) image[c * plane + p] = bg[c];
                    alpha[p] = 0.f; depth[p] = 0.f;
                }
                if (pix_state) { pix_state[4 * p] = C[0]; pix_state[4 * p + 1] = C[1]; pix_state[4 * p + 2] = C[2]; pix_state[4 * p + 3] = Ds; }
                if (n_consumed) n_consumed[p] = pos - s0;
            }
    }
}

/* Back-to-front reverse pass per pixel (double accumulators per pixel, fp32 forward replay).
 * Accumulates into g_means2d[n*2], g_conics[n*4], g_depths[n], g_colors[n*3], g_opac[n] (double,
 * caller-zeroed) -- one private copy per thread would cost too much memory at 1M splats, so the
 * accumulation uses atomics on doubles. */
void oracle_raster_bwd(int W, int H, const int32_t* entry_ids, const int32_t* ranges, const float* means2d,
                       const float* conics, const float* depths, const float* colors, const float* opac,
                       const float* bg, int tile_first, int tile_count,
                       const float* g_image, const float* g_alpha, const float* g_depth,
                       double* g_means2d, double* g_conics, double* g_depths, double* g_colors, double* g_opac) {
    const int tiles_x = (W + TILE - 1) / TILE;
    const int64_t plane = (int64_t)W * H;
#pragma omp parallel
    {
        int cap = 1024;
        int* kid = (int*)malloc(sizeof(int) * cap);
        Eval* kev = (Eval*)malloc(sizeof(Eval) * cap);
        float* kT = (float*)malloc(sizeof(float) * cap);
#pragma omp for schedule(dynamic, 4)
        for (int tile = tile_first; tile < tile_first + tile_count; ++tile) {
            const int tx = tile % tiles_x, ty = tile / tiles_x;
            const int s0 = ranges[2 * tile], s1 = ranges[2 * tile + 1];
            for (int yy = ty * TILE; yy < ty * TILE + TILE && yy < H; ++yy)
                for (int xx = tx * TILE; xx < tx * TILE + TILE && xx < W; ++xx) {
                    /* forward replay, recording contributors */
                    float A = 0.f, C[3] = {bg[0], bg[1], bg[2]}, Ds = 0.f;
                    int m = 0;
                    for (int pos = s0; pos < s1; ++pos) {
                        const int i = entry_ids[pos];
                        Eval ev;
                        if (!eval_splat((float)xx, (float)yy, means2d + 2 * (int64_t)i, conics + 4 * (int64_t)i, opac[i], A, &ev))
                            continue;
                        if (m == cap) {
                            cap *= 2;
                            kid = (int*)realloc(kid, sizeof(int) * cap);
                            kev = (Eval*)realloc(kev, sizeof(Eval) * cap);
                            kT = (float*)realloc(kT, sizeof(float) * cap);
                        }
                        kid[m] = i; kev[m] = ev; kT[m] = 1.0f - A; ++m;
                        for (int c = 0; c < 3; ++c) C[c] += ev.contrib * colors[3 * (int64_t)i + c];
                        A = A + ev.contrib;
                        Ds += ev.contrib * depths[i];
                        if (A >= 0.995f) break;
                    }
                    if (m == 0) continue;
                    const int64_t p = (int64_t)yy * W + xx;
                    double gC[3];
                    for (int c = 0; c < 3; ++c) {
                        float pre = C[c] + (1.0f - A) * bg[c];
                        gC[c] = (pre >= 0.f && pre <= 1.f) ? (double)g_image[c * plane + p] : 0.0;
                    }
                    const double den = (double)(A + 1e-6f);
                    const double gDs = (double)g_depth[p] / den;
                    double gA = ((A >= 0.f && A <= 1.f) ? (double)g_alpha[p] : 0.0) - (double)g_depth[p] * (double)Ds / (den * den);
                    for (int c = 0; c < 3; ++c) gA -= gC[c] * (double)bg[c];
                    /* back to front: S = sum over later contributors of T_j a_j v_j */
                    double S = 0.0;
                    for (int k = m - 1; k >= 0; --k) {
                        const int64_t i = kid[k];
                        const Eval* ev = &kev[k];
                        const double T = kT[k], a = ev->a;
                        const double v = gC[0] * colors[3 * i] + gC[1] * colors[3 * i + 1] + gC[2] * colors[3 * i + 2] + gDs * depths[i] + gA;
                        const double g_a = T * v - (S == 0.0 ? 0.0 : S / (1.0 - a));
                        S += T * a * v;
                        const double ta = T * a;
                        double add_col[3] = {ta * gC[0], ta * gC[1], ta * gC[2]};
                        double add_z = ta * gDs, add_op = 0.0, g_s = 0.0;
                        const float u = opac[i] * ev->w;
                        if (u >= 0.f && u <= 1.f) {
                            add_op = g_a * ev->w;
                            if (ev->e <= 1.f) g_s = -0.5 * ev->w * (g_a * opac[i]);
                        }
                        const double dx = ev->dx, dy = ev->dy;
                        const float* q = conics + 4 * i;
                        const double qs = (double)q[1] + (double)q[2];
#pragma omp atomic
                        g_colors[3 * i] += add_col[0];
#pragma omp atomic
                        g_colors[3 * i + 1] += add_col[1];
#pragma omp atomic
                        g_colors[3 * i + 2] += add_col[2];
#pragma omp atomic
                        g_depths[i] += add_z;
#pragma omp atomic
                        g_opac[i] += add_op;
                        if (g_s != 0.0) {
#pragma omp atomic
                            g_conics[4 * i] += dx * dx * g_s;
#pragma omp atomic
                            g_conics[4 * i + 1] += dx * dy * g_s;
#pragma omp atomic
                            g_conics[4 * i + 2] += dx * dy * g_s;
#pragma omp atomic
                            g_conics[4 * i + 3] += dy * dy * g_s;
#pragma omp atomic
                            g_means2d[2 * i] -= g_s * (2.0 * dx * q[0] + qs * dy);
#pragma omp atomic
                            g_means2d[2 * i + 1] -= g_s * (2.0 * dy * q[3] + qs * dx);
                        }
                    }
                }
        }
        free(kid); free(kev); free(kT);
    }
}

/* Reverse of oracle_project (parameter mode when cov3d == NULL).  Double arithmetic on top of the
 * fp32 forward values; outputs double. */
void oracle_project_bwd(int64_t n, const float* xyz, const float* scaling_log, const float* rotation, const float* cov3d,
                        const float* opacity, int opacity_is_logit, const float* feat0, int64_t feat_stride,
                        const float* cam, const double* g_means2d, const double* g_conics, const double* g_depths,
                        const double* g_colors, const double* g_opac,
                        double* g_xyz, double* g_scaling, double* g_rotation, double* g_cov3d, double* g_opacity,
                        double* g_feat0 /* [n,3] */) {
    const float* Rv = cam;
    const double fx = cam[12], fy = cam[13];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        Splat g;
        if (cov3d) memcpy(g.S, cov3d + i * 9, 9 * sizeof(float));
        else cov_from_params(scaling_log + i * 3, rotation + i * 4, &g);
        project_one(cam, xyz + i * 3, &g, 0.f, 1e30f);
        double o = opacity_is_logit ? (double)sigmoidf_(opacity[i]) : 0.0;
        g_opacity[i] = opacity_is_logit ? g_opac[i] * o * (1.0 - o) : g_opac[i];
        for (int c = 0; c < 3; ++c) {
            double col = sigmoidf_(feat0[i * feat_stride + c]);
            g_feat0[i * 3 + c] = g_colors[i * 3 + c] * col * (1.0 - col);
        }
        const double Q[4] = {g.q00, g.q01, g.q10, g.q11};
        const double* gQ = g_conics + i * 4;
        /* G2 = -Q^T gQ Q^T */
        double T_[4] = {Q[0] * gQ[0] + Q[2] * gQ[2], Q[0] * gQ[1] + Q[2] * gQ[3], Q[1] * gQ[0] + Q[3] * gQ[2], Q[1] * gQ[1] + Q[3] * gQ[3]};
        double G2[4] = {-(T_[0] * Q[0] + T_[1] * Q[1]), -(T_[0] * Q[2] + T_[1] * Q[3]), -(T_[2] * Q[0] + T_[3] * Q[1]), -(T_[2] * Q[2] + T_[3] * Q[3])};
        const double J[6] = {g.J00, 0, g.J02, 0, g.J11, g.J12};
        double GJ[6], gM[9], JMt[6], JM[6], gJ[6];
        for (int c = 0; c < 3; ++c) { GJ[c] = G2[0] * J[c] + G2[1] * J[3 + c]; GJ[3 + c] = G2[2] * J[c] + G2[3] * J[3 + c]; }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) gM[r * 3 + c] = J[r] * GJ[c] + J[3 + r] * GJ[3 + c];
        for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) {
            JMt[r * 3 + c] = J[r * 3] * g.M[c * 3] + J[r * 3 + 1] * g.M[c * 3 + 1] + J[r * 3 + 2] * g.M[c * 3 + 2];
            JM[r * 3 + c] = J[r * 3] * g.M[c] + J[r * 3 + 1] * g.M[3 + c] + J[r * 3 + 2] * g.M[6 + c];
        }
        for (int c = 0; c < 3; ++c) {
            gJ[c] = G2[0] * JMt[c] + G2[1] * JMt[3 + c] + G2[0] * JM[c] + G2[2] * JM[3 + c];
            gJ[3 + c] = G2[2] * JMt[c] + G2[3] * JMt[3 + c] + G2[1] * JM[c] + G2[3] * JM[3 + c];
        }
        const double iz = g.iz, iz2 = iz * iz, iz3 = iz2 * iz, X = g.X, Y = g.Y;
        const double gmx = g_means2d[i * 2], gmy = g_means2d[i * 2 + 1];
        const double gX = gmx * fx * iz + gJ[2] * (-fx * iz2);
        const double gY = gmy * (-fy * iz) + gJ[5] * (fy * iz2);
        const double gZ = g_depths[i] + gmx * (-fx * X * iz2) + gmy * (fy * Y * iz2) + gJ[0] * (-fx * iz2) + gJ[2] * (2 * fx * X * iz3)
                          + gJ[4] * (fy * iz2) + gJ[5] * (-2 * fy * Y * iz3);
        for (int c = 0; c < 3; ++c) g_xyz[i * 3 + c] = Rv[c] * gX + Rv[3 + c] * gY + Rv[6 + c] * gZ;
        double tmp[9], gS[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) tmp[r * 3 + c] = Rv[r] * gM[c] + Rv[3 + r] * gM[3 + c] + Rv[6 + r] * gM[6 + c];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) gS[r * 3 + c] = tmp[r * 3] * Rv[c] + tmp[r * 3 + 1] * Rv[3 + c] + tmp[r * 3 + 2] * Rv[6 + c];
        if (cov3d) { for (int k = 0; k < 9; ++k) g_cov3d[i * 9 + k] = gS[k]; continue; }
        double gR[9];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
            double acc = 0;
            for (int k = 0; k < 3; ++k) acc += (gS[r * 3 + k] + gS[k * 3 + r]) * g.R[k * 3 + c];
            gR[r * 3 + c] = acc * (double)g.sig[c] * (double)g.sig[c];
        }
        for (int k = 0; k < 3; ++k) {
            double acc = 0;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) acc += g.R[r * 3 + k] * gS[r * 3 + c] * g.R[c * 3 + k];
            g_scaling[i * 3 + k] = 2.0 * acc * (double)g.sig[k] * (double)g.sig[k];
        }
        const double w = g.q[0], x = g.q[1], y = g.q[2], z = g.q[3];
        double gq[4];
        gq[0] = 2 * (-z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
        gq[1] = 2 * (y * gR[1] + z * gR[2] + y * gR[3] - 2 * x * gR[4] - w * gR[5] + z * gR[6] + w * gR[7] - 2 * x * gR[8]);
        gq[2] = 2 * (-2 * y * gR[0] + x * gR[1] + w * gR[2] + x * gR[3] + z * gR[5] - w * gR[6] + z * gR[7] - 2 * y * gR[8]);
        gq[3] = 2 * (-2 * z * gR[0] - w * gR[1] + x * gR[2] + w * gR[3] - 2 * z * gR[4] + y * gR[5] + x * gR[6] + y * gR[7]);
        const double dot = w * gq[0] + x * gq[1] + y * gq[2] + z * gq[3];
        const double qh[4] = {w, x, y, z};
        for (int k = 0; k < 4; ++k) g_rotation[i * 4 + k] = (gq[k] - qh[k] * dot) / (double)g.qn;
    }
}
