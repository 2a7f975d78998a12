/*
 * nmr_oracle.c -- CPU restatement of the 3D-SDN / neural_renderer rasteriser kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the sm_100a
 * kernels in 3d-sdn_b200/csrc/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path never calls into oracle/.
 *
 * What it restates (all citations are /root/reference/geometric/neural_renderer/):
 *   orc_face_inv            rasterize.py:240-277   (K1a, per-face barycentric inverse)
 *   orc_visibility          rasterize.py:279-360   (K1b, the deterministic "safe" z-buffer)
 *   orc_texture_sampling    rasterize.py:362-435   (K2)
 *   orc_backward_pixel      rasterize.py:514-745   (K6, NMR edge-crossing gradient)
 *   orc_backward_textures   rasterize.py:747-789   (K7)
 *   orc_backward_depth      rasterize.py:791-844   (K8)
 *
 * Generalisation: the reference renders squares (image_size); here the raster is
 * H rows (y) by W columns (x).  With H == W every expression reduces to the
 * reference's.  `is` in the reference becomes W for x-quantities and H for
 * y-quantities.
 *
 * Arithmetic contract (SURVEY.md section 7 "hard parts"): IEEE float32, NO fused
 * multiply-add (compile with -ffp-contract=off), double wherever an un-suffixed
 * literal in the reference's CUDA-C forces a double sub-expression.  Where a
 * double sub-expression is provably equal to the float one (products by 0.5 /
 * 2.0, quotients of two floats: double rounding is innocuous for p=24 -> 53) the
 * float form is used and the comment says so.
 *
 * Pinning: oracle/build_ref.py compiles the reference's own CUDA-C kernel bodies
 * (extracted at build time from /root/reference, never committed) for the host
 * into oracle/_ref/; tests/test_oracle_vs_ref.py checks this file against them
 * bit-for-bit on square rasters, and tests/golden/ holds vectors they produced.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <limits.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* CUDA float/double -> int conversion: round toward zero, saturate, NaN -> 0. */
static inline int f2i(double v) {
    if (v != v) return 0;
    if (v >= 2147483647.0) return INT_MAX;
    if (v <= -2147483648.0) return INT_MIN;
    return (int)v;
}
/* CUDA fmax/fmin semantics (NaN loses). */
static inline double dmax(double a, double b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
static inline double dmin(double a, double b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
static inline float fmaxc(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
static inline float fminc(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* rasterize.py:258 : p = 0.5 * (v * is + is - 1).  float*int, +int, -1 are float
 * ops; the 0.5* is a double product that is exact, so the value is the float one. */
static inline float to_pixel(float v, int is) {
    float t = v * (float)is;
    t = t + (float)is;
    t = t - 1.0f;
    return 0.5f * t;
}

/* rasterize.py:252,307,537 back-face test. */
static inline int is_backside(const float *f) {
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* K1a  rasterize.py:240-277.  faces, inv: [BF, 9].  inv must be zero-initialised
 * by the caller (xp.zeros_like, :240); back faces are left untouched (:252-253). */
void orc_face_inv(const float *faces, int BF, int H, int W, float *inv) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < BF; i++) {
        const float *face = faces + (size_t)i * 9;
        float *out = inv + (size_t)i * 9;
        if (is_backside(face)) continue;
        float p[3][2];
        for (int n = 0; n < 3; n++) {
            p[n][0] = to_pixel(face[3 * n + 0], W);
            p[n][1] = to_pixel(face[3 * n + 1], H);
        }
        float fi[9] = {
            p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
            p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
            p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
        float den = (p[2][0] * (p[0][1] - p[1][1]) +
                     p[0][0] * (p[1][1] - p[2][1])) +
                    p[1][0] * (p[2][1] - p[0][1]);
        for (int k = 0; k < 9; k++) out[k] = fi[k] / den;
    }
}

/* K1b  rasterize.py:279-360.  One iteration per sample, all faces in index order,
 * strict '<' on depth => lowest face index among equal minima.
 * Buffers are pre-initialised by the caller exactly as rasterize.py:475-493
 * (face_index = -1, depth = far, weight = 0, face_inv_map = 0).
 * near / far are the Python values string-substituted at :353-354, i.e. C double
 * literals compared against the float zp in double. */
void orc_visibility(const float *faces, const float *faces_inv, int B, int F, int H, int W,
                    double near, double far, int return_depth,
                    int32_t *face_index_map, float *weight_map, float *depth_map, float *face_inv_map) {
    const long npix = (long)B * H * W;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < npix; i++) {
        const int bn = (int)(i / ((long)H * W));
        const int pn = (int)(i % ((long)H * W));
        const int yi = pn / W;
        const int xi = pn % W;
        /* :292-293 (2.*yi + 1 - is)/is in double then to float == float quotient of
         * the two exactly-representable integers. */
        const float yp = (float)((2. * yi + 1 - H) / H);
        const float xp = (float)((2. * xi + 1 - W) / W);
        const float xif = (float)xi, yif = (float)yi;

        const float *face = faces + (size_t)bn * F * 9 - 9;
        const float *finv = faces_inv + (size_t)bn * F * 9 - 9;
        float depth_min = (float)far;
        int face_index_min = -1;
        float weight_min[3] = {0, 0, 0};
        float face_inv_min[9] = {0};
        for (int fn = 0; fn < F; fn++) {
            face += 9;
            finv += 9;
            if (is_backside(face)) continue;
            if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
                ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
                ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7]))) continue;
            float w[3];
            for (int k = 0; k < 3; k++) {
                float t = finv[3 * k + 0] * xif;
                float u = finv[3 * k + 1] * yif;
                t = t + u;
                w[k] = t + finv[3 * k + 2];
            }
            float w_sum = 0;
            for (int k = 0; k < 3; k++) {
                w[k] = (float)dmin(dmax((double)w[k], 0.), 1.);
                w_sum += w[k];
            }
            for (int k = 0; k < 3; k++) w[k] /= w_sum;
            float s = w[0] / face[2];
            s = s + w[1] / face[5];
            s = s + w[2] / face[8];
            const float zp = (float)(1. / (double)s);
            if ((double)zp <= near || far <= (double)zp) continue;
            if (zp < depth_min) {
                depth_min = zp;
                face_index_min = fn;
                for (int k = 0; k < 3; k++) weight_min[k] = w[k];
                if (return_depth) for (int k = 0; k < 9; k++) face_inv_min[k] = finv[k];
            }
        }
        if (0 <= face_index_min) {
            depth_map[i] = depth_min;
            face_index_map[i] = face_index_min;
            for (int k = 0; k < 3; k++) weight_map[3 * i + k] = weight_min[k];
            if (return_depth) for (int k = 0; k < 9; k++) face_inv_map[9 * i + k] = face_inv_min[k];
        }
    }
}

/* K2  rasterize.py:362-435.  batch_bug != 0 reproduces :390 (face pointer without
 * the bn*nf offset); 0 uses the per-image face (identical at B == 1). */
void orc_texture_sampling(const float *faces, const float *textures, const int32_t *face_index_map,
                          const float *weight_map, const float *depth_map,
                          int B, int F, int H, int W, int ts, double eps, int batch_bug,
                          float *rgb_map, int32_t *sampling_index_map, float *sampling_weight_map) {
    const long npix = (long)B * H * W;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < npix; i++) {
        const int face_index = face_index_map[i];
        if (face_index < 0) continue;
        const int bn = (int)(i / ((long)H * W));
        const float *face = batch_bug ? faces + (size_t)face_index * 9
                                      : faces + ((size_t)bn * F + face_index) * 9;
        const float *texture = textures + ((size_t)bn * F + face_index) * ts * ts * ts * 3;
        float *pixel = rgb_map + i * 3;
        const float *weight = weight_map + i * 3;
        const float depth = depth_map[i];
        int32_t *sidx = sampling_index_map + i * 8;
        float *sw = sampling_weight_map + i * 8;

        float tif[3];
        for (int k = 0; k < 3; k++) {
            /* :401  weight * (ts - 1 - eps) * (depth / z): float quotient, double products */
            const float q = depth / face[3 * k + 2];
            tif[k] = (float)(((double)weight[k] * ((double)(ts - 1) - eps)) * (double)q);
        }
        float new_pixel[3] = {0, 0, 0};
        for (int pn = 0; pn < 8; pn++) {
            float w = 1;
            int tii[3];
            for (int k = 0; k < 3; k++) {
                const int ti = f2i(tif[k]);
                const float frac = tif[k] - (float)ti;
                if ((pn >> k) % 2 == 0) {
                    w *= 1.0f - frac;
                    tii[k] = ti;
                } else {
                    w *= frac;
                    tii[k] = ti + 1;
                }
            }
            const int isc = tii[0] * ts * ts + tii[1] * ts + tii[2];
            for (int k = 0; k < 3; k++) new_pixel[k] += w * texture[isc * 3 + k];
            sidx[pn] = isc;
            sw[pn] = w;
        }
        for (int k = 0; k < 3; k++) pixel[k] = new_pixel[k];
    }
}

/* helper for K6: dist = a/b * t * 2. / is, then +-eps in double (:646-647). */
static inline float k6_dist(float a, float b, float t, int is, double eps) {
    float d = a / b;
    d = d * t;
    /* "* 2. / is" is a double expression of a float: x*2 exact, quotient of two
     * floats rounded once more to float == float division. */
    d = (float)(((double)d * 2.) / (double)is);
    d = (0 < d) ? (float)((double)d + eps) : (float)((double)d - eps);
    return d;
}

/* K6  rasterize.py:514-745.  grad_faces [B,F,9] must be zero-initialised (:848);
 * this kernel OVERWRITES every face's 9 entries (:733) except back faces (:537). */
void orc_backward_pixel(const float *faces, const int32_t *face_index_map, const float *rgb_map,
                        const float *alpha_map, const float *grad_rgb_map, const float *grad_alpha_map,
                        int B, int F, int H, int W, double eps, int return_rgb, int return_alpha,
                        float *grad_faces) {
    if (!return_rgb && !return_alpha) return;
    const int BF = B * F;
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < BF; i++) {
        const int bn = i / F;
        const int fn = i % F;
        const float *face = faces + (size_t)i * 9;
        float grad_face[9] = {0};
        if (is_backside(face)) continue;
        const long base = (long)bn * H * W;

        for (int edge_num = 0; edge_num < 3; edge_num++) {
            int pi[3];
            float pp[3][2];
            for (int n = 0; n < 3; n++) pi[n] = (edge_num + n) % 3;
            for (int n = 0; n < 3; n++) {
                pp[n][0] = to_pixel(face[3 * pi[n] + 0], W);
                pp[n][1] = to_pixel(face[3 * pi[n] + 1], H);
            }
            for (int axis = 0; axis < 2; axis++) {
                /* is0: raster size along d0, is1: along d1 */
                const int is0 = (axis == 0) ? W : H;
                const int is1 = (axis == 0) ? H : W;
                float p[3][2];
                for (int n = 0; n < 3; n++) for (int d = 0; d < 2; d++) p[n][d] = pp[n][(d + axis) % 2];

                int direction;
                if (axis == 0) direction = (p[0][0] < p[1][0]) ? -1 : 1;
                else direction = (p[0][0] < p[1][0]) ? 1 : -1;

                const int d0_from = f2i(dmax((double)ceilf(fminc(p[0][0], p[1][0])), 0.));
                const int d0_to = f2i(dmin((double)fmaxc(p[0][0], p[1][0]), is0 - 1.));
                for (int d0 = d0_from; d0 <= d0_to; d0++) {
                    const float d0f = (float)d0;
                    float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]);
                    d1_cross = d1_cross * (d0f - p[0][0]);
                    d1_cross = d1_cross + p[0][1];
                    int d1_in;
                    if (0 < direction) d1_in = f2i(floorf(d1_cross)); else d1_in = f2i(ceilf(d1_cross));
                    if (d1_in < 0 || is1 <= d1_in) continue;
                    const int d1_out = d1_in + direction;
                    if (d1_out < 0 || is1 <= d1_out) continue;

                    long map_index_in, map_index_out;
                    if (axis == 0) {
                        map_index_in = base + (long)d1_in * W + d0;
                        map_index_out = base + (long)d1_out * W + d0;
                    } else {
                        map_index_in = base + (long)d0 * W + d1_in;
                        map_index_out = base + (long)d0 * W + d1_out;
                    }
                    float alpha_in = 0, alpha_out = 0;
                    const float *rgb_in = 0, *rgb_out = 0;
                    if (return_alpha) {
                        alpha_in = alpha_map[map_index_in];
                        alpha_out = alpha_map[map_index_out];
                    }
                    if (return_rgb) {
                        rgb_in = rgb_map + map_index_in * 3;
                        rgb_out = rgb_map + map_index_out * 3;
                    }
                    const long map_offset = (axis == 0) ? W : 1;

                    /* out sweep :601-656 */
                    if (face_index_map[map_index_in] == fn) {
                        const int d1_limit = (0 < direction) ? is1 - 1 : 0;
                        const int d1_from = imax(imin(d1_out, d1_limit), 0);
                        const int d1_to = imin(imax(d1_out, d1_limit), is1 - 1);
                        long mi = (axis == 0) ? base + (long)d1_from * W + d0 : base + (long)d0 * W + d1_from;
                        for (int d1 = d1_from; d1 <= d1_to; d1++, mi += map_offset) {
                            float diff_grad = 0;
                            if (return_alpha) diff_grad += (alpha_map[mi] - alpha_in) * grad_alpha_map[mi];
                            if (return_rgb)
                                for (int k = 0; k < 3; k++)
                                    diff_grad += (rgb_map[mi * 3 + k] - rgb_in[k]) * grad_rgb_map[mi * 3 + k];
                            if (diff_grad <= 0) continue;
                            const float t = (float)d1 - d1_cross;
                            if (p[1][0] != d0f) {
                                const float dist = k6_dist(p[1][0] - p[0][0], p[1][0] - d0f, t, is1, eps);
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != d0f) {
                                const float dist = k6_dist(p[1][0] - p[0][0], d0f - p[0][0], t, is1, eps);
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }

                    /* in sweep :659-727 */
                    {
                        float d0_cross2;
                        if ((d0f - p[0][0]) * (d0f - p[2][0]) < 0) {
                            d0_cross2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]);
                            d0_cross2 = d0_cross2 * (d0f - p[0][0]);
                            d0_cross2 = d0_cross2 + p[0][1];
                        } else {
                            d0_cross2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]);
                            d0_cross2 = d0_cross2 * (d0f - p[2][0]);
                            d0_cross2 = d0_cross2 + p[2][1];
                        }
                        int d1_limit;
                        if (0 < direction) d1_limit = f2i(ceilf(d0_cross2)); else d1_limit = f2i(floorf(d0_cross2));
                        const int d1_from = imax(imin(d1_in, d1_limit), 0);
                        const int d1_to = imin(imax(d1_in, d1_limit), is1 - 1);
                        long mi = (axis == 0) ? base + (long)d1_from * W + d0 : base + (long)d0 * W + d1_from;
                        for (int d1 = d1_from; d1 <= d1_to; d1++, mi += map_offset) {
                            if (face_index_map[mi] != fn) continue;
                            float diff_grad = 0;
                            if (return_alpha) diff_grad += (alpha_map[mi] - alpha_out) * grad_alpha_map[mi];
                            if (return_rgb)
                                for (int k = 0; k < 3; k++)
                                    diff_grad += (rgb_map[mi * 3 + k] - rgb_out[k]) * grad_rgb_map[mi * 3 + k];
                            if (diff_grad <= 0) continue;
                            const float t = (float)d1 - d1_cross;
                            if (p[1][0] != d0f) {
                                const float dist = k6_dist(p[1][0] - p[0][0], p[1][0] - d0f, t, is1, eps);
                                grad_face[pi[0] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                            if (p[0][0] != d0f) {
                                const float dist = k6_dist(p[1][0] - p[0][0], d0f - p[0][0], t, is1, eps);
                                grad_face[pi[1] * 3 + (1 - axis)] -= diff_grad / dist;
                            }
                        }
                    }
                }
            }
        }
        for (int k = 0; k < 9; k++) grad_faces[(size_t)i * 9 + k] = grad_face[k];
    }
}

/* K7  rasterize.py:747-789 (serial: the reference uses float atomics). */
void orc_backward_textures(const int32_t *face_index_map, const float *sampling_weight_map,
                           const int32_t *sampling_index_map, const float *grad_rgb_map,
                           int B, int F, int H, int W, int ts, float *grad_textures) {
    /* images are independent (each owns its slice of grad_textures): parallel over bn, serial inside */
#pragma omp parallel for schedule(dynamic, 1)
    for (int bn = 0; bn < B; bn++)
    for (long i = (long)bn * H * W; i < (long)(bn + 1) * H * W; i++) {
        const int face_index = face_index_map[i];
        if (face_index < 0) continue;
        float *gt = grad_textures + ((size_t)bn * F + face_index) * ts * ts * ts * 3;
        for (int pn = 0; pn < 8; pn++) {
            const float w = sampling_weight_map[i * 8 + pn];
            const int isc = sampling_index_map[i * 8 + pn];
            for (int k = 0; k < 3; k++) gt[isc * 3 + k] += w * grad_rgb_map[i * 3 + k];
        }
    }
}

/* K8  rasterize.py:791-844 (serial: float atomics in the reference).  The factor
 * `is / 2` at :833 converts a pixel-space derivative to NDC: W/2 for x, H/2 for y. */
void orc_backward_depth(const float *faces, const float *depth_map, const int32_t *face_index_map,
                        const float *face_inv_map, const float *weight_map, const float *grad_depth_map,
                        int B, int F, int H, int W, float *grad_faces) {
    /* images are independent (each owns its slice of grad_faces): parallel over bn, serial inside */
#pragma omp parallel for schedule(dynamic, 1)
    for (int bn = 0; bn < B; bn++)
    for (long i = (long)bn * H * W; i < (long)(bn + 1) * H * W; i++) {
        const int fn = face_index_map[i];
        if (fn < 0) continue;
        const float *face = faces + ((size_t)bn * F + fn) * 9;
        const float depth = depth_map[i];
        const float depth2 = depth * depth;
        const float *finv = face_inv_map + i * 9;
        const float *weight = weight_map + i * 3;
        const float gd = grad_depth_map[i];
        float *gf = grad_faces + ((size_t)bn * F + fn) * 9;
        for (int k = 0; k < 3; k++) {
            const float zk = face[3 * k + 2];
            gf[3 * k + 2] += gd * weight[k] * depth2 / (zk * zk);
        }
        float tmp[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) tmp[k] += -finv[3 * l + k] / face[3 * l + 2];
        for (int k = 0; k < 3; k++) for (int l = 0; l < 2; l++) {
            const int is = (l == 0) ? W : H;
            gf[3 * k + l] += -gd * tmp[l] * weight[k] * depth2 * (float)is / 2.0f;
        }
    }
}

/* K1u  rasterize.py:102-236: the per-face scanline kernel 3D-SDN actually runs (scripts/env.sh:11 sets
 * NEURAL_RENDERER_UNSAFE=1).  On the GPU it is racy (spin lock per pixel); executed face by face in index order it
 * is deterministic and keeps the first face among equal depths (strict '<'), like the safe path -- but it decides
 * "inside" with ceil/floor scanline bounds instead of the three half-plane tests, so a thin set of edge samples
 * differs.  Used as the comparator / CPU timing arm, never as the parity definition.  Parallel over images. */
void orc_visibility_scanline(const float *faces, int B, int F, int H, int W, double near, double far,
                             int return_depth, int32_t *face_index_map, float *weight_map, float *depth_map,
                             float *face_inv_map) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int bn = 0; bn < B; bn++) {
        for (int fn = 0; fn < F; fn++) {
            const float *face = faces + ((size_t)bn * F + fn) * 9;
            if (is_backside(face)) continue;
            int pi[3];
            if (face[0] < face[3]) {
                pi[0] = (face[6] < face[0]) ? 2 : 0;
                pi[2] = (face[3] < face[6]) ? 2 : 1;
            } else {
                pi[0] = (face[6] < face[3]) ? 2 : 1;
                pi[2] = (face[0] < face[6]) ? 2 : 0;
            }
            pi[1] = 0;
            for (int k = 0; k < 3; k++) if (pi[0] != k && pi[2] != k) pi[1] = k;
            float p[3][3];
            for (int n = 0; n < 3; n++) {
                p[n][0] = to_pixel(face[3 * pi[n] + 0], W);
                p[n][1] = to_pixel(face[3 * pi[n] + 1], H);
                p[n][2] = face[3 * pi[n] + 2];
            }
            if (p[0][0] == p[2][0]) continue;
            float fi[9] = {
                p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
            const float den = (p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1])) +
                              p[1][0] * (p[2][1] - p[0][1]);
            for (int k = 0; k < 9; k++) fi[k] /= den;
            const int xi_min = f2i(dmax((double)ceilf(p[0][0]), 0.));
            const int xi_max = f2i(dmin((double)p[2][0], W - 1.));
            for (int xi = xi_min; xi <= xi_max; xi++) {
                const float xif = (float)xi;
                float yi1, yi2;
                if (xif <= p[1][0]) {
                    if (p[1][0] - p[0][0] != 0) {
                        yi1 = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]);
                        yi1 = yi1 * (xif - p[0][0]);
                        yi1 = yi1 + p[0][1];
                    } else yi1 = p[1][1];
                } else {
                    if (p[2][0] - p[1][0] != 0) {
                        yi1 = (p[2][1] - p[1][1]) / (p[2][0] - p[1][0]);
                        yi1 = yi1 * (xif - p[1][0]);
                        yi1 = yi1 + p[1][1];
                    } else yi1 = p[1][1];
                }
                yi2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]);
                yi2 = yi2 * (xif - p[0][0]);
                yi2 = yi2 + p[0][1];
                const int yi_min = f2i(dmax(0., (double)ceilf(fminc(yi1, yi2))));
                const int yi_max = f2i(dmin((double)fmaxc(yi1, yi2), H - 1.));
                for (int yi = yi_min; yi <= yi_max; yi++) {
                    const long index = (long)bn * H * W + (long)yi * W + xi;
                    const float yif = (float)yi;
                    float w[3], w_sum = 0;
                    for (int k = 0; k < 3; k++) {
                        float t = fi[3 * k + 0] * xif;
                        float u = fi[3 * k + 1] * yif;
                        t = t + u;
                        w[k] = t + fi[3 * k + 2];
                    }
                    for (int k = 0; k < 3; k++) {
                        w[k] = (float)dmin(dmax((double)w[k], 0.), 1.);
                        w_sum += w[k];
                    }
                    for (int k = 0; k < 3; k++) w[k] /= w_sum;
                    float s = w[0] / p[0][2];
                    s = s + w[1] / p[1][2];
                    s = s + w[2] / p[2][2];
                    const float zp = (float)(1. / (double)s);
                    if ((double)zp <= near || far <= (double)zp) continue;
                    if (zp < depth_map[index]) {
                        depth_map[index] = zp;
                        face_index_map[index] = fn;
                        for (int k = 0; k < 3; k++) weight_map[3 * index + pi[k]] = w[k];
                        if (return_depth)
                            for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
                                face_inv_map[9 * index + 3 * pi[l] + k] = fi[3 * l + k];
                    }
                }
            }
        }
    }
}
