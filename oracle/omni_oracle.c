/*
 * omni_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * Plain-C, single-threaded restatement of the arithmetic on the reference's stage 01-03 hot path.
 * The reference (/root/reference/image_processor, pure Python) delegates this arithmetic to
 * un-vendored third-party libraries: OpenCV (opencv-python-headless 4.13.0.92 in this image) and
 * NumPy 2.3.5.  Each function below restates the published algorithm of the library call made at
 * the cited reference call site, and is pinned (tests/test_oracle_*.py) against
 *   (a) the library call itself on seeded inputs, and
 *   (b) golden vectors produced by importing and running the reference's own stage functions
 *       (tests/golden/, generator tools/make_golden.py).
 *
 * Build: oracle/build.py  (gcc -O2 -ffp-contract=off -shared -fPIC).  -ffp-contract=off matters:
 * the float paths must round every product and sum separately, as NumPy / OpenCV do.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "omni_tables.inc"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* 01_resize.py:20  cv2.resize(img, (new_w,new_h), interpolation=cv2.INTER_AREA), shrink only   */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int di, si; float alpha; } orc_tab_t;

/* OpenCV computeResizeAreaTab: per destination index the list of (source index, weight). */
static int orc_area_tab(int ssize, int dsize, double scale, orc_tab_t *tab)
{
    int k = 0;
    for (int dx = 0; dx < dsize; dx++) {
        double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        double cw = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) {
            tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cw);
        }
        for (int sx = sx1; sx < sx2; sx++) {
            tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cw);
        }
        if (fsx2 - sx2 > 1e-3) {
            double a = fsx2 - sx2; if (a > 1.0) a = 1.0; if (a > cw) a = cw;
            tab[k].di = dx; tab[k].si = sx2; tab[k++].alpha = (float)(a / cw);
        }
    }
    return k;
}

static inline uint8_t orc_sat_u8_f(float v)
{
    /* cv::saturate_cast<uchar>(float): cvRound (round-half-even) then clamp */
    long r = lrintf(v);
    return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

/* src: Hs x Ws x 3 u8 contiguous, dst: Hd x Wd x 3.  Returns 0, or -1 on bad sizes. */
ORC_API int orc_resize_area_u8c3(const uint8_t *src, int Hs, int Ws, uint8_t *dst, int Hd, int Wd)
{
    if (Hd <= 0 || Wd <= 0 || Hd > Hs || Wd > Ws) return -1;
    const int cn = 3;
    double scale_x = (double)Ws / Wd, scale_y = (double)Hs / Hd;
    int isx = (int)lrint(scale_x), isy = (int)lrint(scale_y);
    int fast = fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16;
    if (fast && isx == 2 && isy == 2) {
        for (int y = 0; y < Hd; y++)
            for (int x = 0; x < Wd; x++)
                for (int c = 0; c < cn; c++) {
                    const uint8_t *p = src + ((size_t)(2 * y) * Ws + 2 * x) * cn + c;
                    dst[((size_t)y * Wd + x) * cn + c] =
                        (uint8_t)((p[0] + p[cn] + p[(size_t)Ws * cn] + p[(size_t)Ws * cn + cn] + 2) >> 2);
                }
        return 0;
    }
    if (fast) {
        float inv = 1.f / (float)(isx * isy);
        for (int y = 0; y < Hd; y++)
            for (int x = 0; x < Wd; x++)
                for (int c = 0; c < cn; c++) {
                    int sum = 0;
                    for (int j = 0; j < isy; j++)
                        for (int i = 0; i < isx; i++)
                            sum += src[((size_t)(y * isy + j) * Ws + (x * isx + i)) * cn + c];
                    dst[((size_t)y * Wd + x) * cn + c] = orc_sat_u8_f((float)sum * inv);
                }
        return 0;
    }
    orc_tab_t *xt = (orc_tab_t *)malloc(sizeof(orc_tab_t) * (size_t)(Ws + 2 * Wd + 2));
    orc_tab_t *yt = (orc_tab_t *)malloc(sizeof(orc_tab_t) * (size_t)(Hs + 2 * Hd + 2));
    int nx = orc_area_tab(Ws, Wd, scale_x, xt), ny = orc_area_tab(Hs, Hd, scale_y, yt);
    float *buf = (float *)malloc(sizeof(float) * (size_t)Wd * cn);
    float *sum = (float *)calloc((size_t)Wd * cn, sizeof(float));
    int prev = -1;
    for (int j = 0; j < ny; j++) {
        int dy = yt[j].di, sy = yt[j].si; float beta = yt[j].alpha;
        const uint8_t *S = src + (size_t)sy * Ws * cn;
        memset(buf, 0, sizeof(float) * (size_t)Wd * cn);
        for (int k = 0; k < nx; k++) {
            float a = xt[k].alpha; const uint8_t *p = S + (size_t)xt[k].si * cn; float *b = buf + (size_t)xt[k].di * cn;
            for (int c = 0; c < cn; c++) b[c] = b[c] + (float)p[c] * a;
        }
        if (dy != prev) {
            if (prev >= 0)
                for (int i = 0; i < Wd * cn; i++) dst[(size_t)prev * Wd * cn + i] = orc_sat_u8_f(sum[i]);
            for (int i = 0; i < Wd * cn; i++) sum[i] = beta * buf[i];
            prev = dy;
        } else {
            for (int i = 0; i < Wd * cn; i++) sum[i] = sum[i] + beta * buf[i];
        }
    }
    if (prev >= 0)
        for (int i = 0; i < Wd * cn; i++) dst[(size_t)prev * Wd * cn + i] = orc_sat_u8_f(sum[i]);
    free(xt); free(yt); free(buf); free(sum);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 02_color_extract.py:35  cv2.cvtColor(img_bgr, cv2.COLOR_BGR2LAB) on 8-bit input             */
/* ------------------------------------------------------------------------------------------ */

static inline int orc_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
static inline uint8_t orc_clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

ORC_API void orc_bgr2lab_u8(const uint8_t *bgr, size_t npix, uint8_t *lab)
{
    for (size_t i = 0; i < npix; i++) {
        int B = OMNI_LAB_GAMMA[bgr[3 * i]], G = OMNI_LAB_GAMMA[bgr[3 * i + 1]], R = OMNI_LAB_GAMMA[bgr[3 * i + 2]];
        int fX = OMNI_LAB_CBRT[orc_descale(R * 1777 + G * 1541 + B * 778, 12)];
        int fY = OMNI_LAB_CBRT[orc_descale(R * 871 + G * 2929 + B * 296, 12)];
        int fZ = OMNI_LAB_CBRT[orc_descale(R * 73 + G * 448 + B * 3575, 12)];
        lab[3 * i]     = orc_clip_u8(orc_descale(296 * fY - 1336934, 15));
        lab[3 * i + 1] = orc_clip_u8(orc_descale(500 * (fX - fY) + 128 * 32768, 15));
        lab[3 * i + 2] = orc_clip_u8(orc_descale(200 * (fY - fZ) + 128 * 32768, 15));
    }
}

/* ------------------------------------------------------------------------------------------ */
/* 02_color_extract.py:53-55  diffs = data[:,None,:]-centers ; d2 = sum(diffs*diffs,2) ; argmin */
/* float32, every op rounded, order (d0^2 + d1^2) + d2^2, first minimum wins                   */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_assign_f32(const uint8_t *px3, size_t npix, const float *centers, int K, uint8_t *labels)
{
    for (size_t i = 0; i < npix; i++) {
        float p0 = (float)px3[3 * i], p1 = (float)px3[3 * i + 1], p2 = (float)px3[3 * i + 2];
        int best = 0; float bd = 0.f;
        for (int k = 0; k < K; k++) {
            volatile float d0 = p0 - centers[3 * k], d1 = p1 - centers[3 * k + 1], d2 = p2 - centers[3 * k + 2];
            volatile float q0 = d0 * d0, q1 = d1 * d1, q2 = d2 * d2;
            volatile float s = q0 + q1;
            float d = s + q2;
            if (k == 0 || d < bd) { bd = d; best = k; }
        }
        labels[i] = (uint8_t)best;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* process_colors.py:69-77 assign_labels: int16 diff, int16 WRAPPING square, wide sum, argmin  */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_assign_i16wrap(const uint8_t *px3, size_t npix, const uint8_t *pal, int K, uint8_t *labels)
{
    for (size_t i = 0; i < npix; i++) {
        int best = 0; long bd = 0;
        for (int k = 0; k < K; k++) {
            long d = 0;
            for (int c = 0; c < 3; c++) {
                int df = (int)px3[3 * i + c] - (int)pal[3 * k + c];
                d += (int16_t)(uint16_t)(df * df);
            }
            if (k == 0 || d < bd) { bd = d; best = k; }
        }
        labels[i] = (uint8_t)best;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* 02_color_extract.py:150  mask = (labels == k_idx).astype(np.uint8) * 255  (after lut remap)  */
/* ------------------------------------------------------------------------------------------ */
ORC_API void orc_onehot(const uint8_t *labels, size_t npix, const uint8_t *lut, int plane, uint8_t *mask)
{
    for (size_t i = 0; i < npix; i++) mask[i] = (lut ? lut[labels[i]] : labels[i]) == plane ? 255 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 02_color_extract.py:138,151-154 and 03_edge_detect.py:23-30                                  */
/* cv2.getStructuringElement + cv2.morphologyEx(OPEN|CLOSE, iterations=n)                       */
/* ------------------------------------------------------------------------------------------ */

/* shape: 0 = MORPH_RECT, 2 = MORPH_ELLIPSE (OpenCV enum values). se: k*k bytes, row-major. */
ORC_API void orc_structuring_element(int shape, int k, uint8_t *se)
{
    int r = k / 2, c = k / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; i++) {
        int j1 = 0, j2 = 0;
        if (shape == 0) { j2 = k; }
        else {
            int dy = i - r;
            if (abs(dy) <= r) {
                int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
                j1 = c - dx > 0 ? c - dx : 0;
                j2 = c + dx + 1 < k ? c + dx + 1 : k;
            }
        }
        for (int j = 0; j < k; j++) se[i * k + j] = (j >= j1 && j < j2) ? 1 : 0;
    }
}

/* one erode (is_dilate=0) or dilate (1); pixels outside the image are ignored; anchor = k/2;
 * the element is used un-reflected for both (OpenCV behaviour). */
static void orc_morph_once(const uint8_t *src, int h, int w, uint8_t *dst, const uint8_t *se, int k, int is_dilate)
{
    int a = k / 2;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v = is_dilate ? 0 : 255;
            for (int i = 0; i < k; i++) {
                int yy = y + i - a; if (yy < 0 || yy >= h) continue;
                for (int j = 0; j < k; j++) {
                    if (!se[i * k + j]) continue;
                    int xx = x + j - a; if (xx < 0 || xx >= w) continue;
                    int p = src[(size_t)yy * w + xx];
                    v = is_dilate ? (p > v ? p : v) : (p < v ? p : v);
                }
            }
            dst[(size_t)y * w + x] = (uint8_t)v;
        }
}

/* op: 0 = OPEN (erode^n then dilate^n), 1 = CLOSE (dilate^n then erode^n); in-place on img. */
ORC_API void orc_morph_openclose(uint8_t *img, int h, int w, const uint8_t *se, int k, int op, int iters)
{
    if (iters <= 0) return;
    uint8_t *tmp = (uint8_t *)malloc((size_t)h * w);
    for (int phase = 0; phase < 2; phase++) {
        int is_dilate = (op == 0) ? phase : 1 - phase;
        for (int it = 0; it < iters; it++) {
            orc_morph_once(img, h, w, tmp, se, k, is_dilate);
            memcpy(img, tmp, (size_t)h * w);
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------ */
/* 03_edge_detect.py:33  cv2.GaussianBlur(mask, (k,k), 0): separable 8.8 fixed point,           */
/* BORDER_REFLECT_101, single rounding at the end                                               */
/* ------------------------------------------------------------------------------------------ */
static inline int orc_reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

ORC_API int orc_gauss_weights(int k, uint16_t *w)
{
    if (k < OMNI_GAUSS_KMIN || k > OMNI_GAUSS_KMAX || !(k & 1)) return -1;
    memcpy(w, OMNI_GAUSS_W + OMNI_GAUSS_OFFS[(k - 3) / 2], sizeof(uint16_t) * (size_t)k);
    return 0;
}

ORC_API void orc_gaussian_blur_u8(const uint8_t *src, int h, int w, uint8_t *dst, const uint16_t *wt, int k)
{
    int r = k / 2;
    uint32_t *hs = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)h * w);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int i = 0; i < k; i++) s += wt[i] * (uint32_t)src[(size_t)y * w + orc_reflect101(x + i - r, w)];
            hs[(size_t)y * w + x] = s;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t s = 0;
            for (int i = 0; i < k; i++) s += wt[i] * hs[(size_t)orc_reflect101(y + i - r, h) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)((s + 32768u) >> 16);
        }
    free(hs);
}

/* ------------------------------------------------------------------------------------------ */
/* 03_edge_detect.py:34  cv2.Canny(blurred, low, high)  (aperture 3, L1 magnitude)              */
/* Sobel with BORDER_REPLICATE, magnitude 0 outside, fixed-point NMS, 8-connected hysteresis    */
/* ------------------------------------------------------------------------------------------ */
static inline int orc_clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* stage outputs (any may be NULL): mag int32 [h*w]; nms u8 [h*w] 0 none / 1 weak / 2 strong */
ORC_API void orc_canny_u8(const uint8_t *src, int h, int w, double t1, double t2, uint8_t *edges,
                          int32_t *mag_out, uint8_t *nms_out)
{
    double lo_d = t1 < t2 ? t1 : t2, hi_d = t1 < t2 ? t2 : t1;
    int low = (int)floor(lo_d), high = (int)floor(hi_d);
    size_t n = (size_t)h * w;
    int16_t *dx = (int16_t *)malloc(sizeof(int16_t) * n), *dy = (int16_t *)malloc(sizeof(int16_t) * n);
    int32_t *mag = (int32_t *)malloc(sizeof(int32_t) * n);
    uint8_t *map = (uint8_t *)calloc(n, 1);
#define PX(yy, xx) ((int)src[(size_t)orc_clampi(yy, 0, h - 1) * w + orc_clampi(xx, 0, w - 1)])
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int gx = (PX(y - 1, x + 1) + 2 * PX(y, x + 1) + PX(y + 1, x + 1)) - (PX(y - 1, x - 1) + 2 * PX(y, x - 1) + PX(y + 1, x - 1));
            int gy = (PX(y + 1, x - 1) + 2 * PX(y + 1, x) + PX(y + 1, x + 1)) - (PX(y - 1, x - 1) + 2 * PX(y - 1, x) + PX(y - 1, x + 1));
            dx[(size_t)y * w + x] = (int16_t)gx; dy[(size_t)y * w + x] = (int16_t)gy;
            mag[(size_t)y * w + x] = abs(gx) + abs(gy);
        }
#undef PX
#define MAG(yy, xx) (((yy) < 0 || (yy) >= h || (xx) < 0 || (xx) >= w) ? 0 : mag[(size_t)(yy) * w + (xx)])
    int *stack = (int *)malloc(sizeof(int) * (n + 1)); size_t sp = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)y * w + x; int m = mag[i];
            if (m <= low) continue;
            int xs = dx[i], ys = dy[i];
            int ax = abs(xs), ay = abs(ys) << 15;
            int tg22x = ax * 13573;
            int ok;
            if (ay < tg22x) ok = m > MAG(y, x - 1) && m >= MAG(y, x + 1);
            else {
                int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) ok = m > MAG(y - 1, x) && m >= MAG(y + 1, x);
                else { int s = (xs ^ ys) < 0 ? -1 : 1; ok = m > MAG(y - 1, x - s) && m > MAG(y + 1, x + s); }
            }
            if (!ok) continue;
            if (m > high) { map[i] = 2; stack[sp++] = (int)i; } else map[i] = 1;
        }
#undef MAG
    if (nms_out) memcpy(nms_out, map, n);
    if (mag_out) memcpy(mag_out, mag, sizeof(int32_t) * n);
    while (sp) {
        int i = stack[--sp]; int y = i / w, x = i % w;
        for (int j = -1; j <= 1; j++)
            for (int q = -1; q <= 1; q++) {
                int yy = y + j, xx = x + q;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                size_t t = (size_t)yy * w + xx;
                if (map[t] == 1) { map[t] = 2; stack[sp++] = (int)t; }
            }
    }
    for (size_t i = 0; i < n; i++) edges[i] = map[i] == 2 ? 255 : 0;
    free(dx); free(dy); free(mag); free(map); free(stack);
}

/* ------------------------------------------------------------------------------------------ */
/* Chains                                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* 02_color_extract.py:146-154: per plane p: (lut[labels]==p)*255 -> RECT-3 OPEN^o -> CLOSE^c.
 * masks: K planes of h*w. */
ORC_API void orc_layer_masks(const uint8_t *labels, int h, int w, const uint8_t *lut, int K,
                             int open_iters, int close_iters, uint8_t *masks)
{
    uint8_t se[9]; orc_structuring_element(0, 3, se);
    size_t n = (size_t)h * w;
    for (int p = 0; p < K; p++) {
        uint8_t *m = masks + (size_t)p * n;
        orc_onehot(labels, n, lut, p, m);
        orc_morph_openclose(m, h, w, se, 3, 0, open_iters);
        orc_morph_openclose(m, h, w, se, 3, 1, close_iters);
    }
}

/* 03_edge_detect.py:23-34 for one layer: ELLIPSE(k_m) OPEN^o, CLOSE^c, GaussianBlur(ks), Canny. */
ORC_API int orc_edge_chain(const uint8_t *mask, int h, int w, int morph_k, int open_iters, int close_iters,
                           int ks, double t1, double t2, uint8_t *edges)
{
    size_t n = (size_t)h * w;
    uint16_t wt[OMNI_GAUSS_KMAX];
    if (orc_gauss_weights(ks, wt)) return -1;
    if (morph_k < 1 || morph_k > 31) return -2;
    uint8_t *se = (uint8_t *)malloc((size_t)morph_k * morph_k);
    orc_structuring_element(2, morph_k, se);
    uint8_t *m = (uint8_t *)malloc(n), *bl = (uint8_t *)malloc(n);
    memcpy(m, mask, n);
    orc_morph_openclose(m, h, w, se, morph_k, 0, open_iters);
    orc_morph_openclose(m, h, w, se, morph_k, 1, close_iters);
    orc_gaussian_blur_u8(m, h, w, bl, wt, ks);
    orc_canny_u8(bl, h, w, t1, t2, edges, NULL, NULL);
    free(se); free(m); free(bl);
    return 0;
}


/* ------------------------------------------------------------------------------------------ */
/* 04_find_contours.py:35-99  thinning_zhangsuen(bin_0_255, layer)  (row "next" of SURVEY 8f)   */
/* ------------------------------------------------------------------------------------------ */
/* The reference works on the bounding box of the non-zero pixels padded by 2 with zero fill
 * outside -- identical to zero padding around the whole image.  Its neighbour names are rotated
 * by 180 degrees against the textbook (P2 = pixel BELOW): _shift(roi, dy, dx)[y, x] = roi[y - dy, x - dx].
 *   P2 = (y+1, x)   P3 = (y+1, x-1)  P4 = (y, x-1)   P5 = (y-1, x-1)
 *   P6 = (y-1, x)   P7 = (y-1, x+1)  P8 = (y, x+1)   P9 = (y+1, x+1)
 * Each iteration: sub-step 1 deletes (A==1, 2<=B<=6, P2*P4*P6==0, P4*P6*P8==0) simultaneously, sub-step 2
 * (P2*P4*P8==0, P2*P6*P8==0) on the result; stops when an iteration removes nothing or after 120 iterations.
 * in/out: h x w u8, in > 0 is foreground, out in {0,255}.  removed (optional): removed[i] = pixels deleted in
 * iteration i+1 (the number the reference logs).  Returns the number of iterations executed. */
static int orc_zs_substep(uint8_t *roi, int h, int w, int step, uint8_t *del)
{
    int n = 0;
#define ZS(y, x) (((y) >= 0 && (y) < h && (x) >= 0 && (x) < w) ? roi[(size_t)(y) * w + (x)] : 0)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            del[(size_t)y * w + x] = 0;
            if (!roi[(size_t)y * w + x]) continue;
            int P2 = ZS(y + 1, x), P3 = ZS(y + 1, x - 1), P4 = ZS(y, x - 1), P5 = ZS(y - 1, x - 1);
            int P6 = ZS(y - 1, x), P7 = ZS(y - 1, x + 1), P8 = ZS(y, x + 1), P9 = ZS(y + 1, x + 1);
            int B = P2 + P3 + P4 + P5 + P6 + P7 + P8 + P9;
            int A = (!P2 && P3) + (!P3 && P4) + (!P4 && P5) + (!P5 && P6) + (!P6 && P7) + (!P7 && P8) + (!P8 && P9) + (!P9 && P2);
            int c = step == 1 ? ((P2 * P4 * P6) == 0 && (P4 * P6 * P8) == 0) : ((P2 * P4 * P8) == 0 && (P2 * P6 * P8) == 0);
            if (A == 1 && B >= 2 && B <= 6 && c) { del[(size_t)y * w + x] = 1; n++; }
        }
#undef ZS
    for (size_t i = 0; i < (size_t)h * w; i++) if (del[i]) roi[i] = 0;
    return n;
}

ORC_API int orc_thin_zhangsuen(const uint8_t *in, int h, int w, uint8_t *out, int max_iter, int32_t *removed)
{
    size_t n = (size_t)h * w;
    uint8_t *roi = (uint8_t *)malloc(n ? n : 1), *del = (uint8_t *)malloc(n ? n : 1);
    for (size_t i = 0; i < n; i++) roi[i] = in[i] > 0;
    int it = 0, changed = 1;
    while (changed && it < max_iter) {
        it++;
        int n1 = orc_zs_substep(roi, h, w, 1, del);
        int n2 = orc_zs_substep(roi, h, w, 2, del);
        if (removed) removed[it - 1] = n1 + n2;
        changed = (n1 + n2) > 0;
    }
    for (size_t i = 0; i < n; i++) out[i] = roi[i] ? 255 : 0;
    free(roi); free(del);
    return it;
}


/* ------------------------------------------------------------------------------------------ */
/* 02_color_extract.py:82-109  legacy swatch extraction                       */
/* ------------------------------------------------------------------------------------------ */
/* Per name i: the swatch colours[i] is tried as RGB (-> BGR reversed) and as-is; cv2.inRange(img, c - tol, c + tol)
 * with the bounds clipped to [0,255]; the candidate with more non-zeros wins (>= favours the reversed one, 02:100-101);
 * RECT-3 open, then close.  img: h x w x 3 (BGR in memory); colors: K x 3 as written in config.json; masks: K planes;
 * choice (optional): 0 = reversed (bgr1), 1 = as-is (bgr2). */
ORC_API void orc_swatch_masks(const uint8_t *img, int h, int w, const int *colors, int K, int tol, uint8_t *masks, int *choice)
{
    size_t n = (size_t)h * w;
    uint8_t se[9];
    orc_structuring_element(0, 3, se);
    uint8_t *m1 = (uint8_t *)malloc(n ? n : 1), *m2 = (uint8_t *)malloc(n ? n : 1);
    for (int i = 0; i < K; i++) {
        int c1[3] = {colors[3 * i + 2], colors[3 * i + 1], colors[3 * i]}, c2[3] = {colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]};
        int lo1[3], hi1[3], lo2[3], hi2[3];
        for (int d = 0; d < 3; d++) {
            lo1[d] = c1[d] - tol < 0 ? 0 : c1[d] - tol; hi1[d] = c1[d] + tol > 255 ? 255 : c1[d] + tol;
            lo2[d] = c2[d] - tol < 0 ? 0 : c2[d] - tol; hi2[d] = c2[d] + tol > 255 ? 255 : c2[d] + tol;
            /* np.array(..., np.uint8) of a value outside [0,255] cannot occur: max(0, .) / min(255, .) were applied first,
             * but a swatch component > 255 + tol or < -tol would wrap there; such configs are rejected by the GPU entry point */
        }
        size_t nz1 = 0, nz2 = 0;
        for (size_t p = 0; p < n; p++) {
            const uint8_t *q = img + 3 * p;
            int a = q[0] >= lo1[0] && q[0] <= hi1[0] && q[1] >= lo1[1] && q[1] <= hi1[1] && q[2] >= lo1[2] && q[2] <= hi1[2];
            int b = q[0] >= lo2[0] && q[0] <= hi2[0] && q[1] >= lo2[1] && q[1] <= hi2[1] && q[2] >= lo2[2] && q[2] <= hi2[2];
            m1[p] = a ? 255 : 0; m2[p] = b ? 255 : 0;
            nz1 += a; nz2 += b;
        }
        uint8_t *out = masks + (size_t)i * n;
        memcpy(out, nz1 >= nz2 ? m1 : m2, n);
        if (choice) choice[i] = nz1 >= nz2 ? 0 : 1;
        orc_morph_openclose(out, h, w, se, 3, 0, 1);
        orc_morph_openclose(out, h, w, se, 3, 1, 1);
    }
    free(m1); free(m2);
}
