/*
 * lumina_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the arithmetic that the reference's page-image hot
 * path executes.  The reference (backend/utils/image_preprocessing.py) is a
 * thin Python layer over Pillow 12.2.0 (libImaging) and OpenCV 4.13.0; those
 * wheels are un-vendored and un-pinned (requirements.txt:20-23), so each
 * function below restates the published algorithm of the library routine the
 * reference call site reaches, and is PINNED by tests/golden/ vectors that
 * were produced by importing and running the real reference module in the
 * build container (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (ocr-system_b200/) never links or calls it.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off; no SIMD intrinsics,
 * no FMA contraction: float32 evaluation order is part of the spec).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* ------------------------------------------------------------------------- *
 * A1. PIL Image.resize(size, LANCZOS)   image_preprocessing.py:110,551
 *     libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
 *     ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc
 * ------------------------------------------------------------------------- */
#define ORC_PREC_BITS 22

static double orc_sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
static double orc_lanczos3(double x) {
    if (-3.0 <= x && x < 3.0) return orc_sinc(x) * orc_sinc(x / 3);
    return 0.0;
}

/* ksize for one axis */
ORC_API int orc_lanczos_ksize(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    double support = 3.0 * fs;
    return (int)ceil(support) * 2 + 1;
}

/* bounds: out_size x {xmin, n}; coeffs: out_size x ksize int32 (22-bit fixed) */
ORC_API void orc_lanczos_coeffs(int in_size, int out_size, int32_t *bounds, int32_t *coeffs) {
    double scale = (double)in_size / out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    double support = 3.0 * fs;
    int ksize = (int)ceil(support) * 2 + 1;
    double *k = (double *)malloc(sizeof(double) * ksize);
    double ss = 1.0 / fs;
    for (int xx = 0; xx < out_size; xx++) {
        double center = (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; x++) {
            double w = orc_lanczos3((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; x++)
            if (ww != 0.0) k[x] /= ww;
        for (int x = xmax; x < ksize; x++) k[x] = 0.0;
        bounds[xx * 2 + 0] = xmin;
        bounds[xx * 2 + 1] = xmax;
        for (int x = 0; x < ksize; x++) {
            double v = k[x];
            coeffs[xx * ksize + x] =
                v < 0 ? (int)(-0.5 + v * (1 << ORC_PREC_BITS)) : (int)(0.5 + v * (1 << ORC_PREC_BITS));
        }
    }
    free(k);
}

/* src: h x w x c (tight), dst: oh x ow x c.  Two passes with a uint8 intermediate (Resample.c ImagingResampleInner).
 * Pass order: horizontal first -- except for very tall images that shrink vertically, where Pillow (12.2.0) runs the
 * vertical pass first: both passes needed, h > 100 * w and oh < h.  The intermediate is rounded to uint8, so the order
 * is visible in the result.  Rule found by probing Image.resize here (boundary exact at h = 100 w + 1 for L and RGB,
 * independent of the target width; oh >= h keeps the horizontal pass first) after the degenerate-shape sweep
 * (tools/sweep_dropin_vs_reference.py --tiny) showed differences on 4 x 2212 -> 3 x 2000 pages. */
static int orc_vertical_first(int h, int w, int oh, int ow) {
    return ow != w && oh != h && (long long)h > 100LL * w && oh < h;
}
static void orc_resize_h(const uint8_t *src, int h, int w, int c, uint8_t *dst, int ow) {
    int kx = orc_lanczos_ksize(w, ow);
    int32_t *bx = (int32_t *)malloc(sizeof(int32_t) * 2 * ow);
    int32_t *cx = (int32_t *)malloc(sizeof(int32_t) * (size_t)kx * ow);
    orc_lanczos_coeffs(w, ow, bx, cx);
    for (int y = 0; y < h; y++) {
        const uint8_t *row = src + (size_t)y * w * c;
        uint8_t *orow = dst + (size_t)y * ow * c;
        for (int xx = 0; xx < ow; xx++) {
            int xmin = bx[xx * 2], n = bx[xx * 2 + 1];
            const int32_t *k = cx + (size_t)xx * kx;
            for (int ch = 0; ch < c; ch++) {
                int ss = 1 << (ORC_PREC_BITS - 1);
                for (int x = 0; x < n; x++) ss += row[(xmin + x) * c + ch] * k[x];
                orow[xx * c + ch] = sat_u8(ss >> ORC_PREC_BITS);
            }
        }
    }
    free(bx); free(cx);
}
static void orc_resize_v(const uint8_t *src, int h, int row_bytes, uint8_t *dst, int oh) {
    int ky = orc_lanczos_ksize(h, oh);
    int32_t *by = (int32_t *)malloc(sizeof(int32_t) * 2 * oh);
    int32_t *cy = (int32_t *)malloc(sizeof(int32_t) * (size_t)ky * oh);
    orc_lanczos_coeffs(h, oh, by, cy);
    for (int yy = 0; yy < oh; yy++) {
        int ymin = by[yy * 2], n = by[yy * 2 + 1];
        const int32_t *k = cy + (size_t)yy * ky;
        uint8_t *orow = dst + (size_t)yy * row_bytes;
        for (int i = 0; i < row_bytes; i++) {
            int ss = 1 << (ORC_PREC_BITS - 1);
            for (int y = 0; y < n; y++) ss += src[(size_t)(ymin + y) * row_bytes + i] * k[y];
            orow[i] = sat_u8(ss >> ORC_PREC_BITS);
        }
    }
    free(by); free(cy);
}
ORC_API void orc_resize_lanczos_u8(const uint8_t *src, int h, int w, int c, uint8_t *dst, int oh, int ow) {
    if (ow == w && oh == h) { memcpy(dst, src, (size_t)h * w * c); return; }
    if (ow == w) { orc_resize_v(src, h, w * c, dst, oh); return; }
    if (oh == h) { orc_resize_h(src, h, w, c, dst, ow); return; }
    if (orc_vertical_first(h, w, oh, ow)) {
        uint8_t *tmp = (uint8_t *)malloc((size_t)oh * w * c);
        orc_resize_v(src, h, w * c, tmp, oh);
        orc_resize_h(tmp, oh, w, c, dst, ow);
        free(tmp);
    } else {
        uint8_t *tmp = (uint8_t *)malloc((size_t)h * ow * c);
        orc_resize_h(src, h, w, c, tmp, ow);
        orc_resize_v(tmp, h, ow * c, dst, oh);
        free(tmp);
    }
}

/* ------------------------------------------------------------------------- *
 * A2. Grayscale.  PIL convert('L') (Convert.c rgb2l)  image_preprocessing.py:169
 *     OpenCV RGB2BGR + BGR2GRAY (color_rgb)           image_preprocessing.py:395-396
 * ------------------------------------------------------------------------- */
ORC_API void orc_gray_pil(const uint8_t *rgb, size_t npx, uint8_t *out) {
    for (size_t i = 0; i < npx; i++)
        out[i] = (uint8_t)((19595u * rgb[i * 3] + 38470u * rgb[i * 3 + 1] + 7471u * rgb[i * 3 + 2] + 0x8000u) >> 16);
}
ORC_API void orc_gray_cv(const uint8_t *rgb, size_t npx, uint8_t *out) {
    for (size_t i = 0; i < npx; i++)
        out[i] = (uint8_t)((9798u * rgb[i * 3] + 19235u * rgb[i * 3 + 1] + 3735u * rgb[i * 3 + 2] + (1u << 14)) >> 15);
}

/* ------------------------------------------------------------------------- *
 * A3. ImageEnhance.Contrast(img).enhance(f)   image_preprocessing.py:143-144
 *     mean = int(ImageStat.Stat(img.convert("L")).mean[0] + 0.5)
 *     out  = Image.blend(const(mean), img, f)   (libImaging/Blend.c)
 * ------------------------------------------------------------------------- */
static inline uint8_t orc_blend_px(uint8_t in1, uint8_t in2, float alpha, int interp) {
    float t = (float)((int)in1 + alpha * (float)((int)in2 - (int)in1));
    if (interp) return (uint8_t)t;
    if (t <= 0.0f) return 0;
    if (t >= 255.0f) return 255;
    return (uint8_t)t;
}
ORC_API int orc_contrast_mean(const uint8_t *img, size_t npx, int c) {
    uint64_t hist[256] = {0};
    for (size_t i = 0; i < npx; i++) {
        uint8_t l = c == 3 ? (uint8_t)((19595u * img[i * 3] + 38470u * img[i * 3 + 1] + 7471u * img[i * 3 + 2] + 0x8000u) >> 16)
                           : img[i];
        hist[l]++;
    }
    double sum = 0.0; /* ImageStat: sum of i*h[i] in Python float, count as int */
    uint64_t isum = 0, cnt = 0;
    for (int i = 0; i < 256; i++) { isum += (uint64_t)i * hist[i]; cnt += hist[i]; }
    sum = (double)isum;
    return (int)(sum / (double)cnt + 0.5);
}
ORC_API void orc_blend_const(const uint8_t *img, size_t nbytes, int mean, float alpha, uint8_t *out) {
    int interp = (alpha >= 0.0f && alpha <= 1.0f);
    for (size_t i = 0; i < nbytes; i++) out[i] = orc_blend_px((uint8_t)mean, img[i], alpha, interp);
}
ORC_API void orc_contrast(const uint8_t *img, int h, int w, int c, float factor, uint8_t *out) {
    int mean = orc_contrast_mean(img, (size_t)h * w, c);
    orc_blend_const(img, (size_t)h * w * c, mean, factor, out);
}

/* ------------------------------------------------------------------------- *
 * A4. ImageEnhance.Sharpness(img).enhance(f)  image_preprocessing.py:157-158
 *     degenerate = img.filter(ImageFilter.SMOOTH) (libImaging/Filter.c 3x3,
 *     kernel (1,1,1,1,5,1,1,1,1)/13, float32, border copied), then blend.
 * ------------------------------------------------------------------------- */
ORC_API void orc_smooth3(const uint8_t *img, int h, int w, int c, uint8_t *out) {
    float k[9];
    const float kk[9] = {1, 1, 1, 1, 5, 1, 1, 1, 1};
    for (int i = 0; i < 9; i++) k[i] = kk[i] / 13.0f; /* ImageFilter scale applied in float32 */
    size_t pitch = (size_t)w * c;
    memcpy(out, img, (size_t)h * pitch);
    for (int y = 1; y < h - 1; y++) {
        const uint8_t *in_1 = img + (y - 1) * pitch, *in0 = img + y * pitch, *in1 = img + (y + 1) * pitch;
        uint8_t *o = out + y * pitch;
        for (int x = 1; x < w - 1; x++)
            for (int ch = 0; ch < c; ch++) {
                int i = x * c + ch;
                float ss = 0.5f;
                ss += (float)in1[i - c] * k[0] + (float)in1[i] * k[1] + (float)in1[i + c] * k[2];
                ss += (float)in0[i - c] * k[3] + (float)in0[i] * k[4] + (float)in0[i + c] * k[5];
                ss += (float)in_1[i - c] * k[6] + (float)in_1[i] * k[7] + (float)in_1[i + c] * k[8];
                o[i] = ss <= 0.0f ? 0 : (ss >= 255.0f ? 255 : (uint8_t)ss);
            }
    }
}
ORC_API void orc_blend2(const uint8_t *in1, const uint8_t *in2, size_t nbytes, float alpha, uint8_t *out) {
    int interp = (alpha >= 0.0f && alpha <= 1.0f);
    for (size_t i = 0; i < nbytes; i++) out[i] = orc_blend_px(in1[i], in2[i], alpha, interp);
}
ORC_API void orc_sharpness(const uint8_t *img, int h, int w, int c, float factor, uint8_t *out) {
    uint8_t *s = (uint8_t *)malloc((size_t)h * w * c);
    orc_smooth3(img, h, w, c, s);
    orc_blend2(s, img, (size_t)h * w * c, factor, out);
    free(s);
}

/* ------------------------------------------------------------------------- *
 * A5. ImageFilter.MedianFilter(3) (RankFilter.c, edge-replicated expand)
 *     image_preprocessing.py:165 ; fixed threshold  image_preprocessing.py:184-185
 * ------------------------------------------------------------------------- */
ORC_API void orc_median3(const uint8_t *img, int h, int w, int c, uint8_t *out) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int ch = 0; ch < c; ch++) {
                uint8_t v[9];
                int n = 0;
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        int yy = clampi(y + dy, 0, h - 1), xx = clampi(x + dx, 0, w - 1);
                        v[n++] = img[((size_t)yy * w + xx) * c + ch];
                    }
                for (int i = 1; i < 9; i++) { /* insertion sort */
                    uint8_t t = v[i];
                    int j = i - 1;
                    while (j >= 0 && v[j] > t) { v[j + 1] = v[j]; j--; }
                    v[j + 1] = t;
                }
                out[((size_t)y * w + x) * c + ch] = v[4];
            }
}
/* L (already gray) -> 0/255 bytes; PIL mode "1" stores 0/255 per pixel */
ORC_API void orc_threshold(const uint8_t *gray, size_t npx, int thr, uint8_t *out) {
    for (size_t i = 0; i < npx; i++) out[i] = gray[i] > thr ? 255 : 0;
}

/* ------------------------------------------------------------------------- *
 * A0. ImageOps.exif_transpose  image_preprocessing.py:173
 *     orientation 2 FLIP_LR, 3 ROT180, 4 FLIP_TB, 5 TRANSPOSE, 6 ROT270,
 *     7 TRANSVERSE, 8 ROT90; else copy.  out dims: swapped for 5..8.
 * ------------------------------------------------------------------------- */
ORC_API void orc_exif_transpose(const uint8_t *img, int h, int w, int c, int orientation, uint8_t *out) {
    int oh = (orientation >= 5 && orientation <= 8) ? w : h;
    int ow = (orientation >= 5 && orientation <= 8) ? h : w;
    for (int y = 0; y < oh; y++)
        for (int x = 0; x < ow; x++) {
            int sy, sx;
            switch (orientation) {
            case 2: sy = y; sx = w - 1 - x; break;
            case 3: sy = h - 1 - y; sx = w - 1 - x; break;
            case 4: sy = h - 1 - y; sx = x; break;
            case 5: sy = x; sx = y; break;                 /* TRANSPOSE */
            case 6: sy = h - 1 - x; sx = y; break;         /* ROTATE_270 */
            case 7: sy = h - 1 - x; sx = w - 1 - y; break; /* TRANSVERSE */
            case 8: sy = x; sx = w - 1 - y; break;         /* ROTATE_90 */
            default: sy = y; sx = x; break;
            }
            memcpy(out + ((size_t)y * ow + x) * c, img + ((size_t)sy * w + sx) * c, c);
        }
}

/* ------------------------------------------------------------------------- *
 * A6. cv2.adaptiveThreshold(gray,255,GAUSSIAN_C,BINARY,11,2)
 *     image_preprocessing.py:486-492.  fp32 separable Gaussian (sigma=2.0,
 *     BORDER_REPLICATE), rint, dst = (src - mean > -C) ? 255 : 0.
 *     Row pass: RowFilter (taps accumulated left to right); column pass:
 *     SymmColumnFilter (centre tap, then k[i]*(S[+i]+S[-i]) outward).  This
 *     order is bit-identical to OpenCV with cv2.setUseOptimized(False); the
 *     AVX/FMA-dispatched build differs from its own plain path on ~1e-6 of
 *     mask pixels (rounding ties), so tests pin against the plain path.
 * ------------------------------------------------------------------------- */
ORC_API void orc_gauss11_kernel(float *k) {
    /* getGaussianKernel(11, sigma<=0 -> 0.3*((11-1)*0.5-1)+0.8 = 2.0), float32 */
    double sigma = 2.0, scale2x = -0.5 / (sigma * sigma), sum = 0.0;
    double t[11];
    for (int i = 0; i < 11; i++) {
        double x = i - 5.0;
        t[i] = exp(scale2x * x * x);
        sum += t[i];
    }
    sum = 1.0 / sum;
    for (int i = 0; i < 11; i++) k[i] = (float)(t[i] * sum);
}
/* dispatch 0: OpenCV's plain path (cv2.setUseOptimized(False), or a CPU without AVX2): every product and sum rounded
 * separately.  dispatch 1: OpenCV's default dispatch on x86 hosts with AVX2 + FMA3 (filter.simd.hpp built for AVX2):
 * the 8-lane vector loops of both passes use fused multiply-add (v_muladd), the row filter's 4-lane step behind them
 * too; the scalar remainders do not.  I.e. with t = w % 8: row pass fused for x < w - (w % 4), column pass fused for
 * x < w - t.  Found by probing cv2.adaptiveThreshold (4.13.0) here: this model reproduces the optimised build on 1080
 * random / text pages and 800 narrow pages (3.8 M tail rows), where the plain order differs on ~10 % of the pages
 * by one pixel (tests/golden/make_adaptive_golden.py pins it). */
ORC_API void orc_adaptive_gauss11_x(const uint8_t *gray, int h, int w, int cval, int dispatch, uint8_t *out) {
    float k[11];
    orc_gauss11_kernel(k);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)h * w);
    const int row_fused_end = dispatch ? w - (w % 4) : 0, col_fused_end = dispatch ? w - (w % 8) : 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float s = 0.0f;
            if (x < row_fused_end)
                for (int i = 0; i < 11; i++) s = fmaf((float)gray[(size_t)y * w + clampi(x + i - 5, 0, w - 1)], k[i], s);
            else
                for (int i = 0; i < 11; i++) s += (float)gray[(size_t)y * w + clampi(x + i - 5, 0, w - 1)] * k[i];
            tmp[(size_t)y * w + x] = s;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            /* SymmColumnFilter: centre tap, then symmetric pairs outward */
            float s = k[5] * tmp[(size_t)y * w + x];
            for (int i = 1; i <= 5; i++) {
                const float pr = tmp[(size_t)clampi(y + i, 0, h - 1) * w + x] + tmp[(size_t)clampi(y - i, 0, h - 1) * w + x];
                if (x < col_fused_end) s = fmaf(k[5 + i], pr, s);
                else s += k[5 + i] * pr;
            }
            int mean = (int)lrintf(s); /* round-half-even */
            if (mean < 0) mean = 0; if (mean > 255) mean = 255;
            out[(size_t)y * w + x] = ((int)gray[(size_t)y * w + x] - mean > -cval) ? 255 : 0;
        }
    free(tmp);
}
ORC_API void orc_adaptive_gauss11(const uint8_t *gray, int h, int w, int cval, uint8_t *out) {
    orc_adaptive_gauss11_x(gray, h, w, cval, 0, out);
}

/* ------------------------------------------------------------------------- *
 * A7. cv2.Canny(gray, 50, 150, apertureSize=3)   image_preprocessing.py:399
 *     Sobel 3x3 replicate border, L1 magnitude, NMS with TG22 fixed point,
 *     hysteresis by 8-connected flood from strong pixels.
 * ------------------------------------------------------------------------- */
ORC_API void orc_canny(const uint8_t *g, int h, int w, int low, int high, uint8_t *edges) {
    size_t n = (size_t)h * w;
    int16_t *dx = (int16_t *)malloc(n * 2), *dy = (int16_t *)malloc(n * 2);
    int *mag = (int *)calloc((size_t)(h + 2) * (w + 2), sizeof(int)); /* zero ring */
    uint8_t *map = (uint8_t *)malloc(n);
#define G(y, x) ((int)g[(size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)])
#define MAG(y, x) mag[(size_t)((y) + 1) * (w + 2) + (x) + 1]
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int gx = (G(y - 1, x + 1) + 2 * G(y, x + 1) + G(y + 1, x + 1)) - (G(y - 1, x - 1) + 2 * G(y, x - 1) + G(y + 1, x - 1));
            int gy = (G(y + 1, x - 1) + 2 * G(y + 1, x) + G(y + 1, x + 1)) - (G(y - 1, x - 1) + 2 * G(y - 1, x) + G(y - 1, x + 1));
            dx[(size_t)y * w + x] = (int16_t)gx;
            dy[(size_t)y * w + x] = (int16_t)gy;
            MAG(y, x) = abs(gx) + abs(gy);
        }
    const int TG22 = 13573; /* (int)(0.4142135623730950488016887242097*(1<<15)+0.5) */
    size_t *stack = (size_t *)malloc(n * sizeof(size_t));
    size_t sp = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int m = MAG(y, x);
            uint8_t r = 1;
            if (m > low) {
                int xs = dx[(size_t)y * w + x], ys = dy[(size_t)y * w + x];
                int ax = abs(xs), ay = abs(ys) << 15;
                int tg22x = ax * TG22;
                int pass = 0;
                if (ay < tg22x) pass = (m > MAG(y, x - 1) && m >= MAG(y, x + 1));
                else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) pass = (m > MAG(y - 1, x) && m >= MAG(y + 1, x));
                    else {
                        int s = (xs ^ ys) < 0 ? -1 : 1;
                        pass = (m > MAG(y - 1, x - s) && m > MAG(y + 1, x + s));
                    }
                }
                if (pass) {
                    if (m > high) { r = 2; stack[sp++] = (size_t)y * w + x; }
                    else r = 0;
                }
            }
            map[(size_t)y * w + x] = r;
        }
    while (sp) {
        size_t p = stack[--sp];
        int y = (int)(p / w), x = (int)(p % w);
        for (int ddy = -1; ddy <= 1; ddy++)
            for (int ddx = -1; ddx <= 1; ddx++) {
                int yy = y + ddy, xx = x + ddx;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                size_t q = (size_t)yy * w + xx;
                if (map[q] == 0) { map[q] = 2; stack[sp++] = q; }
            }
    }
    for (size_t i = 0; i < n; i++) edges[i] = map[i] == 2 ? 255 : 0;
#undef G
#undef MAG
    free(dx); free(dy); free(mag); free(map); free(stack);
}

/* ------------------------------------------------------------------------- *
 * A8. cv2.HoughLinesP(edges, 1, pi/180, 100, minLineLength=100, maxLineGap=10)
 *     image_preprocessing.py:402-407 -> hough.cpp HoughLinesProbabilistic
 *     (progressive probabilistic Hough, cv::RNG seeded with (uint64)-1).
 *     lines_out: max_lines x 4 int32 (x0,y0,x1,y1); returns line count.
 * ------------------------------------------------------------------------- */
static inline int orc_cvround_f(float v) { return (int)lrintf(v); }
static inline int orc_cvround_d(double v) { return (int)lrint(v); }

ORC_API int orc_ppht(const uint8_t *edges, int h, int w, double rho_d, double theta_d, int threshold,
                     int line_length, int line_gap, int32_t *lines_out, int max_lines) {
    float rho = (float)rho_d, theta = (float)theta_d;
    float irho = 1 / rho;
    int numangle = orc_cvround_d(M_PI / theta);
    int numrho = orc_cvround_d(((w + h) * 2 + 1) / rho);
    int32_t *accum = (int32_t *)calloc((size_t)numangle * numrho, sizeof(int32_t));
    uint8_t *mask = (uint8_t *)malloc((size_t)h * w);
    float *ttab = (float *)malloc(sizeof(float) * numangle * 2);
    for (int n = 0; n < numangle; n++) {
        ttab[n * 2] = (float)(cos((double)n * theta) * irho);
        ttab[n * 2 + 1] = (float)(sin((double)n * theta) * irho);
    }
    int32_t *nz = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)h * w);
    int count = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            if (edges[(size_t)y * w + x]) { mask[(size_t)y * w + x] = 1; nz[count * 2] = x; nz[count * 2 + 1] = y; count++; }
            else mask[(size_t)y * w + x] = 0;
        }
    uint64_t state = (uint64_t)-1;
    int nlines = 0;
    const int shift = 16;
    for (; count > 0; count--) {
        state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
        int idx = (int)((uint32_t)state % (uint32_t)count);
        int max_val = threshold - 1, max_n = 0;
        int j = nz[idx * 2], i = nz[idx * 2 + 1];
        int le_x[2] = {0, 0}, le_y[2] = {0, 0};
        nz[idx * 2] = nz[(count - 1) * 2];
        nz[idx * 2 + 1] = nz[(count - 1) * 2 + 1];
        if (!mask[(size_t)i * w + j]) continue;
        int32_t *adata = accum;
        for (int n = 0; n < numangle; n++, adata += numrho) {
            int r = orc_cvround_f((float)j * ttab[n * 2] + (float)i * ttab[n * 2 + 1]);
            r += (numrho - 1) / 2;
            int val = ++adata[r];
            if (max_val < val) { max_val = val; max_n = n; }
        }
        if (max_val < threshold) continue;
        float a = -ttab[max_n * 2 + 1], b = ttab[max_n * 2];
        int x0 = j, y0 = i, dx0, dy0, xflag;
        if (fabsf(a) > fabsf(b)) {
            xflag = 1;
            dx0 = a > 0 ? 1 : -1;
            dy0 = orc_cvround_d(b * (1 << shift) / fabs(a));
            y0 = (y0 << shift) + (1 << (shift - 1));
        } else {
            xflag = 0;
            dy0 = b > 0 ? 1 : -1;
            dx0 = orc_cvround_d(a * (1 << shift) / fabs(b));
            x0 = (x0 << shift) + (1 << (shift - 1));
        }
        for (int k = 0; k < 2; k++) {
            int gap = 0, x = x0, y = y0, dx = dx0, dy = dy0;
            if (k > 0) { dx = -dx; dy = -dy; }
            for (;; x += dx, y += dy) {
                int i1, j1;
                if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
                if (j1 < 0 || j1 >= w || i1 < 0 || i1 >= h) break;
                if (mask[(size_t)i1 * w + j1]) { gap = 0; le_y[k] = i1; le_x[k] = j1; }
                else if (++gap > line_gap) break;
            }
        }
        int good = abs(le_x[1] - le_x[0]) >= line_length || abs(le_y[1] - le_y[0]) >= line_length;
        for (int k = 0; k < 2; k++) {
            int x = x0, y = y0, dx = dx0, dy = dy0;
            if (k > 0) { dx = -dx; dy = -dy; }
            for (;; x += dx, y += dy) {
                int i1, j1;
                if (xflag) { j1 = x; i1 = y >> shift; } else { j1 = x >> shift; i1 = y; }
                uint8_t *m = mask + (size_t)i1 * w + j1;
                if (*m) {
                    if (good) {
                        adata = accum;
                        for (int n = 0; n < numangle; n++, adata += numrho) {
                            int r = orc_cvround_f((float)j1 * ttab[n * 2] + (float)i1 * ttab[n * 2 + 1]);
                            r += (numrho - 1) / 2;
                            adata[r]--;
                        }
                    }
                    *m = 0;
                }
                if (i1 == le_y[k] && j1 == le_x[k]) break;
            }
        }
        if (good) {
            if (nlines < max_lines) {
                lines_out[nlines * 4 + 0] = le_x[0]; lines_out[nlines * 4 + 1] = le_y[0];
                lines_out[nlines * 4 + 2] = le_x[1]; lines_out[nlines * 4 + 3] = le_y[1];
            }
            nlines++;
        }
    }
    free(accum); free(mask); free(ttab); free(nz);
    return nlines;
}

/* deskew angle: image_preprocessing.py:414-428 (arctan2 -> degrees -> fold -> median) */
static int orc_cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
ORC_API double orc_median_angle(const int32_t *lines, int nlines) {
    if (nlines <= 0) return 0.0;
    double *ang = (double *)malloc(sizeof(double) * nlines);
    for (int i = 0; i < nlines; i++) {
        double dy = (double)(lines[i * 4 + 3] - lines[i * 4 + 1]), dx = (double)(lines[i * 4 + 2] - lines[i * 4 + 0]);
        double a = atan2(dy, dx) * (180.0 / M_PI); /* np.degrees */
        if (a < -45) a += 90; else if (a > 45) a -= 90;
        ang[i] = a;
    }
    qsort(ang, nlines, sizeof(double), orc_cmp_double);
    double m = (nlines & 1) ? ang[nlines / 2] : (ang[nlines / 2 - 1] + ang[nlines / 2]) / 2.0; /* np.mean of the two */
    free(ang);
    return m;
}

/* ------------------------------------------------------------------------- *
 * A9. cv2.getRotationMatrix2D + cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE)
 *     image_preprocessing.py:442-450 -> imgwarp.cpp WarpAffineInvoker +
 *     remapBicubic<FixedPtCast<int,uchar,15>> with the 32x32 int16 table.
 * ------------------------------------------------------------------------- */
ORC_API void orc_rotation_matrix(double cx, double cy, double angle_deg, double scale, double *M) {
    double ang = angle_deg * (M_PI / 180.0);
    double alpha = cos(ang) * scale, beta = sin(ang) * scale;
    M[0] = alpha; M[1] = beta; M[2] = (1 - alpha) * cx - beta * cy;
    M[3] = -beta; M[4] = alpha; M[5] = beta * cx + (1 - alpha) * cy;
}

#define ORC_TAB_SZ 32
static int16_t orc_cubic_tab[ORC_TAB_SZ * ORC_TAB_SZ][4][4];
static int orc_cubic_tab_ready = 0;

static void orc_cubic_coeffs(float x, float *c) {
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}
ORC_API void orc_cubic_table(int16_t *tab_out /* 1024*16 */) {
    if (!orc_cubic_tab_ready) {
        float t1[ORC_TAB_SZ][4];
        for (int i = 0; i < ORC_TAB_SZ; i++) orc_cubic_coeffs((float)i * (1.0f / ORC_TAB_SZ), t1[i]);
        for (int i = 0; i < ORC_TAB_SZ; i++)
            for (int j = 0; j < ORC_TAB_SZ; j++) {
                int16_t(*it)[4] = orc_cubic_tab[i * ORC_TAB_SZ + j];
                int isum = 0;
                for (int k1 = 0; k1 < 4; k1++) {
                    float vy = t1[i][k1];
                    for (int k2 = 0; k2 < 4; k2++) {
                        float v = vy * t1[j][k2];
                        int iv = (int)lrintf(v * 32768.0f);
                        it[k1][k2] = (int16_t)(iv > 32767 ? 32767 : (iv < -32768 ? -32768 : iv));
                        isum += it[k1][k2];
                    }
                }
                if (isum != 32768) {
                    int diff = isum - 32768;
                    int ksize2 = 2, Mk1 = ksize2, Mk2 = ksize2, mk1 = ksize2, mk2 = ksize2;
                    for (int k1 = ksize2; k1 < ksize2 + 2; k1++)
                        for (int k2 = ksize2; k2 < ksize2 + 2; k2++) {
                            if (it[k1][k2] < it[mk1][mk2]) { mk1 = k1; mk2 = k2; }
                            else if (it[k1][k2] > it[Mk1][Mk2]) { Mk1 = k1; Mk2 = k2; }
                        }
                    if (diff < 0) it[Mk1][Mk2] = (int16_t)(it[Mk1][Mk2] - diff);
                    else it[mk1][mk2] = (int16_t)(it[mk1][mk2] - diff);
                }
            }
        orc_cubic_tab_ready = 1;
    }
    if (tab_out) memcpy(tab_out, orc_cubic_tab, sizeof(orc_cubic_tab));
}

/* M = forward 2x3 matrix as passed to cv2.warpAffine (no WARP_INVERSE_MAP) */
ORC_API void orc_warp_affine_cubic_u8(const uint8_t *src, int h, int w, int c, const double *Mfwd, uint8_t *dst) {
    orc_cubic_table(NULL);
    double M[6];
    double D = Mfwd[0] * Mfwd[4] - Mfwd[1] * Mfwd[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = Mfwd[4] * D, A22 = Mfwd[0] * D;
    M[0] = A11; M[1] = Mfwd[1] * (-D); M[3] = Mfwd[3] * (-D); M[4] = A22;
    double b1 = -M[0] * Mfwd[2] - M[1] * Mfwd[5];
    double b2 = -M[3] * Mfwd[2] - M[4] * Mfwd[5];
    M[2] = b1; M[5] = b2;
    const int AB_BITS = 10, AB_SCALE = 1 << AB_BITS, INTER_BITS = 5, INTER_TAB = 1 << INTER_BITS;
    const int round_delta = AB_SCALE / INTER_TAB / 2;
    int *adelta = (int *)malloc(sizeof(int) * w), *bdelta = (int *)malloc(sizeof(int) * w);
    for (int x = 0; x < w; x++) {
        adelta[x] = (int)lrint(M[0] * x * AB_SCALE); /* saturate_cast<int>(double) = cvRound */
        bdelta[x] = (int)lrint(M[3] * x * AB_SCALE);
    }
    for (int y = 0; y < h; y++) {
        int X0 = (int)lrint((M[1] * y + M[2]) * AB_SCALE) + round_delta;
        int Y0 = (int)lrint((M[4] * y + M[5]) * AB_SCALE) + round_delta;
        for (int x = 0; x < w; x++) {
            int X = (X0 + adelta[x]) >> (AB_BITS - INTER_BITS);
            int Y = (Y0 + bdelta[x]) >> (AB_BITS - INTER_BITS);
            /* xy stored as saturate_cast<short>; pages < 32768 so no saturation in range */
            int sx = clampi(X >> INTER_BITS, -32768, 32767) - 1, sy = clampi(Y >> INTER_BITS, -32768, 32767) - 1;
            int ai = (Y & (INTER_TAB - 1)) * INTER_TAB + (X & (INTER_TAB - 1));
            const int16_t(*wt)[4] = orc_cubic_tab[ai];
            for (int ch = 0; ch < c; ch++) {
                int sum = 0;
                for (int k1 = 0; k1 < 4; k1++) {
                    int yy = clampi(sy + k1, 0, h - 1);
                    for (int k2 = 0; k2 < 4; k2++) {
                        int xx = clampi(sx + k2, 0, w - 1);
                        sum += src[((size_t)yy * w + xx) * c + ch] * wt[k1][k2];
                    }
                }
                dst[((size_t)y * w + x) * c + ch] = sat_u8((sum + (1 << 14)) >> 15);
            }
        }
    }
    free(adelta); free(bdelta);
}

/* ------------------------------------------------------------------------- *
 * B3 [upstream PaddleOCR DetResizeForTest + NormalizeImage + ToCHWImage].
 *     cv2.resize(INTER_LINEAR) on uint8 (resize.cpp: 11-bit fixed-point
 *     coefficients, HResize to int32 then VResizeLinear), then
 *     (x*(1/255) - mean)/std in float32, HWC -> CHW.   parity unpinned by the
 *     reference (not in its tree); pinned against cv2.resize in tests.
 * ------------------------------------------------------------------------- */
ORC_API void orc_det_target_size(int h, int w, int limit, int *rh, int *rw) {
    /* limit_type = 'max' */
    double ratio = 1.0;
    int mx = h > w ? h : w;
    if (mx > limit) ratio = (double)limit / mx;
    int a = (int)(h * ratio), b = (int)(w * ratio);
    /* python round(): half-to-even */
    a = (int)(lrint(a / 32.0) * 32); b = (int)(lrint(b / 32.0) * 32);
    *rh = a < 32 ? 32 : a;
    *rw = b < 32 ? 32 : b;
}
ORC_API void orc_resize_linear_u8(const uint8_t *src, int h, int w, int c, uint8_t *dst, int oh, int ow) {
    const int COEF_BITS = 11, ONE = 1 << COEF_BITS;
    double sx = (double)w / ow, sy = (double)h / oh; /* inv_scale = ow/w; scale = 1/inv_scale */
    double inv_x = (double)ow / w, inv_y = (double)oh / h;
    sx = 1. / inv_x; sy = 1. / inv_y;
    int *xofs = (int *)malloc(sizeof(int) * ow), *yofs = (int *)malloc(sizeof(int) * oh);
    short *ax = (short *)malloc(sizeof(short) * 2 * ow), *ay = (short *)malloc(sizeof(short) * 2 * oh);
    for (int dxi = 0; dxi < ow; dxi++) {
        float fx = (float)((dxi + 0.5) * sx - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        if (s < 0) { fx = 0; s = 0; }
        if (s >= w - 1) { fx = 0; s = w - 1; }
        xofs[dxi] = s;
        float c0 = 1.f - fx, c1 = fx;
        ax[dxi * 2] = (short)lrintf(c0 * ONE); /* saturate_cast<short>(float) = cvRound */
        ax[dxi * 2 + 1] = (short)lrintf(c1 * ONE);
    }
    for (int dyi = 0; dyi < oh; dyi++) {
        float fy = (float)((dyi + 0.5) * sy - 0.5);
        int s = (int)floorf(fy);
        fy -= s;
        yofs[dyi] = s;
        float c0 = 1.f - fy, c1 = fy;
        ay[dyi * 2] = (short)lrintf(c0 * ONE);
        ay[dyi * 2 + 1] = (short)lrintf(c1 * ONE);
    }
    int *r0 = (int *)malloc(sizeof(int) * ow * c), *r1 = (int *)malloc(sizeof(int) * ow * c);
    for (int dyi = 0; dyi < oh; dyi++) {
        int sy0 = clampi(yofs[dyi], 0, h - 1), sy1 = clampi(yofs[dyi] + 1, 0, h - 1);
        const uint8_t *s0 = src + (size_t)sy0 * w * c, *s1 = src + (size_t)sy1 * w * c;
        for (int dxi = 0; dxi < ow; dxi++) {
            int x0 = xofs[dxi], x1 = x0 + 1 < w ? x0 + 1 : w - 1;
            for (int ch = 0; ch < c; ch++) {
                r0[dxi * c + ch] = s0[x0 * c + ch] * ax[dxi * 2] + s0[x1 * c + ch] * ax[dxi * 2 + 1];
                r1[dxi * c + ch] = s1[x0 * c + ch] * ax[dxi * 2] + s1[x1 * c + ch] * ax[dxi * 2 + 1];
            }
        }
        short b0 = ay[dyi * 2], b1 = ay[dyi * 2 + 1];
        uint8_t *o = dst + (size_t)dyi * ow * c;
        for (int i = 0; i < ow * c; i++)
            o[i] = sat_u8((((b0 * (r0[i] >> 4)) >> 16) + ((b1 * (r1[i] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(yofs); free(ax); free(ay); free(r0); free(r1);
}
/* img HWC u8 (oh x ow x 3) -> CHW float32 */
ORC_API void orc_normalize_chw(const uint8_t *img, int h, int w, const float *mean, const float *std, float scale, float *out) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int ch = 0; ch < 3; ch++) {
                float v = (float)img[((size_t)y * w + x) * 3 + ch] * scale;
                out[((size_t)ch * h + y) * w + x] = (v - mean[ch]) / std[ch];
            }
}

/* ------------------------------------------------------------------------- *
 * B2 [upstream PaddleOCR CTCLabelDecode]: argmax (first max), max prob,
 *     collapse repeats, drop blank 0, conf = float32 mean of kept max probs.
 *     idx_out: N x T int32 kept class ids (-1 padded); len_out: N; conf_out: N
 * ------------------------------------------------------------------------- */
/* numpy float32 add.reduce: n<8 sequential; <=128: 8 accumulators, then
 * ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder sequential; else recurse. */
static float orc_np_pairwise_sum(const float *a, int n) {
    if (n < 8) {
        float r = 0.0f;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int q = 0; q < 8; q++) r[q] = a[q];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int q = 0; q < 8; q++) r[q] += a[i + q];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return orc_np_pairwise_sum(a, n2) + orc_np_pairwise_sum(a + n2, n - n2);
}
ORC_API void orc_ctc_greedy(const float *probs, int n, int t, int c, int32_t *idx_out, int32_t *pos_out,
                            int32_t *len_out, float *conf_out) {
    for (int b = 0; b < n; b++) {
        int prev = -1, len = 0;
        float kept[4096];
        for (int s = 0; s < t; s++) {
            const float *p = probs + ((size_t)b * t + s) * c;
            int best = 0;
            float bv = p[0];
            if (bv != bv) { /* numpy argmax: first NaN wins */ }
            else
                for (int k = 1; k < c; k++) {
                    if (p[k] != p[k]) { best = k; bv = p[k]; break; }
                    if (p[k] > bv) { bv = p[k]; best = k; }
                }
            int keep = (s == 0 || best != prev) && best != 0;
            prev = best;
            if (keep) {
                idx_out[(size_t)b * t + len] = best;
                pos_out[(size_t)b * t + len] = s;
                if (len < 4096) kept[len] = bv;
                len++;
            }
        }
        for (int s = len; s < t; s++) { idx_out[(size_t)b * t + s] = -1; pos_out[(size_t)b * t + s] = -1; }
        len_out[b] = len;
        /* np.mean(float32) = pairwise add.reduce / n */
        if (len == 0) conf_out[b] = 0.0f;
        else {
            float sum = orc_np_pairwise_sum(kept, len);
            conf_out[b] = sum / (float)len;
        }
    }
}

/* ------------------------------------------------------------------------- *
 * Synthetic page generator (host build of include/lumina_synth.h; the CUDA
 * build of the same header must produce identical bytes).
 * ------------------------------------------------------------------------- */
#include "../include/lumina_synth.h"
ORC_API void orc_synth_page(uint8_t *dst, int h, int w, uint64_t seed) {
    lsyn_page_t p;
    lsyn_page_init(&p, h, w, seed);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int ch = 0; ch < 3; ch++) dst[((size_t)y * w + x) * 3 + ch] = lsyn_pixel(&p, x, y, ch);
}
ORC_API double orc_synth_skew_deg(int h, int w, uint64_t seed) {
    lsyn_page_t p;
    lsyn_page_init(&p, h, w, seed);
    return atan2((double)p.sinq, (double)p.cosq) * 180.0 / M_PI;
}

/* host builds of the synthetic DB probability map / CTC posterior generators (include/lumina_synth.h) */
ORC_API void orc_synth_prob_map(float *dst, int h, int w, uint64_t seed64) {
    const uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32)) * 2654435761u + 777u;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[(size_t)y * w + x] = lsyn_prob(w, seed, x, y);
}
ORC_API void orc_synth_ctc(float *dst, int n, int T, int C, uint64_t crop0, uint32_t seed) {
    for (int i = 0; i < n; i++)
        for (int t = 0; t < T; t++) {
            uint32_t win, tie;
            const uint32_t nn = (uint32_t)(crop0 + (uint64_t)i);
            lsyn_ctc_step(seed, C, nn, t, &win, &tie);
            float *o = dst + ((size_t)i * T + t) * C;
            for (int c = 0; c < C; c++) o[c] = lsyn_ctc_value(seed, T, nn, t, (uint32_t)c, win, tie);
        }
}
