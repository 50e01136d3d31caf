/*
 * lumina_synth.h -- deterministic synthetic A4 text page, integer arithmetic
 * only, so the SAME bytes come out of the CUDA generator
 * (lumina_synth_pages_u8) and of this header compiled for the host
 * (oracle/lumina_oracle.c -> orc_synth_page).  Workload input for tests and
 * bench (paper tint 235-255, noise sigma~4, ~55 text lines
 * of dark glyph boxes with punched counters, 3 rulings, global skew within
 * +-3 degrees).  The reference ships no fixtures (.MISSING_LARGE_BLOBS), hence
 * synthetic pages.
 */
#ifndef LUMINA_SYNTH_H
#define LUMINA_SYNTH_H
#include <stdint.h>

#ifdef __CUDACC__
#define LSYN_HD __host__ __device__ __forceinline__
#else
#define LSYN_HD static inline
#endif

LSYN_HD uint32_t lsyn_mix(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
LSYN_HD uint32_t lsyn_hash3(uint32_t a, uint32_t b, uint32_t c) {
    return lsyn_mix(a * 0x9e3779b1u ^ lsyn_mix(b * 0x85ebca77u ^ lsyn_mix(c + 0x27d4eb2fu)));
}

typedef struct {
    int32_t h, w;
    int32_t cosq, sinq;      /* Q16 rotation (skew) */
    int32_t top, left, right, bottom;
    int32_t pitch, gh, cw;   /* line pitch, glyph height, cell width */
    int32_t rule[3];
    int32_t tint[3];
    uint32_t seed;
} lsyn_page_t;

LSYN_HD void lsyn_page_init(lsyn_page_t *p, int h, int w, uint64_t seed64) {
    uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32)) * 2654435761u + 12345u;
    p->h = h; p->w = w; p->seed = seed;
    /* skew: tan(a/2) = k/4096, k in [-107,107]  (|a| <= 3 deg) */
    int32_t k = (int32_t)(lsyn_hash3(seed, 1, 0) % 215u) - 107;
    int64_t t2 = (int64_t)k * k;               /* Q24 */
    int64_t den = (4096LL * 4096LL) + t2;
    p->cosq = (int32_t)((((4096LL * 4096LL) - t2) << 16) / den);
    p->sinq = (int32_t)(((2LL * k * 4096LL) << 16) / den);
    /* geometry scales with the page height (A4 300 dpi = 3508) */
    int32_t s = h; /* numerators over 3508 */
    p->top = 200 * s / 3508; p->bottom = h - 200 * s / 3508;
    p->left = 180 * s / 3508; p->right = w - 180 * s / 3508;
    p->pitch = (56 + (int32_t)(lsyn_hash3(seed, 2, 0) % 12u)) * s / 3508;
    if (p->pitch < 8) p->pitch = 8;
    p->gh = (26 + (int32_t)(lsyn_hash3(seed, 3, 0) % 10u)) * s / 3508;
    if (p->gh < 4) p->gh = 4;
    if (p->gh > p->pitch - 3) p->gh = p->pitch - 3;
    p->cw = 34 * s / 3508; if (p->cw < 6) p->cw = 6;
    for (int i = 0; i < 3; i++) {
        p->rule[i] = p->top + (int32_t)(lsyn_hash3(seed, 4, i) % (uint32_t)(p->bottom - p->top > 1 ? p->bottom - p->top : 1));
        p->tint[i] = 235 + (int32_t)(lsyn_hash3(seed, 5, i) % 21u);
    }
}

/* ink coverage of page-space point (u,v): returns 0..255 gray of the ink, or -1 for paper */
LSYN_HD int lsyn_ink(const lsyn_page_t *p, int32_t u, int32_t v) {
    if (u < p->left || u >= p->right || v < p->top || v >= p->bottom) return -1;
    int thick = p->h >= 2000 ? 1 : 0;
    for (int i = 0; i < 3; i++) {
        int32_t d = v - p->rule[i];
        if (d >= -thick && d <= 1) return 20;
    }
    int32_t li = (v - p->top) / p->pitch, vy = (v - p->top) % p->pitch;
    if (vy >= p->gh) return -1;
    int32_t ci = (u - p->left) / p->cw, ux = (u - p->left) % p->cw;
    uint32_t hsh = lsyn_hash3(p->seed, 100u + (uint32_t)li, (uint32_t)ci);
    if ((hsh & 7u) == 0u) return -1;                      /* word gap */
    /* short last lines of "paragraphs" */
    uint32_t lh = lsyn_hash3(p->seed, 7u, (uint32_t)li);
    if ((lh & 3u) == 0u && u > p->left + (int32_t)((lh >> 8) % (uint32_t)(p->right - p->left))) return -1;
    int32_t gw = p->cw * (12 + (int32_t)((hsh >> 3) % 19u)) / 34; /* 12..30 of 34 */
    if (gw < 2) gw = 2;
    if (ux >= gw) return -1;
    int32_t bx = p->cw * 5 / 34 + 1, by = p->gh * 6 / 30 + 1;
    if (gw >= p->cw * 18 / 34 && ux >= bx && ux < gw - bx && vy >= by && vy < p->gh - by) return -1; /* counter */
    return (int)((hsh >> 12) % 71u);
}

/* one output byte: pixel (x,y), channel ch */
LSYN_HD uint8_t lsyn_pixel(const lsyn_page_t *p, int32_t x, int32_t y, int ch) {
    /* 2x2 supersampling at quarter-pixel offsets, rotation about the centre in Q16 */
    int32_t cx4 = p->w * 2, cy4 = p->h * 2; /* centre in quarter px */
    int32_t acc = 0;
    for (int s = 0; s < 4; s++) {
        int32_t qx = x * 4 + 1 + 2 * (s & 1) - cx4, qy = y * 4 + 1 + (s & 2) - cy4;
        int64_t ru = (int64_t)qx * p->cosq - (int64_t)qy * p->sinq;
        int64_t rv = (int64_t)qx * p->sinq + (int64_t)qy * p->cosq;
        int32_t u = (int32_t)((ru >> 16) + cx4) >> 2, v = (int32_t)((rv >> 16) + cy4) >> 2;
        int ink = lsyn_ink(p, u, v);
        acc += ink < 0 ? p->tint[ch] : ink;
    }
    int32_t val = (acc + 2) >> 2;
    uint32_t nh = lsyn_hash3(p->seed ^ 0xa5a5a5a5u, (uint32_t)(y * p->w + x), (uint32_t)ch);
    int32_t ns = (int32_t)((nh & 255u) + ((nh >> 8) & 255u) + ((nh >> 16) & 255u) + (nh >> 24)) - 510;
    val += (ns * 7) >> 8; /* sigma ~ 4 */
    return (uint8_t)(val < 0 ? 0 : (val > 255 ? 255 : val));
}

/* ---- synthetic DB probability map ---------------------------------------------------
 * One text box per 60x30 cell (960x960 -> 512 cells, ~496 boxes): width 18-52, height 10-16, half of them
 * sheared by up to ~7 degrees, filled U(0.75,0.99) with a softened 1-px border (x0.55, still above thresh 0.3),
 * one in eight of the larger ones with an interior hole (a hole contour for findContours); background
 * U(0,0.12).  Integer decisions, one float multiply per pixel: identical floats on host and device. */
LSYN_HD float lsyn_prob(int w, uint32_t seed, int x, int y) {
    const uint32_t px = lsyn_hash3(seed ^ 0x1234567u, (uint32_t)(y * w + x), 0u);
    const float bg = (float)(px & 0xffffu) * (0.12f / 65536.0f);
    const int ci = x / 60, cj = y / 30;
    const uint32_t hc = lsyn_hash3(seed, 0x5000u + (uint32_t)cj, (uint32_t)ci);
    if ((hc & 31u) == 0u) return bg;                               /* empty cell */
    const int bw = 18 + (int)((hc >> 5) % 35u), bh = 10 + (int)((hc >> 11) % 7u);
    const int cx = ci * 60 + 30 + (int)((hc >> 14) % 5u) - 2, cy = cj * 30 + 15 + (int)((hc >> 17) % 5u) - 2;
    const int t = ((hc >> 31) & 1u) ? (int)((hc >> 20) % 33u) - 16 : 0;   /* shear, Q7 */
    const int dx = x - cx, dy = y - cy;
    int u = dx * 128 + dy * t, v = dy * 128 - dx * t;
    if (u < 0) u = -u;
    if (v < 0) v = -v;
    if (u > bw * 64 || v > bh * 64) return bg;
    if (((hc >> 27) & 7u) == 0u && bw >= 24 && bh >= 12) {         /* interior hole */
        const int hx = 1 + (int)((hc >> 30) & 1u);
        if (dx >= -hx && dx <= hx && dy >= -1 && dy <= 1) return bg * 0.75f;
    }
    float val = 0.75f + (float)(px >> 16) * (0.24f / 65536.0f);
    if (u > bw * 64 - 128 || v > bh * 64 - 128) val = val * 0.55f;
    return val;
}

/* ---- synthetic CTC posteriors --------------------------------------------------------
 * [n][T][C] float32: per step a peaked distribution (winner 0.9, the rest U(0, 2^-15)), class 0 = blank one
 * step in five, one step in four repeats the previous step's class, one in sixteen plants an exact tie with a
 * second class (numpy's first-index rule decides). */
LSYN_HD uint32_t lsyn_ctc_base(uint32_t seed, int C, uint32_t n, int t) {
    const uint32_t hb = lsyn_hash3(seed, n, (uint32_t)t);
    return (hb % 5u == 0u) ? 0u : 1u + (hb >> 3) % (uint32_t)(C - 1);
}
LSYN_HD void lsyn_ctc_step(uint32_t seed, int C, uint32_t n, int t, uint32_t *win, uint32_t *tie) {
    const uint32_t hb = lsyn_hash3(seed, n, (uint32_t)t);
    uint32_t wn = (t > 0 && ((hb >> 24) & 3u) == 0u) ? lsyn_ctc_base(seed, C, n, t - 1) : lsyn_ctc_base(seed, C, n, t);
    *win = wn;
    *tie = (((hb >> 26) & 15u) == 0u) ? (wn + 1u + (hb >> 8) % 97u) % (uint32_t)C : wn;
}
LSYN_HD float lsyn_ctc_value(uint32_t seed, int T, uint32_t n, int t, uint32_t c, uint32_t win, uint32_t tie) {
    if (c == win || c == tie) return 0.9f;
    const uint32_t hv = lsyn_hash3(seed ^ 0x9e3779b9u, n * (uint32_t)T + (uint32_t)t, c);
    return (float)(hv & 0xffffu) * (1.0f / 2147483648.0f);
}
#endif
