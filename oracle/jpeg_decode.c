/*
 * CPU ORACLE -- test infrastructure only (never linked into the product path).
 *
 * Sequential restatement of the baseline-JPEG *decode* that `Image.open(...)` performs for the reference's
 * `load_image` / `load_image_bytes` (backend/utils/image_preprocessing.py:57-75).  The codec itself is a
 * third-party dependency absent from /root/reference: Pillow (requirements.txt:20, unpinned; this image has
 * Pillow 12.2.0) -> libjpeg-turbo (API 6.2).  Its published integer pipeline is restated here:
 *   jdmarker.c   marker parsing (SOF0/SOF1 8-bit, DQT, DHT, DRI, SOS; JFIF / Adobe colour-space rule)
 *   jdhuff.c     sequential Huffman decoding, HUFF_EXTEND, DC prediction, restart intervals
 *   jidctint.c   dequantise + "islow" 8x8 inverse DCT (CONST_BITS 13, PASS1_BITS 2), level shift, range limit
 *   jdsample.c   fullsize / h2v1 / h2v2 "fancy" (triangle) upsampling, plain replication when the
 *                down-sampled width is <= 2; jdmainct.c context rows (top row / last real row replicated)
 *   jdcolor.c    YCbCr -> RGB with the 16-bit fixed-point tables
 * PINNED: tests/test_oracle_jpeg_decode.py compares this file with live Pillow decodes (np.asarray(Image.open))
 * on every sampling mode, odd sizes, optimised / standard tables and restart intervals.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static const uint8_t kZigzag[64 + 16] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
    13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45,
    38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

typedef struct {
    int present;
    uint8_t bits[17];
    uint8_t vals[256];
    /* jdhuff.c jpeg_make_d_derived_tbl */
    int32_t maxcode[18];
    int32_t valoffset[17];
} HuffTbl;

typedef struct {
    int id, h, v, tq, td, ta;
    int wblocks, hblocks; /* padded to whole MCUs (interleaved) */
    int dw, dh;           /* downsampled_width / height */
    int16_t *coef;        /* [hblocks][wblocks][64] natural order */
    uint8_t *plane;       /* [hblocks*8][wblocks*8] */
} Comp;

typedef struct {
    const uint8_t *p;
    size_t len, pos;
    uint32_t acc;
    int nbits;
    int hit_marker;
} Bits;

static void derive(HuffTbl *t) {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        if (t->bits[l]) {
            t->valoffset[l] = k - code;
            k += t->bits[l];
            code += t->bits[l];
            t->maxcode[l] = code - 1;
        } else
            t->maxcode[l] = -1;
        code <<= 1;
    }
    t->maxcode[17] = 0xFFFFF;
}

static int next_bit(Bits *b) {
    if (b->nbits == 0) {
        uint32_t c = 0;
        if (!b->hit_marker && b->pos < b->len) {
            c = b->p[b->pos];
            if (c == 0xFF) {
                if (b->pos + 1 < b->len && b->p[b->pos + 1] == 0x00)
                    b->pos += 2;
                else {
                    b->hit_marker = 1; /* jdhuff.c: feed zeros once a marker is reached */
                    c = 0;
                }
            } else
                b->pos++;
        }
        b->acc = c;
        b->nbits = 8;
    }
    b->nbits--;
    return (b->acc >> b->nbits) & 1;
}

static int get_bits(Bits *b, int n) {
    int v = 0;
    while (n--) v = (v << 1) | next_bit(b);
    return v;
}

static int decode_sym(Bits *b, const HuffTbl *t) {
    int code = next_bit(b), l = 1;
    while (l <= 16 && code > t->maxcode[l]) {
        code = (code << 1) | next_bit(b);
        l++;
    }
    if (l > 16) return 0; /* jdhuff.c: corrupt data decodes as a zero symbol */
    return t->vals[(code + t->valoffset[l]) & 0xFF];
}

static int huff_extend(int r, int s) { return r < (1 << (s - 1)) ? r + (int)((~0u) << s) + 1 : r; }

#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

static uint8_t clamp_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* jidctint.c jpeg_idct_islow (the column / row zero shortcuts give the same values as the full formula) */
static void idct_islow(const int16_t *in, const uint16_t *q, uint8_t *out, int pitch) {
    int32_t ws[64];
    for (int c = 0; c < 8; c++) {
#define DQ(r) ((int32_t)(int16_t)(in[(r)*8 + c] * q[(r)*8 + c]))
        int32_t z2 = DQ(2), z3 = DQ(6);
        int32_t z1 = (z2 + z3) * FIX_0_541196100;
        int32_t tmp2 = z1 + z3 * (-FIX_1_847759065);
        int32_t tmp3 = z1 + z2 * FIX_0_765366865;
        z2 = DQ(0);
        z3 = DQ(4);
        int32_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
        int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = DQ(7);
        tmp1 = DQ(5);
        tmp2 = DQ(3);
        tmp3 = DQ(1);
        z1 = tmp0 + tmp3;
        z2 = tmp1 + tmp2;
        z3 = tmp0 + tmp2;
        int32_t z4 = tmp1 + tmp3;
        int32_t z5 = (z3 + z4) * FIX_1_175875602;
        tmp0 *= FIX_0_298631336;
        tmp1 *= FIX_2_053119869;
        tmp2 *= FIX_3_072711026;
        tmp3 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223;
        z2 *= -FIX_2_562915447;
        z3 *= -FIX_1_961570560;
        z4 *= -FIX_0_390180644;
        z3 += z5;
        z4 += z5;
        tmp0 += z1 + z3;
        tmp1 += z2 + z4;
        tmp2 += z2 + z3;
        tmp3 += z1 + z4;
        ws[0 * 8 + c] = DESCALE(tmp10 + tmp3, 11);
        ws[7 * 8 + c] = DESCALE(tmp10 - tmp3, 11);
        ws[1 * 8 + c] = DESCALE(tmp11 + tmp2, 11);
        ws[6 * 8 + c] = DESCALE(tmp11 - tmp2, 11);
        ws[2 * 8 + c] = DESCALE(tmp12 + tmp1, 11);
        ws[5 * 8 + c] = DESCALE(tmp12 - tmp1, 11);
        ws[3 * 8 + c] = DESCALE(tmp13 + tmp0, 11);
        ws[4 * 8 + c] = DESCALE(tmp13 - tmp0, 11);
#undef DQ
    }
    for (int r = 0; r < 8; r++) {
        const int32_t *w = ws + r * 8;
        int32_t z2 = w[2], z3 = w[6];
        int32_t z1 = (z2 + z3) * FIX_0_541196100;
        int32_t tmp2 = z1 + z3 * (-FIX_1_847759065);
        int32_t tmp3 = z1 + z2 * FIX_0_765366865;
        int32_t tmp0 = (w[0] + w[4]) * 8192, tmp1 = (w[0] - w[4]) * 8192;
        int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7];
        tmp1 = w[5];
        tmp2 = w[3];
        tmp3 = w[1];
        z1 = tmp0 + tmp3;
        z2 = tmp1 + tmp2;
        z3 = tmp0 + tmp2;
        int32_t z4 = tmp1 + tmp3;
        int32_t z5 = (z3 + z4) * FIX_1_175875602;
        tmp0 *= FIX_0_298631336;
        tmp1 *= FIX_2_053119869;
        tmp2 *= FIX_3_072711026;
        tmp3 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223;
        z2 *= -FIX_2_562915447;
        z3 *= -FIX_1_961570560;
        z4 *= -FIX_0_390180644;
        z3 += z5;
        z4 += z5;
        tmp0 += z1 + z3;
        tmp1 += z2 + z4;
        tmp2 += z2 + z3;
        tmp3 += z1 + z4;
        uint8_t *o = out + (size_t)r * pitch;
        o[0] = clamp_u8(128 + DESCALE(tmp10 + tmp3, 18));
        o[7] = clamp_u8(128 + DESCALE(tmp10 - tmp3, 18));
        o[1] = clamp_u8(128 + DESCALE(tmp11 + tmp2, 18));
        o[6] = clamp_u8(128 + DESCALE(tmp11 - tmp2, 18));
        o[2] = clamp_u8(128 + DESCALE(tmp12 + tmp1, 18));
        o[5] = clamp_u8(128 + DESCALE(tmp12 - tmp1, 18));
        o[3] = clamp_u8(128 + DESCALE(tmp13 + tmp0, 18));
        o[4] = clamp_u8(128 + DESCALE(tmp13 - tmp0, 18));
    }
}

/* jdsample.c: one full-resolution row `y` of a chroma component (dw x dh real samples, plane pitch `pitch`) */
static void upsample_row(const Comp *c, int hs, int vs, int y, int out_w, uint8_t *out) {
    int pitch = c->wblocks * 8, dw = c->dw, dh = c->dh;
    if (hs == 1 && vs == 1) {
        memcpy(out, c->plane + (size_t)y * pitch, out_w);
        return;
    }
    if (hs == 2 && vs == 1) {
        const uint8_t *in = c->plane + (size_t)y * pitch;
        if (dw <= 2) {
            for (int x = 0; x < out_w; x++) out[x] = in[x >> 1];
            return;
        }
        for (int x = 0; x < out_w; x++) {
            int i = x >> 1, v = in[i];
            if (x & 1)
                out[x] = (uint8_t)(i == dw - 1 ? v : (v * 3 + in[i + 1] + 2) >> 2);
            else
                out[x] = (uint8_t)(i == 0 ? v : (v * 3 + in[i - 1] + 1) >> 2);
        }
        return;
    }
    /* h2v2 */
    int r0 = y >> 1;
    const uint8_t *in0 = c->plane + (size_t)(r0 < dh ? r0 : dh - 1) * pitch;
    if (dw <= 2) {
        for (int x = 0; x < out_w; x++) out[x] = in0[x >> 1];
        return;
    }
    int r1 = (y & 1) ? r0 + 1 : r0 - 1;
    if (r1 < 0) r1 = 0;
    if (r1 > dh - 1) r1 = dh - 1;
    const uint8_t *in1 = c->plane + (size_t)r1 * pitch;
    for (int x = 0; x < out_w; x++) {
        int i = x >> 1;
        int cur = in0[i] * 3 + in1[i];
        if (x & 1) {
            if (i == dw - 1)
                out[x] = (uint8_t)((cur * 4 + 7) >> 4);
            else
                out[x] = (uint8_t)((cur * 3 + in0[i + 1] * 3 + in1[i + 1] + 7) >> 4);
        } else {
            if (i == 0)
                out[x] = (uint8_t)((cur * 4 + 8) >> 4);
            else
                out[x] = (uint8_t)((cur * 3 + in0[i - 1] * 3 + in1[i - 1] + 8) >> 4);
        }
    }
}

static int rd16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

/* Returns 0 on success; -1 malformed; -4 a JPEG this restatement does not cover (progressive, CMYK, ...).
 * out == NULL: only *w, *h, *c are filled. */
EXPORT int orc_jpeg_decode(const uint8_t *f, size_t len, uint8_t *out, int *w_out, int *h_out, int *c_out) {
    uint16_t qt[4][64];
    int qt_present[4] = {0, 0, 0, 0};
    HuffTbl dc[4], ac[4];
    Comp comp[3];
    memset(dc, 0, sizeof dc);
    memset(ac, 0, sizeof ac);
    memset(comp, 0, sizeof comp);
    int W = 0, H = 0, nc = 0, ri = 0, saw_jfif = 0, saw_adobe = 0, adobe_transform = 0, have_sof = 0;
    size_t pos = 2, scan_pos = 0;
    if (len < 4 || f[0] != 0xFF || f[1] != 0xD8) return -1;
    for (;;) {
        while (pos < len && f[pos] != 0xFF) pos++;
        while (pos < len && f[pos] == 0xFF) pos++;
        if (pos >= len) return -1;
        int m = f[pos++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return -1;
        if (pos + 2 > len) return -1;
        int L = rd16(f + pos);
        if (L < 2 || pos + L > len) return -1;
        const uint8_t *s = f + pos + 2;
        int n = L - 2;
        if (m == 0xE0 && n >= 5 && !memcmp(s, "JFIF\0", 5)) saw_jfif = 1;
        if (m == 0xEE && n >= 12 && !memcmp(s, "Adobe", 5)) {
            saw_adobe = 1;
            adobe_transform = s[11];
        }
        if (m == 0xDB) {
            int i = 0;
            while (i < n) {
                int pq = s[i] >> 4, tq = s[i] & 15;
                i++;
                if (tq > 3 || i + (pq ? 128 : 64) > n) return -1;
                for (int k = 0; k < 64; k++) {
                    qt[tq][kZigzag[k]] = (uint16_t)(pq ? rd16(s + i + 2 * k) : s[i + k]);
                }
                qt_present[tq] = 1;
                i += pq ? 128 : 64;
            }
        } else if (m == 0xC4) {
            int i = 0;
            while (i < n) {
                if (i + 17 > n) return -1;
                int tc = s[i] >> 4, th = s[i] & 15, cnt = 0;
                if (tc > 1 || th > 3) return -1;
                HuffTbl *t = tc ? &ac[th] : &dc[th];
                t->bits[0] = 0;
                for (int l = 1; l <= 16; l++) cnt += (t->bits[l] = s[i + l]);
                i += 17;
                if (cnt > 256 || i + cnt > n) return -1;
                memset(t->vals, 0, 256);
                memcpy(t->vals, s + i, cnt);
                i += cnt;
                t->present = 1;
                derive(t);
            }
        } else if (m == 0xDD) {
            if (n < 2) return -1;
            ri = rd16(s);
        } else if (m == 0xC0 || m == 0xC1) {
            if (n < 6 || have_sof) return -1;
            if (s[0] != 8) return -4;
            H = rd16(s + 1);
            W = rd16(s + 3);
            nc = s[5];
            if (W == 0 || H == 0) return -1;
            if (nc != 1 && nc != 3) return -4;
            if (n < 6 + 3 * nc) return -1;
            for (int i = 0; i < nc; i++) {
                comp[i].id = s[6 + 3 * i];
                comp[i].h = s[7 + 3 * i] >> 4;
                comp[i].v = s[7 + 3 * i] & 15;
                comp[i].tq = s[8 + 3 * i];
                if (comp[i].tq > 3) return -1;
            }
            have_sof = 1;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return -4; /* progressive / lossless / arithmetic */
        } else if (m == 0xDA) {
            if (!have_sof || n < 1) return -1;
            int ns = s[0];
            if (ns != nc) return -4; /* non-interleaved multi-scan baseline files */
            if (n < 1 + 2 * ns + 3) return -1;
            for (int i = 0; i < ns; i++) {
                if (s[1 + 2 * i] != comp[i].id) return -4;
                comp[i].td = s[2 + 2 * i] >> 4;
                comp[i].ta = s[2 + 2 * i] & 15;
                if (comp[i].td > 3 || comp[i].ta > 3) return -1;
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) return -4;
            scan_pos = pos + L;
            break;
        }
        pos += L;
    }
    /* jdapimin.c default_decompress_parms: colour space of 3-component files */
    if (nc == 3) {
        int ycc = 1;
        if (saw_jfif)
            ycc = 1;
        else if (saw_adobe)
            ycc = adobe_transform != 0;
        else if (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B')
            ycc = 0;
        if (!ycc) return -4;
        if (comp[1].h != 1 || comp[1].v != 1 || comp[2].h != 1 || comp[2].v != 1) return -4;
        if (!((comp[0].h == 1 && comp[0].v == 1) || (comp[0].h == 2 && comp[0].v == 1) ||
              (comp[0].h == 2 && comp[0].v == 2)))
            return -4;
    }
    *w_out = W;
    *h_out = H;
    *c_out = nc;
    if (!out) return 0;
    int hmax = nc == 1 ? 1 : comp[0].h, vmax = nc == 1 ? 1 : comp[0].v;
    if (nc == 1) comp[0].h = comp[0].v = 1; /* a single-component scan is never interleaved */
    int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
    for (int i = 0; i < nc; i++) {
        Comp *c = &comp[i];
        if (!qt_present[c->tq] || !dc[c->td].present || !ac[c->ta].present) return -1;
        c->wblocks = mcux * c->h;
        c->hblocks = mcuy * c->v;
        c->dw = (W * c->h + hmax - 1) / hmax;
        c->dh = (H * c->v + vmax - 1) / vmax;
        c->coef = (int16_t *)calloc((size_t)c->wblocks * c->hblocks * 64, 2);
        c->plane = (uint8_t *)malloc((size_t)c->wblocks * c->hblocks * 64);
        if (!c->coef || !c->plane) return -1;
    }
    /* ---- jdhuff.c decode_mcu ---- */
    Bits b = {f, len, scan_pos, 0, 0, 0};
    int pred[3] = {0, 0, 0}, togo = ri;
    for (int my = 0; my < mcuy; my++)
        for (int mx = 0; mx < mcux; mx++) {
            if (ri && togo == 0) {
                /* process_restart: drop partial byte, expect RSTn */
                b.nbits = 0;
                if (b.hit_marker || (b.pos + 1 < b.len && b.p[b.pos] == 0xFF && b.p[b.pos + 1] >= 0xD0 &&
                                     b.p[b.pos + 1] <= 0xD7)) {
                    b.pos += 2;
                    b.hit_marker = 0;
                }
                pred[0] = pred[1] = pred[2] = 0;
                togo = ri;
            }
            for (int i = 0; i < nc; i++) {
                Comp *c = &comp[i];
                for (int by = 0; by < c->v; by++)
                    for (int bx = 0; bx < c->h; bx++) {
                        int16_t *blk = c->coef + ((size_t)(my * c->v + by) * c->wblocks + mx * c->h + bx) * 64;
                        int s = decode_sym(&b, &dc[c->td]);
                        if (s) {
                            int r = get_bits(&b, s);
                            s = huff_extend(r, s);
                        }
                        pred[i] += s;
                        blk[0] = (int16_t)pred[i];
                        for (int k = 1; k < 64; k++) {
                            int rs = decode_sym(&b, &ac[c->ta]);
                            int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                k += r;
                                r = get_bits(&b, s);
                                blk[kZigzag[k]] = (int16_t)huff_extend(r, s); /* k <= 78: 16 spare entries */
                            } else {
                                if (r != 15) break;
                                k += 15;
                            }
                        }
                    }
            }
            togo--;
        }
    /* ---- jidctint.c ---- */
    for (int i = 0; i < nc; i++) {
        Comp *c = &comp[i];
        int pitch = c->wblocks * 8;
        for (int by = 0; by < c->hblocks; by++)
            for (int bx = 0; bx < c->wblocks; bx++)
                idct_islow(c->coef + ((size_t)by * c->wblocks + bx) * 64, qt[c->tq],
                           c->plane + (size_t)by * 8 * pitch + bx * 8, pitch);
    }
    /* ---- jdsample.c + jdcolor.c ---- */
    if (nc == 1) {
        for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, comp[0].plane + (size_t)y * comp[0].wblocks * 8, W);
    } else {
        uint8_t *cb = (uint8_t *)malloc(W + 16), *cr = (uint8_t *)malloc(W + 16);
        int hs = comp[0].h, vs = comp[0].v;
        for (int y = 0; y < H; y++) {
            const uint8_t *yy = comp[0].plane + (size_t)y * comp[0].wblocks * 8;
            upsample_row(&comp[1], hs, vs, y, W, cb);
            upsample_row(&comp[2], hs, vs, y, W, cr);
            uint8_t *o = out + (size_t)y * W * 3;
            for (int x = 0; x < W; x++) {
                int Y = yy[x], u = cb[x] - 128, v = cr[x] - 128;
                /* jdcolor.c build_ycc_rgb_table: FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 32768 */
                int r = Y + ((91881 * v + 32768) >> 16);
                int g = Y + ((-22554 * u + 32768 + (-46802) * v) >> 16);
                int bl = Y + ((116130 * u + 32768) >> 16);
                o[3 * x] = clamp_u8(r);
                o[3 * x + 1] = clamp_u8(g);
                o[3 * x + 2] = clamp_u8(bl);
            }
        }
        free(cb);
        free(cr);
    }
    for (int i = 0; i < nc; i++) {
        free(comp[i].coef);
        free(comp[i].plane);
    }
    return 0;
}
