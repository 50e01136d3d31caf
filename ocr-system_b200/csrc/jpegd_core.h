// jpegd_core.h -- baseline-JPEG decode building blocks shared by the CUDA kernels (k_jpegd.cu) and the host
// unit-test build (tests/jpegd_host.cpp).  No CUDA runtime calls in here.
//
// Reference boundary: backend/utils/image_preprocessing.py:57-75 (`load_image` / `load_image_bytes`:
// Image.open(...) [+ convert]) and backend/services/ocr_service.py:494-496,716-718 (files / bytes are what the
// engine is handed).  The decoder behind Pillow is libjpeg-turbo; the arithmetic restated here is its
// jdhuff.c / jidctint.c (islow) / jdsample.c (fancy upsampling) / jdcolor.c pipeline, so the raster is the
// same bytes `np.asarray(Image.open(f))` holds.
//
// Parallel entropy decoding.  A Huffman stream has no random access, but it re-synchronises by itself: a
// decoder started at a wrong bit position usually falls into step with the true symbol sequence after a few
// dozen symbols.  The scan (0xFF00 stuffing and RSTn markers removed, kept as big-endian 32-bit words) is cut
// into sub-sequences of `sub_bits` bits; thread i decodes sub-sequence i from a guessed state, and hands its
// exit state (bit position, block slot inside the MCU, zig-zag index) to thread i+1, which re-decodes from it
// until no exit state changes any more.  Then an exclusive scan of the per-sub-sequence block counts gives every
// thread the index of its first block and a last pass writes the coefficients.  DC predictions are a prefix sum
// over the written differences.  A restart marker is a hard synchronisation point (byte aligned, state reset,
// absolute MCU number known).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define JD_HD __host__ __device__ __forceinline__
#else
#define JD_HD inline
#endif

#define JD_LUT_BITS 10
#define JD_MAX_TABLES 6 /* distinct (class, id) tables a 3-component scan can select */
#define JD_MAX_SLOTS 6  /* blocks per MCU: 4:2:0 -> Y Y Y Y Cb Cr */

#define JD_SUB_BITS 6   /* second-level tables resolve code bits 11..16 */
#ifndef JD_MAX_SUB
#define JD_MAX_SUB 16   /* 10-bit prefixes with longer codes a table may have before it falls back to the ladder */
#endif

struct JdHuff {
    /* first level, indexed by the next 10 bits: (length << 8) | symbol for codes of <= 10 bits; 0x8000 | n for a
     * prefix of longer codes (n = second-level table, 0xFF = walk the maxcode ladder); 0 = no such code */
    uint16_t lut[1 << JD_LUT_BITS];
    uint16_t lut2[JD_MAX_SUB << JD_SUB_BITS]; /* indexed by n * 64 + the following 6 bits: (length << 8) | symbol, or 0 */
    int32_t maxcode[18];                      /* jdhuff.c jpeg_make_d_derived_tbl: largest code of length l, -1 if none */
    int32_t valoff[17];                       /* vals index = code + valoff[l] */
    uint8_t vals[256];
};

struct JdPageHdr {
    uint32_t scan_off; /* first entropy-coded byte, relative to the start of the batch blob */
    uint32_t scan_len; /* bytes from scan_off to the end of the file (the device finds the terminating marker) */
    uint32_t restart_interval; /* MCUs; 0 = none */
    uint32_t stream_word_off;  /* where this page's unstuffed stream starts in the workspace (32-bit words) */
    uint32_t cta_off;          /* first CTA (chunk of sub-sequences) of this page in the entropy kernel's grid */
    uint32_t n_cta;            /* chunks reserved (from the stuffed length: an upper bound) */
    uint32_t rst_off;          /* first restart-position slot of this page */
    uint32_t n_rst_max;
    uint16_t mcux, mcuy;
    uint8_t ncomp, hs, vs, bpm, ntab, pad0, pad1, pad2;
    uint8_t slot_comp[8], slot_dc[8], slot_ac[8]; /* per block slot of an MCU: component, DC table, AC table */
    /* DC differences / predictions are kept per component: block (mcu m, slot s) -> slot_dcbase[s] + m * slot_cnt[s] */
    uint8_t slot_cnt[8];
    uint32_t slot_dcbase[8];
    uint32_t dc_off[4];   /* first element of component c (multiples of 8), dc_off[ncomp] = elements per page */
    uint32_t tile_off;    /* first 16 KB tile of this page in the unstuff kernels' grid */
    uint32_t n_tiles;
};

struct JdPage {
    JdPageHdr h;
    uint16_t qt[3][64]; /* natural (row-major) order, per component */
    JdHuff tab[JD_MAX_TABLES];
};

struct JdState {
    uint32_t p;  /* bit position in the unstuffed stream */
    uint32_t sk; /* slot | k << 8   (k = 0: the next symbol is a DC code) */
};

static const uint8_t kJdZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                      30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ------------------------------------------------------------------------------------------------------------
// Bit window over the big-endian word stream: peek32(p) = the 32 bits that start at bit p.
// Word i of a page's stream.  (A layout transposed inside chunks of 256 sub-sequences -- neighbouring threads reading
// neighbouring words -- was measured slower on B200: the 4-byte scatter writes of the unstuff pass leave partially
// written L2 sectors and the decoder's lanes drift apart by several words; the linear layout keeps each thread on its
// own 32-byte sectors, which it re-uses 8 times.)
#define JD_CHUNK_LOG2 8
JD_HD uint32_t jd_word_index(uint32_t i, uint32_t swl) {
    (void)swl;
    return i;
}

struct JdBits {
    const uint32_t *w;
    uint32_t swl;
    uint32_t wi, w0, w1, w2; /* words wi, wi+1 and (loaded one refill ahead of its use) wi+2 */
    JD_HD uint32_t ld(uint32_t i) const { return w[jd_word_index(i, swl)]; }
    JD_HD void init(const uint32_t *words, uint32_t sub_words_log2, uint32_t p) {
        w = words;
        swl = sub_words_log2;
        wi = p >> 5;
        w0 = ld(wi);
        w1 = ld(wi + 1);
        w2 = ld(wi + 2);
    }
    JD_HD uint32_t peek32(uint32_t p) {
        uint32_t i = p >> 5;
        if (i != wi) {
            if (i == wi + 1) {
                w0 = w1;
                w1 = w2;
            } else {
                w0 = ld(i);
                w1 = ld(i + 1);
            }
            w2 = ld(i + 2);
            wi = i;
        }
        uint32_t s = p & 31;
#ifdef __CUDA_ARCH__
        return __funnelshift_l(w1, w0, s);
#else
        return s ? (w0 << s) | (w1 >> (32 - s)) : w0;
#endif
    }
};

// One Huffman symbol from the 32-bit window v: returns (length << 8) | symbol.  An impossible prefix (only
// reachable while a thread is still out of step, or in a corrupt file) decodes as symbol 0 of length 16, as
// libjpeg's "corrupt data" path does.
JD_HD uint32_t jd_symbol(const JdHuff &t, uint32_t v) {
    uint32_t e = t.lut[v >> (32 - JD_LUT_BITS)];
    if (e & 0x8000u) {
        const uint32_t n = e & 0xFFu;
        if (n != 0xFFu) {
            e = t.lut2[(n << JD_SUB_BITS) | ((v >> (32 - JD_LUT_BITS - JD_SUB_BITS)) & ((1u << JD_SUB_BITS) - 1u))];
        } else {
            e = 0;
            for (int l = JD_LUT_BITS + 1; l <= 16; l++) {
                int32_t code = (int32_t)(v >> (32 - l));
                if (code <= t.maxcode[l]) {
                    e = ((uint32_t)l << 8) | t.vals[(code + t.valoff[l]) & 255];
                    break;
                }
            }
        }
    }
    return e ? e : (16u << 8);
}

JD_HD int jd_extend(uint32_t r, int s) { return (int)r < (1 << (s - 1)) ? (int)r - (1 << s) + 1 : (int)r; }

// What one thread learns about its sub-sequence.
struct JdSubResult {
    JdState exit;     /* first symbol boundary at or behind the end of the sub-sequence */
    int32_t nblocks;  /* blocks completed since the entry (or since the last restart marker crossed) */
    int32_t abs_base; /* block index right behind the last restart marker crossed, -1 if none */
};

// Decodes the symbols that start in [entry.p, end_bit).  `rst` = restart boundaries of this page (bit positions
// in the unstuffed stream, ascending, n_rst of them); total_bits = length of the stream.  With WRITE, block
// `blk0 + completed` receives its AC coefficients in natural order (coef, 64 int16 per block, pre-zeroed) and its
// DC *difference* (dcdiff, one int16 per block, laid out per component: see JdPageHdr); blocks >= nblk_total are dropped.
template <bool WRITE>
JD_HD JdSubResult jd_decode_sub(const JdPageHdr &pg, const JdHuff *tabs, const uint8_t *zz, const uint32_t *words,
                                uint32_t sub_words_log2, uint32_t total_bits, const uint32_t *rst, int n_rst, JdState entry, uint32_t end_bit,
                                int32_t blk0, int32_t nblk_total, int16_t *coef, int16_t *dcdiff) {
    JdSubResult res;
    uint32_t p = entry.p, slot = entry.sk & 255, k = entry.sk >> 8;
    int32_t nb = 0, abs_base = -1;
    if (end_bit > total_bits) end_bit = total_bits;
    if (p >= total_bits) {
        res.exit = entry;
        res.nblocks = 0;
        res.abs_base = -1;
        return res;
    }
    // first restart boundary strictly behind p (a state that sits on a boundary is already past the marker)
    int ri = 0;
    {
        int lo = 0, hi = n_rst;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (rst[mid] <= p) lo = mid + 1; else hi = mid;
        }
        ri = lo;
    }
    uint32_t limit = ri < n_rst ? rst[ri] : total_bits;
    const uint32_t bpm = pg.bpm;
    const int32_t blocks_per_interval = (int32_t)(pg.restart_interval * bpm);
    JdBits bits;
    bits.init(words, sub_words_log2, p);
    int32_t blk = blk0;
    /* (dc table | ac table << 4) of every block slot in one register: no memory access at a block end */
    unsigned long long slotmap = 0;
    for (uint32_t i = 0; i < bpm; i++) slotmap |= (unsigned long long)(pg.slot_dc[i] | (pg.slot_ac[i] << 4)) << (8 * i);
    uint32_t tsel = (uint32_t)(slotmap >> (8 * slot));
    const JdHuff *t_dc = tabs + (tsel & 15u), *t_ac = tabs + ((tsel >> 4) & 15u);
    while (p < end_bit) {
        const uint32_t v = bits.peek32(p);
        const bool is_dc = k == 0;
        const uint32_t e = jd_symbol(is_dc ? *t_dc : *t_ac, v);
        const uint32_t len = e >> 8, sym = e & 255, s = sym & 15, r = is_dc ? 0 : sym >> 4;
        const uint32_t tot = len + s;
        if (p + tot > limit) {
            if (limit >= total_bits) { /* ran into the end of the scan */
                p = total_bits;
                break;
            }
            /* restart marker: discard the padding bits, reset, the MCU number is known again */
            p = limit;
            ri++;
            limit = ri < n_rst ? rst[ri] : total_bits;
            slot = 0;
            k = 0;
            abs_base = ri * blocks_per_interval;
            nb = 0;
            blk = abs_base;
            tsel = (uint32_t)slotmap;
            t_dc = tabs + (tsel & 15u);
            t_ac = tabs + ((tsel >> 4) & 15u);
            continue;
        }
        p += tot;
        if (s | (uint32_t)is_dc) {
            const int val = s ? jd_extend((v << len) >> (32 - s), (int)s) : 0;
            k += r;
            if (WRITE && blk < nblk_total) {
                if (is_dc) {
                    /* index from the block number alone, so that a corrupt stream (block number and slot out of
                     * step) can never write outside the page's DC array */
                    const uint32_t m = (uint32_t)blk / bpm, sl = (uint32_t)blk - m * bpm;
                    dcdiff[pg.slot_dcbase[sl] + m * pg.slot_cnt[sl]] = (int16_t)val;
                }
                else coef[(size_t)blk * 64 + (k < 64 ? zz[k] : 63)] = (int16_t)val;
            }
            k++;
        } else {
            k = r == 15 ? k + 16 : 64;
        }
        if (k >= 64) {
            k = 0;
            nb++;
            blk++;
            if (++slot == bpm) slot = 0;
            tsel = (uint32_t)(slotmap >> (8 * slot));
            t_dc = tabs + (tsel & 15u);
            t_ac = tabs + ((tsel >> 4) & 15u);
        }
    }
    res.exit.p = p;
    res.exit.sk = slot | (k << 8);
    res.nblocks = nb;
    res.abs_base = abs_base;
    return res;
}

// ------------------------------------------------------------------------------------------------------------
// jidctint.c jpeg_idct_islow on one block: in = 64 coefficients (natural order, in[0] = DC), q = quantiser.
// out[r * pitch + c] = range-limited sample.  (SIMD builds multiply coefficient and quantiser in 16 bits.)
#define JD_FIX_0_298631336 2446
#define JD_FIX_0_390180644 3196
#define JD_FIX_0_541196100 4433
#define JD_FIX_0_765366865 6270
#define JD_FIX_0_899976223 7373
#define JD_FIX_1_175875602 9633
#define JD_FIX_1_501321110 12299
#define JD_FIX_1_847759065 15137
#define JD_FIX_1_961570560 16069
#define JD_FIX_2_053119869 16819
#define JD_FIX_2_562915447 20995
#define JD_FIX_3_072711026 25172

JD_HD void jd_idct_1d(int32_t i0, int32_t i1, int32_t i2, int32_t i3, int32_t i4, int32_t i5, int32_t i6, int32_t i7,
                      int shift, int32_t *o) {
    int32_t z2 = i2, z3 = i6;
    int32_t z1 = (z2 + z3) * JD_FIX_0_541196100;
    int32_t tmp2 = z1 + z3 * (-JD_FIX_1_847759065);
    int32_t tmp3 = z1 + z2 * JD_FIX_0_765366865;
    int32_t tmp0 = (i0 + i4) * 8192, tmp1 = (i0 - i4) * 8192;
    int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = i7;
    tmp1 = i5;
    tmp2 = i3;
    tmp3 = i1;
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3;
    int32_t z5 = (z3 + z4) * JD_FIX_1_175875602;
    tmp0 *= JD_FIX_0_298631336;
    tmp1 *= JD_FIX_2_053119869;
    tmp2 *= JD_FIX_3_072711026;
    tmp3 *= JD_FIX_1_501321110;
    z1 *= -JD_FIX_0_899976223;
    z2 *= -JD_FIX_2_562915447;
    z3 *= -JD_FIX_1_961570560;
    z4 *= -JD_FIX_0_390180644;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    const int32_t rnd = 1 << (shift - 1);
    o[0] = (tmp10 + tmp3 + rnd) >> shift;
    o[7] = (tmp10 - tmp3 + rnd) >> shift;
    o[1] = (tmp11 + tmp2 + rnd) >> shift;
    o[6] = (tmp11 - tmp2 + rnd) >> shift;
    o[2] = (tmp12 + tmp1 + rnd) >> shift;
    o[5] = (tmp12 - tmp1 + rnd) >> shift;
    o[3] = (tmp13 + tmp0 + rnd) >> shift;
    o[4] = (tmp13 - tmp0 + rnd) >> shift;
}

JD_HD uint32_t jd_clamp_u8(int v) { return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

// jdcolor.c ycc_rgb_convert: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802,
// FIX(0.34414) = 22554, ONE_HALF = 32768 (the Cb-to-G table carries the rounding constant).
JD_HD void jd_ycc_to_rgb(int y, int cb, int cr, uint32_t &r, uint32_t &g, uint32_t &b) {
    int u = cb - 128, v = cr - 128;
    r = jd_clamp_u8(y + ((91881 * v + 32768) >> 16));
    g = jd_clamp_u8(y + ((-22554 * u + 32768 - 46802 * v) >> 16));
    b = jd_clamp_u8(y + ((116130 * u + 32768) >> 16));
}

// jdsample.c: full-resolution chroma sample x of output row y from a (dw x dh real samples) plane.
// mode 0 = 1x1, 1 = h2v1, 2 = h2v2; fancy (triangle) filters unless dw <= 2 (jinit_upsampler).
JD_HD int jd_upsample_at(const uint8_t *plane, int pitch, int dw, int dh, int mode, int x, int y) {
    if (mode == 0) return plane[(size_t)y * pitch + x];
    int i = x >> 1;
    if (mode == 1) {
        const uint8_t *in = plane + (size_t)y * pitch;
        int v = in[i];
        if (dw <= 2) return v;
        if (x & 1) return i == dw - 1 ? v : (v * 3 + in[i + 1] + 2) >> 2;
        return i == 0 ? v : (v * 3 + in[i - 1] + 1) >> 2;
    }
    int r0 = y >> 1;
    if (r0 > dh - 1) r0 = dh - 1;
    const uint8_t *in0 = plane + (size_t)r0 * pitch;
    if (dw <= 2) return in0[i];
    int r1 = (y & 1) ? (y >> 1) + 1 : (y >> 1) - 1;
    if (r1 < 0) r1 = 0;
    if (r1 > dh - 1) r1 = dh - 1;
    const uint8_t *in1 = plane + (size_t)r1 * pitch;
    int cur = in0[i] * 3 + in1[i];
    if (x & 1) return i == dw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + in0[i + 1] * 3 + in1[i + 1] + 7) >> 4;
    return i == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + in0[i - 1] * 3 + in1[i - 1] + 8) >> 4;
}

// ------------------------------------------------------------------------------------------------------------
// Host side: marker parsing and table construction (jdmarker.c / jdhuff.c jpeg_make_d_derived_tbl).
struct JdInfo {
    int width, height, ncomp, hs, vs; /* hs, vs: luma sampling factors (1 or 2) */
};

static inline int jd_rd16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

static inline void jd_build_huff(const uint8_t *bits /*[17]*/, const uint8_t *vals, JdHuff *t) {
    memset(t, 0, sizeof *t);
    memcpy(t->vals, vals, 256);
    int code = 0, k = 0, nsub = 0;
    for (int l = 1; l <= 16; l++) {
        if (bits[l]) {
            t->valoff[l] = k - code;
            for (int i = 0; i < bits[l]; i++, k++, code++) {
                if (code >= (1 << l)) continue; /* over-subscribed table: jdhuff.c rejects it; never matched here */
                const uint16_t ent = (uint16_t)((l << 8) | vals[k & 255]);
                if (l <= JD_LUT_BITS) {
                    int first = code << (JD_LUT_BITS - l), cnt = 1 << (JD_LUT_BITS - l);
                    for (int j = 0; j < cnt; j++) t->lut[first + j] = ent;
                } else {
                    int prefix = code >> (l - JD_LUT_BITS);
                    uint16_t &slot = t->lut[prefix];
                    if (slot == 0) slot = (uint16_t)(0x8000 | (nsub < JD_MAX_SUB ? nsub++ : 0xFF));
                    if (!(slot & 0x8000)) continue; /* prefix already taken by a shorter code: malformed */
                    int n = slot & 0xFF;
                    if (n == 0xFF) continue;
                    int rest = l - JD_LUT_BITS; /* 1..6 bits behind the prefix */
                    int first = (code & ((1 << rest) - 1)) << (JD_SUB_BITS - rest), cnt = 1 << (JD_SUB_BITS - rest);
                    for (int j = 0; j < cnt; j++) t->lut2[(n << JD_SUB_BITS) + first + j] = ent;
                }
            }
            t->maxcode[l] = code - 1;
        } else
            t->maxcode[l] = -1;
        code <<= 1;
    }
    t->maxcode[17] = 0xFFFFF;
}

// Parses the headers of one file.  0 = ok (info + page filled except the workspace offsets and scan_off being
// relative to `f`); -1 = malformed; -4 = a JPEG outside the device decoder's subset (progressive, arithmetic,
// 12-bit, CMYK / RGB colour spaces, 4:4:0 and other sampling grids, multi-scan baseline files): the caller
// decodes such a file with the host codec, as the reference does for every file.
static inline int jd_parse(const uint8_t *f, size_t len, JdInfo *info, JdPage *pg) {
    uint8_t hbits[2][4][17], hvals[2][4][256];
    int hpresent[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    uint16_t qt[4][64];
    int qpresent[4] = {0, 0, 0, 0};
    int cid[3] = {0, 0, 0}, ch[3] = {1, 1, 1}, cv[3] = {1, 1, 1}, ctq[3] = {0, 0, 0}, ctd[3] = {0, 0, 0}, cta[3] = {0, 0, 0};
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
        int L = jd_rd16(f + pos);
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
                for (int k = 0; k < 64; k++) qt[tq][kJdZigzag[k]] = (uint16_t)(pq ? jd_rd16(s + i + 2 * k) : s[i + k]);
                qpresent[tq] = 1;
                i += pq ? 128 : 64;
            }
        } else if (m == 0xC4) {
            int i = 0;
            while (i < n) {
                if (i + 17 > n) return -1;
                int tc = s[i] >> 4, th = s[i] & 15, cnt = 0;
                if (tc > 1 || th > 3) return -1;
                hbits[tc][th][0] = 0;
                for (int l = 1; l <= 16; l++) cnt += (hbits[tc][th][l] = s[i + l]);
                i += 17;
                if (cnt > 256 || i + cnt > n) return -1;
                memset(hvals[tc][th], 0, 256);
                memcpy(hvals[tc][th], s + i, cnt);
                if (tc == 0)
                    for (int j = 0; j < cnt; j++)
                        if (s[i + j] > 15) return -1; /* jdhuff.c: JERR_BAD_HUFF_TABLE */
                i += cnt;
                hpresent[tc][th] = 1;
            }
        } else if (m == 0xDD) {
            if (n < 2) return -1;
            ri = jd_rd16(s);
        } else if (m == 0xC0 || m == 0xC1) {
            if (n < 6 || have_sof) return -1;
            if (s[0] != 8) return -4;
            H = jd_rd16(s + 1);
            W = jd_rd16(s + 3);
            nc = s[5];
            if (W == 0 || H == 0) return -1;
            if (nc != 1 && nc != 3) return -4;
            if (n < 6 + 3 * nc) return -1;
            for (int i = 0; i < nc; i++) {
                cid[i] = s[6 + 3 * i];
                ch[i] = s[7 + 3 * i] >> 4;
                cv[i] = s[7 + 3 * i] & 15;
                ctq[i] = s[8 + 3 * i];
                if (ctq[i] > 3) return -1;
            }
            have_sof = 1;
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return -4;
        } else if (m == 0xDA) {
            if (!have_sof || n < 1) return -1;
            int ns = s[0];
            if (ns != nc) return -4;
            if (n < 1 + 2 * ns + 3) return -1;
            for (int i = 0; i < ns; i++) {
                if (s[1 + 2 * i] != cid[i]) return -4;
                ctd[i] = s[2 + 2 * i] >> 4;
                cta[i] = s[2 + 2 * i] & 15;
                if (ctd[i] > 3 || cta[i] > 3) return -1;
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0) return -4;
            scan_pos = pos + L;
            break;
        }
        pos += L;
    }
    if (nc == 3) {
        int ycc = 1; /* jdapimin.c default_decompress_parms */
        if (saw_jfif)
            ycc = 1;
        else if (saw_adobe)
            ycc = adobe_transform != 0;
        else if (cid[0] == 'R' && cid[1] == 'G' && cid[2] == 'B')
            ycc = 0;
        if (!ycc) return -4;
        if (ch[1] != 1 || cv[1] != 1 || ch[2] != 1 || cv[2] != 1) return -4;
        if (!((ch[0] == 1 && cv[0] == 1) || (ch[0] == 2 && cv[0] == 1) || (ch[0] == 2 && cv[0] == 2))) return -4;
    } else {
        ch[0] = cv[0] = 1; /* a single-component scan is never interleaved */
    }
    if (scan_pos >= len || (uint64_t)(len - scan_pos) >= (1ull << 28)) return -1;
    info->width = W;
    info->height = H;
    info->ncomp = nc;
    info->hs = ch[0];
    info->vs = cv[0];
    memset(pg, 0, offsetof(JdPage, tab));
    JdPageHdr *ph = &pg->h;
    ph->scan_off = (uint32_t)scan_pos;
    ph->scan_len = (uint32_t)(len - scan_pos);
    ph->restart_interval = (uint32_t)ri;
    ph->mcux = (uint16_t)((W + 8 * ch[0] - 1) / (8 * ch[0]));
    ph->mcuy = (uint16_t)((H + 8 * cv[0] - 1) / (8 * cv[0]));
    ph->ncomp = (uint8_t)nc;
    ph->hs = (uint8_t)ch[0];
    ph->vs = (uint8_t)cv[0];
    /* de-duplicate the (class, id) tables the scan selects */
    int map[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}}, ntab = 0, slot = 0;
    for (int i = 0; i < nc; i++) {
        if (!qpresent[ctq[i]] || !hpresent[0][ctd[i]] || !hpresent[1][cta[i]]) return -1;
        memcpy(pg->qt[i], qt[ctq[i]], sizeof qt[0]);
        for (int cls = 0; cls < 2; cls++) {
            int id = cls ? cta[i] : ctd[i];
            if (map[cls][id] < 0) {
                if (ntab >= JD_MAX_TABLES) return -4;
                jd_build_huff(hbits[cls][id], hvals[cls][id], &pg->tab[ntab]);
                map[cls][id] = ntab++;
            }
        }
        int nblk = (i == 0) ? ch[0] * cv[0] : 1;
        for (int b = 0; b < nblk; b++, slot++) {
            ph->slot_comp[slot] = (uint8_t)i;
            ph->slot_dc[slot] = (uint8_t)map[0][ctd[i]];
            ph->slot_ac[slot] = (uint8_t)map[1][cta[i]];
        }
    }
    ph->bpm = (uint8_t)slot;
    ph->ntab = (uint8_t)ntab;
    {
        const uint32_t nmcu = (uint32_t)ph->mcux * ph->mcuy;
        uint32_t off = 0;
        for (int i = 0, sl = 0; i < nc; i++) {
            const int cnt = (i == 0) ? ch[0] * cv[0] : 1;
            ph->dc_off[i] = off;
            for (int b = 0; b < cnt; b++, sl++) {
                ph->slot_cnt[sl] = (uint8_t)cnt;
                ph->slot_dcbase[sl] = off + (uint32_t)b;
            }
            off += (nmcu * (uint32_t)cnt + 7u) & ~7u;
        }
        ph->dc_off[nc] = off;
    }
    return 0;
}
