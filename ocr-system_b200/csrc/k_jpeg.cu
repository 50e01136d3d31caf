// k_jpeg.cu -- baseline JPEG encoder for batches of RGB pages, byte-compatible with what the reference gets
// from Pillow / libjpeg-turbo.
//
// Reference: backend/utils/image_preprocessing.py:496-557 (compress_for_azure: image.save(format='JPEG',
// quality=q, optimize=True) in a quality loop) and :331-347 (image_to_bytes, optimize off).  The codec is the
// third-party libjpeg(-turbo) behind Pillow's JpegEncode.c; every arithmetic step below restates its
// integer pipeline so that the file is the same byte stream:
//   jccolor.c   rgb_ycc_convert      16-bit fixed-point RGB -> YCbCr
//   jcsample.c  h2v2_downsample      2x2 chroma average with the alternating 1,2 bias, edges replicated
//   jfdctint.c  jpeg_fdct_islow      13-bit integer DCT, two passes, output scaled by 8
//   jcdctmgr.c  quantize             round-half-up division by 8*q
//   jcparam.c   jpeg_set_quality     Annex K tables scaled by quality (baseline: clamp to 1..255)
//   jccoefct.c  compress_data        dummy blocks at the right / bottom edge (AC 0, DC of the previous block)
//   jchuff.c    encode_one_block, jpeg_gen_optimal_table, jpeg_make_c_derived_tbl, byte stuffing, 1-padding
//   jcmarker.c  SOI APP0(JFIF 1.1) DQT DQT SOF0 DHTx4 SOS ... EOI
// What is parallel: blocks (DCT, symbol statistics, code lengths, emission at prefix-summed bit offsets),
// pages (Huffman table construction, scans, byte stuffing).
#include <string.h>

#include <vector>

#include "common.cuh"

namespace lumina {

// ---- constant tables --------------------------------------------------------------------------
__constant__ uint8_t c_zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const uint8_t h_zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
static const uint8_t std_qlum[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const uint8_t std_qchr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                     99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 Huffman tables: bits[1..16] then the symbols
static const uint8_t std_dc_lum_bits[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t std_dc_chr_bits[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t std_dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t std_ac_lum_bits[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t std_ac_lum_vals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t std_ac_chr_bits[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t std_ac_chr_vals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct JpegGeom {
    int h, w, mcu_w, mcu_h, ybw, ybh, blocks;  // ybw/ybh: real luma blocks per row / column
};
static JpegGeom jpeg_geom(int h, int w) {
    JpegGeom g;
    g.h = h; g.w = w;
    g.mcu_w = (w + 15) / 16; g.mcu_h = (h + 15) / 16;
    g.ybw = (w + 7) / 8; g.ybh = (h + 7) / 8;
    g.blocks = g.mcu_w * g.mcu_h * 6;
    return g;
}

// one Huffman table in encoder form (jchuff.c c_derived_tbl) plus the DHT payload
struct JpegHuff {
    uint16_t code[256];
    uint8_t size[256];
    uint8_t bits[17];
    uint8_t vals[256];
    int32_t nvals;
};
constexpr int JT_DC0 = 0, JT_AC0 = 1, JT_DC1 = 2, JT_AC1 = 3;  // table slots per page

struct JpegQuant {
    uint16_t q8[2][64];  // 8 * quantiser, natural order, [0] luma [1] chroma
};

// ---- colour conversion (jccolor.c) + sampling --------------------------------------------------
#define JFIX(x) ((int)((x) * 65536.0 + 0.5))
__device__ __forceinline__ int jpg_y(int r, int g, int b) {
    return (JFIX(0.29900) * r + JFIX(0.58700) * g + JFIX(0.11400) * b + 32768) >> 16;
}
__device__ __forceinline__ int jpg_cb(int r, int g, int b) {
    return (-JFIX(0.16874) * r - JFIX(0.33126) * g + JFIX(0.50000) * b + (128 << 16) + 32768 - 1) >> 16;
}
__device__ __forceinline__ int jpg_cr(int r, int g, int b) {
    return (JFIX(0.50000) * r - JFIX(0.41869) * g - JFIX(0.08131) * b + (128 << 16) + 32768 - 1) >> 16;
}

// jfdctint.c, one 8-point pass.  FIRST: out0/out4 shifted up by PASS1_BITS, others descaled by CONST-PASS1;
// second pass: out0/out4 descaled by PASS1_BITS, others by CONST+PASS1.
template <bool FIRST>
__device__ __forceinline__ void jpg_fdct8(int d[8]) {
    constexpr int CB = 13, P1 = 2;
    const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
    const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    constexpr int SH = FIRST ? CB - P1 : CB + P1;
    constexpr int RND = 1 << (SH - 1);
    if (FIRST) {
        d[0] = (tmp10 + tmp11) << P1;
        d[4] = (tmp10 - tmp11) << P1;
    } else {
        d[0] = (tmp10 + tmp11 + (1 << (P1 - 1))) >> P1;
        d[4] = (tmp10 - tmp11 + (1 << (P1 - 1))) >> P1;
    }
    int z1 = (tmp12 + tmp13) * 4433;
    d[2] = (z1 + tmp13 * 6270 + RND) >> SH;
    d[6] = (z1 + tmp12 * (-15137) + RND) >> SH;
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * 9633;
    const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    d[7] = (t4 + z1 + z3 + RND) >> SH;
    d[5] = (t5 + z2 + z4 + RND) >> SH;
    d[3] = (t6 + z2 + z3 + RND) >> SH;
    d[1] = (t7 + z1 + z4 + RND) >> SH;
}

// block b of a page -> (component class, dummy?, position)
struct JpegBlk {
    int comp;    // 0 Y, 1 Cb, 2 Cr
    int bx, by;  // block coordinates inside the component
    bool dummy;
};
__device__ __forceinline__ JpegBlk jpg_block(int b, const JpegGeom &g) {
    JpegBlk k;
    const int mcu = b / 6, s = b - mcu * 6;
    const int my = mcu / g.mcu_w, mx = mcu - my * g.mcu_w;
    if (s < 4) {
        k.comp = 0; k.bx = mx * 2 + (s & 1); k.by = my * 2 + (s >> 1);
        k.dummy = k.bx >= g.ybw || k.by >= g.ybh;
    } else {
        k.comp = s - 3; k.bx = mx; k.by = my; k.dummy = false;
    }
    return k;
}

// ---- K1: colour conversion + sampling + forward DCT, one thread per block --------------------------
__global__ void __launch_bounds__(128) jpeg_fdct_kernel(const uint8_t *__restrict__ rgb, int16_t *__restrict__ coef, JpegGeom g) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.blocks) return;
    const int page = blockIdx.y;
    const uint8_t *src = rgb + (size_t)page * g.h * g.w * 3;
    int16_t *out = coef + ((size_t)page * g.blocks + b) * 64;
    const JpegBlk k = jpg_block(b, g);
    if (k.dummy) return;  // never read: later passes derive dummy blocks from the geometry
    int ws[64];
    if (k.comp == 0) {
#pragma unroll 1
        for (int y = 0; y < 8; y++) {
            const int py = min(k.by * 8 + y, g.h - 1);
            const uint8_t *row = src + (size_t)py * g.w * 3;
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const int px = min(k.bx * 8 + x, g.w - 1);
                ws[y * 8 + x] = jpg_y(row[px * 3], row[px * 3 + 1], row[px * 3 + 2]) - 128;
            }
        }
    } else {
#pragma unroll 1
        for (int y = 0; y < 8; y++) {
            // bottom edge (jcprepct.c): the colour buffer is padded to a 2-row group only, the rest of the iMCU
            // replicates the last DOWNSAMPLED row; the right edge is padded in full resolution (jcsample.c)
            const int cy = min(k.by * 8 + y, (g.h + 1) / 2 - 1);
            const int py0 = min(cy * 2, g.h - 1), py1 = min(cy * 2 + 1, g.h - 1);
            const uint8_t *r0 = src + (size_t)py0 * g.w * 3, *r1 = src + (size_t)py1 * g.w * 3;
#pragma unroll
            for (int x = 0; x < 8; x++) {
                const int cx = k.bx * 8 + x;
                const int px0 = min(cx * 2, g.w - 1), px1 = min(cx * 2 + 1, g.w - 1);
                int s;
                if (k.comp == 1)
                    s = jpg_cb(r0[px0 * 3], r0[px0 * 3 + 1], r0[px0 * 3 + 2]) + jpg_cb(r0[px1 * 3], r0[px1 * 3 + 1], r0[px1 * 3 + 2]) +
                        jpg_cb(r1[px0 * 3], r1[px0 * 3 + 1], r1[px0 * 3 + 2]) + jpg_cb(r1[px1 * 3], r1[px1 * 3 + 1], r1[px1 * 3 + 2]);
                else
                    s = jpg_cr(r0[px0 * 3], r0[px0 * 3 + 1], r0[px0 * 3 + 2]) + jpg_cr(r0[px1 * 3], r0[px1 * 3 + 1], r0[px1 * 3 + 2]) +
                        jpg_cr(r1[px0 * 3], r1[px0 * 3 + 1], r1[px0 * 3 + 2]) + jpg_cr(r1[px1 * 3], r1[px1 * 3 + 1], r1[px1 * 3 + 2]);
                ws[y * 8 + x] = ((s + ((cx & 1) ? 2 : 1)) >> 2) - 128;   // h2v2_downsample: bias 1,2,1,2,...
            }
        }
    }
#pragma unroll
    for (int y = 0; y < 8; y++) jpg_fdct8<true>(ws + y * 8);
#pragma unroll
    for (int x = 0; x < 8; x++) {
        int d[8];
#pragma unroll
        for (int y = 0; y < 8; y++) d[y] = ws[y * 8 + x];
        jpg_fdct8<false>(d);
#pragma unroll
        for (int y = 0; y < 8; y++) ws[y * 8 + x] = d[y];
    }
#pragma unroll
    for (int i = 0; i < 64; i += 2)
        *reinterpret_cast<uint32_t *>(out + i) = (uint32_t)(uint16_t)ws[i] | ((uint32_t)(uint16_t)ws[i + 1] << 16);
}

// jcdctmgr.c quantize(): round half up on the magnitude
__device__ __forceinline__ int jpg_quant(int c, int q8) {
    if (c < 0) return -((-c + (q8 >> 1)) / q8);
    return (c + (q8 >> 1)) / q8;
}

// quantised DC of the last real block of the same component that precedes block b in scan order
// (dummy blocks copy that DC, jccoefct.c), or 0 at the start of the scan
__device__ int jpg_dc_pred(const int16_t *page_coef, int b, const JpegGeom &g, const JpegQuant &Q) {
    const int mcu = b / 6, s = b - mcu * 6;
    if (s >= 4) {
        if (mcu == 0) return 0;
        return jpg_quant(page_coef[(size_t)((mcu - 1) * 6 + s) * 64], Q.q8[1][0]);
    }
    int pb = b - 1;
    for (;;) {
        if (pb < 0) return 0;
        const int ps = pb % 6;
        if (ps >= 4) { pb -= 1; continue; }        // chroma slot: keep walking back (only from s == 0)
        if (!jpg_block(pb, g).dummy) return jpg_quant(page_coef[(size_t)pb * 64], Q.q8[0][0]);
        pb -= 1;
    }
}

__device__ __forceinline__ int jpg_nbits(int v) { return 32 - __clz(v); }  // v >= 0

// Walk the symbols of block b (jchuff.c encode_one_block); F(table slot, symbol, extra bits value, extra bit count)
template <class F>
__device__ __forceinline__ void jpg_walk(const int16_t *page_coef, int b, const JpegGeom &g, const JpegQuant &Q, F &&emit) {
    const JpegBlk k = jpg_block(b, g);
    const int tc = k.comp == 0 ? 0 : 1;
    const int dct = tc ? JT_DC1 : JT_DC0, act = tc ? JT_AC1 : JT_AC0;
    if (k.dummy) {  // DC = previous DC (difference 0), AC all zero
        emit(dct, 0, 0, 0);
        emit(act, 0, 0, 0);
        return;
    }
    const int16_t *c = page_coef + (size_t)b * 64;
    const uint16_t *q8 = Q.q8[tc];
    {
        const int diff = jpg_quant(c[0], q8[0]) - jpg_dc_pred(page_coef, b, g, Q);
        int t = diff, t2 = diff;
        if (t < 0) { t = -t; t2--; }
        const int nb = jpg_nbits(t);
        emit(dct, nb, nb ? (t2 & ((1 << nb) - 1)) : 0, nb);
    }
    int r = 0;
#pragma unroll 1
    for (int i = 1; i < 64; i++) {
        const int z = c_zz[i];
        const int v = jpg_quant(c[z], q8[z]);
        if (v == 0) { r++; continue; }
        while (r > 15) { emit(act, 0xF0, 0, 0); r -= 16; }
        int t = v, t2 = v;
        if (t < 0) { t = -t; t2--; }
        const int nb = jpg_nbits(t);
        emit(act, (r << 4) + nb, t2 & ((1 << nb) - 1), nb);
        r = 0;
    }
    if (r > 0) emit(act, 0, 0, 0);
}

// ---- K2: symbol statistics (optimize=True), one thread per block, CTA-private histograms -------------
__global__ void __launch_bounds__(256) jpeg_hist_kernel(const int16_t *__restrict__ coef, uint32_t *__restrict__ hist, JpegGeom g, JpegQuant Q) {
    __shared__ uint32_t sh[4 * 256];
    for (int i = threadIdx.x; i < 4 * 256; i += 256) sh[i] = 0;
    __syncthreads();
    const int page = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int16_t *pc = coef + (size_t)page * g.blocks * 64;
    if (b < g.blocks) jpg_walk(pc, b, g, Q, [&](int tbl, int sym, int, int) { atomicAdd(&sh[tbl * 256 + sym], 1u); });
    __syncthreads();
    uint32_t *hp = hist + (size_t)page * 4 * 257;
    for (int i = threadIdx.x; i < 4 * 256; i += 256)
        if (sh[i]) atomicAdd(&hp[(i >> 8) * 257 + (i & 255)], sh[i]);
}

// jchuff.c jpeg_make_c_derived_tbl: canonical codes from bits / vals
__device__ void jpg_derive(JpegHuff &T) {
    uint8_t huffsize[257];
    uint16_t huffcode[257];
    int p = 0;
    for (int l = 1; l <= 16; l++)
        for (int i = 0; i < T.bits[l]; i++) huffsize[p++] = (uint8_t)l;
    huffsize[p] = 0;
    const int lastp = p;
    unsigned code = 0;
    int si = huffsize[0];
    p = 0;
    while (huffsize[p]) {
        while (huffsize[p] == si) { huffcode[p++] = (uint16_t)code; code++; }
        code <<= 1;
        si++;
    }
    for (int i = 0; i < 256; i++) { T.code[i] = 0; T.size[i] = 0; }
    for (p = 0; p < lastp; p++) { T.code[T.vals[p]] = huffcode[p]; T.size[T.vals[p]] = huffsize[p]; }
    T.nvals = lastp;
}

// ---- K3: optimal Huffman tables (jchuff.c jpeg_gen_optimal_table), one warp per (page, table) ---------
// The two "smallest frequency, larger symbol on ties" searches of every merge step are warp-parallel (each
// lane scans 9 of the 257 entries, then a shuffle reduction on (frequency, -symbol) keys); the code-length
// chains, the length limiting and the canonical code assignment are sequential and run on lane 0.
__device__ __forceinline__ unsigned long long jpg_warp_min(unsigned long long k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, k, o);
        k = t < k ? t : k;
    }
    return k;
}
__global__ void __launch_bounds__(32) jpeg_opt_table_kernel(const uint32_t *__restrict__ hist, JpegHuff *__restrict__ tables) {
    __shared__ uint32_t freq[257 + 31];
    __shared__ int codesize[257], others[257];
    const int t = blockIdx.x, lane = threadIdx.x;  // t = page * 4 + slot
    const uint32_t *hp = hist + (size_t)t * 257;
    JpegHuff &T = tables[t];
    for (int i = lane; i < 257 + 31; i += 32) freq[i] = i < 256 ? hp[i] : (i == 256 ? 1u : 0u);   // [256]: reserves the all-ones code
    for (int i = lane; i < 257; i += 32) { codesize[i] = 0; others[i] = -1; }
    __syncwarp();
    constexpr unsigned long long NONE = ~0ull;
    for (;;) {
        // key = frequency << 16 | (65535 - symbol): minimum = smallest frequency, larger symbol on ties
        // (libjpeg's scan keeps the LAST index among equal minima; frequencies above 1e9 are never picked)
        unsigned long long k1 = NONE;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const int i = lane + 32 * j;
            const uint32_t f = freq[i];
            if (f && f <= 1000000000u) {
                const unsigned long long k = ((unsigned long long)f << 16) | (unsigned long long)(65535 - i);
                k1 = k < k1 ? k : k1;
            }
        }
        k1 = jpg_warp_min(k1);
        const int c1 = k1 == NONE ? -1 : 65535 - (int)(k1 & 0xffffu);
        unsigned long long k2 = NONE;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const int i = lane + 32 * j;
            const uint32_t f = freq[i];
            if (f && f <= 1000000000u && i != c1) {
                const unsigned long long k = ((unsigned long long)f << 16) | (unsigned long long)(65535 - i);
                k2 = k < k2 ? k : k2;
            }
        }
        k2 = jpg_warp_min(k2);
        if (k2 == NONE) break;
        if (lane == 0) {
            int a = c1, b = 65535 - (int)(k2 & 0xffffu);
            freq[a] += freq[b];
            freq[b] = 0;
            codesize[a]++;
            while (others[a] >= 0) { a = others[a]; codesize[a]++; }
            others[a] = b;
            codesize[b]++;
            while (others[b] >= 0) { b = others[b]; codesize[b]++; }
        }
        __syncwarp();
    }
    if (lane != 0) return;
    uint8_t bits[33];
    for (int i = 0; i <= 32; i++) bits[i] = 0;
    for (int i = 0; i <= 256; i++)
        if (codesize[i]) bits[codesize[i] > 32 ? 32 : codesize[i]]++;
    for (int i = 32; i > 16; i--) {
        while (bits[i] > 0) {
            int j = i - 2;
            while (bits[j] == 0) j--;
            bits[i] -= 2; bits[i - 1]++; bits[j + 1] += 2; bits[j]--;
        }
    }
    int i = 16;
    while (bits[i] == 0) i--;
    bits[i]--;
    for (int l = 0; l <= 16; l++) T.bits[l] = bits[l];
    int p = 0;
    for (int l = 1; l <= 32; l++)
        for (int j = 0; j <= 255; j++)
            if (codesize[j] == l) T.vals[p++] = (uint8_t)j;
    for (; p < 256; p++) T.vals[p] = 0;
    jpg_derive(T);
}

// standard tables: derive codes once, replicate per page
__global__ void __launch_bounds__(32) jpeg_std_table_kernel(JpegHuff *__restrict__ tables, int n_tables) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tables) return;
    jpg_derive(tables[t]);
}

// ---- K4: entropy-coded bits per block -----------------------------------------------------------------
__global__ void __launch_bounds__(256) jpeg_len_kernel(const int16_t *__restrict__ coef, const JpegHuff *__restrict__ tables,
                                                       uint32_t *__restrict__ lens, JpegGeom g, JpegQuant Q) {
    __shared__ uint8_t ssize[4 * 256];
    const int page = blockIdx.y;
    const JpegHuff *T = tables + (size_t)page * 4;
    for (int i = threadIdx.x; i < 4 * 256; i += 256) ssize[i] = T[i >> 8].size[i & 255];
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.blocks) return;
    const int16_t *pc = coef + (size_t)page * g.blocks * 64;
    uint32_t n = 0;
    jpg_walk(pc, b, g, Q, [&](int tbl, int sym, int, int nb) { n += ssize[tbl * 256 + sym] + nb; });
    lens[(size_t)page * g.blocks + b] = n;
}

// ---- K5: exclusive scan of the block lengths, one CTA per page -------------------------------------------
__global__ void __launch_bounds__(1024) jpeg_scan_kernel(const uint32_t *__restrict__ lens, unsigned long long *__restrict__ offs,
                                                         unsigned long long *__restrict__ total_bits, int blocks) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    const int page = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *lp = lens + (size_t)page * blocks;
    unsigned long long *op = offs + (size_t)page * blocks;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < blocks; base += 1024) {
        const int i = base + tid;
        const unsigned long long v = i < blocks ? lp[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = wsum[lane];
            unsigned long long s = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            wsum[lane] = s - w;
        }
        __syncthreads();
        const unsigned long long c = carry;
        if (i < blocks) op[i] = c + wsum[warp] + inc - v;
        __syncthreads();
        if (tid == 1023) carry = c + wsum[warp] + inc;
        __syncthreads();
    }
    if (tid == 0) total_bits[page] = carry;
}

// MSB-first bit writer into a zero-initialised word buffer (neighbouring blocks share boundary words)
struct JpgBits {
    uint32_t *words;
    unsigned long long pos;  // absolute bit position
    unsigned long long buf;
    int nb;
    __device__ __forceinline__ void flush(int n) {  // write the n (<= 32) oldest pending bits
        const uint32_t v = (uint32_t)((buf >> (nb - n)) & (n == 32 ? 0xffffffffu : ((1u << n) - 1u)));
        const unsigned long long wi = pos >> 5;
        const int sh = (int)(pos & 31);
        // logical big-endian word: bit 31 first
        const int room = 32 - sh;
        if (n <= room) {
            const uint32_t w = v << (room - n);
            atomicOr(&words[wi], __byte_perm(w, 0, 0x0123));
        } else {
            const uint32_t w0 = v >> (n - room), w1 = v << (32 - (n - room));
            atomicOr(&words[wi], __byte_perm(w0, 0, 0x0123));
            atomicOr(&words[wi + 1], __byte_perm(w1, 0, 0x0123));
        }
        pos += n;
        nb -= n;
    }
    __device__ __forceinline__ void put(uint32_t code, int n) {
        if (n == 0) return;
        buf = (buf << n) | code;
        nb += n;
        if (nb >= 32) flush(32);
    }
    __device__ __forceinline__ void finish() {
        if (nb > 0) flush(nb);
    }
};

// ---- K6: emission, one thread per block -----------------------------------------------------------------
__global__ void __launch_bounds__(256) jpeg_emit_kernel(const int16_t *__restrict__ coef, const JpegHuff *__restrict__ tables,
                                                        const unsigned long long *__restrict__ offs,
                                                        const unsigned long long *__restrict__ total_bits, uint8_t *__restrict__ packed,
                                                        size_t packed_stride, JpegGeom g, JpegQuant Q) {
    __shared__ uint16_t scode[4 * 256];
    __shared__ uint8_t ssize[4 * 256];
    const int page = blockIdx.y;
    const JpegHuff *T = tables + (size_t)page * 4;
    for (int i = threadIdx.x; i < 4 * 256; i += 256) { scode[i] = T[i >> 8].code[i & 255]; ssize[i] = T[i >> 8].size[i & 255]; }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.blocks) return;
    const int16_t *pc = coef + (size_t)page * g.blocks * 64;
    JpgBits W;
    W.words = reinterpret_cast<uint32_t *>(packed + (size_t)page * packed_stride);
    W.pos = offs[(size_t)page * g.blocks + b];
    W.buf = 0; W.nb = 0;
    jpg_walk(pc, b, g, Q, [&](int tbl, int sym, int val, int nb) {
        W.put(scode[tbl * 256 + sym], ssize[tbl * 256 + sym]);
        W.put((uint32_t)val, nb);
    });
    if (b == g.blocks - 1) {  // jchuff.c flush_bits: pad the last byte with ones
        const int pad = (int)((8 - (total_bits[page] & 7)) & 7);
        W.put((1u << pad) - 1u, pad);
    }
    W.finish();
}

// ---- K7: byte stuffing (0xFF -> 0xFF 0x00), one CTA per page ------------------------------------------------
__global__ void __launch_bounds__(1024) jpeg_stuff_kernel(const uint8_t *__restrict__ packed, size_t packed_stride,
                                                          const unsigned long long *__restrict__ total_bits, uint8_t *__restrict__ out,
                                                          size_t out_stride, unsigned long long *__restrict__ out_bytes) {
    __shared__ uint32_t wsum[32];
    __shared__ unsigned long long carry;
    const int page = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *src = packed + (size_t)page * packed_stride;
    uint8_t *dst = out + (size_t)page * out_stride;
    const unsigned long long nbytes = (total_bits[page] + 7) >> 3;
    if (tid == 0) carry = 0;
    __syncthreads();
    constexpr int PER = 16;
    for (unsigned long long base = 0; base < nbytes; base += 1024ull * PER) {
        const unsigned long long p0 = base + (unsigned long long)tid * PER;
        uint8_t v[PER];
        uint32_t cnt = 0, have = 0;
        if (p0 < nbytes) {
            have = (uint32_t)min((unsigned long long)PER, nbytes - p0);
            if (have == PER) {
                const uint4 q = *reinterpret_cast<const uint4 *>(src + p0);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < PER; i++) v[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
            } else {
                for (uint32_t i = 0; i < PER; i++) v[i] = i < have ? src[p0 + i] : 0;
            }
            for (uint32_t i = 0; i < have; i++) cnt += 1 + (v[i] == 0xFF);
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = wsum[lane];
            uint32_t s = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            wsum[lane] = s - w;
        }
        __syncthreads();
        const unsigned long long c = carry;
        unsigned long long o = c + wsum[warp] + inc - cnt;
        for (uint32_t i = 0; i < have; i++) {
            if (o < out_stride) dst[o] = v[i];
            o++;
            if (v[i] == 0xFF) {
                if (o < out_stride) dst[o] = 0;
                o++;
            }
        }
        __syncthreads();
        if (tid == 1023) carry = c + wsum[warp] + inc;
        __syncthreads();
    }
    if (tid == 0) out_bytes[page] = carry;
}

// ---- workspace -------------------------------------------------------------------------------------------
struct JpegLayout {
    size_t coef_off, lens_off, offs_off, hist_off, tables_off, totals_off, packed_off, stuffed_off, total;
    size_t packed_stride, stuffed_stride;
};
static size_t jalign(size_t v) { return (v + 255) & ~(size_t)255; }
static JpegLayout jpeg_layout(int n, const JpegGeom &g) {
    JpegLayout L;
    size_t off = 0;
    L.coef_off = off; off = jalign(off + (size_t)n * g.blocks * 64 * 2);
    L.lens_off = off; off = jalign(off + (size_t)n * g.blocks * 4);
    L.offs_off = off; off = jalign(off + (size_t)n * g.blocks * 8);
    L.hist_off = off; off = jalign(off + (size_t)n * 4 * 257 * 4);
    L.tables_off = off; off = jalign(off + (size_t)n * 4 * sizeof(JpegHuff));
    L.totals_off = off; off = jalign(off + (size_t)n * 2 * 8);
    // worst case of one block: 27 + 63 * 26 bits = 209 bytes
    L.packed_stride = jalign((size_t)g.blocks * 210 + 64);
    L.stuffed_stride = L.packed_stride + L.packed_stride / 4;   // stuffing beyond +25 % is reported as an error
    L.packed_off = off; off = jalign(off + (size_t)n * L.packed_stride);
    L.stuffed_off = off; off = jalign(off + (size_t)n * L.stuffed_stride);
    L.total = off;
    return L;
}

static void jpeg_quant_tables(int quality, uint8_t q[2][64]) {
    // jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
    if (quality <= 0) quality = 1;
    if (quality > 100) quality = 100;
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < 64; i++) {
            long v = ((long)(t ? std_qchr[i] : std_qlum[i]) * scale + 50L) / 100L;
            if (v <= 0) v = 1;
            if (v > 255) v = 255;
            q[t][i] = (uint8_t)v;
        }
}

static size_t put_seg(uint8_t *o, size_t p, uint8_t marker, const uint8_t *payload, size_t len) {
    o[p++] = 0xFF; o[p++] = marker;
    o[p++] = (uint8_t)((len + 2) >> 8); o[p++] = (uint8_t)((len + 2) & 255);
    memcpy(o + p, payload, len);
    return p + len;
}

}  // namespace lumina

using namespace lumina;

LUMINA_API size_t lumina_jpeg_workspace_bytes(int n, int h, int w) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    return jpeg_layout(n, jpeg_geom(h, w)).total;
}

// flags: bit 0 = optimize (per-image optimal Huffman tables), bit 1 = reuse the DCT coefficients the previous
// call left in this workspace (same pages, another quality)
LUMINA_API int lumina_jpeg_encode_rgb(const uint8_t *d_rgb, int n, int h, int w, int quality, int flags, uint8_t *h_out,
                                      size_t out_stride, int64_t *h_sizes, void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(d_rgb && h_sizes && d_workspace, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE(h <= 65535 && w <= 65535, "JPEG dimensions are 16-bit");
    LUMINA_REQUIRE((((uintptr_t)d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
    const JpegGeom g = jpeg_geom(h, w);
    const JpegLayout L = jpeg_layout(n, g);
    if (workspace_bytes < L.total) return set_error(LUMINA_E_NOMEM, "jpeg workspace too small: need %zu bytes", L.total);
    cudaStream_t st = as_stream(stream);
    uint8_t *ws = (uint8_t *)d_workspace;
    int16_t *coef = (int16_t *)(ws + L.coef_off);
    uint32_t *lens = (uint32_t *)(ws + L.lens_off);
    unsigned long long *offs = (unsigned long long *)(ws + L.offs_off);
    uint32_t *hist = (uint32_t *)(ws + L.hist_off);
    JpegHuff *tables = (JpegHuff *)(ws + L.tables_off);
    unsigned long long *totals = (unsigned long long *)(ws + L.totals_off);  // [n] bits, [n] stuffed bytes
    const bool optimize = (flags & 1) != 0, reuse = (flags & 2) != 0;

    uint8_t qt[2][64];
    jpeg_quant_tables(quality, qt);
    JpegQuant Q;
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < 64; i++) Q.q8[t][i] = (uint16_t)(qt[t][i] * 8);

    const dim3 bgrid((unsigned)div_up(g.blocks, 256), (unsigned)n);
    LUMINA_REQUIRE(n <= 65535, "batch too large for grid");
    if (!reuse) {
        jpeg_fdct_kernel<<<dim3((unsigned)div_up(g.blocks, 128), (unsigned)n), 128, 0, st>>>(d_rgb, coef, g);
        LUMINA_KERNEL_CHECK("jpeg_fdct_kernel");
    }
    std::vector<JpegHuff> host_tables((size_t)n * 4);
    if (optimize) {
        LUMINA_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)n * 4 * 257 * 4, st));
        jpeg_hist_kernel<<<bgrid, 256, 0, st>>>(coef, hist, g, Q);
        LUMINA_KERNEL_CHECK("jpeg_hist_kernel");
        jpeg_opt_table_kernel<<<n * 4, 32, 0, st>>>(hist, tables);
        LUMINA_KERNEL_CHECK("jpeg_opt_table_kernel");
    } else {
        JpegHuff std4[4];
        memset(std4, 0, sizeof(std4));
        memcpy(std4[JT_DC0].bits, std_dc_lum_bits, 17); memcpy(std4[JT_DC0].vals, std_dc_vals, 12);
        memcpy(std4[JT_AC0].bits, std_ac_lum_bits, 17); memcpy(std4[JT_AC0].vals, std_ac_lum_vals, 162);
        memcpy(std4[JT_DC1].bits, std_dc_chr_bits, 17); memcpy(std4[JT_DC1].vals, std_dc_vals, 12);
        memcpy(std4[JT_AC1].bits, std_ac_chr_bits, 17); memcpy(std4[JT_AC1].vals, std_ac_chr_vals, 162);
        for (int i = 0; i < n; i++) memcpy(&host_tables[(size_t)i * 4], std4, sizeof(std4));
        LUMINA_CUDA_TRY(cudaMemcpyAsync(tables, host_tables.data(), host_tables.size() * sizeof(JpegHuff), cudaMemcpyHostToDevice, st));
        jpeg_std_table_kernel<<<div_up(n * 4, 32), 32, 0, st>>>(tables, n * 4);
        LUMINA_KERNEL_CHECK("jpeg_std_table_kernel");
        LUMINA_CUDA_TRY(cudaStreamSynchronize(st));  // host_tables is reused below
    }
    jpeg_len_kernel<<<bgrid, 256, 0, st>>>(coef, tables, lens, g, Q);
    LUMINA_KERNEL_CHECK("jpeg_len_kernel");
    jpeg_scan_kernel<<<n, 1024, 0, st>>>(lens, offs, totals, g.blocks);
    LUMINA_KERNEL_CHECK("jpeg_scan_kernel");
    // the packed stream is OR-ed into place: clear what it can reach (bounded by the stride)
    LUMINA_CUDA_TRY(cudaMemsetAsync(ws + L.packed_off, 0, (size_t)n * L.packed_stride, st));
    jpeg_emit_kernel<<<bgrid, 256, 0, st>>>(coef, tables, offs, totals, ws + L.packed_off, L.packed_stride, g, Q);
    LUMINA_KERNEL_CHECK("jpeg_emit_kernel");
    jpeg_stuff_kernel<<<n, 1024, 0, st>>>(ws + L.packed_off, L.packed_stride, totals, ws + L.stuffed_off, L.stuffed_stride, totals + n);
    LUMINA_KERNEL_CHECK("jpeg_stuff_kernel");
    std::vector<unsigned long long> tot((size_t)n * 2);
    LUMINA_CUDA_TRY(cudaMemcpyAsync(tot.data(), totals, tot.size() * 8, cudaMemcpyDeviceToHost, st));
    LUMINA_CUDA_TRY(cudaMemcpyAsync(host_tables.data(), tables, host_tables.size() * sizeof(JpegHuff), cudaMemcpyDeviceToHost, st));
    LUMINA_CUDA_TRY(cudaStreamSynchronize(st));

    // ---- headers (jcmarker.c order, as Pillow writes them) and the copy of every stream that fits ----
    for (int i = 0; i < n; i++) {
        if (tot[(size_t)n + i] > L.stuffed_stride)
            return set_error(LUMINA_E_NOMEM, "page %d: entropy-coded data exceeds the workspace bound", i);
        uint8_t hd[1024];
        size_t p = 0;
        hd[p++] = 0xFF; hd[p++] = 0xD8;
        static const uint8_t jfif[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
        p = put_seg(hd, p, 0xE0, jfif, 14);
        for (int t = 0; t < 2; t++) {
            uint8_t dq[65];
            dq[0] = (uint8_t)t;
            for (int k = 0; k < 64; k++) dq[1 + k] = qt[t][h_zz[k]];
            p = put_seg(hd, p, 0xDB, dq, 65);
        }
        const uint8_t sof[15] = {8, (uint8_t)(h >> 8), (uint8_t)(h & 255), (uint8_t)(w >> 8), (uint8_t)(w & 255), 3,
                                 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
        p = put_seg(hd, p, 0xC0, sof, 15);
        static const int order[4] = {JT_DC0, JT_AC0, JT_DC1, JT_AC1};
        static const uint8_t tcth[4] = {0x00, 0x10, 0x01, 0x11};
        for (int k = 0; k < 4; k++) {
            const JpegHuff &T = host_tables[(size_t)i * 4 + order[k]];
            uint8_t dh[1 + 16 + 256];
            dh[0] = tcth[k];
            int nv = 0;
            for (int l = 1; l <= 16; l++) { dh[l] = T.bits[l]; nv += T.bits[l]; }
            memcpy(dh + 17, T.vals, (size_t)nv);
            p = put_seg(hd, p, 0xC4, dh, 17 + (size_t)nv);
        }
        static const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
        p = put_seg(hd, p, 0xDA, sos, 10);
        const size_t body = (size_t)tot[(size_t)n + i];
        const size_t fsize = p + body + 2;
        h_sizes[i] = (int64_t)fsize;
        if (h_out && fsize <= out_stride) {
            uint8_t *o = h_out + (size_t)i * out_stride;
            memcpy(o, hd, p);
            LUMINA_CUDA_TRY(cudaMemcpyAsync(o + p, ws + L.stuffed_off + (size_t)i * L.stuffed_stride, body, cudaMemcpyDeviceToHost, st));
            o[p + body] = 0xFF; o[p + body + 1] = 0xD9;
        }
    }
    LUMINA_CUDA_TRY(cudaStreamSynchronize(st));
    return LUMINA_OK;
}
