// k_resize.cu -- PIL Image.resize(LANCZOS) for uint8 pages, bit-exact.
//
// Replaces Pillow libImaging/Resample.c (precompute_coeffs,
// normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc) reached from
// backend/utils/image_preprocessing.py:110 (resize_if_needed) and :551.
//
// Arithmetic: per axis support = 3*max(scale,1), coefficients
// in double on the HOST (glibc sin, so no device-ulp can flip a fixed-point
// coefficient), normalised, converted to 22-bit fixed point; horizontal pass
// first, rounded to uint8, then vertical pass on that uint8 intermediate.
//
// B200 design: one fused kernel.  A CTA owns a strip of TOW output columns and a
// segment of output rows of one page and slides down the page: input rows are
// staged chunk-wise in shared memory with 128-bit coalesced loads, the
// horizontal pass writes a uint8 ring buffer in shared memory (the 14.9 MB/page
// intermediate never touches HBM), and the vertical pass emits every output row
// whose tap window is complete.  HBM traffic = input once (+halo) + output once.
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include <cuda.h>   // CUtensorMap (the encoder is fetched with cudaGetDriverEntryPoint: no libcuda link)
#include <cudaTypedefs.h>
#include <string.h>

#include "common.cuh"

struct lumina_resize_plan {
    int in_h, in_w, out_h, out_w;
    int kx, ky;                 // taps per output column / row
    int32_t *d_bx, *d_cx;       // [out_w][2], [out_w][kx]
    int32_t *d_by, *d_cy;       // [out_h][2], [out_h][ky]
    // dp4a path: coefficients split into byte planes (c = c0 + 256*c1 + 65536*c2, c2 signed), 4 taps per word
    uint32_t *d_cxp;            // [out_w][3 planes][kxw]            tap t -> word t/4, byte t%4
    uint32_t *d_cyp;            // [out_h][3 planes][kyw]            tap t -> byte (ymin & 3) + t  (row-group phase)
    int kxw, kyw;
    int max_seg_px;             // widest input column span of any TOW-column strip
    // tensor-core path (int8 mma.sync): per 8-column tile the B fragments of the three coefficient byte planes
    uint32_t *d_bfrag;          // [tiles][ksteps][3 planes][2][32 lanes]
    int32_t *d_kb;              // [tiles] first staged byte (multiple of 4, relative to the strip's xs16) of the tile's K window
    int ksteps;                 // K window of a tile in units of 32 input pixels (0: path not available)
    int imma_span;              // bytes of a staged plane row the A fragments may touch
    // tensor-core vertical pass: per tile of 8 output rows the B fragments over a K window of 64 intermediate rows
    uint32_t *d_vfrag;          // [out_h/8 tiles][2 k-steps][3 planes][2][32 lanes]
    int32_t *d_vg0;             // [tiles][2] first ring row-group (input row / 4) of the tile's K window, last tap row + 1
    int vtiles;                 // 0: path not available (a tile's taps do not fit 64 rows)
    int vsteps;                 // K steps of 32 ring rows a vertical tile needs (1: every tile's taps fit 32 rows; 2)
    int device;
};

namespace lumina {

constexpr int PREC_BITS = 22;
constexpr int TOW = 64;   // output columns per CTA strip
constexpr int RB = 16;    // input rows per staged chunk
constexpr int RING = 64;  // intermediate ring rows (>= ky + RB)

static double sinc_d(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
static double lanczos3_d(double x) { return (-3.0 <= x && x < 3.0) ? sinc_d(x) * sinc_d(x / 3) : 0.0; }

static int ksize_for(int in_size, int out_size) {
    double scale = (double)in_size / out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(3.0 * fs) * 2 + 1;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc
static void host_coeffs(int in_size, int out_size, std::vector<int32_t> &bounds, std::vector<int32_t> &coeffs) {
    double scale = (double)in_size / out_size;
    double fs = scale < 1.0 ? 1.0 : scale;
    double support = 3.0 * fs;
    int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    coeffs.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    double ss = 1.0 / fs;
    for (int xx = 0; xx < out_size; xx++) {
        double center = (xx + 0.5) * scale, ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; x++) {
            double w = lanczos3_d((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; x++)
            if (ww != 0.0) k[x] /= ww;
        bounds[xx * 2] = xmin;
        bounds[xx * 2 + 1] = xmax;
        for (int x = 0; x < xmax; x++)
            coeffs[(size_t)xx * ksize + x] =
                k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PREC_BITS)) : (int)(0.5 + k[x] * (1 << PREC_BITS));
    }
}

struct ResizeParams {
    const uint8_t *src;
    uint8_t *dst;
    const int32_t *bx, *cx, *by, *cy;
    int in_h, in_w, out_h, out_w, kx, ky;
    int rows_per_seg;  // output rows per CTA
    int segb;          // bytes per staged input row in shared memory (multiple of 16)
    size_t src_total;  // total bytes of the src batch (over-read guard)
};

__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)min(max(v >> PREC_BITS, 0), 255); }
// same value, one VIMNMX.RELU: max(min(v >> 22, 255), 0)
__device__ __forceinline__ int clip8r(int v) { return __vimin_s32_relu(v >> PREC_BITS, 255); }

// C channels, KX = compile-time bound on horizontal taps (kx <= KX)
template <int C, int KX>
__global__ void __launch_bounds__(256) resize_strip_kernel(const ResizeParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *inbuf = smem;                          // [RB][segb]
    uint8_t *ring = smem + (size_t)RB * p.segb;     // [RING][TOW*C]
    __shared__ int rowoff[RB];                      // byte offset of column xs inside each staged row

    constexpr int ROWB = TOW * C;  // intermediate bytes per row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = warp & 1, rw = warp >> 1;  // column group, row interleave (4)
    const int ox0 = blockIdx.x * TOW;
    const int oy0 = blockIdx.y * p.rows_per_seg;
    const int oy1 = min(oy0 + p.rows_per_seg, p.out_h);
    const int page = blockIdx.z;
    const size_t pitch = (size_t)p.in_w * C;
    const uint8_t *src = p.src + (size_t)page * p.in_h * pitch;
    uint8_t *dst = p.dst + (size_t)page * p.out_h * p.out_w * C;

    // strip geometry
    const int ox_last = min(ox0 + TOW, p.out_w) - 1;
    const int xs = p.bx[ox0 * 2];
    const int xe = p.bx[ox_last * 2] + p.bx[ox_last * 2 + 1];
    const int ys = p.by[oy0 * 2];
    const int ye = p.by[(oy1 - 1) * 2] + p.by[(oy1 - 1) * 2 + 1];

    // this lane's output column + coefficients (registers)
    const int ox = ox0 + cg * 32 + lane;
    const bool col_ok = ox < p.out_w;
    int coef[KX];
    int xoff = 0;
    {
        int xmin = col_ok ? p.bx[ox * 2] : xs;
        xoff = (xmin - xs) * C;
#pragma unroll
        for (int t = 0; t < KX; t++) coef[t] = (col_ok && t < p.kx) ? p.cx[(size_t)ox * p.kx + t] : 0;
    }

    const uint8_t *src_end = p.src + p.src_total;
    int next_oy = oy0;  // next output row to emit

    for (int r0 = ys; r0 < ye; r0 += RB) {
        const int nrows = min(RB, ye - r0);
        // ---- stage rows [r0, r0+nrows) x bytes [xs*C, xe*C) (aligned down) ----
        {
            const int seg_bytes = (xe - xs) * C;
            for (int r = warp; r < nrows; r += 8) {
                const uint8_t *g = src + (size_t)(r0 + r) * pitch + (size_t)xs * C;
                const int mis = (int)((uintptr_t)g & 15);
                const uint8_t *ga = g - mis;
                if (lane == 0) rowoff[r] = mis;
                const int nvec = (mis + seg_bytes + 15) >> 4;
                uint8_t *s = inbuf + (size_t)r * p.segb;
                for (int v = lane; v < nvec; v += 32) {
                    const uint8_t *gp = ga + (size_t)v * 16;
                    uint4 val;
                    if (gp >= p.src && gp + 16 <= src_end) val = ldg_stream_u4(gp);
                    else {
                        uint32_t t4[4] = {0, 0, 0, 0};
                        for (int b = 0; b < 16; b++)
                            if (gp + b >= p.src && gp + b < src_end) t4[b >> 2] |= (uint32_t)gp[b] << (8 * (b & 3));
                        val = make_uint4(t4[0], t4[1], t4[2], t4[3]);
                    }
                    *reinterpret_cast<uint4 *>(s + (size_t)v * 16) = val;
                }
            }
        }
        __syncthreads();
        // ---- horizontal pass: lane = output column, rows interleaved over 4 warps ----
        if (col_ok) {
            constexpr int NW = (KX * C + 3) / 4;  // aligned words covering KX taps
            for (int r = rw; r < nrows; r += 4) {
                const int off = rowoff[r] + xoff;
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(inbuf + (size_t)r * p.segb + (off & ~3));
                const int sh = (off & 3) * 8;
                uint32_t a[NW + 1];
#pragma unroll
                for (int i = 0; i <= NW; i++) a[i] = wp[i];
#pragma unroll
                for (int i = 0; i < NW; i++) a[i] = __funnelshift_r(a[i], a[i + 1], sh);
                int acc[C];
#pragma unroll
                for (int ch = 0; ch < C; ch++) acc[ch] = 1 << (PREC_BITS - 1);
#pragma unroll
                for (int t = 0; t < KX; t++)
#pragma unroll
                    for (int ch = 0; ch < C; ch++) {
                        const int bi = t * C + ch;
                        // one PRMT per tap (byte -> zero-extended int) instead of SHF+LOP3: the kernel is ALU-pipe bound
                        acc[ch] += (int)__byte_perm(a[bi >> 2], 0u, 0x4440u | (unsigned)(bi & 3)) * coef[t];
                    }
                uint8_t *o = ring + (size_t)((r0 + r - ys) & (RING - 1)) * ROWB + (cg * 32 + lane) * C;
#pragma unroll
                for (int ch = 0; ch < C; ch++) o[ch] = clip8(acc[ch]);
            }
        }
        __syncthreads();
        // ---- vertical pass: emit all output rows whose window is now complete ----
        const int rows_done = r0 + nrows;
        int oy_end = next_oy;
        while (oy_end < oy1 && p.by[oy_end * 2] + p.by[oy_end * 2 + 1] <= rows_done) oy_end++;
        constexpr int WPR = ROWB / 4;  // words per intermediate row
        const int ntask = (oy_end - next_oy) * WPR;
        for (int task = tid; task < ntask; task += 256) {
            const int oy = next_oy + task / WPR, wj = task % WPR;
            const int ymin = p.by[oy * 2], n = p.by[oy * 2 + 1];
            const int32_t *k = p.cy + (size_t)oy * p.ky;
            int a0 = 1 << (PREC_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
            for (int j = 0; j < n; j++) {
                const uint32_t w =
                    *reinterpret_cast<const uint32_t *>(ring + (size_t)((ymin + j - ys) & (RING - 1)) * ROWB + wj * 4);
                const int kk = __ldg(k + j);
                a0 += (int)__byte_perm(w, 0u, 0x4440u) * kk;
                a1 += (int)__byte_perm(w, 0u, 0x4441u) * kk;
                a2 += (int)__byte_perm(w, 0u, 0x4442u) * kk;
                a3 += (int)__byte_perm(w, 0u, 0x4443u) * kk;
            }
            const int bcol = ox0 * C + wj * 4;  // byte column in the output row
            const int row_bytes = p.out_w * C;
            uint8_t *o = dst + (size_t)oy * row_bytes + bcol;
            if (bcol + 0 < row_bytes) o[0] = clip8(a0);
            if (bcol + 1 < row_bytes) o[1] = clip8(a1);
            if (bcol + 2 < row_bytes) o[2] = clip8(a2);
            if (bcol + 3 < row_bytes) o[3] = clip8(a3);
        }
        next_oy = oy_end;
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// dp4a variant (RGB, 16-byte aligned rows): the exact 22-bit x 8-bit MACs as byte-plane dot products.
// A coefficient is c = c0 + 256*c1 + 65536*c2 (c0, c1 unsigned bytes, c2 signed), so
//   sum_t c[t]*p[t] = dp4a(p, c0) + 256*dp4a(p, c1) + 65536*dp4a(p, c2)   -- still pure integer, bit-exact,
// and one dp4a consumes 4 taps of one channel.  That needs 4 consecutive taps of a channel in one word:
// rows are staged PLANAR (R | G | B planes, de-interleaved with 6 PRMT per 4 pixels while staging) for the
// horizontal pass, and the intermediate ring keeps 4 consecutive ROWS per word for the vertical pass.
// Per (row, output column): 54 dp4a + 18 funnel shifts instead of 69 IMAD + 69 PRMT.
// ---------------------------------------------------------------------------------------------
constexpr int RINGG = 16;  // ring row-groups of 4 rows (>= (ky + RB + 6) / 4)

// unsigned pixel bytes x signed coefficient bytes (the intrinsic only offers s8*s8 and u8*u8)
__device__ __forceinline__ int dp4a_u8s8(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct ResizeDp4aParams {
    const uint8_t *src;
    uint8_t *dst;
    const int32_t *bx, *by;
    const uint32_t *cxp, *cyp;
    const uint32_t *bfrag;
    const int32_t *kb;
    const uint32_t *vfrag;
    const int32_t *vg0;
    int in_h, in_w, out_h, out_w, kxw, kyw;
    int rows_per_seg;
    int segpx;         // staged pixels per row and plane (multiple of 16, + slack)
    int rawpitch;      // bulk-staged variant: bytes between staged interleaved rows (== 48 mod 128: conflict-free fragments)
    int copy_bytes;    // bulk-staged variant: bytes fetched per row (multiple of 16)
    size_t src_total;
};

template <int KXW>
__global__ void __launch_bounds__(256, 4) resize_strip_dp4a_kernel(const ResizeDp4aParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int ROWB = TOW * 3;
    uint8_t *inbuf = smem;                                             // [RB][3][segpx]
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem + (size_t)RB * 3 * p.segpx);  // [RINGG][ROWB] words of 4 rows

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = warp & 1, rw = warp >> 1;
    const int ox0 = blockIdx.x * TOW;
    const int oy0 = blockIdx.y * p.rows_per_seg;
    const int oy1 = min(oy0 + p.rows_per_seg, p.out_h);
    const int page = blockIdx.z;
    const size_t pitch = (size_t)p.in_w * 3;
    const uint8_t *src = p.src + (size_t)page * p.in_h * pitch;
    uint8_t *dst = p.dst + (size_t)page * p.out_h * p.out_w * 3;

    const int ox_last = min(ox0 + TOW, p.out_w) - 1;
    const int xs = p.bx[ox0 * 2];
    const int xe = p.bx[ox_last * 2] + p.bx[ox_last * 2 + 1];
    const int xs16 = xs & ~15;
    const int ngroups = (xe - xs16 + 15) >> 4;  // 16-pixel groups staged per row
    const int ys = p.by[oy0 * 2];
    const int ye = p.by[(oy1 - 1) * 2] + p.by[(oy1 - 1) * 2 + 1];

    const int ox = ox0 + cg * 32 + lane;
    const bool col_ok = ox < p.out_w;
    uint32_t cw[3][KXW];
    int poff = 0;  // byte offset of this lane's first tap inside a plane row
    {
        const int xmin = col_ok ? p.bx[ox * 2] : xs;
        poff = xmin - xs16;
#pragma unroll
        for (int pl = 0; pl < 3; pl++)
#pragma unroll
            for (int j = 0; j < KXW; j++)
                cw[pl][j] = (col_ok && j < p.kxw) ? p.cxp[((size_t)ox * 3 + pl) * p.kxw + j] : 0u;
    }
    const uint8_t *src_end = p.src + p.src_total;
    int next_oy = oy0;
    __shared__ int s_oy_end;

    for (int r0 = ys; r0 < ye; r0 += RB) {
        const int nrows = min(RB, ye - r0);
        // ---- stage + de-interleave: a thread moves 16 pixels (48 B in, 3 x 16 B out) ----
        for (int t = tid; t < nrows * ngroups; t += 256) {
            const int r = t / ngroups, g = t - r * ngroups;
            const uint8_t *gp = src + (size_t)(r0 + r) * pitch + (size_t)(xs16 + g * 16) * 3;
            uint32_t w[12];
            if (gp + 48 <= src_end) {
                const uint4 a = ldg_stream_u4(gp), b = ldg_stream_u4(gp + 16), c = ldg_stream_u4(gp + 32);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
                w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    uint32_t v = 0;
                    for (int b = 0; b < 4; b++)
                        if (gp + i * 4 + b < src_end) v |= (uint32_t)gp[i * 4 + b] << (8 * b);
                    w[i] = v;
                }
            }
            uint32_t R[4], G[4], B[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
                R[q] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
                G[q] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
                B[q] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
            }
            uint8_t *sr = inbuf + (size_t)r * 3 * p.segpx + g * 16;
            *reinterpret_cast<uint4 *>(sr) = make_uint4(R[0], R[1], R[2], R[3]);
            *reinterpret_cast<uint4 *>(sr + p.segpx) = make_uint4(G[0], G[1], G[2], G[3]);
            *reinterpret_cast<uint4 *>(sr + 2 * p.segpx) = make_uint4(B[0], B[1], B[2], B[3]);
        }
        __syncthreads();
        // ---- horizontal pass ----
        if (col_ok) {
            for (int r = rw; r < nrows; r += 4) {
                const int arow = r0 + r;  // absolute input row
                uint32_t *rg = ring + (size_t)((arow >> 2) & (RINGG - 1)) * ROWB + (cg * 32 + lane) * 3;
                const int sh = (poff & 3) * 8;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const uint32_t *wp = reinterpret_cast<const uint32_t *>(inbuf + ((size_t)r * 3 + ch) * p.segpx + (poff & ~3));
                    uint32_t a[KXW + 1];
#pragma unroll
                    for (int i = 0; i <= KXW; i++) a[i] = wp[i];
                    int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                    for (int i = 0; i < KXW; i++) {
                        const uint32_t v = __funnelshift_r(a[i], a[i + 1], sh);
                        s0 = (int)__dp4a(v, cw[0][i], (uint32_t)s0);
                        s1 = (int)__dp4a(v, cw[1][i], (uint32_t)s1);
                        s2 = dp4a_u8s8(v, cw[2][i], s2);  // signed top plane
                    }
                    // exact modulo 2^32; the true sum fits int32
                    const int acc = (int)((1u << (PREC_BITS - 1)) + (uint32_t)s0 + ((uint32_t)s1 << 8) + ((uint32_t)s2 << 16));
                    reinterpret_cast<uint8_t *>(rg + ch)[arow & 3] = clip8(acc);
                }
            }
        }
        if (tid == 255) {  // which output rows have their whole tap window in the ring after this chunk
            const int rows_done = r0 + nrows;
            int oe = next_oy;
            while (oe < oy1 && p.by[oe * 2] + p.by[oe * 2 + 1] <= rows_done) oe++;
            s_oy_end = oe;
        }
        __syncthreads();
        // ---- vertical pass: a thread owns 4 adjacent byte columns of one output row (one 128-bit ring
        // load per 4 taps x 4 columns, the row's coefficient words shared by the 4 columns) ----
        const int oy_end = s_oy_end;
        constexpr int COL4 = ROWB / 4;
        const int ntask = (oy_end - next_oy) * COL4;
        const int row_bytes = p.out_w * 3;
        for (int task = tid; task < ntask; task += 256) {
            const int orow = task / COL4, c4 = task - orow * COL4;
            const int oy = next_oy + orow;
            const int bcol = ox0 * 3 + c4 * 4;
            if (bcol >= row_bytes) continue;
            const int ymin = p.by[oy * 2];
            const uint32_t *k = p.cyp + (size_t)oy * 3 * p.kyw;
            int s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
            int rgrp = (ymin >> 2) & (RINGG - 1);
            for (int j = 0; j < p.kyw; j++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(ring + (size_t)rgrp * ROWB + c4 * 4);
                const uint32_t k0 = __ldg(k + j), k1 = __ldg(k + p.kyw + j), k2 = __ldg(k + 2 * p.kyw + j);
                const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    s0[q] = (int)__dp4a(vv[q], k0, (uint32_t)s0[q]);
                    s1[q] = (int)__dp4a(vv[q], k1, (uint32_t)s1[q]);
                    s2[q] = dp4a_u8s8(vv[q], k2, s2[q]);
                }
                rgrp = (rgrp + 1) & (RINGG - 1);
            }
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                o[q] = clip8((int)((1u << (PREC_BITS - 1)) + (uint32_t)s0[q] + ((uint32_t)s1[q] << 8) + ((uint32_t)s2[q] << 16)));
            uint8_t *dp = dst + (size_t)oy * row_bytes + bcol;
            if (bcol + 4 <= row_bytes && (((uintptr_t)dp) & 3) == 0) {
                *reinterpret_cast<uint32_t *>(dp) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            } else if (bcol + 4 <= row_bytes && (((uintptr_t)dp) & 1) == 0) {
                *reinterpret_cast<uint16_t *>(dp) = (uint16_t)(o[0] | (o[1] << 8));
                *reinterpret_cast<uint16_t *>(dp + 2) = (uint16_t)(o[2] | (o[3] << 8));
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (bcol + q < row_bytes) dp[q] = (uint8_t)o[q];
            }
        }
        next_oy = oy_end;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core variant of the horizontal pass.  PIL's horizontal resample IS a banded matrix product: 16 staged
// rows x the K input pixels under an 8-column tile (A, u8, straight from the de-interleaved plane rows in shared
// memory) times the tile's coefficients (B, the same three byte planes as the dp4a form: two unsigned, one signed),
// accumulated in int32 -- exact, so the result is the same bytes.  One warp owns one 8-column tile of the strip and
// issues mma.sync.m16n8k32 (u8 x u8 / u8 x s8 -> s32): KSTEPS x 3 MMAs + 4 x KSTEPS shared-memory loads per channel
// replace 16 x 8 x (18 dp4a + 6 funnel shifts + 7 loads).  Staging and the vertical pass are the dp4a kernel's.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_u8u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// first MMA of a chain: C = 0 (the zero register, no accumulator initialisation)
__device__ __forceinline__ void mma_u8s8_z(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(0));
}

template <int KSTEPS>
__global__ void __launch_bounds__(256, 4) resize_strip_imma_kernel(const ResizeDp4aParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int ROWB = TOW * 3;
    uint8_t *inbuf = smem;                                             // [RB][3][segpx]
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem + (size_t)RB * 3 * p.segpx);  // [RINGG][ROWB] words of 4 rows

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ox0 = blockIdx.x * TOW;
    const int oy0 = blockIdx.y * p.rows_per_seg;
    const int oy1 = min(oy0 + p.rows_per_seg, p.out_h);
    const int page = blockIdx.z;
    const size_t pitch = (size_t)p.in_w * 3;
    const uint8_t *src = p.src + (size_t)page * p.in_h * pitch;
    uint8_t *dst = p.dst + (size_t)page * p.out_h * p.out_w * 3;

    const int ox_last = min(ox0 + TOW, p.out_w) - 1;
    const int xs = p.bx[ox0 * 2];
    const int xe = p.bx[ox_last * 2] + p.bx[ox_last * 2 + 1];
    const int xs16 = xs & ~15;
    const int ngroups = (xe - xs16 + 15) >> 4;  // 16-pixel groups staged per row
    const int ys = p.by[oy0 * 2];
    const int ye = p.by[(oy1 - 1) * 2] + p.by[(oy1 - 1) * 2 + 1];

    // this warp's 8-column tile: B fragments (3 byte planes x KSTEPS x 2 registers) and the start of its K window
    const int tile = blockIdx.x * (TOW / 8) + warp;
    const bool tile_ok = tile * 8 < p.out_w;
    uint32_t bf[KSTEPS][3][2];
    int kb = 0;
    if (tile_ok) {
        kb = p.kb[tile];
#pragma unroll
        for (int st = 0; st < KSTEPS; st++)
#pragma unroll
            for (int pl = 0; pl < 3; pl++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++) bf[st][pl][hf] = p.bfrag[((((size_t)tile * KSTEPS + st) * 3 + pl) * 2 + hf) * 32 + lane];
    }
    const int grp = lane >> 2, tq = lane & 3;
    const uint8_t *src_end = p.src + p.src_total;
    int next_oy = oy0;
    __shared__ int s_oy_end;

    for (int r0 = ys; r0 < ye; r0 += RB) {
        const int nrows = min(RB, ye - r0);
        // ---- stage + de-interleave: a thread moves 16 pixels (48 B in, 3 x 16 B out) ----
        for (int t = tid; t < nrows * ngroups; t += 256) {
            const int r = t / ngroups, g = t - r * ngroups;
            const uint8_t *gp = src + (size_t)(r0 + r) * pitch + (size_t)(xs16 + g * 16) * 3;
            uint32_t w[12];
            if (gp + 48 <= src_end) {
                const uint4 a = ldg_stream_u4(gp), b = ldg_stream_u4(gp + 16), c = ldg_stream_u4(gp + 32);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
                w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    uint32_t v = 0;
                    for (int b = 0; b < 4; b++)
                        if (gp + i * 4 + b < src_end) v |= (uint32_t)gp[i * 4 + b] << (8 * b);
                    w[i] = v;
                }
            }
            uint32_t R[4], G[4], B[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
                R[q] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
                G[q] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
                B[q] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
            }
            uint8_t *sr = inbuf + (size_t)r * 3 * p.segpx + g * 16;
            *reinterpret_cast<uint4 *>(sr) = make_uint4(R[0], R[1], R[2], R[3]);
            *reinterpret_cast<uint4 *>(sr + p.segpx) = make_uint4(G[0], G[1], G[2], G[3]);
            *reinterpret_cast<uint4 *>(sr + 2 * p.segpx) = make_uint4(B[0], B[1], B[2], B[3]);
        }
        __syncthreads();
        // ---- horizontal pass: 16 rows x 8 columns x K per warp on the tensor cores ----
        if (tile_ok) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                int acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0}, acc2[4] = {0, 0, 0, 0};
                const uint8_t *r_lo = inbuf + ((size_t)grp * 3 + ch) * p.segpx + kb + tq * 4;
                const uint8_t *r_hi = inbuf + ((size_t)(grp + 8) * 3 + ch) * p.segpx + kb + tq * 4;
#pragma unroll
                for (int st = 0; st < KSTEPS; st++) {
                    uint32_t af[4];
                    af[0] = *reinterpret_cast<const uint32_t *>(r_lo + st * 32);
                    af[1] = *reinterpret_cast<const uint32_t *>(r_hi + st * 32);
                    af[2] = *reinterpret_cast<const uint32_t *>(r_lo + st * 32 + 16);
                    af[3] = *reinterpret_cast<const uint32_t *>(r_hi + st * 32 + 16);
                    mma_u8u8(acc0, af, bf[st][0][0], bf[st][0][1]);
                    mma_u8u8(acc1, af, bf[st][1][0], bf[st][1][1]);
                    mma_u8s8(acc2, af, bf[st][2][0], bf[st][2][1]);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int row = grp + (i >> 1) * 8;            // row of the chunk
                    const int col = warp * 8 + tq * 2 + (i & 1);   // column of the strip
                    if (row < nrows && ox0 + col < p.out_w) {
                        const int arow = r0 + row;
                        // exact modulo 2^32; the true sum fits int32
                        const int acc = (int)((1u << (PREC_BITS - 1)) + (uint32_t)acc0[i] + ((uint32_t)acc1[i] << 8) + ((uint32_t)acc2[i] << 16));
                        uint32_t *rg = ring + (size_t)((arow >> 2) & (RINGG - 1)) * ROWB + col * 3;
                        reinterpret_cast<uint8_t *>(rg + ch)[arow & 3] = clip8(acc);
                    }
                }
            }
        }
        if (tid == 255) {  // which output rows have their whole tap window in the ring after this chunk
            const int rows_done = r0 + nrows;
            int oe = next_oy;
            while (oe < oy1 && p.by[oe * 2] + p.by[oe * 2 + 1] <= rows_done) oe++;
            s_oy_end = oe;
        }
        __syncthreads();
        // ---- vertical pass: a thread owns 4 adjacent byte columns of one output row (one 128-bit ring
        // load per 4 taps x 4 columns, the row's coefficient words shared by the 4 columns) ----
        const int oy_end = s_oy_end;
        constexpr int COL4 = ROWB / 4;
        const int ntask = (oy_end - next_oy) * COL4;
        const int row_bytes = p.out_w * 3;
        for (int task = tid; task < ntask; task += 256) {
            const int orow = task / COL4, c4 = task - orow * COL4;
            const int oy = next_oy + orow;
            const int bcol = ox0 * 3 + c4 * 4;
            if (bcol >= row_bytes) continue;
            const int ymin = p.by[oy * 2];
            const uint32_t *k = p.cyp + (size_t)oy * 3 * p.kyw;
            int s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
            int rgrp = (ymin >> 2) & (RINGG - 1);
            for (int j = 0; j < p.kyw; j++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(ring + (size_t)rgrp * ROWB + c4 * 4);
                const uint32_t k0 = __ldg(k + j), k1 = __ldg(k + p.kyw + j), k2 = __ldg(k + 2 * p.kyw + j);
                const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    s0[q] = (int)__dp4a(vv[q], k0, (uint32_t)s0[q]);
                    s1[q] = (int)__dp4a(vv[q], k1, (uint32_t)s1[q]);
                    s2[q] = dp4a_u8s8(vv[q], k2, s2[q]);
                }
                rgrp = (rgrp + 1) & (RINGG - 1);
            }
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                o[q] = clip8((int)((1u << (PREC_BITS - 1)) + (uint32_t)s0[q] + ((uint32_t)s1[q] << 8) + ((uint32_t)s2[q] << 16)));
            uint8_t *dp = dst + (size_t)oy * row_bytes + bcol;
            if (bcol + 4 <= row_bytes && (((uintptr_t)dp) & 3) == 0) {
                *reinterpret_cast<uint32_t *>(dp) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
            } else if (bcol + 4 <= row_bytes && (((uintptr_t)dp) & 1) == 0) {
                *reinterpret_cast<uint16_t *>(dp) = (uint16_t)(o[0] | (o[1] << 8));
                *reinterpret_cast<uint16_t *>(dp + 2) = (uint16_t)(o[2] | (o[3] << 8));
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (bcol + q < row_bytes) dp[q] = (uint8_t)o[q];
            }
        }
        next_oy = oy_end;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Bulk-staged variant (the shipped one).  What limited the kernel above was not arithmetic but the serial phases of a
// chunk: every CTA waited for its own staging loads (23 % of the stall samples), de-interleaved them through shared
// memory, and only then started the MMAs.  Here one thread fetches the INTERLEAVED rows of chunk i+2 with
// cp.async.bulk (global -> shared, completion on an mbarrier) while chunks i and i+1 are computed, and nobody
// de-interleaves into shared memory at all: a lane reads the three words that hold its four pixels and PRMTs them
// into the R, G and B fragment words directly (row pitch == 48 mod 128 bytes keeps the 32 lanes on 32 banks).  The
// three coefficient byte planes share one accumulator: c = ((c2 << 8) + c1 << 8) + c0 modulo 2^32, the MMAs of plane
// p accumulating on top of the shifted sum of the planes above -- the true sum fits int32, so this is exact.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RS_DONE;\n"
        "bra RS_WAIT;\n"
        "RS_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int RINGV = 32;    // ring row-groups of the bulk kernel: the 64-row K window of a tile + the chunk being written
constexpr uint32_t kRound = 1u << (PREC_BITS - 1);   // 0.5 in fixed point, added with the second byte shift of the plane sum
constexpr int ROWP = TOW * 3 + 8;  // ring pitch in words (== 8 mod 32: the vertical fragments' 4 groups x 8 columns hit 32 banks)

template <int KSTEPS, int VSTEPS, bool TMA>
__global__ void __launch_bounds__(256, 4) resize_strip_bulk_kernel(const ResizeDp4aParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int ROWB = TOW * 3;
    constexpr int STAGES = 2;
    uint8_t *raw = smem;                                                                  // [STAGES][RB][rawpitch]
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem + (size_t)STAGES * RB * p.rawpitch);  // [RINGV][ROWP] words of 4 rows
    __shared__ __align__(8) unsigned long long s_full[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ox0 = blockIdx.x * TOW;
    const int oy0 = blockIdx.y * p.rows_per_seg;      // multiple of 8: vertical tiles are global
    const int oy1 = min(oy0 + p.rows_per_seg, p.out_h);
    const int page = blockIdx.z;
    const size_t pitch = (size_t)p.in_w * 3;
    const uint8_t *src = p.src + (size_t)page * p.in_h * pitch;
    uint8_t *dst = p.dst + (size_t)page * p.out_h * p.out_w * 3;
    const int xs16 = p.bx[ox0 * 2] & ~15;
    const int ys = p.by[oy0 * 2];
    const int ye = p.by[(oy1 - 1) * 2] + p.by[(oy1 - 1) * 2 + 1];
    const int nchunks = (ye - ys + RB - 1) / RB;
    const uint8_t *src_end = p.src + p.src_total;

    // TMA == true: one thread fetches a whole chunk (16 rows x rawpitch bytes of the batch seen as a 2-D tensor of 32-bit
    // words, row pitch = page row pitch) with a single cp.async.bulk.tensor; words past a row end or past the batch end
    // are zero-filled by the copy engine.  TMA == false (rows wider than a 256-element box): every warp issues the bulk
    // copies of two rows (a copy instruction is warp-uniform, so n rows issued by one warp are n serial iterations) and
    // arrives on the chunk's barrier with its byte count; the last rows of the batch are cut at its end.  What stays
    // stale, zero or foreign in shared memory only meets zero coefficients.
    auto issue = [&](int c) {
        const int buf = c % STAGES;
        const int r0 = ys + c * RB, nrows = min(RB, ye - r0);
        const uint32_t bar = rs_smem_u32(&s_full[buf]);
        if (TMA) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(RB * p.rawpitch)) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(rs_smem_u32(raw + (size_t)buf * RB * p.rawpitch)), "l"(&tmap), "r"((xs16 * 3) >> 2),
                               "r"(page * p.in_h + r0), "r"(bar) : "memory");
            }
            return;
        }
        const int row = warp * 2 + (lane & 1);
        const uint8_t *gp = src + (size_t)(r0 + row) * pitch + (size_t)xs16 * 3;
        const long long left = src_end - gp;
        const uint32_t bytes = row < nrows ? (uint32_t)(left < p.copy_bytes ? left : p.copy_bytes) : 0u;
        const uint32_t total = bytes + __shfl_xor_sync(0xffffffffu, bytes, 1);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
        if (lane < 2 && bytes) {
            const uint32_t d = rs_smem_u32(raw + ((size_t)buf * RB + row) * p.rawpitch);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(d), "l"(gp), "r"(bytes), "r"(bar) : "memory");
        }
    };
    if (tid == 0) {
        for (int i = 0; i < STAGES; i++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rs_smem_u32(&s_full[i])), "r"(TMA ? 1 : 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (TMA) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    __syncthreads();
    for (int c = 0; c < STAGES && c < nchunks; c++) issue(c);

    // this warp's 8-column tile: B fragments (3 byte planes x KSTEPS x 2 registers) and the start of its K window
    const int tile = blockIdx.x * (TOW / 8) + warp;
    const bool tile_ok = tile * 8 < p.out_w;
    uint32_t bf[KSTEPS][3][2];
    int kb = 0;
    if (tile_ok) {
        kb = p.kb[tile];
#pragma unroll
        for (int st = 0; st < KSTEPS; st++)
#pragma unroll
            for (int pl = 0; pl < 3; pl++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++) bf[st][pl][hf] = p.bfrag[((((size_t)tile * KSTEPS + st) * 3 + pl) * 2 + hf) * 32 + lane];
    }
    const int grp = lane >> 2, tq = lane & 3;
    const int rawpitch_w = p.rawpitch >> 2;
    const int row_bytes = p.out_w * 3;
    int next_oy = oy0;           // first output row not yet written (a multiple of 8 until the segment's last tile)
    int2 vmeta = __ldg(reinterpret_cast<const int2 *>(p.vg0) + (oy0 >> 3));   // {first ring row-group, last tap row + 1} of the pending tile

    for (int c = 0; c < nchunks; c++) {
        const int r0 = ys + c * RB;
        const int nrows = min(RB, ye - r0);
        const int buf = c % STAGES;
        rs_mbar_wait(rs_smem_u32(&s_full[buf]), (uint32_t)((c / STAGES) & 1));
        // ---- horizontal pass: 16 rows x 8 columns x K per warp on the tensor cores, A straight from the interleaved rows ----
        if (tile_ok) {
            const uint32_t *lo = reinterpret_cast<const uint32_t *>(raw + ((size_t)buf * RB + grp) * p.rawpitch) + 3 * ((kb >> 2) + tq);
            const uint32_t *hi = lo + 8 * rawpitch_w;
            uint32_t af[3][KSTEPS][4];
#pragma unroll
            for (int st = 0; st < KSTEPS; st++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
#pragma unroll
                    for (int up = 0; up < 2; up++) {
                        const uint32_t *q = (up ? hi : lo) + 3 * (st * 8 + hf * 4);
                        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                        af[0][st][hf * 2 + up] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
                        af[1][st][hf * 2 + up] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
                        af[2][st][hf * 2 + up] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
                    }
            // ring byte addresses of this lane's two rows (columns / channels are immediate offsets from them)
            const int arow_lo = r0 + grp, arow_hi = arow_lo + 8;
            uint8_t *ring_b = reinterpret_cast<uint8_t *>(ring);
            uint8_t *rp_lo = ring_b + ((size_t)((arow_lo >> 2) & (RINGV - 1)) * ROWP + (warp * 8 + tq * 2) * 3) * 4 + (arow_lo & 3);
            uint8_t *rp_hi = ring_b + ((size_t)((arow_hi >> 2) & (RINGV - 1)) * ROWP + (warp * 8 + tq * 2) * 3) * 4 + (arow_hi & 3);
            // rows past the last chunk's end are computed from stale staging and stored too: their ring slots belong to
            // rows beyond this segment's last tap, which no tile reads with a non-zero coefficient
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                int acc[4];
                mma_u8s8_z(acc, af[ch][0], bf[0][2][0], bf[0][2][1]);
#pragma unroll
                for (int st = 1; st < KSTEPS; st++) mma_u8s8(acc, af[ch][st], bf[st][2][0], bf[st][2][1]);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = (int)((uint32_t)acc[i] << 8);
#pragma unroll
                for (int st = 0; st < KSTEPS; st++) mma_u8u8(acc, af[ch][st], bf[st][1][0], bf[st][1][1]);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = (int)(((uint32_t)acc[i] << 8) + kRound);   // the rounding term rides on the shift
#pragma unroll
                for (int st = 0; st < KSTEPS; st++) mma_u8u8(acc, af[ch][st], bf[st][0][0], bf[st][0][1]);
                // columns past out_w (last strip) land in ring columns nobody stores from
#pragma unroll
                for (int i = 0; i < 4; i++) ((i >> 1) ? rp_hi : rp_lo)[((i & 1) * 3 + ch) * 4] = (uint8_t)clip8r(acc[i]);   // exact modulo 2^32; the true sum fits int32
            }
        }
        __syncthreads();   // the ring is complete, and nobody reads this chunk's staged rows any more
        if (c + STAGES < nchunks) issue(c + STAGES);
        // ---- vertical pass on the tensor cores: tiles of 8 output rows (N) x 16 byte columns (M) x 32 * VSTEPS ring rows (K).
        // A = the ring words themselves (4 consecutive rows of one byte column), B = the tile's coefficient fragments.
        // No barrier after it: the next chunk's horizontal pass writes ring rows this pass reads only against zero
        // coefficients (the ring holds 128 rows, a tile's taps lie within the last 80).
        const int rows_done = r0 + nrows;
        while (next_oy < oy1 && vmeta.y <= rows_done) {       // the tile's last tap row is in the ring
            const int vt = next_oy >> 3;
            const int g0 = vmeta.x;
            uint32_t vb[VSTEPS][3][2];
#pragma unroll
            for (int st = 0; st < VSTEPS; st++)
#pragma unroll
                for (int pl = 0; pl < 3; pl++)
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) vb[st][pl][hf] = __ldg(p.vfrag + ((((size_t)vt * 2 + st) * 3 + pl) * 2 + hf) * 32 + lane);
            const int col0 = ox0 * 3;
            uint8_t *dtile = dst + (size_t)(next_oy + tq * 2) * row_bytes + col0 + grp;
            const bool full_rows = next_oy + 8 <= oy1;
            // 12 column tiles over 8 warps: the four extra ones alternate between the warp halves from tile to tile
            const int extra = (((warp >> 2) ^ vt) & 1) ? -1 : 8 + (warp & 3);
#pragma unroll 1
            for (int rep = 0; rep < 2; rep++) {
                const int mt = rep ? extra : warp;
                if (mt < 0) break;
                const int mb = mt * 16;
                uint32_t va[VSTEPS][4];
#pragma unroll
                for (int st = 0; st < VSTEPS; st++) {
                    const uint32_t *g_lo = ring + (size_t)((g0 + st * 8 + tq) & (RINGV - 1)) * ROWP + mb + grp;
                    const uint32_t *g_hi = ring + (size_t)((g0 + st * 8 + 4 + tq) & (RINGV - 1)) * ROWP + mb + grp;
                    va[st][0] = g_lo[0]; va[st][1] = g_lo[8]; va[st][2] = g_hi[0]; va[st][3] = g_hi[8];
                }
                int acc[4];
                mma_u8s8_z(acc, va[0], vb[0][2][0], vb[0][2][1]);
#pragma unroll
                for (int st = 1; st < VSTEPS; st++) mma_u8s8(acc, va[st], vb[st][2][0], vb[st][2][1]);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = (int)((uint32_t)acc[i] << 8);
#pragma unroll
                for (int st = 0; st < VSTEPS; st++) mma_u8u8(acc, va[st], vb[st][1][0], vb[st][1][1]);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = (int)(((uint32_t)acc[i] << 8) + kRound);
#pragma unroll
                for (int st = 0; st < VSTEPS; st++) mma_u8u8(acc, va[st], vb[st][0][0], vb[st][0][1]);
                // lane (grp, tq) holds byte columns grp, grp + 8 of output rows 2 tq, 2 tq + 1: one address per column tile
                uint8_t *o = dtile + mb;
                if (full_rows && col0 + mb + 16 <= row_bytes) {
                    o[0] = (uint8_t)clip8r(acc[0]); o[row_bytes] = (uint8_t)clip8r(acc[1]);
                    o[8] = (uint8_t)clip8r(acc[2]); o[row_bytes + 8] = (uint8_t)clip8r(acc[3]);
                } else {
                    const bool c0 = col0 + mb + grp < row_bytes, c1 = col0 + mb + grp + 8 < row_bytes;
                    const bool r0ok = next_oy + tq * 2 < oy1, r1ok = next_oy + tq * 2 + 1 < oy1;
                    if (c0 && r0ok) o[0] = (uint8_t)clip8r(acc[0]);
                    if (c0 && r1ok) o[row_bytes] = (uint8_t)clip8r(acc[1]);
                    if (c1 && r0ok) o[8] = (uint8_t)clip8r(acc[2]);
                    if (c1 && r1ok) o[row_bytes + 8] = (uint8_t)clip8r(acc[3]);
                }
            }
            next_oy = min(next_oy + 8, oy1);
            if (next_oy < oy1) {
                vmeta = __ldg(reinterpret_cast<const int2 *>(p.vg0) + (next_oy >> 3));
                // warm L1 with the next tile's fragments (12 lines of 128 bytes): they are needed a chunk or two from now
                if (lane < 12) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.vfrag + ((size_t)(next_oy >> 3) * 12 + lane) * 32));
            }
        }
    }
}

// one axis only (the other is identity) or tap counts beyond the unrolled
// variants: plain two-pass kernels through an HBM intermediate.
template <int C>
__global__ void __launch_bounds__(256) resize_h_generic_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                               const int32_t *__restrict__ bx, const int32_t *__restrict__ cx,
                                                               int kx, int rows_total, int in_w, int out_w) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (ox >= out_w || row >= rows_total) return;
    const uint8_t *s = src + (size_t)row * in_w * C;
    const int xmin = bx[ox * 2], n = bx[ox * 2 + 1];
    int acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) acc[ch] = 1 << (PREC_BITS - 1);
    for (int t = 0; t < n; t++) {
        const int kk = cx[(size_t)ox * kx + t];
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[ch] += (int)__ldg(s + (size_t)(xmin + t) * C + ch) * kk;
    }
#pragma unroll
    for (int ch = 0; ch < C; ch++) dst[((size_t)row * out_w + ox) * C + ch] = clip8(acc[ch]);
}

__global__ void __launch_bounds__(256) resize_v_generic_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                               const int32_t *__restrict__ by, const int32_t *__restrict__ cy,
                                                               int ky, int in_h, int out_h, int row_bytes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y, page = blockIdx.z;
    if (i >= row_bytes) return;
    const uint8_t *s = src + (size_t)page * in_h * row_bytes;
    const int ymin = by[oy * 2], n = by[oy * 2 + 1];
    int acc = 1 << (PREC_BITS - 1);
    for (int j = 0; j < n; j++) acc += (int)__ldg(s + (size_t)(ymin + j) * row_bytes + i) * cy[(size_t)oy * ky + j];
    dst[((size_t)page * out_h + oy) * row_bytes + i] = clip8(acc);
}

}  // namespace lumina

using namespace lumina;

LUMINA_API int lumina_resize_plan_create(int in_h, int in_w, int out_h, int out_w, lumina_resize_plan **plan) {
    LUMINA_REQUIRE(plan != nullptr, "null plan pointer");
    LUMINA_REQUIRE(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "sizes must be positive");
    std::vector<int32_t> bx, cx, by, cy;
    host_coeffs(in_w, out_w, bx, cx);
    host_coeffs(in_h, out_h, by, cy);
    lumina_resize_plan *pl = new lumina_resize_plan();
    pl->in_h = in_h; pl->in_w = in_w; pl->out_h = out_h; pl->out_w = out_w;
    pl->kx = ksize_for(in_w, out_w); pl->ky = ksize_for(in_h, out_h);
    pl->d_bx = pl->d_cx = pl->d_by = pl->d_cy = nullptr;
    pl->d_cxp = pl->d_cyp = nullptr;
    pl->kxw = (pl->kx + 3) / 4;
    pl->kyw = (pl->ky + 3 + 3) / 4;  // a window may start at byte 1..3 of its first row-group
    std::vector<uint32_t> cxp((size_t)out_w * 3 * pl->kxw, 0u), cyp((size_t)out_h * 3 * pl->kyw, 0u);
    auto put = [](std::vector<uint32_t> &tab, size_t base, int words, int pos, int32_t c) {
        // c = c0 + 256*c1 + 65536*c2 with c0,c1 in [0,255], c2 = c >> 16 (arithmetic)
        const uint32_t b[3] = {(uint32_t)c & 0xffu, ((uint32_t)c >> 8) & 0xffu, (uint32_t)(c >> 16) & 0xffu};
        for (int plane = 0; plane < 3; plane++) tab[base + (size_t)plane * words + pos / 4] |= b[plane] << (8 * (pos % 4));
    };
    for (int x = 0; x < out_w; x++)
        for (int t = 0; t < bx[x * 2 + 1]; t++) put(cxp, (size_t)x * 3 * pl->kxw, pl->kxw, t, cx[(size_t)x * pl->kx + t]);
    for (int y = 0; y < out_h; y++)
        for (int t = 0; t < by[y * 2 + 1]; t++)
            put(cyp, (size_t)y * 3 * pl->kyw, pl->kyw, (by[y * 2] & 3) + t, cy[(size_t)y * pl->ky + t]);
    int msp = 0;
    for (int ox0 = 0; ox0 < out_w; ox0 += TOW) {
        int last = (ox0 + TOW < out_w ? ox0 + TOW : out_w) - 1;
        int span = bx[last * 2] + bx[last * 2 + 1] - bx[ox0 * 2];
        if (span > msp) msp = span;
    }
    pl->max_seg_px = msp;
    // tensor-core path: per 8-column tile the K window [kb, kb + 32 * ksteps) of staged plane bytes and the B
    // fragments of mma.m16n8k32 (lane l holds column l >> 2, rows (l & 3) * 4 .. + 3 and the same + 16)
    pl->d_bfrag = nullptr; pl->d_kb = nullptr; pl->ksteps = 0; pl->imma_span = 0;
    std::vector<uint32_t> bfrag;
    std::vector<int32_t> kbv;
    {
        const int tiles = (out_w + 7) / 8;
        kbv.assign(tiles, 0);
        int maxk = 0, span = 0;
        for (int t = 0; t < tiles; t++) {
            const int strip0 = (t * 8 / TOW) * TOW;
            const int xs16 = bx[strip0 * 2] & ~15;
            const int kb = (bx[t * 8 * 2] - xs16) & ~3;
            kbv[t] = kb;
            for (int n = 0; n < 8 && t * 8 + n < out_w; n++) {
                const int ox = t * 8 + n;
                const int need = bx[ox * 2] - xs16 - kb + bx[ox * 2 + 1];
                maxk = need > maxk ? need : maxk;
            }
        }
        const int ks = (maxk + 31) / 32;
        if (ks >= 1 && ks <= 3) {
            pl->ksteps = ks;
            bfrag.assign((size_t)tiles * ks * 3 * 2 * 32, 0u);
            for (int t = 0; t < tiles; t++) {
                const int strip0 = (t * 8 / TOW) * TOW;
                const int xs16 = bx[strip0 * 2] & ~15;
                if (kbv[t] + ks * 32 > span) span = kbv[t] + ks * 32;
                for (int lane = 0; lane < 32; lane++) {
                    const int ox = t * 8 + (lane >> 2);
                    if (ox >= out_w) continue;
                    const int off = bx[ox * 2] - xs16 - kbv[t], ntap = bx[ox * 2 + 1];
                    for (int st = 0; st < ks; st++)
                        for (int hf = 0; hf < 2; hf++)
                            for (int e = 0; e < 4; e++) {
                                const int kk = st * 32 + hf * 16 + (lane & 3) * 4 + e, tap = kk - off;
                                if (tap < 0 || tap >= ntap) continue;
                                const int32_t c = cx[(size_t)ox * pl->kx + tap];
                                const uint32_t bb[3] = {(uint32_t)c & 0xffu, ((uint32_t)c >> 8) & 0xffu, (uint32_t)(c >> 16) & 0xffu};
                                for (int plane = 0; plane < 3; plane++)
                                    bfrag[((((size_t)t * ks + st) * 3 + plane) * 2 + hf) * 32 + lane] |= bb[plane] << (8 * e);
                            }
                }
            }
            pl->imma_span = span;
        }
    }
    // tensor-core vertical pass: tiles of 8 output rows; K = 64 intermediate rows from the 4-row group that holds the
    // tile's first tap (B fragment of mma.m16n8k32: lane l holds output row l >> 2, K rows (l & 3) * 4 .. + 3 and + 16)
    pl->d_vfrag = nullptr; pl->d_vg0 = nullptr; pl->vtiles = 0; pl->vsteps = 2;
    std::vector<uint32_t> vfrag;
    std::vector<int32_t> vg0;
    {
        const int tiles = (out_h + 7) / 8;
        vg0.assign((size_t)tiles * 2, 0);
        bool fits = true, fits32 = true;
        for (int t = 0; t < tiles; t++) {
            const int g0 = by[t * 8 * 2] >> 2;
            const int last = (t * 8 + 7 < out_h ? t * 8 + 7 : out_h - 1);
            vg0[t * 2] = g0;
            vg0[t * 2 + 1] = by[last * 2] + by[last * 2 + 1];   // the tile is complete once this many rows are in the ring
            for (int r = 0; r < 8 && t * 8 + r < out_h; r++) {
                const int oy = t * 8 + r;
                if (by[oy * 2] < 4 * g0 || by[oy * 2] + by[oy * 2 + 1] > 4 * g0 + 64) fits = false;
                if (by[oy * 2] + by[oy * 2 + 1] > 4 * g0 + 32) fits32 = false;
            }
        }
        if (fits) {
            pl->vtiles = tiles;
            pl->vsteps = fits32 ? 1 : 2;   // scales below ~2.2: the second K step would only meet zero coefficients
            vfrag.assign((size_t)tiles * 2 * 3 * 2 * 32, 0u);
            for (int t = 0; t < tiles; t++)
                for (int lane = 0; lane < 32; lane++) {
                    const int oy = t * 8 + (lane >> 2);
                    if (oy >= out_h) continue;
                    const int ymin = by[oy * 2], ntap = by[oy * 2 + 1];
                    for (int st = 0; st < 2; st++)
                        for (int hf = 0; hf < 2; hf++)
                            for (int e = 0; e < 4; e++) {
                                const int row = 4 * vg0[t * 2] + st * 32 + hf * 16 + (lane & 3) * 4 + e, tap = row - ymin;
                                if (tap < 0 || tap >= ntap) continue;
                                const int32_t c = cy[(size_t)oy * pl->ky + tap];
                                const uint32_t bb[3] = {(uint32_t)c & 0xffu, ((uint32_t)c >> 8) & 0xffu, (uint32_t)(c >> 16) & 0xffu};
                                for (int plane = 0; plane < 3; plane++)
                                    vfrag[((((size_t)t * 2 + st) * 3 + plane) * 2 + hf) * 32 + lane] |= bb[plane] << (8 * e);
                            }
                }
        }
    }
    cudaGetDevice(&pl->device);
    auto up = [](int32_t **d, const std::vector<int32_t> &h) -> cudaError_t {
        cudaError_t e = cudaMalloc((void **)d, h.size() * sizeof(int32_t));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    };
    auto upu = [](uint32_t **d, const std::vector<uint32_t> &h) -> cudaError_t {
        cudaError_t e2 = cudaMalloc((void **)d, h.size() * sizeof(uint32_t));
        if (e2 != cudaSuccess) return e2;
        return cudaMemcpy(*d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    };
    cudaError_t e;
    if ((e = up(&pl->d_bx, bx)) != cudaSuccess || (e = up(&pl->d_cx, cx)) != cudaSuccess ||
        (e = up(&pl->d_by, by)) != cudaSuccess || (e = up(&pl->d_cy, cy)) != cudaSuccess ||
        (e = upu(&pl->d_cxp, cxp)) != cudaSuccess || (e = upu(&pl->d_cyp, cyp)) != cudaSuccess ||
        (pl->ksteps && ((e = upu(&pl->d_bfrag, bfrag)) != cudaSuccess || (e = up(&pl->d_kb, kbv)) != cudaSuccess)) ||
        (pl->vtiles && ((e = upu(&pl->d_vfrag, vfrag)) != cudaSuccess || (e = up(&pl->d_vg0, vg0)) != cudaSuccess))) {
        lumina_resize_plan_destroy(pl);
        return set_error(LUMINA_E_CUDA, "resize plan upload failed: %s", cudaGetErrorString(e));
    }
    *plan = pl;
    return LUMINA_OK;
}

LUMINA_API void lumina_resize_plan_destroy(lumina_resize_plan *pl) {
    if (!pl) return;
    cudaFree(pl->d_bx); cudaFree(pl->d_cx); cudaFree(pl->d_by); cudaFree(pl->d_cy);
    cudaFree(pl->d_cxp); cudaFree(pl->d_cyp);
    cudaFree(pl->d_bfrag); cudaFree(pl->d_kb);
    cudaFree(pl->d_vfrag); cudaFree(pl->d_vg0);
    delete pl;
}

// Pillow (12.2.0) runs the VERTICAL pass first on very tall images that shrink vertically: both passes needed,
// in_h > 100 * in_w and out_h < in_h (rule and evidence: oracle/lumina_oracle.c orc_vertical_first).  The uint8
// intermediate makes the order visible, so such geometries (strips a few pixels wide) take the generic kernels in that order.
static bool vertical_first(const lumina_resize_plan *pl) {
    return pl->in_w != pl->out_w && pl->in_h != pl->out_h && (long long)pl->in_h > 100LL * pl->in_w && pl->out_h < pl->in_h;
}

static bool fused_ok(const lumina_resize_plan *pl) {
    return pl->kx <= 32 && pl->ky + RB <= RING && pl->in_w != pl->out_w && pl->in_h != pl->out_h && !vertical_first(pl);
}

LUMINA_API size_t lumina_resize_workspace_bytes(const lumina_resize_plan *pl, int n, int c) {
    if (!pl || fused_ok(pl)) return 0;
    if (vertical_first(pl)) return (size_t)n * pl->out_h * pl->in_w * c;
    return (size_t)n * pl->in_h * pl->out_w * c;  // HBM intermediate of the generic two-pass path
}

template <int C, int KX>
static int launch_strip(const lumina_resize_plan *pl, const uint8_t *src, uint8_t *dst, int n, cudaStream_t st) {
    ResizeParams p;
    p.src = src; p.dst = dst;
    p.bx = pl->d_bx; p.cx = pl->d_cx; p.by = pl->d_by; p.cy = pl->d_cy;
    p.in_h = pl->in_h; p.in_w = pl->in_w; p.out_h = pl->out_h; p.out_w = pl->out_w;
    p.kx = pl->kx; p.ky = pl->ky;
    p.src_total = (size_t)n * pl->in_h * pl->in_w * C;
    // bytes per staged row: span + KX taps of slack (zero coefficients) + 16 B misalignment, 16-B multiple
    p.segb = (((pl->max_seg_px + KX) * C + 16 + 4) + 15) & ~15;
    const int strips = div_up(pl->out_w, TOW);
    // choose the row-segment count so the grid covers the 148 SMs several times
    int segs = 1;
    while ((long long)strips * segs * n < 4LL * kNumSMs * 4 && pl->out_h / (segs * 2) >= 64) segs *= 2;
    p.rows_per_seg = div_up(pl->out_h, segs);
    segs = div_up(pl->out_h, p.rows_per_seg);
    size_t smem = (size_t)RB * p.segb + (size_t)RING * TOW * C;
    auto kern = resize_strip_kernel<C, KX>;
    if (smem > 48 * 1024) LUMINA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LUMINA_REQUIRE(n <= 65535 && segs <= 65535, "batch too large for grid");
    kern<<<dim3(strips, segs, n), 256, smem, st>>>(p);
    LUMINA_KERNEL_CHECK("resize_strip_kernel");
    return LUMINA_OK;
}


template <int KXW>
static int launch_strip_dp4a(const lumina_resize_plan *pl, const uint8_t *src, uint8_t *dst, int n, cudaStream_t st) {
    ResizeDp4aParams p;
    p.src = src; p.dst = dst; p.bx = pl->d_bx; p.by = pl->d_by; p.cxp = pl->d_cxp; p.cyp = pl->d_cyp;
    p.in_h = pl->in_h; p.in_w = pl->in_w; p.out_h = pl->out_h; p.out_w = pl->out_w; p.kxw = pl->kxw; p.kyw = pl->kyw;
    p.src_total = (size_t)n * pl->in_h * pl->in_w * 3;
    // staged pixels per plane row: strip span rounded out to 16-pixel groups + the words a lane may read past its taps
    p.segpx = ((pl->max_seg_px + 15 + 16 + KXW * 4 + 8) + 15) & ~15;
    const int strips = div_up(pl->out_w, TOW);
    int segs = 1;
    while ((long long)strips * segs * n < 4LL * kNumSMs * 4 && pl->out_h / (segs * 2) >= 64) segs *= 2;
    p.rows_per_seg = div_up(pl->out_h, segs);
    segs = div_up(pl->out_h, p.rows_per_seg);
    const size_t smem = (size_t)RB * 3 * p.segpx + (size_t)RINGG * TOW * 3 * 4;
    auto kern = resize_strip_dp4a_kernel<KXW>;
    if (smem > 48 * 1024) LUMINA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LUMINA_REQUIRE(n <= 65535 && segs <= 65535, "batch too large for grid");
    kern<<<dim3(strips, segs, n), 256, smem, st>>>(p);
    LUMINA_KERNEL_CHECK("resize_strip_dp4a_kernel");
    return LUMINA_OK;
}

template <int KSTEPS, int VSTEPS>
static int launch_strip_bulk(const lumina_resize_plan *pl, const uint8_t *src, uint8_t *dst, int n, cudaStream_t st) {
    ResizeDp4aParams p;
    p.src = src; p.dst = dst; p.bx = pl->d_bx; p.by = pl->d_by; p.cxp = pl->d_cxp; p.cyp = pl->d_cyp;
    p.bfrag = pl->d_bfrag; p.kb = pl->d_kb;
    p.in_h = pl->in_h; p.in_w = pl->in_w; p.out_h = pl->out_h; p.out_w = pl->out_w; p.kxw = pl->kxw; p.kyw = pl->kyw;
    p.src_total = (size_t)n * pl->in_h * pl->in_w * 3;
    p.segpx = 0;
    p.vfrag = pl->d_vfrag; p.vg0 = pl->d_vg0;
    // a staged row holds every byte an A fragment reads: imma_span pixels from the strip's xs16
    p.copy_bytes = (pl->imma_span * 3 + 15) & ~15;
    p.rawpitch = p.copy_bytes;
    while (p.rawpitch % 128 != 48) p.rawpitch += 16;
    const int strips = div_up(pl->out_w, TOW);
    int segs = 1;
    while ((long long)strips * segs * n < 4LL * kNumSMs * 4 && pl->out_h / (segs * 2) >= 64) segs *= 2;
    p.rows_per_seg = (div_up(pl->out_h, segs) + 7) & ~7;   // vertical tiles of 8 output rows are global
    segs = div_up(pl->out_h, p.rows_per_seg);
    const size_t smem = (size_t)2 * RB * p.rawpitch + (size_t)RINGV * ROWP * 4;
    LUMINA_REQUIRE(n <= 65535 && segs <= 65535, "batch too large for grid");
    // the batch as a 2-D tensor of 32-bit words [n * in_h rows][in_w * 3 / 4], box = one staged chunk (16 rows x rawpitch bytes)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    bool tma = p.rawpitch <= 1024 && (pl->in_w * 3) % 16 == 0 && !getenv("LUMINA_RESIZE_NO_TMA");
    if (tma) {
        static PFN_cuTensorMapEncodeTiled encode = [] {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
            return (PFN_cuTensorMapEncodeTiled)fn;
        }();
        const cuuint64_t dims[2] = {(cuuint64_t)(pl->in_w * 3 / 4), (cuuint64_t)n * (cuuint64_t)pl->in_h};
        const cuuint64_t strides[1] = {(cuuint64_t)pl->in_w * 3};
        const cuuint32_t box[2] = {(cuuint32_t)(p.rawpitch / 4), (cuuint32_t)RB};
        const cuuint32_t estr[2] = {1, 1};
        tma = encode && encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(src), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (tma) {
        auto kern = resize_strip_bulk_kernel<KSTEPS, VSTEPS, true>;
        if (smem > 48 * 1024) LUMINA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(strips, segs, n), 256, smem, st>>>(p, tmap);
    } else {
        auto kern = resize_strip_bulk_kernel<KSTEPS, VSTEPS, false>;
        if (smem > 48 * 1024) LUMINA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(strips, segs, n), 256, smem, st>>>(p, tmap);
    }
    LUMINA_KERNEL_CHECK("resize_strip_bulk_kernel");
    return LUMINA_OK;
}

template <int KSTEPS>
static int launch_strip_imma(const lumina_resize_plan *pl, const uint8_t *src, uint8_t *dst, int n, cudaStream_t st) {
    ResizeDp4aParams p;
    p.src = src; p.dst = dst; p.bx = pl->d_bx; p.by = pl->d_by; p.cxp = pl->d_cxp; p.cyp = pl->d_cyp;
    p.bfrag = pl->d_bfrag; p.kb = pl->d_kb;
    p.in_h = pl->in_h; p.in_w = pl->in_w; p.out_h = pl->out_h; p.out_w = pl->out_w; p.kxw = pl->kxw; p.kyw = pl->kyw;
    p.src_total = (size_t)n * pl->in_h * pl->in_w * 3;
    // staged pixels per plane row: the strip's span rounded out to 16-pixel groups, and every byte an A fragment reads
    int seg = pl->max_seg_px + 15 + 16 + 8;
    if (pl->imma_span + 16 > seg) seg = pl->imma_span + 16;
    p.segpx = (seg + 15) & ~15;
    if ((p.segpx / 4) % 8 == 0) p.segpx += 16;   // plane rows 3 * segpx apart: keep the 8 fragment rows on distinct banks
    const int strips = div_up(pl->out_w, TOW);
    int segs = 1;
    while ((long long)strips * segs * n < 4LL * kNumSMs * 4 && pl->out_h / (segs * 2) >= 64) segs *= 2;
    p.rows_per_seg = div_up(pl->out_h, segs);
    segs = div_up(pl->out_h, p.rows_per_seg);
    const size_t smem = (size_t)RB * 3 * p.segpx + (size_t)RINGG * TOW * 3 * 4;
    auto kern = resize_strip_imma_kernel<KSTEPS>;
    if (smem > 48 * 1024) LUMINA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LUMINA_REQUIRE(n <= 65535 && segs <= 65535, "batch too large for grid");
    kern<<<dim3(strips, segs, n), 256, smem, st>>>(p);
    LUMINA_KERNEL_CHECK("resize_strip_imma_kernel");
    return LUMINA_OK;
}

static bool dp4a_ok(const lumina_resize_plan *pl, const uint8_t *src, int c) {
    // measured on B200 (64 A4 pages): dp4a 1.66 ms vs IMAD 2.45 ms at 23 taps (-> 960), 3.16 vs 4.27 ms at 13 taps (-> 2000)
    return c == 3 && (pl->in_w % 16) == 0 && (((uintptr_t)src) & 15) == 0 && pl->kxw <= 8 && pl->kx >= 9 &&
           (pl->ky + RB + 6) / 4 + 1 <= RINGG && !getenv("LUMINA_RESIZE_IMAD");
}

LUMINA_API int lumina_resize_lanczos_u8(const lumina_resize_plan *pl, const uint8_t *d_src, uint8_t *d_dst, int n, int c,
                                        void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(pl && d_src && d_dst, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0, "empty batch");
    cudaStream_t st = as_stream(stream);
    if (pl->in_w == pl->out_w && pl->in_h == pl->out_h) {
        LUMINA_CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, (size_t)n * pl->in_h * pl->in_w * c, cudaMemcpyDeviceToDevice, st));
        return LUMINA_OK;
    }
    if (fused_ok(pl) && dp4a_ok(pl, d_src, c) && pl->ksteps && !getenv("LUMINA_RESIZE_DP4A")) {
        // measured on B200 (64 A4 pages -> 678x960): tensor-core horizontal pass vs dp4a, see DESIGN.md
        if (pl->ksteps <= 2 && pl->vtiles && !getenv("LUMINA_RESIZE_STAGED")) {
            // measured on B200: bulk-staged interleaved rows vs staged + de-interleaved planes, see DESIGN.md
            if (pl->ksteps == 1) return pl->vsteps == 1 ? launch_strip_bulk<1, 1>(pl, d_src, d_dst, n, st) : launch_strip_bulk<1, 2>(pl, d_src, d_dst, n, st);
            return pl->vsteps == 1 ? launch_strip_bulk<2, 1>(pl, d_src, d_dst, n, st) : launch_strip_bulk<2, 2>(pl, d_src, d_dst, n, st);
        }
        if (pl->ksteps == 1) return launch_strip_imma<1>(pl, d_src, d_dst, n, st);
        if (pl->ksteps == 2) return launch_strip_imma<2>(pl, d_src, d_dst, n, st);
        return launch_strip_imma<3>(pl, d_src, d_dst, n, st);
    }
    if (fused_ok(pl) && dp4a_ok(pl, d_src, c)) {
        if (pl->kxw <= 2) return launch_strip_dp4a<2>(pl, d_src, d_dst, n, st);
        if (pl->kxw <= 4) return launch_strip_dp4a<4>(pl, d_src, d_dst, n, st);
        if (pl->kxw <= 6) return launch_strip_dp4a<6>(pl, d_src, d_dst, n, st);
        return launch_strip_dp4a<8>(pl, d_src, d_dst, n, st);
    }
    if (fused_ok(pl)) {
        if (c == 3) {
            if (pl->kx <= 8) return launch_strip<3, 8>(pl, d_src, d_dst, n, st);
            if (pl->kx <= 16) return launch_strip<3, 16>(pl, d_src, d_dst, n, st);
            if (pl->kx <= 24) return launch_strip<3, 24>(pl, d_src, d_dst, n, st);
            return launch_strip<3, 32>(pl, d_src, d_dst, n, st);
        } else {
            if (pl->kx <= 8) return launch_strip<1, 8>(pl, d_src, d_dst, n, st);
            if (pl->kx <= 16) return launch_strip<1, 16>(pl, d_src, d_dst, n, st);
            if (pl->kx <= 24) return launch_strip<1, 24>(pl, d_src, d_dst, n, st);
            return launch_strip<1, 32>(pl, d_src, d_dst, n, st);
        }
    }
    // generic two-pass path
    const size_t need = lumina_resize_workspace_bytes(pl, n, c);
    if (vertical_first(pl)) {
        if (!d_workspace || workspace_bytes < need)
            return set_error(LUMINA_E_NOMEM, "resize workspace too small: need %zu bytes", need);
        const int row_bytes = pl->in_w * c;
        LUMINA_REQUIRE(pl->out_h <= 65535 && n <= 65535, "image too tall for grid");
        resize_v_generic_kernel<<<dim3(div_up(row_bytes, 256), pl->out_h, n), 256, 0, st>>>(d_src, (uint8_t *)d_workspace, pl->d_by, pl->d_cy,
                                                                                          pl->ky, pl->in_h, pl->out_h, row_bytes);
        LUMINA_KERNEL_CHECK("resize_v_generic_kernel");
        const long long rows_v = (long long)n * pl->out_h;
        for (long long r0 = 0; r0 < rows_v; r0 += 65535) {
            const int rows = (int)((rows_v - r0 < 65535) ? rows_v - r0 : 65535);
            dim3 grid(div_up(pl->out_w, 256), rows);
            const uint8_t *s = (const uint8_t *)d_workspace + (size_t)r0 * pl->in_w * c;
            uint8_t *d = d_dst + (size_t)r0 * pl->out_w * c;
            if (c == 3) resize_h_generic_kernel<3><<<grid, 256, 0, st>>>(s, d, pl->d_bx, pl->d_cx, pl->kx, rows, pl->in_w, pl->out_w);
            else resize_h_generic_kernel<1><<<grid, 256, 0, st>>>(s, d, pl->d_bx, pl->d_cx, pl->kx, rows, pl->in_w, pl->out_w);
            LUMINA_KERNEL_CHECK("resize_h_generic_kernel");
        }
        return LUMINA_OK;
    }
    const uint8_t *hsrc = d_src;
    const long long rows_total = (long long)n * pl->in_h;
    LUMINA_REQUIRE(rows_total <= 65535LL * 32768LL, "batch too large");
    if (pl->in_w != pl->out_w) {
        uint8_t *hdst = (pl->in_h == pl->out_h) ? d_dst : (uint8_t *)d_workspace;
        if (pl->in_h != pl->out_h) {
            if (!d_workspace || workspace_bytes < need)
                return set_error(LUMINA_E_NOMEM, "resize workspace too small: need %zu bytes", need);
        }
        // rows_total may exceed 65535: loop in slabs of 65535 rows
        for (long long r0 = 0; r0 < rows_total; r0 += 65535) {
            int rows = (int)((rows_total - r0 < 65535) ? rows_total - r0 : 65535);
            dim3 grid(div_up(pl->out_w, 256), rows);
            const uint8_t *s = d_src + (size_t)r0 * pl->in_w * c;
            uint8_t *d = hdst + (size_t)r0 * pl->out_w * c;
            if (c == 3) resize_h_generic_kernel<3><<<grid, 256, 0, st>>>(s, d, pl->d_bx, pl->d_cx, pl->kx, rows, pl->in_w, pl->out_w);
            else resize_h_generic_kernel<1><<<grid, 256, 0, st>>>(s, d, pl->d_bx, pl->d_cx, pl->kx, rows, pl->in_w, pl->out_w);
            LUMINA_KERNEL_CHECK("resize_h_generic_kernel");
        }
        hsrc = hdst;
    }
    if (pl->in_h != pl->out_h) {
        const int row_bytes = pl->out_w * c;
        LUMINA_REQUIRE(pl->out_h <= 65535 && n <= 65535, "image too tall for grid");
        dim3 grid(div_up(row_bytes, 256), pl->out_h, n);
        resize_v_generic_kernel<<<grid, 256, 0, st>>>(hsrc, d_dst, pl->d_by, pl->d_cy, pl->ky, pl->in_h, pl->out_h, row_bytes);
        LUMINA_KERNEL_CHECK("resize_v_generic_kernel");
    }
    return LUMINA_OK;
}
