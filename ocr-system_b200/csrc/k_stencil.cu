// k_stencil.cu -- neighbourhood kernels on uint8 pages:
//   sharpness  = Blend(SMOOTH3x3(img), img, f)   Pillow Filter.c + Blend.c   (:157-158)
//   contrast+sharpness fused (LUT applied on the fly, one read / one write)  (:613-618)
//   median 3x3 edge-replicated                  Pillow RankFilter.c         (:165)
//   adaptive threshold Gaussian 11x11, C        OpenCV thresh.cpp           (:486-492)
// Line numbers: backend/utils/image_preprocessing.py.
//
// The 3x3 kernels treat a page as a flat byte stream: a thread owns 4 consecutive
// bytes (one aligned 32-bit store) and fetches three 12-byte windows (rows y-1, y,
// y+1) with aligned 32-bit loads + funnel shifts, so any width / channel count is
// coalesced.  float32 evaluation order follows the libraries exactly; the file is
// compiled with -fmad=false and uses explicit _rn intrinsics where order matters.
#include "common.cuh"

namespace lumina {

struct Flat {
    const uint8_t *base;  // 4-byte aligned batch base
    long long total;      // total bytes in the batch
};

__device__ __forceinline__ uint32_t flat_word(const Flat &f, long long wi) {
    if (wi < 0) return 0u;
    const long long b = wi * 4;
    if (b + 4 <= f.total) return __ldg(reinterpret_cast<const uint32_t *>(f.base) + wi);
    uint32_t v = 0;
    for (int k = 0; k < 4; k++)
        if (b + k < f.total) v |= (uint32_t)__ldg(f.base + b + k) << (8 * k);
    return v;
}

// 12 bytes starting at flat byte index `start` (any alignment, may be negative)
__device__ __forceinline__ void load12(const Flat &f, long long start, uint32_t w[3]) {
    const long long wi = start >> 2;  // floor
    const int sh = (int)(start & 3) * 8;
    const uint32_t a0 = flat_word(f, wi), a1 = flat_word(f, wi + 1), a2 = flat_word(f, wi + 2);
    const uint32_t a3 = sh ? flat_word(f, wi + 3) : 0u;
    w[0] = __funnelshift_r(a0, a1, sh);
    w[1] = __funnelshift_r(a1, a2, sh);
    w[2] = __funnelshift_r(a2, a3, sh);
}
__device__ __forceinline__ int win_byte(const uint32_t w[3], int k) { return (int)((w[k >> 2] >> (8 * (k & 3))) & 0xffu); }

__device__ __forceinline__ uint32_t blend_px_f(int in1, int in2, float alpha, bool interp) {
    float t = __fadd_rn((float)in1, __fmul_rn(alpha, (float)(in2 - in1)));
    if (interp) return (uint32_t)(int)t;
    if (t <= 0.0f) return 0u;
    if (t >= 255.0f) return 255u;
    return (uint32_t)(int)t;
}

// MODE 0: sharpness.  MODE 1: contrast LUT then sharpness.  MODE 2: median.
template <int C, int MODE>
__global__ void __launch_bounds__(256) stencil3_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int h,
                                                       int w, long long total, const int32_t *__restrict__ mean,
                                                       float calpha, float salpha, float k1, float k5) {
    __shared__ uint8_t lut[256];
    const int page = blockIdx.y;
    const long long pitch = (long long)w * C, page_bytes = pitch * h;
    const long long page_base = (long long)page * page_bytes;
    if (MODE == 1) {
        const bool ci = calpha >= 0.0f && calpha <= 1.0f;
        lut[threadIdx.x] = (uint8_t)blend_px_f(mean[page], threadIdx.x, calpha, ci);
        __syncthreads();
    }
    const Flat f{src, total};
    const int lead = (int)(page_base & 3);
    const bool si = salpha >= 0.0f && salpha <= 1.0f;
    const long long nthreads = (page_bytes + lead + 3) >> 2;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nthreads; t += (long long)gridDim.x * blockDim.x) {
        const long long q0 = t * 4 - lead;  // page-relative byte index of this thread's word
        const long long g0 = page_base + q0;
        uint32_t up[3], mid[3], dn[3];
        load12(f, g0 - pitch - 4, up);
        load12(f, g0 - 4, mid);
        load12(f, g0 + pitch - 4, dn);
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                up[i] = pack4(lut[byte_of(up[i], 0)], lut[byte_of(up[i], 1)], lut[byte_of(up[i], 2)], lut[byte_of(up[i], 3)]);
                mid[i] = pack4(lut[byte_of(mid[i], 0)], lut[byte_of(mid[i], 1)], lut[byte_of(mid[i], 2)], lut[byte_of(mid[i], 3)]);
                dn[i] = pack4(lut[byte_of(dn[i], 0)], lut[byte_of(dn[i], 1)], lut[byte_of(dn[i], 2)], lut[byte_of(dn[i], 3)]);
            }
        }
        long long y = q0 >= 0 ? q0 / pitch : 0;
        long long xb = q0 >= 0 ? q0 - y * pitch : q0;
        uint32_t outw = 0;
        bool valid[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long q = q0 + j;
            valid[j] = q >= 0 && q < page_bytes;
            long long yy = y, xx = xb + j;
            while (xx >= pitch) { xx -= pitch; yy += 1; }   // a 4-byte group spans several rows when a row has fewer than 4 bytes
            const int kc = 4 + j;
            const int center = win_byte(mid, kc);
            uint32_t o;
            if (MODE == 2) {
                // RankFilter: ImagingExpand replicates edges
                const bool top = yy == 0, bot = yy == h - 1, lft = xx < C, rgt = xx >= pitch - C;
                const int kl = lft ? kc : kc - C, kr = rgt ? kc : kc + C;
                uint32_t ru[3], rd[3];
#pragma unroll
                for (int i = 0; i < 3; i++) { ru[i] = top ? mid[i] : up[i]; rd[i] = bot ? mid[i] : dn[i]; }
                int v[9] = {win_byte(ru, kl), win_byte(ru, kc), win_byte(ru, kr), win_byte(mid, kl), center,
                            win_byte(mid, kr), win_byte(rd, kl), win_byte(rd, kc), win_byte(rd, kr)};
#define CSWAP(a, b) { const int lo_ = min(v[a], v[b]); v[b] = max(v[a], v[b]); v[a] = lo_; }
                // 19-exchange median-of-9 network
                CSWAP(1, 2) CSWAP(4, 5) CSWAP(7, 8) CSWAP(0, 1) CSWAP(3, 4) CSWAP(6, 7) CSWAP(1, 2) CSWAP(4, 5)
                CSWAP(7, 8) CSWAP(0, 3) CSWAP(5, 8) CSWAP(4, 7) CSWAP(3, 6) CSWAP(1, 4) CSWAP(2, 5) CSWAP(4, 7)
                CSWAP(4, 2) CSWAP(6, 4) CSWAP(4, 2)
#undef CSWAP
                o = (uint32_t)v[4];
            } else {
                const bool interior = yy > 0 && yy < h - 1 && xx >= C && xx < pitch - C;
                if (interior) {
                    // Filter.c 3x3: ss = 0.5; ss += row(y+1); ss += row(y); ss += row(y-1)
                    float ss = 0.5f;
                    ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn((float)win_byte(dn, kc - C), k1),
                                                           __fmul_rn((float)win_byte(dn, kc), k1)),
                                                 __fmul_rn((float)win_byte(dn, kc + C), k1)));
                    ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn((float)win_byte(mid, kc - C), k1),
                                                           __fmul_rn((float)center, k5)),
                                                 __fmul_rn((float)win_byte(mid, kc + C), k1)));
                    ss = __fadd_rn(ss, __fadd_rn(__fadd_rn(__fmul_rn((float)win_byte(up, kc - C), k1),
                                                           __fmul_rn((float)win_byte(up, kc), k1)),
                                                 __fmul_rn((float)win_byte(up, kc + C), k1)));
                    const int sm = ss <= 0.0f ? 0 : (ss >= 255.0f ? 255 : (int)ss);
                    o = blend_px_f(sm, center, salpha, si);
                } else {
                    o = blend_px_f(center, center, salpha, si);  // smooth copies the border -> blend(x,x)=x
                }
            }
            outw |= o << (8 * j);
        }
        uint8_t *d = dst + g0;
        if (valid[0] && valid[3]) *reinterpret_cast<uint32_t *>(d) = outw;
        else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (valid[j]) d[j] = (uint8_t)(outw >> (8 * j));
        }
    }
}

// ---------------------------------------------------------------------------
// adaptive threshold: fp32 separable Gaussian (RowFilter left-to-right, then
// SymmColumnFilter centre + symmetric pairs), rint, LUT compare.  One CTA per
// 64x32 tile; gray tile (+5 halo) and the row-filtered floats live in shared memory.
// Register blocking: a row-filter item is 4 consecutive outputs of one row (14 tile bytes = 4 aligned word
// loads, each byte converted once instead of once per tap it meets; one 128-bit store), a column-filter item is
// 4 consecutive rows of one column (14 floats for 4 outputs instead of 11 each).  The tile is filled row by
// row (warp = row, lane = byte column), so no index division anywhere.  Every product and sum is a separate
// fp32 operation in OpenCV's order -- the blocking only shares operands.
// ---------------------------------------------------------------------------
constexpr int AT_TW = 64, AT_TH = 32, AT_R = 5;
constexpr int AT_ROWS = AT_TH + 2 * AT_R;          // 42 staged rows
constexpr int AT_COLS = AT_TW + 2 * AT_R;          // 74 staged byte columns
constexpr int AT_TP = 80;                          // tile pitch in bytes (word-aligned rows, >= 76: the last item reads 16 bytes from column 60)
constexpr int AT_FP = AT_TW + 4;                   // float pitch (16-byte aligned rows)
struct Gauss11 { float k[11]; };

// FUSED = OpenCV's default dispatch on x86 hosts with AVX2 + FMA3: the 8-lane vector loops of both filter passes (and
// the row filter's 4-lane step behind them) use fused multiply-add, the scalar remainders do not: row pass fused for
// x < w - (w % 4), column pass fused for x < w - (w % 8) (oracle/lumina_oracle.c orc_adaptive_gauss11_x has the
// evidence).  FUSED = false is OpenCV's plain path.  The file is compiled with -fmad=false, so only the explicit
// __fmaf_rn below is ever fused.
template <int C, bool FUSED>
__global__ void __launch_bounds__(256) adaptive_gauss11_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                               int h, int w, int cval, const Gauss11 g) {
    __shared__ __align__(16) uint8_t tile[AT_ROWS][AT_TP];
    __shared__ __align__(16) float rowf[AT_ROWS][AT_FP];
    const int page = blockIdx.z;
    const int x0 = blockIdx.x * AT_TW, y0 = blockIdx.y * AT_TH;
    const uint8_t *s = src + (size_t)page * h * w * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int ty = warp; ty < AT_ROWS; ty += 8) {
        const int yy = min(max(y0 + ty - AT_R, 0), h - 1);
        const uint8_t *row = s + (size_t)yy * w * C;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int tx = lane + 32 * j;
            if (tx < AT_COLS) {
                const int xx = min(max(x0 + tx - AT_R, 0), w - 1);
                const uint8_t *px = row + (size_t)xx * C;
                uint32_t v;
                if (C == 3) v = (19595u * __ldg(px) + 38470u * __ldg(px + 1) + 7471u * __ldg(px + 2) + 0x8000u) >> 16;
                else v = __ldg(px);
                tile[ty][tx] = (uint8_t)v;
            }
        }
    }
    __syncthreads();
    // row filter: item = (row, 4 output columns); 42 x 16 items
    for (int i = threadIdx.x; i < AT_ROWS * (AT_TW / 4); i += 256) {
        const int ty = i >> 4, q = i & 15;
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(&tile[ty][q * 4]);   // word-aligned (pitch 80), not 16-byte
        const uint32_t wd[4] = {tw[0], tw[1], tw[2], tw[3]};
        float xf[14];
#pragma unroll
        for (int b = 0; b < 14; b++) xf[b] = (float)((wd[b >> 2] >> (8 * (b & 3))) & 255u);
        float o[4];
        const int row_fused_end = w - (w & 3);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            float acc = __fmul_rn(xf[e], g.k[0]);
            if (FUSED && x0 + q * 4 + e < row_fused_end) {
#pragma unroll
                for (int k = 1; k < 11; k++) acc = __fmaf_rn(xf[e + k], g.k[k], acc);
            } else {
#pragma unroll
                for (int k = 1; k < 11; k++) acc = __fadd_rn(acc, __fmul_rn(xf[e + k], g.k[k]));
            }
            o[e] = acc;
        }
        *reinterpret_cast<float4 *>(&rowf[ty][q * 4]) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // column filter + compare: item = (column, 4 output rows); 64 x 8 items
    for (int i = threadIdx.x; i < AT_TW * (AT_TH / 4); i += 256) {
        const int tx = i & 63, ty0 = (i >> 6) * 4;
        const int x = x0 + tx;
        if (x >= w) continue;
        float rf[14];
#pragma unroll
        for (int r = 0; r < 14; r++) rf[r] = rowf[ty0 + r][tx];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int y = y0 + ty0 + e;
            if (y >= h) break;
            float acc = __fmul_rn(g.k[5], rf[e + AT_R]);
            if (FUSED && x < w - (w & 7)) {
#pragma unroll
                for (int k = 1; k <= 5; k++) acc = __fmaf_rn(g.k[5 + k], __fadd_rn(rf[e + AT_R + k], rf[e + AT_R - k]), acc);
            } else {
#pragma unroll
                for (int k = 1; k <= 5; k++)
                    acc = __fadd_rn(acc, __fmul_rn(g.k[5 + k], __fadd_rn(rf[e + AT_R + k], rf[e + AT_R - k])));
            }
            int m = __float2int_rn(acc);
            m = min(max(m, 0), 255);
            const int sv = tile[ty0 + e + AT_R][tx + AT_R];
            dst[((size_t)page * h + y) * w + x] = (sv - m > -cval) ? 255 : 0;
        }
    }
}

}  // namespace lumina

using namespace lumina;

template <int MODE>
static int launch_stencil3(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, const int32_t *d_mean,
                           float calpha, float salpha, void *stream, const char *name) {
    LUMINA_REQUIRE(d_src && d_dst, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE((((uintptr_t)d_src) & 3) == 0 && (((uintptr_t)d_dst) & 3) == 0, "buffers must be 4-byte aligned");
    LUMINA_REQUIRE(d_src != d_dst, "in-place stencil not supported");
    const long long page_bytes = (long long)h * w * c, total = page_bytes * n;
    const float k1 = 1.0f / 13.0f, k5 = 5.0f / 13.0f;  // ImageFilter.SMOOTH (1,1,1,1,5,1,1,1,1)/13 in float32
    int bx = (int)((page_bytes / 4 + 255) / 256);
    int cap = (kNumSMs * 16 + n - 1) / n;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    LUMINA_REQUIRE(n <= 65535, "batch too large for grid");
    dim3 grid(bx, n);
    cudaStream_t st = as_stream(stream);
    if (c == 3) stencil3_kernel<3, MODE><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, total, d_mean, calpha, salpha, k1, k5);
    else stencil3_kernel<1, MODE><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, total, d_mean, calpha, salpha, k1, k5);
    LUMINA_KERNEL_CHECK(name);
    return LUMINA_OK;
}

LUMINA_API int lumina_sharpness_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, float factor,
                                   void *stream) {
    return launch_stencil3<0>(d_src, d_dst, n, h, w, c, nullptr, 1.0f, factor, stream, "stencil3_kernel<sharp>");
}
LUMINA_API int lumina_contrast_sharpness_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                                            const int32_t *d_mean, float contrast_factor, float sharp_factor,
                                            void *stream) {
    LUMINA_REQUIRE(d_mean != nullptr, "null mean pointer");
    return launch_stencil3<1>(d_src, d_dst, n, h, w, c, d_mean, contrast_factor, sharp_factor, stream,
                              "stencil3_kernel<contrast+sharp>");
}
LUMINA_API int lumina_median3_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, void *stream) {
    return launch_stencil3<2>(d_src, d_dst, n, h, w, c, nullptr, 1.0f, 1.0f, stream, "stencil3_kernel<median>");
}

LUMINA_API int lumina_adaptive_gauss11_ex_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, int cval,
                                             int cv_dispatch, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE(cv_dispatch == LUMINA_CV_PLAIN || cv_dispatch == LUMINA_CV_AVX2, "cv_dispatch must be LUMINA_CV_PLAIN or LUMINA_CV_AVX2");
    // cv::getGaussianKernel(11, sigma = 0.3*((11-1)*0.5-1)+0.8 = 2.0) -> float32 taps
    Gauss11 g;
    {
        const double sigma = 2.0, s2 = -0.5 / (sigma * sigma);
        double t[11], sum = 0.0;
        for (int i = 0; i < 11; i++) { double x = i - 5.0; t[i] = exp(s2 * x * x); sum += t[i]; }
        sum = 1.0 / sum;
        for (int i = 0; i < 11; i++) g.k[i] = (float)(t[i] * sum);
    }
    dim3 grid(div_up(w, AT_TW), div_up(h, AT_TH), n);
    LUMINA_REQUIRE(grid.y <= 65535 && n <= 65535, "image too large for grid");
    cudaStream_t st = as_stream(stream);
    const bool fused = cv_dispatch == LUMINA_CV_AVX2;
    if (c == 3 && fused) adaptive_gauss11_kernel<3, true><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, cval, g);
    else if (c == 3) adaptive_gauss11_kernel<3, false><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, cval, g);
    else if (fused) adaptive_gauss11_kernel<1, true><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, cval, g);
    else adaptive_gauss11_kernel<1, false><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, cval, g);
    LUMINA_KERNEL_CHECK("adaptive_gauss11_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_adaptive_gauss11_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, int cval,
                                          void *stream) {
    return lumina_adaptive_gauss11_ex_u8(d_src, d_dst, n, h, w, c, cval, LUMINA_CV_PLAIN, stream);
}
