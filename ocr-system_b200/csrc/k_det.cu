// k_det.cu -- [upstream PaddleOCR] DetResizeForTest(limit_side_len, 'max') +
// NormalizeImage(scale, mean, std, 'hwc') + ToCHWImage in one pass.
//
// Not in the reference tree: it is the "normalize" step
// BASELINE.json's north_star names, restated from PaddleOCR.  The uint8 stage is
// cv2.resize(INTER_LINEAR) bit-exact (resize.cpp: 11-bit fixed-point taps,
// HResizeLinear to int32, VResizeLinear ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2);
// the float stage is (x*scale - mean[c]) / std[c] in float32 without contraction.
// One CTA per 128 x 32 output tile: the horizontal taps of its 128 columns, the vertical taps of its 32 rows and
// the 3 x 256 possible normalised values (x*scale - mean[c]) / std[c] are computed once into shared memory (the
// float stage has only 256 inputs per channel, so the division leaves the per-pixel path; each entry is the same
// three IEEE operations as before).  A thread then produces 4 consecutive pixels of a row per step: 128-bit stores
// into each channel plane (CHW).  Source bytes come through L1.
#include <math.h>

#include "common.cuh"

namespace lumina {

struct DetNorm { float mean[3], stdv[3], scale; };
constexpr int DT_W = 64, DT_H = 32, DT_PX = 2;   // 2 pixels per thread per step: 40 registers, twice the resident warps of the 4-pixel form

__global__ void __launch_bounds__(256) det_resize_normalize_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst,
                                                                   int h, int w, int oh, int ow, double sx, double sy,
                                                                   const DetNorm nm) {
    __shared__ float lut[3][256];
    __shared__ int2 xtab[DT_W];   // {x0 * 3, a0 | a1 << 16}
    __shared__ int xoff1[DT_W];   // x1 * 3
    __shared__ int4 ytab[DT_H];   // {y0, y1, b0, b1}
    constexpr int ONE = 2048;
    const int tid = threadIdx.x;
    const int bx0 = blockIdx.x * DT_W, by0 = blockIdx.y * DT_H;
    const int page = blockIdx.z;
#pragma unroll
    for (int ch = 0; ch < 3; ch++)
        lut[ch][tid] = __fdiv_rn(__fadd_rn(__fmul_rn((float)tid, nm.scale), -nm.mean[ch]), nm.stdv[ch]);
    if (tid < DT_W) {
        const int x = min(bx0 + tid, ow - 1);
        float fx = (float)__dadd_rn(__dmul_rn((double)x + 0.5, sx), -0.5);
        int x0 = (int)floorf(fx);
        fx = __fadd_rn(fx, -(float)x0);
        if (x0 < 0) { fx = 0.f; x0 = 0; }
        if (x0 >= w - 1) { fx = 0.f; x0 = w - 1; }
        const int x1 = min(x0 + 1, w - 1);
        const int a0 = __float2int_rn(__fmul_rn(__fadd_rn(1.f, -fx), (float)ONE)), a1 = __float2int_rn(__fmul_rn(fx, (float)ONE));
        xtab[tid] = make_int2(x0 * 3, a0 | (a1 << 16));   // 0 <= a0, a1 <= 2048
        xoff1[tid] = x1 * 3;
    } else if (tid < DT_W + DT_H) {
        const int y = min(by0 + tid - DT_W, oh - 1);
        float fy = (float)__dadd_rn(__dmul_rn((double)y + 0.5, sy), -0.5);
        const int y0r = (int)floorf(fy);
        fy = __fadd_rn(fy, -(float)y0r);
        const int b0 = __float2int_rn(__fmul_rn(__fadd_rn(1.f, -fy), (float)ONE)), b1 = __float2int_rn(__fmul_rn(fy, (float)ONE));
        ytab[tid - DT_W] = make_int4(min(max(y0r, 0), h - 1), min(max(y0r + 1, 0), h - 1), b0, b1);
    }
    __syncthreads();
    const uint8_t *s = src + (size_t)page * h * w * 3;
    const size_t plane = (size_t)oh * ow;
    float *o_page = dst + (size_t)page * 3 * plane;
    const int q = tid & 31, tx = q * DT_PX;      // DT_PX consecutive columns
    const int x = bx0 + tx;
    if (x >= ow) return;
    const bool vec = (ow & 1) == 0 && x + DT_PX <= ow && ((((uintptr_t)dst) & 7) == 0);
    int xo0[DT_PX], xo1[DT_PX], a0[DT_PX], a1[DT_PX];
#pragma unroll
    for (int e = 0; e < DT_PX; e++) {
        const int2 t = xtab[tx + e];
        xo0[e] = t.x; xo1[e] = xoff1[tx + e];
        a0[e] = t.y & 0xffff; a1[e] = t.y >> 16;
    }
    for (int ty = tid >> 5; ty < DT_H; ty += 8) {
        const int y = by0 + ty;
        if (y >= oh) break;
        const int4 yt = ytab[ty];
        const uint8_t *r0 = s + (size_t)yt.x * w * 3, *r1 = s + (size_t)yt.y * w * 3;
        float out[3][DT_PX];
#pragma unroll
        for (int e = 0; e < DT_PX; e++)
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const int S0 = (int)__ldg(r0 + xo0[e] + ch) * a0[e] + (int)__ldg(r0 + xo1[e] + ch) * a1[e];
                const int S1 = (int)__ldg(r1 + xo0[e] + ch) * a0[e] + (int)__ldg(r1 + xo1[e] + ch) * a1[e];
                int v = (((yt.z * (S0 >> 4)) >> 16) + ((yt.w * (S1 >> 4)) >> 16) + 2) >> 2;
                v = min(max(v, 0), 255);
                out[ch][e] = lut[ch][v];
            }
        float *o = o_page + (size_t)y * ow + x;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            if (vec) *reinterpret_cast<float2 *>(o + ch * plane) = make_float2(out[ch][0], out[ch][1]);
            else {
#pragma unroll
                for (int e = 0; e < DT_PX; e++)
                    if (x + e < ow) o[ch * plane + e] = out[ch][e];
            }
        }
    }
}

}  // namespace lumina

using namespace lumina;

// upstream DetResizeForTest.resize_image_type0: the three limit types differ only in the ratio
//   "max" (0): shrink when the longer side exceeds the limit      "min" (1): enlarge when the shorter side is below it
//   "resize_long" (2): the longer side always becomes the limit
// then int(h * ratio), int(w * ratio), each rounded to a multiple of 32 (python round(): half to even), at least 32.
LUMINA_API int lumina_det_target_size_ex(int h, int w, int limit, int limit_type, int *out_h, int *out_w) {
    LUMINA_REQUIRE(out_h && out_w, "null pointer");
    LUMINA_REQUIRE(h > 0 && w > 0 && limit > 0, "empty image or limit");
    LUMINA_REQUIRE(limit_type >= 0 && limit_type <= 2, "limit_type must be 0 (max), 1 (min) or 2 (resize_long)");
    double ratio = 1.0;
    const int mx = h > w ? h : w, mn = h < w ? h : w;
    if (limit_type == 0) { if (mx > limit) ratio = (double)limit / mx; }
    else if (limit_type == 1) { if (mn < limit) ratio = (double)limit / mn; }
    else ratio = (double)limit / mx;
    int a = (int)(h * ratio), b = (int)(w * ratio);
    a = (int)(lrint(a / 32.0) * 32);  // python round(): half to even
    b = (int)(lrint(b / 32.0) * 32);
    *out_h = a < 32 ? 32 : a;
    *out_w = b < 32 ? 32 : b;
    return LUMINA_OK;
}

LUMINA_API void lumina_det_target_size(int h, int w, int limit, int *out_h, int *out_w) {
    if (lumina_det_target_size_ex(h, w, limit, 0, out_h, out_w) != LUMINA_OK && out_h && out_w) *out_h = *out_w = 32;
}

LUMINA_API int lumina_det_resize_normalize(const uint8_t *d_src, float *d_dst, int n, int h, int w, int oh, int ow,
                                           const float *h_mean3, const float *h_std3, float scale, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst && h_mean3 && h_std3, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "empty batch");
    DetNorm nm;
    for (int i = 0; i < 3; i++) { nm.mean[i] = h_mean3[i]; nm.stdv[i] = h_std3[i]; }
    nm.scale = scale;
    // cv::resize: inv_scale = dsize/ssize (double), scale = 1/inv_scale
    const double sx = 1.0 / ((double)ow / w), sy = 1.0 / ((double)oh / h);
    dim3 grid(div_up(ow, DT_W), div_up(oh, DT_H), n);
    LUMINA_REQUIRE(grid.y <= 65535 && n <= 65535, "image too large for grid");
    det_resize_normalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_src, d_dst, h, w, oh, ow, sx, sy, nm);
    LUMINA_KERNEL_CHECK("det_resize_normalize_kernel");
    return LUMINA_OK;
}
