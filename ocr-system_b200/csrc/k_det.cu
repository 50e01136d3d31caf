// k_det.cu -- [upstream PaddleOCR] DetResizeForTest(limit_side_len, 'max') +
// NormalizeImage(scale, mean, std, 'hwc') + ToCHWImage in one pass.
//
// Not in the reference tree: it is the "normalize" step
// BASELINE.json's north_star names, restated from PaddleOCR.  The uint8 stage is
// cv2.resize(INTER_LINEAR) bit-exact (resize.cpp: 11-bit fixed-point taps,
// HResizeLinear to int32, VResizeLinear ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2);
// the float stage is (x*scale - mean[c]) / std[c] in float32 without contraction.
// One thread per output pixel: 4 source pixels in (L1-cached), 3 floats out (CHW).
#include <math.h>

#include "common.cuh"

namespace lumina {

struct DetNorm { float mean[3], stdv[3], scale; };

__global__ void __launch_bounds__(256) det_resize_normalize_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst,
                                                                   int h, int w, int oh, int ow, double sx, double sy,
                                                                   const DetNorm nm) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int page = blockIdx.z;
    if (x >= ow || y >= oh) return;
    constexpr int ONE = 2048;
    // horizontal taps
    float fx = (float)__dadd_rn(__dmul_rn((double)x + 0.5, sx), -0.5);
    int x0 = (int)floorf(fx);
    fx = __fadd_rn(fx, -(float)x0);
    if (x0 < 0) { fx = 0.f; x0 = 0; }
    if (x0 >= w - 1) { fx = 0.f; x0 = w - 1; }
    const int x1 = min(x0 + 1, w - 1);
    const int a0 = __float2int_rn(__fmul_rn(__fadd_rn(1.f, -fx), (float)ONE)), a1 = __float2int_rn(__fmul_rn(fx, (float)ONE));
    // vertical taps
    float fy = (float)__dadd_rn(__dmul_rn((double)y + 0.5, sy), -0.5);
    const int y0r = (int)floorf(fy);
    fy = __fadd_rn(fy, -(float)y0r);
    const int b0 = __float2int_rn(__fmul_rn(__fadd_rn(1.f, -fy), (float)ONE)), b1 = __float2int_rn(__fmul_rn(fy, (float)ONE));
    const int y0 = min(max(y0r, 0), h - 1), y1 = min(max(y0r + 1, 0), h - 1);
    const uint8_t *s = src + (size_t)page * h * w * 3;
    const uint8_t *r0 = s + (size_t)y0 * w * 3, *r1 = s + (size_t)y1 * w * 3;
    float *o = dst + (size_t)page * 3 * oh * ow + (size_t)y * ow + x;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const int S0 = (int)__ldg(r0 + x0 * 3 + ch) * a0 + (int)__ldg(r0 + x1 * 3 + ch) * a1;
        const int S1 = (int)__ldg(r1 + x0 * 3 + ch) * a0 + (int)__ldg(r1 + x1 * 3 + ch) * a1;
        int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        const float f = __fdiv_rn(__fadd_rn(__fmul_rn((float)v, nm.scale), -nm.mean[ch]), nm.stdv[ch]);
        o[(size_t)ch * oh * ow] = f;
    }
}

}  // namespace lumina

using namespace lumina;

LUMINA_API void lumina_det_target_size(int h, int w, int limit, int *out_h, int *out_w) {
    double ratio = 1.0;
    const int mx = h > w ? h : w;
    if (mx > limit) ratio = (double)limit / mx;
    int a = (int)(h * ratio), b = (int)(w * ratio);
    a = (int)(lrint(a / 32.0) * 32);  // python round(): half to even
    b = (int)(lrint(b / 32.0) * 32);
    *out_h = a < 32 ? 32 : a;
    *out_w = b < 32 ? 32 : b;
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
    dim3 grid(div_up(ow, 64), div_up(oh, 4), n);
    LUMINA_REQUIRE(grid.y <= 65535 && n <= 65535, "image too large for grid");
    det_resize_normalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_src, d_dst, h, w, oh, ow, sx, sy, nm);
    LUMINA_KERNEL_CHECK("det_resize_normalize_kernel");
    return LUMINA_OK;
}
