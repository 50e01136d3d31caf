// k_binarize.cu -- global (Otsu) and local (Sauvola) binarisation of page planes.
//
// BASELINE.json's north_star names "integral-image Sauvola/Otsu binarization" beside the reference's own two
// binarisers (image_preprocessing.py:175-185 fixed threshold, :462-494 cv2.adaptiveThreshold -- k_point.cu /
// k_stencil.cu).  The reference itself never calls these two, so they are additional operators behind
// the same C-ABI, not replacements of a reference call site:
//   Otsu     -- oracle = cv2.threshold(gray, 0, 255, THRESH_BINARY | THRESH_OTSU) (getThreshVal_Otsu_8u in
//               modules/imgproc/src/thresh.cpp, restated below in double, same operation order): histogram kernel
//               (per-warp private shared-memory histograms), one thread per page for the 256-step recurrence, then
//               the comparison.  Bit-equal to cv2 (threshold and mask).
//   Sauvola  -- T = m * (1 + k * (s / R - 1)) over a (2r+1)^2 window clipped to the page, m and s from the exact
//               integer sums of x and x^2.  No library in this image implements it (skimage is absent): parity is
//               against the float64 NumPy restatement in oracle/ (integral images in int64) -- "parity unpinned"
//               by the reference.  One CTA per 64x32 output tile: tile + halo staged in shared memory, the two
//               integral images built THERE (uint32: a tile sum never exceeds 2^32), four look-ups per sum, so the
//               page is read once (plus halo) and written once; no integral image ever goes to HBM.
// Compiled with -fmad=false: the double expressions must round like NumPy's / OpenCV's separate operations.
#include "common.cuh"

namespace lumina {

// ---- Otsu ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hist256_kernel(const uint8_t *__restrict__ src, size_t px_per_page,
                                                      unsigned int *__restrict__ hist) {
    __shared__ unsigned int sh[8][256];
    const int page = blockIdx.y, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const uint8_t *p = src + (size_t)page * px_per_page;
    const size_t nvec = px_per_page / 16;
    const bool aligned = (((uintptr_t)p) & 15) == 0;
    if (aligned) {
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
            const uint4 v = ldg_stream_u4(p + i * 16);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int b = 0; b < 4; b++) atomicAdd(&sh[warp][(w[q] >> (8 * b)) & 255u], 1u);
        }
    }
    const size_t tail0 = aligned ? nvec * 16 : 0;
    for (size_t i = tail0 + (size_t)blockIdx.x * 256 + threadIdx.x; i < px_per_page; i += (size_t)gridDim.x * 256)
        atomicAdd(&sh[warp][p[i]], 1u);
    __syncthreads();
    unsigned int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
    if (s) atomicAdd(&hist[(size_t)page * 256 + threadIdx.x], s);
}

// getThreshVal_Otsu_8u (thresh.cpp): everything in double, in this order.
__global__ void otsu_threshold_kernel(const unsigned int *__restrict__ hist, size_t px_per_page, int n,
                                      int32_t *__restrict__ thresh) {
    const int page = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n) return;
    const unsigned int *h = hist + (size_t)page * 256;
    const double scale = 1. / (double)px_per_page;
    double mu = 0;
    for (int i = 0; i < 256; i++) mu += (double)i * (double)h[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
    for (int i = 0; i < 256; i++) {
        const double p_i = (double)h[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        const double q2 = 1. - q1;
        if (fmin(q1, q2) < eps || fmax(q1, q2) > 1. - eps) continue;
        mu1 = (mu1 + (double)i * p_i) / q1;
        const double mu2 = (mu - q1 * mu1) / q2;
        const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) {
            max_sigma = sigma;
            max_val = (double)i;
        }
    }
    thresh[page] = (int32_t)max_val;
}

__global__ void __launch_bounds__(256) threshold_pages_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                              size_t px_per_page, const int32_t *__restrict__ thresh) {
    const int page = blockIdx.y;
    const int t = thresh[page];
    const uint8_t *p = src + (size_t)page * px_per_page;
    uint8_t *o = dst + (size_t)page * px_per_page;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < px_per_page; i += (size_t)gridDim.x * 256)
        o[i] = (int)p[i] > t ? 255 : 0;
}

// ---- Sauvola ---------------------------------------------------------------------------------------------
constexpr int SV_TW = 64, SV_TH = 32, SV_RMAX = 24;

__global__ void __launch_bounds__(256) sauvola_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int h,
                                                      int w, int r, double k, double R) {
    extern __shared__ __align__(16) unsigned char sv_smem[];
    const int SW = SV_TW + 2 * r + 1, SHH = SV_TH + 2 * r + 1;   // integral images carry a zero row / column in front
    uint32_t *I1 = reinterpret_cast<uint32_t *>(sv_smem);          // [SHH][SW] sums of x
    uint32_t *I2 = I1 + (size_t)SHH * SW;                          // sums of x^2
    const int page = blockIdx.z;
    const uint8_t *P = src + (size_t)page * h * w;
    uint8_t *O = dst + (size_t)page * h * w;
    const int x0 = blockIdx.x * SV_TW - r, y0 = blockIdx.y * SV_TH - r;   // page coordinates of the staged tile's origin
    const int tw = SW - 1, th = SHH - 1;
    // stage: value (0 outside the page) into I1[y+1][x+1], its square into I2; first row / column zero
    for (int i = threadIdx.x; i < SHH * SW; i += 256) {
        const int ty = i / SW, tx = i - ty * SW;
        uint32_t v = 0;
        if (ty > 0 && tx > 0) {
            const int gx = x0 + tx - 1, gy = y0 + ty - 1;
            if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = P[(size_t)gy * w + gx];
        }
        I1[i] = v;
        I2[i] = v * v;
    }
    __syncthreads();
    // row prefix sums: one warp per row, 32 elements per step with a carry
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int ty = 1 + warp; ty <= th; ty += 8) {
        uint32_t c1 = 0, c2 = 0;
        for (int b = 1; b <= tw; b += 32) {
            const int tx = b + lane;
            uint32_t a1 = tx <= tw ? I1[ty * SW + tx] : 0u, a2 = tx <= tw ? I2[ty * SW + tx] : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u1 = __shfl_up_sync(0xffffffffu, a1, d), u2 = __shfl_up_sync(0xffffffffu, a2, d);
                if (lane >= d) { a1 += u1; a2 += u2; }
            }
            a1 += c1; a2 += c2;
            if (tx <= tw) { I1[ty * SW + tx] = a1; I2[ty * SW + tx] = a2; }
            c1 = __shfl_sync(0xffffffffu, a1, 31);
            c2 = __shfl_sync(0xffffffffu, a2, 31);
        }
    }
    __syncthreads();
    // column prefix sums: one thread per column
    for (int tx = 1 + threadIdx.x; tx <= tw; tx += 256) {
        uint32_t a1 = 0, a2 = 0;
        for (int ty = 1; ty <= th; ty++) {
            a1 += I1[ty * SW + tx]; a2 += I2[ty * SW + tx];
            I1[ty * SW + tx] = a1; I2[ty * SW + tx] = a2;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SV_TW * SV_TH; i += 256) {
        const int oy = i / SV_TW, ox = i - oy * SV_TW;
        const int gx = blockIdx.x * SV_TW + ox, gy = blockIdx.y * SV_TH + oy;
        if (gx >= w || gy >= h) continue;
        // window clipped to the page, in tile coordinates (+1 for the zero border): [xa, xb) x [ya, yb)
        const int wx0 = max(gx - r, 0), wx1 = min(gx + r, w - 1), wy0 = max(gy - r, 0), wy1 = min(gy + r, h - 1);
        const int xa = wx0 - x0, xb = wx1 - x0 + 1, ya = wy0 - y0, yb = wy1 - y0 + 1;
        const uint32_t s1 = I1[yb * SW + xb] - I1[ya * SW + xb] - I1[yb * SW + xa] + I1[ya * SW + xa];
        const uint32_t s2 = I2[yb * SW + xb] - I2[ya * SW + xb] - I2[yb * SW + xa] + I2[ya * SW + xa];
        const double cnt = (double)((wx1 - wx0 + 1) * (wy1 - wy0 + 1));
        const double m = (double)s1 / cnt;
        double var = (double)s2 / cnt - m * m;
        if (var < 0.) var = 0.;
        const double s = sqrt(var);
        const double T = m * (1. + k * (s / R - 1.));
        const uint8_t v = P[(size_t)gy * w + gx];
        O[(size_t)gy * w + gx] = (double)v > T ? 255 : 0;
    }
}

}  // namespace lumina

using namespace lumina;

LUMINA_API int lumina_otsu_u8(const uint8_t *d_gray, uint8_t *d_dst, int n, int h, int w, int32_t *d_thresh,
                              uint32_t *d_hist_scratch, void *stream) {
    LUMINA_REQUIRE(d_gray && d_thresh && d_hist_scratch, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && n <= 65535, "bad batch");
    cudaStream_t st = as_stream(stream);
    const size_t px = (size_t)h * w;
    LUMINA_CUDA_TRY(cudaMemsetAsync(d_hist_scratch, 0, (size_t)n * 256 * 4, st));
    int bx = (int)((px / 16 + 255) / 256);
    const int cap = (kNumSMs * 8 + n - 1) / n;
    bx = bx > cap ? cap : (bx < 1 ? 1 : bx);
    hist256_kernel<<<dim3(bx, n), 256, 0, st>>>(d_gray, px, d_hist_scratch);
    LUMINA_KERNEL_CHECK("hist256_kernel");
    otsu_threshold_kernel<<<(n + 63) / 64, 64, 0, st>>>(d_hist_scratch, px, n, d_thresh);
    LUMINA_KERNEL_CHECK("otsu_threshold_kernel");
    if (d_dst) {
        threshold_pages_kernel<<<dim3(bx, n), 256, 0, st>>>(d_gray, d_dst, px, d_thresh);
        LUMINA_KERNEL_CHECK("threshold_pages_kernel");
    }
    return LUMINA_OK;
}

LUMINA_API int lumina_sauvola_u8(const uint8_t *d_gray, uint8_t *d_dst, int n, int h, int w, int window, double k, double R,
                                 void *stream) {
    LUMINA_REQUIRE(d_gray && d_dst, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && n <= 65535, "bad batch");
    LUMINA_REQUIRE(window >= 3 && (window & 1) == 1 && window / 2 <= SV_RMAX, "window must be odd, 3 .. 49");
    LUMINA_REQUIRE(R > 0., "R must be positive");
    const int r = window / 2;
    const size_t smem = (size_t)2 * (SV_TW + 2 * r + 1) * (SV_TH + 2 * r + 1) * 4;
    LUMINA_CUDA_TRY(cudaFuncSetAttribute(sauvola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sauvola_kernel<<<dim3(div_up(w, SV_TW), div_up(h, SV_TH), n), 256, smem, as_stream(stream)>>>(d_gray, d_dst, h, w, r, k, R);
    LUMINA_KERNEL_CHECK("sauvola_kernel");
    return LUMINA_OK;
}
