// k_warp.cu -- cv2.getRotationMatrix2D + cv2.warpAffine(INTER_CUBIC,
// BORDER_REPLICATE) on uint8 pages, bit-exact (the deskew rotation).
//
// Replaces OpenCV imgwarp.cpp (WarpAffineInvoker + remapBicubic with the 32x32
// int16 weight table) reached from backend/utils/image_preprocessing.py:442-450.
// Arithmetic: SURVEY App. A9 -- matrix inverted in double on the host exactly as
// OpenCV does, AB_BITS=10 fixed-point coordinates, 5 fractional bits per axis,
// 16 int16 weights per (fy,fx) with OpenCV's sum-to-32768 fix-up on the
// lower-right 2x2 taps, (sum + 2^14) >> 15.
#include <math.h>

#include <mutex>

#include "common.cuh"

namespace lumina {

constexpr int TAB_SZ = 32;
static int16_t h_cubic_tab[TAB_SZ * TAB_SZ * 16];
static std::once_flag cubic_once;

static void cubic_coeffs(float x, float *c) {
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

static void build_cubic_table() {
    float t1[TAB_SZ][4];
    for (int i = 0; i < TAB_SZ; i++) cubic_coeffs((float)i * (1.0f / TAB_SZ), t1[i]);
    for (int i = 0; i < TAB_SZ; i++)
        for (int j = 0; j < TAB_SZ; j++) {
            int16_t *it = h_cubic_tab + (size_t)(i * TAB_SZ + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; k1++)
                for (int k2 = 0; k2 < 4; k2++) {
                    const float v = t1[i][k1] * t1[j][k2];
                    int iv = (int)lrintf(v * 32768.0f);
                    iv = iv > 32767 ? 32767 : (iv < -32768 ? -32768 : iv);
                    it[k1 * 4 + k2] = (int16_t)iv;
                    isum += iv;
                }
            if (isum != 32768) {
                const int diff = isum - 32768;
                int Mk = 2 * 4 + 2, mk = 2 * 4 + 2;  // search the {2,3}x{2,3} taps (OpenCV's ksize2 quirk)
                for (int k1 = 2; k1 < 4; k1++)
                    for (int k2 = 2; k2 < 4; k2++) {
                        const int q = k1 * 4 + k2;
                        if (it[q] < it[mk]) mk = q;
                        else if (it[q] > it[Mk]) Mk = q;
                    }
                if (diff < 0) it[Mk] = (int16_t)(it[Mk] - diff);
                else it[mk] = (int16_t)(it[mk] - diff);
            }
        }
}

// per-device copy of the table (init-once cache; the only global mutable state)
static std::mutex tab_mutex;
static int16_t *d_tab_for_device[64] = {nullptr};

static int16_t *device_cubic_table(cudaError_t *err) {
    std::call_once(cubic_once, build_cubic_table);
    int dev = 0;
    *err = cudaGetDevice(&dev);
    if (*err != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(tab_mutex);
    if (!d_tab_for_device[dev]) {
        int16_t *d = nullptr;
        *err = cudaMalloc((void **)&d, sizeof(h_cubic_tab));
        if (*err != cudaSuccess) return nullptr;
        *err = cudaMemcpy(d, h_cubic_tab, sizeof(h_cubic_tab), cudaMemcpyHostToDevice);
        if (*err != cudaSuccess) { cudaFree(d); return nullptr; }
        d_tab_for_device[dev] = d;
    }
    return d_tab_for_device[dev];
}

constexpr int WARP_CHUNK = 32;  // pages per launch (matrices travel as kernel arguments)
struct WarpMats {
    double m[WARP_CHUNK][6];  // INVERTED matrices
    uint8_t apply[WARP_CHUNK];
};

template <int C>
__global__ void __launch_bounds__(256) warp_affine_cubic_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                                int h, int w, const int16_t *__restrict__ tab,
                                                                const WarpMats mats) {
    const int page = blockIdx.z;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const uint8_t *s = src + (size_t)page * h * w * C;
    uint8_t *d = dst + ((size_t)page * h * w + (size_t)y * w + x) * C;
    if (!mats.apply[page]) {
#pragma unroll
        for (int ch = 0; ch < C; ch++) d[ch] = __ldg(s + ((size_t)y * w + x) * C + ch);
        return;
    }
    const double *M = mats.m[page];
    constexpr int AB_BITS = 10, INTER_BITS = 5, INTER_TAB = 32;
    constexpr double AB_SCALE = 1024.0;
    constexpr int round_delta = 1024 / 32 / 2;
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), AB_SCALE));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), AB_SCALE));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], (double)y), M[2]), AB_SCALE)) + round_delta;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], (double)y), M[5]), AB_SCALE)) + round_delta;
    const int X = (X0 + adelta) >> (AB_BITS - INTER_BITS), Y = (Y0 + bdelta) >> (AB_BITS - INTER_BITS);
    const int sx = min(max(X >> INTER_BITS, -32768), 32767) - 1, sy = min(max(Y >> INTER_BITS, -32768), 32767) - 1;
    const int ai = (Y & (INTER_TAB - 1)) * INTER_TAB + (X & (INTER_TAB - 1));
    const uint4 *tw = reinterpret_cast<const uint4 *>(tab + (size_t)ai * 16);
    const uint4 w0 = __ldg(tw), w1 = __ldg(tw + 1);
    const uint32_t wp[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    int sum[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) sum[ch] = 0;
    if (C == 3 && sx >= 0 && sx + 4 < w) {
        // interior: the 4 x 3 bytes of a row are contiguous -- four aligned 32-bit loads + funnel shifts instead of
        // twelve byte loads (the one pixel of margin keeps the 16-byte read inside the row)
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) {
            const int yy = min(max(sy + k1, 0), h - 1);
            const uintptr_t addr = (uintptr_t)(s + ((size_t)yy * w + sx) * 3);
            const uint32_t *wp32 = reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3);
            const uint32_t a0 = __ldg(wp32), a1 = __ldg(wp32 + 1), a2 = __ldg(wp32 + 2), a3 = __ldg(wp32 + 3);
            const int sh = (int)(addr & 3) * 8;
            const uint32_t v0 = __funnelshift_r(a0, a1, sh), v1 = __funnelshift_r(a1, a2, sh), v2 = __funnelshift_r(a2, a3, sh);
            // the row's 12 bytes R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3 -> one word of four taps per channel (2 PRMT each),
            // then the four int16 weights of the row (two packed words) against it: dp2a.lo = taps 0,1, dp2a.hi = taps 2,3
            const uint32_t cw[3] = {__byte_perm(__byte_perm(v0, v1, 0x0630), v2, 0x5210),
                                    __byte_perm(__byte_perm(v0, v1, 0x0741), v2, 0x6210),
                                    __byte_perm(__byte_perm(v0, v1, 0x0052), v2, 0x7410)};
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(sum[ch]) : "r"(wp[k1 * 2]), "r"(cw[ch]));
                asm("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(sum[ch]) : "r"(wp[k1 * 2 + 1]), "r"(cw[ch]));
            }
        }
#pragma unroll
        for (int ch = 0; ch < C; ch++) d[ch] = (uint8_t)min(max((sum[ch] + (1 << 14)) >> 15, 0), 255);
        return;
    }
    int xx[4];
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) xx[k2] = min(max(sx + k2, 0), w - 1) * C;
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        const int yy = min(max(sy + k1, 0), h - 1);
        const uint8_t *row = s + (size_t)yy * w * C;
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) {
            const int q = k1 * 4 + k2;
            const int wt = (int)(short)((wp[q >> 1] >> (16 * (q & 1))) & 0xffffu);
#pragma unroll
            for (int ch = 0; ch < C; ch++) sum[ch] += (int)__ldg(row + xx[k2] + ch) * wt;
        }
    }
#pragma unroll
    for (int ch = 0; ch < C; ch++) d[ch] = (uint8_t)min(max((sum[ch] + (1 << 14)) >> 15, 0), 255);
}

}  // namespace lumina

using namespace lumina;

LUMINA_API void lumina_rotation_matrix_host(double cx, double cy, double angle_deg, double scale, double *m) {
    const double ang = angle_deg * (M_PI / 180.0);
    const double alpha = cos(ang) * scale, beta = sin(ang) * scale;
    m[0] = alpha; m[1] = beta; m[2] = (1 - alpha) * cx - beta * cy;
    m[3] = -beta; m[4] = alpha; m[5] = beta * cx + (1 - alpha) * cy;
}

// cv::warpAffine: invert the forward matrix in double (imgwarp.cpp, invertAffineTransform's arithmetic)
LUMINA_API void lumina_invert_affine_host(const double *F, double *M) {
    double D = F[0] * F[4] - F[1] * F[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = F[4] * D, A22 = F[0] * D;
    M[0] = A11; M[1] = F[1] * (-D); M[3] = F[3] * (-D); M[4] = A22;
    const double b1 = -M[0] * F[2] - M[1] * F[5];
    const double b2 = -M[3] * F[2] - M[4] * F[5];
    M[2] = b1; M[5] = b2;
}

LUMINA_API int lumina_warp_affine_cubic_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                                           const double *h_m6, const uint8_t *h_apply, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst && h_m6, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE(d_src != d_dst, "in-place warp not supported");
    cudaError_t err;
    const int16_t *tab = device_cubic_table(&err);
    if (!tab) return set_error(LUMINA_E_CUDA, "cubic table upload failed: %s", cudaGetErrorString(err));
    cudaStream_t st = as_stream(stream);
    for (int p0 = 0; p0 < n; p0 += WARP_CHUNK) {
        const int np = n - p0 < WARP_CHUNK ? n - p0 : WARP_CHUNK;
        WarpMats mats;
        memset(&mats, 0, sizeof(mats));
        for (int i = 0; i < np; i++) {
            lumina_invert_affine_host(h_m6 + (size_t)(p0 + i) * 6, mats.m[i]);
            mats.apply[i] = h_apply ? h_apply[p0 + i] : 1;
        }
        dim3 grid(div_up(w, 32), div_up(h, 8), np);
        LUMINA_REQUIRE(grid.y <= 65535, "image too tall for grid");
        const size_t off = (size_t)p0 * h * w * c;
        if (c == 3) warp_affine_cubic_kernel<3><<<grid, 256, 0, st>>>(d_src + off, d_dst + off, h, w, tab, mats);
        else warp_affine_cubic_kernel<1><<<grid, 256, 0, st>>>(d_src + off, d_dst + off, h, w, tab, mats);
        LUMINA_KERNEL_CHECK("warp_affine_cubic_kernel");
    }
    return LUMINA_OK;
}
