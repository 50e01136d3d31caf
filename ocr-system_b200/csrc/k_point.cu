// k_point.cu -- per-pixel kernels: grayscale (PIL / OpenCV constants), fixed
// binarize, contrast (mean + blend LUT), EXIF transpose.  All HBM-bound byte
// streams: 128-bit coalesced loads/stores, 16 pixels per thread.
//
// Reference call sites (backend/utils/image_preprocessing.py):
//   convert('L') :169,184,481   cvtColor RGB2BGR/BGR2GRAY :395-396
//   ImageEnhance.Contrast :143-144   point(>thr,'1') :185   exif_transpose :173
#include "common.cuh"

namespace lumina {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---------------------------------------------------------------------------
// gray: 16 px / thread: 3 x uint4 in, 1 x uint4 out.  MODE 0 = PIL, 1 = OpenCV.
// ---------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ uint32_t gray_of(uint32_t r, uint32_t g, uint32_t b) {
    if (MODE == 0) return (19595u * r + 38470u * g + 7471u * b + 0x8000u) >> 16;
    return (9798u * r + 19235u * g + 3735u * b + (1u << 14)) >> 15;
}

// POST: 0 = gray, 1 = gray > thr ? 255 : 0
template <int MODE, int POST>
__global__ void __launch_bounds__(256) gray16_kernel(const uint8_t *__restrict__ rgb, uint8_t *__restrict__ out,
                                                     size_t npx, int thr) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * 16;
    if (p0 >= npx) return;
    if (p0 + 16 <= npx) {
        const uint4 *src = reinterpret_cast<const uint4 *>(rgb + p0 * 3);
        uint4 a = ldg_stream_u4(src), b = ldg_stream_u4(src + 1), c = ldg_stream_u4(src + 2);
        uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int px = q * 4 + j, bi = px * 3;
                uint32_t r = byte_of(w[bi >> 2], bi & 3), g = byte_of(w[(bi + 1) >> 2], (bi + 1) & 3),
                         bl = byte_of(w[(bi + 2) >> 2], (bi + 2) & 3);
                uint32_t y = gray_of<MODE>(r, g, bl);
                if (POST == 1) y = (int)y > thr ? 255u : 0u;
                v |= y << (8 * j);
            }
            o[q] = v;
        }
        stg_stream_u4(out + p0, make_uint4(o[0], o[1], o[2], o[3]));
    } else {
        for (size_t p = p0; p < npx; p++) {
            uint32_t y = gray_of<MODE>(rgb[p * 3], rgb[p * 3 + 1], rgb[p * 3 + 2]);
            if (POST == 1) y = (int)y > thr ? 255u : 0u;
            out[p] = (uint8_t)y;
        }
    }
}

__global__ void __launch_bounds__(256) threshold16_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                          size_t n, int thr) {
    size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 >= n) return;
    if (p0 + 16 <= n) {
        uint4 a = ldg_stream_u4(in + p0);
        uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) v |= ((int)byte_of(w[q], j) > thr ? 255u : 0u) << (8 * j);
            w[q] = v;
        }
        stg_stream_u4(out + p0, make_uint4(w[0], w[1], w[2], w[3]));
    } else {
        for (size_t p = p0; p < n; p++) out[p] = in[p] > thr ? 255 : 0;
    }
}

// ---------------------------------------------------------------------------
// contrast mean: exact integer sum of PIL-L over each page (== sum i*hist[i]),
// warp-shuffle reduce, one 64-bit atomic per block.  mean = int(sum/cnt + 0.5)
// ---------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) lsum_kernel(const uint8_t *__restrict__ src, size_t px_per_page,
                                                   unsigned long long *__restrict__ sums) {
    const int page = blockIdx.y;
    const uint8_t *base = src + (size_t)page * px_per_page * C;
    // pages are not necessarily 16B aligned (px_per_page*C arbitrary): peel to alignment
    size_t nbytes = px_per_page * C;
    uint32_t acc = 0;
    unsigned long long total = 0;
    size_t mis = (16 - ((uintptr_t)base & 15)) & 15;
    if (C == 3) mis = 0;  // RGB path below uses pixel granularity, handled with scalar loads at the ends
    if (C == 1) {
        size_t head = mis < nbytes ? mis : nbytes;
        size_t nvec = (nbytes - head) / 16;
        const uint4 *v = reinterpret_cast<const uint4 *>(base + head);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
            uint4 a = ldg_stream_u4(v + i);
            acc += __dp4a(a.x, 0x01010101u, 0u) + __dp4a(a.y, 0x01010101u, 0u) + __dp4a(a.z, 0x01010101u, 0u) +
                   __dp4a(a.w, 0x01010101u, 0u);
            if (acc > 0xf0000000u) { total += acc; acc = 0; }
        }
        if (blockIdx.x == 0) {
            for (size_t i = threadIdx.x; i < head; i += blockDim.x) acc += base[i];
            for (size_t i = head + nvec * 16 + threadIdx.x; i < nbytes; i += blockDim.x) acc += base[i];
        }
    } else {
        // 16 px (48 B) per iteration when the page base is 16B aligned, else scalar
        bool aligned = (((uintptr_t)base) & 15) == 0;
        size_t nvec = aligned ? px_per_page / 16 : 0;
        const uint4 *v = reinterpret_cast<const uint4 *>(base);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
            uint4 a = ldg_stream_u4(v + i * 3), b = ldg_stream_u4(v + i * 3 + 1), c = ldg_stream_u4(v + i * 3 + 2);
            uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int px = 0; px < 16; px++) {
                int bi = px * 3;
                acc += gray_of<0>(byte_of(w[bi >> 2], bi & 3), byte_of(w[(bi + 1) >> 2], (bi + 1) & 3),
                                  byte_of(w[(bi + 2) >> 2], (bi + 2) & 3));
            }
        }
        for (size_t p = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < px_per_page;
             p += (size_t)gridDim.x * blockDim.x)
            acc += gray_of<0>(base[p * 3], base[p * 3 + 1], base[p * 3 + 2]);
    }
    total += acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_down_sync(0xffffffffu, total, o);
    __shared__ unsigned long long wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) s += wsum[i];
        atomicAdd(&sums[page], s);
    }
}

__global__ void mean_from_sum_kernel(const unsigned long long *__restrict__ sums, size_t px_per_page, int n,
                                     int32_t *__restrict__ mean) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mean[i] = (int)(__ddiv_rn((double)sums[i], (double)px_per_page) + 0.5);
}

// Blend.c: out = (UINT8) clamp(in1 + alpha*(in2-in1)) in float32, no FMA contraction
__device__ __forceinline__ uint32_t blend_px(int in1, int in2, float alpha, bool interp) {
    float t = __fadd_rn((float)in1, __fmul_rn(alpha, (float)(in2 - in1)));
    if (interp) return (uint32_t)(int)t;
    if (t <= 0.0f) return 0u;
    if (t >= 255.0f) return 255u;
    return (uint32_t)(int)t;
}

// contrast apply: per-block 256-entry LUT of blend(mean, x) in shared memory, 16 B / thread
__global__ void __launch_bounds__(256) contrast_apply_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                             size_t bytes_per_page, const int32_t *__restrict__ mean,
                                                             float alpha) {
    __shared__ uint8_t lut[256];
    const int page = blockIdx.y;
    const bool interp = alpha >= 0.0f && alpha <= 1.0f;
    lut[threadIdx.x] = (uint8_t)blend_px(mean[page], threadIdx.x, alpha, interp);
    __syncthreads();
    const uint8_t *s = src + (size_t)page * bytes_per_page;
    uint8_t *d = dst + (size_t)page * bytes_per_page;
    const bool aligned = ((((uintptr_t)s) | ((uintptr_t)d)) & 15) == 0;
    size_t nvec = aligned ? bytes_per_page / 16 : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 a = ldg_stream_u4(s + i * 16);
        uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
            w[q] = pack4(lut[byte_of(w[q], 0)], lut[byte_of(w[q], 1)], lut[byte_of(w[q], 2)], lut[byte_of(w[q], 3)]);
        stg_stream_u4(d + i * 16, make_uint4(w[0], w[1], w[2], w[3]));
    }
    for (size_t i = nvec * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < bytes_per_page;
         i += (size_t)gridDim.x * blockDim.x)
        d[i] = lut[s[i]];
}

// EXIF transpose: one thread per destination pixel (C bytes); reads are gathered
// through L1 for the transposing cases.
template <int C>
__global__ void __launch_bounds__(256) exif_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int h,
                                                   int w, int orientation) {
    const int oh = (orientation >= 5) ? w : h, ow = (orientation >= 5) ? h : w;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= ow || y >= oh) return;
    int sy, sx;
    switch (orientation) {
        case 2: sy = y; sx = w - 1 - x; break;
        case 3: sy = h - 1 - y; sx = w - 1 - x; break;
        case 4: sy = h - 1 - y; sx = x; break;
        case 5: sy = x; sx = y; break;
        case 6: sy = h - 1 - x; sx = y; break;
        case 7: sy = h - 1 - x; sx = w - 1 - y; break;
        case 8: sy = x; sx = w - 1 - y; break;
        default: sy = y; sx = x; break;
    }
    const size_t page = (size_t)blockIdx.z * h * w;
    const uint8_t *s = src + (page + (size_t)sy * w + sx) * C;
    uint8_t *d = dst + (page + (size_t)y * ow + x) * C;
#pragma unroll
    for (int k = 0; k < C; k++) d[k] = __ldg(s + k);
}

// Pillow NEAREST resize (Geometry.c ImagingScaleAffine): dst[y][x] = src[ytab[y]][xtab[x]], single-byte pixels
__global__ void __launch_bounds__(256) resize_nearest_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int in_h,
                                                             int in_w, int out_h, int out_w, const int32_t *__restrict__ xtab,
                                                             const int32_t *__restrict__ ytab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out_w) return;
    const int sx = xtab[x], sy = ytab[y];
    const size_t page = blockIdx.z;
    const uint8_t v = (sx >= 0 && sy >= 0) ? __ldg(src + (page * in_h + sy) * (size_t)in_w + sx) : (uint8_t)0;
    dst[(page * out_h + y) * (size_t)out_w + x] = v;
}

}  // namespace lumina

using namespace lumina;

LUMINA_API int lumina_abi_version(void) { return 4; }  // 3: round 2 (JPEG decode, db_postprocess_ex, Otsu / Sauvola, skew estimate); 4: deskew decision from caller-computed angles, adaptive threshold dispatch modes, line-list copy
LUMINA_API const char *lumina_last_error_string(void) { return g_err; }
LUMINA_API uint64_t lumina_launch_count(void) { return g_launches.load(); }

// The first `lines` segments of every page of a HoughLinesP result [n][stride][4] int32 to (pinned) host memory
// [n][lines][4]: one strided copy on the stream (cudaMemcpy2DAsync), so that the few hundred segments a text page
// yields travel instead of the whole max_lines capacity -- and no library kernel packs them first.
LUMINA_API int lumina_copy_lines_to_host(const int32_t *d_lines, int n, int stride, int lines, int32_t *h_lines, void *stream) {
    LUMINA_REQUIRE(d_lines && h_lines, "null pointer");
    LUMINA_REQUIRE(n > 0 && stride > 0 && lines > 0 && lines <= stride, "bad line-list shape");
    LUMINA_CUDA_TRY(cudaMemcpy2DAsync(h_lines, (size_t)lines * 16, d_lines, (size_t)stride * 16, (size_t)lines * 16, (size_t)n,
                                      cudaMemcpyDeviceToHost, as_stream(stream)));
    return LUMINA_OK;
}

LUMINA_API void lumina_target_size(int width, int height, int max_dim, int *out_w, int *out_h) {
    // image_preprocessing.py:94-105 -- python float division then int() truncation
    if ((width > height ? width : height) <= max_dim) { *out_w = width; *out_h = height; return; }
    if (width > height) { *out_w = max_dim; *out_h = (int)(height * ((double)max_dim / width)); }
    else { *out_h = max_dim; *out_w = (int)(width * ((double)max_dim / height)); }
}

static int gray_launch(int mode, int post, const uint8_t *d_rgb, uint8_t *d_out, size_t npx, int thr, cudaStream_t st) {
    LUMINA_REQUIRE(d_rgb && d_out, "null pointer");
    if (npx == 0) return LUMINA_OK;
    LUMINA_REQUIRE((((uintptr_t)d_rgb) & 15) == 0 && (((uintptr_t)d_out) & 15) == 0, "buffers must be 16-byte aligned");
    size_t threads = (npx + 15) / 16;
    dim3 grid((unsigned)((threads + 255) / 256));
    if (mode == 0 && post == 0) gray16_kernel<0, 0><<<grid, 256, 0, st>>>(d_rgb, d_out, npx, thr);
    else if (mode == 1 && post == 0) gray16_kernel<1, 0><<<grid, 256, 0, st>>>(d_rgb, d_out, npx, thr);
    else gray16_kernel<0, 1><<<grid, 256, 0, st>>>(d_rgb, d_out, npx, thr);
    LUMINA_KERNEL_CHECK("gray16_kernel");
    return LUMINA_OK;
}

// Pillow resizes mode "P" / "1" images with NEAREST whatever filter is asked for (Image.resize); the source index of
// output x is (int)(a/2 + x*a) with the products accumulated by repeated addition in double (ImagingScaleAffine)
LUMINA_API void lumina_nearest_table_host(int in_size, int out_size, int32_t *h_tab) {
    const double a = (double)in_size / (double)out_size;
    double xo = 0.0 + a * 0.5;
    for (int x = 0; x < out_size; x++) {
        const int xin = xo < 0.0 ? -1 : (int)xo;
        h_tab[x] = (xin >= 0 && xin < in_size) ? xin : -1;
        xo += a;
    }
}

LUMINA_API int lumina_resize_nearest_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int in_h, int in_w, int out_h, int out_w,
                                        const int32_t *d_xtab, const int32_t *d_ytab, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst && d_xtab && d_ytab, "null pointer");
    LUMINA_REQUIRE(n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && out_h <= 65535 && n <= 65535, "bad geometry");
    dim3 grid(div_up(out_w, 256), out_h, n);
    resize_nearest_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_src, d_dst, in_h, in_w, out_h, out_w, d_xtab, d_ytab);
    LUMINA_KERNEL_CHECK("resize_nearest_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_rgb2gray_pil_u8(const uint8_t *d_rgb, uint8_t *d_gray, size_t npx, void *stream) {
    return gray_launch(0, 0, d_rgb, d_gray, npx, 0, as_stream(stream));
}
LUMINA_API int lumina_rgb2gray_cv_u8(const uint8_t *d_rgb, uint8_t *d_gray, size_t npx, void *stream) {
    return gray_launch(1, 0, d_rgb, d_gray, npx, 0, as_stream(stream));
}
// ---- ingest: Pillow keeps RGB images as 4 bytes per pixel (R, G, B, pad); drop the pad byte --------------
// 16 pixels per thread: four 128-bit loads, three 128-bit stores
__global__ void __launch_bounds__(256) rgbx_to_rgb16_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t npx) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p0 = t * 16;
    if (p0 >= npx) return;
    if (p0 + 16 <= npx) {
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 v = ldg_stream_u4(src + p0 * 4 + k * 16);
            w[k * 4] = v.x; w[k * 4 + 1] = v.y; w[k * 4 + 2] = v.z; w[k * 4 + 3] = v.w;
        }
        uint32_t o[12];
#pragma unroll
        for (int q = 0; q < 4; q++) {   // 4 pixels (r g b x)*4 -> 3 words
            const uint32_t a = w[q * 4], b = w[q * 4 + 1], c = w[q * 4 + 2], d = w[q * 4 + 3];
            o[q * 3 + 0] = __byte_perm(a, b, 0x4210);            // r0 g0 b0 r1
            o[q * 3 + 1] = __byte_perm(b, c, 0x5421);            // g1 b1 r2 g2
            o[q * 3 + 2] = __byte_perm(c, d, 0x6542);            // b2 r3 g3 b3
        }
#pragma unroll
        for (int k = 0; k < 3; k++) stg_stream_u4(dst + p0 * 3 + k * 16, make_uint4(o[k * 4], o[k * 4 + 1], o[k * 4 + 2], o[k * 4 + 3]));
    } else {
        for (size_t p = p0; p < npx; p++) {
            dst[p * 3] = src[p * 4]; dst[p * 3 + 1] = src[p * 4 + 1]; dst[p * 3 + 2] = src[p * 4 + 2];
        }
    }
}

// npx pixels of R,G,B,pad (Pillow's in-memory layout of mode "RGB") -> tightly packed R,G,B
LUMINA_API int lumina_rgbx_to_rgb_u8(const uint8_t *d_src, uint8_t *d_dst, size_t npx, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst, "null pointer");
    if (npx == 0) return LUMINA_OK;
    LUMINA_REQUIRE((((uintptr_t)d_src) & 15) == 0 && (((uintptr_t)d_dst) & 15) == 0, "buffers must be 16-byte aligned");
    const size_t threads = (npx + 15) / 16;
    rgbx_to_rgb16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(d_src, d_dst, npx);
    LUMINA_KERNEL_CHECK("rgbx_to_rgb16_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_binarize_u8(const uint8_t *d_src, uint8_t *d_dst, size_t npx, int c, int threshold, void *stream) {
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    if (c == 3) return gray_launch(0, 1, d_src, d_dst, npx, threshold, as_stream(stream));
    LUMINA_REQUIRE(d_src && d_dst, "null pointer");
    if (npx == 0) return LUMINA_OK;
    LUMINA_REQUIRE((((uintptr_t)d_src) & 15) == 0 && (((uintptr_t)d_dst) & 15) == 0, "buffers must be 16-byte aligned");
    size_t threads = (npx + 15) / 16;
    threshold16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(d_src, d_dst, npx, threshold);
    LUMINA_KERNEL_CHECK("threshold16_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_contrast_mean_u8(const uint8_t *d_src, int n, int h, int w, int c, uint64_t *d_sum_scratch,
                                       int32_t *d_mean, void *stream) {
    LUMINA_REQUIRE(d_src && d_sum_scratch && d_mean, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    cudaStream_t st = as_stream(stream);
    LUMINA_CUDA_TRY(cudaMemsetAsync(d_sum_scratch, 0, sizeof(uint64_t) * n, st));
    size_t px = (size_t)h * w;
    int bx = (int)((px / 16 + 255) / 256);
    int cap = (kNumSMs * 8 + n - 1) / n;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(bx, n);
    if (c == 3) lsum_kernel<3><<<grid, 256, 0, st>>>(d_src, px, (unsigned long long *)d_sum_scratch);
    else lsum_kernel<1><<<grid, 256, 0, st>>>(d_src, px, (unsigned long long *)d_sum_scratch);
    LUMINA_KERNEL_CHECK("lsum_kernel");
    mean_from_sum_kernel<<<(n + 127) / 128, 128, 0, st>>>((unsigned long long *)d_sum_scratch, px, n, d_mean);
    LUMINA_KERNEL_CHECK("mean_from_sum_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_contrast_apply_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                                        const int32_t *d_mean, float factor, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst && d_mean, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    size_t bytes = (size_t)h * w * c;
    int bx = (int)((bytes / 16 + 255) / 256);
    int cap = (kNumSMs * 8 + n - 1) / n;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    contrast_apply_kernel<<<dim3(bx, n), 256, 0, as_stream(stream)>>>(d_src, d_dst, bytes, d_mean, factor);
    LUMINA_KERNEL_CHECK("contrast_apply_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_exif_transpose_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                                        int orientation, void *stream) {
    LUMINA_REQUIRE(d_src && d_dst, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    cudaStream_t st = as_stream(stream);
    if (orientation < 2 || orientation > 8) {
        LUMINA_CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, (size_t)n * h * w * c, cudaMemcpyDeviceToDevice, st));
        return LUMINA_OK;
    }
    const int oh = orientation >= 5 ? w : h, ow = orientation >= 5 ? h : w;
    LUMINA_REQUIRE(oh <= 65535 && n <= 65535, "image too tall for grid");
    dim3 grid((ow + 255) / 256, oh, n);
    if (c == 3) exif_kernel<3><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, orientation);
    else exif_kernel<1><<<grid, 256, 0, st>>>(d_src, d_dst, h, w, orientation);
    LUMINA_KERNEL_CHECK("exif_kernel");
    return LUMINA_OK;
}
