// k_canny.cu -- cv2.Canny(gray, low, high, apertureSize=3, L2gradient=False),
// bit-exact, with the RGB->BGR->GRAY conversion of deskew fused in front.
//
// Replaces OpenCV canny.cpp / color_rgb reached from
// backend/utils/image_preprocessing.py:394-399.  Arithmetic: SURVEY App. A2/A7.
//
//   canny_nms_kernel : tile in shared memory (gray +2 halo, |dx|+|dy| +1 halo),
//                      Sobel 3x3 replicate border, magnitude zero outside the
//                      image, NMS with the TG22 fixed-point sector test
//                      -> map {0 candidate, 1 suppressed, 2 strong}
//   hysteresis       : "candidate 8-connected to a strong pixel" is order
//                      independent, so it is a union-find labelling (ccl.cuh) of
//                      the non-suppressed pixels + a strong flag on each root.
#include "ccl.cuh"

namespace lumina {

constexpr int CN_TW = 64, CN_TH = 16;

template <int C>
__global__ void __launch_bounds__(256) canny_nms_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ map, int h,
                                                        int w, int low, int high) {
    __shared__ uint8_t g[CN_TH + 4][CN_TW + 4];
    __shared__ short sdx[CN_TH + 2][CN_TW + 2];
    __shared__ short sdy[CN_TH + 2][CN_TW + 2];
    const int page = blockIdx.z;
    const int x0 = blockIdx.x * CN_TW, y0 = blockIdx.y * CN_TH;
    const uint8_t *s = src + (size_t)page * h * w * C;
    for (int i = threadIdx.x; i < (CN_TH + 4) * (CN_TW + 4); i += 256) {
        const int ty = i / (CN_TW + 4), tx = i % (CN_TW + 4);
        const int yy = min(max(y0 + ty - 2, 0), h - 1), xx = min(max(x0 + tx - 2, 0), w - 1);
        const uint8_t *px = s + ((size_t)yy * w + xx) * C;
        uint32_t v;
        if (C == 3) v = (9798u * __ldg(px) + 19235u * __ldg(px + 1) + 3735u * __ldg(px + 2) + (1u << 14)) >> 15;
        else v = __ldg(px);
        g[ty][tx] = (uint8_t)v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (CN_TH + 2) * (CN_TW + 2); i += 256) {
        const int ty = i / (CN_TW + 2), tx = i % (CN_TW + 2);
        const int gy = ty + 1, gx = tx + 1;  // centre in g[][]
        const int dx = (g[gy - 1][gx + 1] + 2 * g[gy][gx + 1] + g[gy + 1][gx + 1]) -
                       (g[gy - 1][gx - 1] + 2 * g[gy][gx - 1] + g[gy + 1][gx - 1]);
        const int dy = (g[gy + 1][gx - 1] + 2 * g[gy + 1][gx] + g[gy + 1][gx + 1]) -
                       (g[gy - 1][gx - 1] + 2 * g[gy - 1][gx] + g[gy - 1][gx + 1]);
        const int yy = y0 + ty - 1, xx = x0 + tx - 1;
        const bool inside = yy >= 0 && yy < h && xx >= 0 && xx < w;
        sdx[ty][tx] = inside ? (short)dx : (short)0;  // magnitude outside the image is 0
        sdy[ty][tx] = inside ? (short)dy : (short)0;
    }
    __syncthreads();
    constexpr int TG22 = 13573;  // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
#define MAGAT(ty, tx) (abs((int)sdx[ty][tx]) + abs((int)sdy[ty][tx]))
    for (int i = threadIdx.x; i < CN_TH * CN_TW; i += 256) {
        const int ty = i / CN_TW, tx = i % CN_TW;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= w || y >= h) continue;
        const int cy = ty + 1, cx = tx + 1;
        const int xs = sdx[cy][cx], ys = sdy[cy][cx];
        const int m = abs(xs) + abs(ys);
        uint8_t r = 1;
        if (m > low) {
            const int ax = abs(xs), ay = abs(ys) << 15;
            const int tg22x = ax * TG22;
            bool pass;
            if (ay < tg22x) pass = m > MAGAT(cy, cx - 1) && m >= MAGAT(cy, cx + 1);
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) pass = m > MAGAT(cy - 1, cx) && m >= MAGAT(cy + 1, cx);
                else {
                    const int sgn = (xs ^ ys) < 0 ? -1 : 1;
                    pass = m > MAGAT(cy - 1, cx - sgn) && m > MAGAT(cy + 1, cx + sgn);
                }
            }
            if (pass) r = m > high ? 2 : 0;
        }
        map[((size_t)page * h + y) * w + x] = r;
    }
#undef MAGAT
}

struct CannyFG {
    const uint8_t *map;
    __device__ __forceinline__ bool operator()(size_t pbase, int idx) const { return map[pbase + idx] != 1; }
};

// flatten + raise the strong flag (bit 2) on the root's map byte.  All writers
// store the same value for a given root (root's own class | 4), so the plain
// byte store is a benign race.
__global__ void __launch_bounds__(256) canny_flatten_kernel(uint8_t *__restrict__ map, int *__restrict__ labels, long long hw,
                                                            long long total_px) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_px) return;
    const int page = (int)(g / hw);
    const int idx = (int)(g - (long long)page * hw);
    int *L = labels + (size_t)page * hw;
    if (L[idx] < 0) return;
    const int root = ccl_find(L, idx);
    L[idx] = root;
    uint8_t *mp = map + (size_t)page * hw;
    if ((mp[idx] & 3) == 2) {
        const uint8_t rv = mp[root];
        if (!(rv & 4)) mp[root] = (uint8_t)((rv & 3) | 4);
    }
}

__global__ void __launch_bounds__(256) canny_final_kernel(const uint8_t *__restrict__ map, const int *__restrict__ labels,
                                                          uint8_t *__restrict__ edges, long long hw, long long total_px) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_px) return;
    const int lab = labels[g];
    uint8_t e = 0;
    if (lab >= 0) {
        const long long page = g / hw;
        if (map[page * hw + lab] & 4) e = 255;
    }
    edges[g] = e;
}

// ---- 4 pixels per thread (any width; needs h*w % 4 == 0 so that pages start on a word) -------------------------
// The batch is one flat pixel stream; a thread owns 4 consecutive pixels (one 32-bit map load, one 128-bit label
// load / store).  Row runs are formed inside flat 32-pixel segments (8 lanes assemble the segment's bits with three
// xor-shuffles); a row start breaks a run.  Only the ~14 % non-suppressed pixels touch labels at all.
__device__ __forceinline__ uint32_t canny_fg4(uint32_t wv) {   // bit k = (byte k & 3) != 1
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) r |= ((((wv >> (8 * k)) & 3u) != 1u) ? 1u : 0u) << k;
    return r;
}

__global__ void __launch_bounds__(256) canny_init4_kernel(const uint8_t *__restrict__ map, int *__restrict__ labels, int w,
                                                          long long hw, long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = g < total_groups;
    const long long gidx = g * 4;                       // flat pixel index over the batch
    const int pidx0 = in ? (int)(gidx % hw) : 0;        // page-relative index of the first pixel
    uint32_t fg = 0, brk = 0;
    if (in) {
        fg = canny_fg4(*reinterpret_cast<const uint32_t *>(map + gidx));
        int x = pidx0 % w;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (x == 0) brk |= 1u << k;
            if (++x == w) x = 0;
        }
    }
    const int sub = (int)(g & 7);
    uint32_t sfg = fg << (4 * sub), sbrk = brk << (4 * sub);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        sfg |= __shfl_xor_sync(0xffffffffu, sfg, o);
        sbrk |= __shfl_xor_sync(0xffffffffu, sbrk, o);
    }
    if (!in) return;
    // run starts: position 0, a row start, or the pixel behind a suppressed one
    const uint32_t starts = (~sfg << 1) | 1u | sbrk;
    int out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int pos = 4 * sub + k;
        if ((sfg >> pos) & 1u) {
            const uint32_t below = starts & ((2u << pos) - 1u);
            const int start = 31 - __clz(below);
            out[k] = pidx0 + k - (pos - start);
        } else {
            out[k] = -1;
        }
    }
    *reinterpret_cast<int4 *>(labels + gidx) = make_int4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) canny_merge4_kernel(const uint8_t *__restrict__ map, int *__restrict__ labels, int h,
                                                           int w, long long hw, long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const long long gidx = g * 4;
    const uint32_t fg = canny_fg4(*reinterpret_cast<const uint32_t *>(map + gidx));
    if (!fg) return;
    const long long page = gidx / hw;
    const int pidx0 = (int)(gidx - page * hw);
    const uint8_t *M = map + page * hw;
    int *L = labels + page * hw;
    const bool seg_start = (g & 7) == 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (!((fg >> k) & 1u)) continue;
        const int idx = pidx0 + k;
        const int y = idx / w, x = idx - y * w;
        const bool left_fg = x > 0 && (k > 0 ? ((fg >> (k - 1)) & 1u) != 0u : (M[idx - 1] & 3) != 1);
        // runs are pre-linked inside a flat 32-pixel segment: the left link is due only at its first pixel
        if (k == 0 && seg_start && left_fg) ccl_union(L, idx, idx - 1);
        if (y > 0) {
            const int up = idx - w;
            if ((M[up] & 3) != 1) {
                const bool upleft_fg = x > 0 && (M[up - 1] & 3) != 1;
                if (!(left_fg && upleft_fg)) ccl_union(L, idx, up);
            } else {
                if (x > 0 && (M[up - 1] & 3) != 1) ccl_union(L, idx, up - 1);
                if (x + 1 < w && (M[up + 1] & 3) != 1) ccl_union(L, idx, up + 1);
            }
        }
    }
}

__global__ void __launch_bounds__(256) canny_flatten4_kernel(uint8_t *__restrict__ map, int *__restrict__ labels, long long hw,
                                                             long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const long long gidx = g * 4;
    const int4 lv = *reinterpret_cast<const int4 *>(labels + gidx);
    if (lv.x < 0 && lv.y < 0 && lv.z < 0 && lv.w < 0) return;   // all suppressed
    const long long page = gidx / hw;
    const int pidx0 = (int)(gidx - page * hw);
    uint8_t *mp = map + page * hw;
    int *L = labels + page * hw;
    const uint32_t mw = *reinterpret_cast<const uint32_t *>(map + gidx);
    const int lab[4] = {lv.x, lv.y, lv.z, lv.w};
    int out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        out[k] = lab[k];
        if (lab[k] < 0) continue;
        if (k > 0 && lab[k] == lab[k - 1]) out[k] = out[k - 1];
        else {
            int a = lab[k], p = a == pidx0 + k ? a : L[a];
            while (p != a) { a = p; p = L[a]; }
            out[k] = a;
        }
        if (((mw >> (8 * k)) & 3u) == 2u) {   // strong pixel: raise the flag on the root (same value from every writer)
            const uint8_t rv = mp[out[k]];
            if (!(rv & 4)) mp[out[k]] = (uint8_t)((rv & 3) | 4);
        }
    }
    *reinterpret_cast<int4 *>(labels + gidx) = make_int4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) canny_final4_kernel(const uint8_t *__restrict__ map, const int *__restrict__ labels,
                                                           uint8_t *__restrict__ edges, long long hw, long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const long long gidx = g * 4;
    const int4 lv = *reinterpret_cast<const int4 *>(labels + gidx);
    const int lab[4] = {lv.x, lv.y, lv.z, lv.w};
    uint32_t e = 0;
    if (lv.x >= 0 || lv.y >= 0 || lv.z >= 0 || lv.w >= 0) {
        const uint8_t *mp = map + (gidx / hw) * hw;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (lab[k] >= 0 && (mp[lab[k]] & 4)) e |= 0xFFu << (8 * k);
    }
    *reinterpret_cast<uint32_t *>(edges + gidx) = e;
}

}  // namespace lumina

using namespace lumina;

LUMINA_API size_t lumina_canny_workspace_bytes(int n, int h, int w) {
    const size_t px = (size_t)n * h * w;
    return ((px + 255) & ~(size_t)255) + px * sizeof(int);  // map (u8) + labels (int32)
}

LUMINA_API int lumina_canny_u8(const uint8_t *d_src, uint8_t *d_edges, int n, int h, int w, int c, int low, int high,
                               void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(d_src && d_edges && d_workspace, "null pointer");
    LUMINA_REQUIRE(c == 1 || c == 3, "c must be 1 or 3");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE((long long)h * w < (1LL << 31), "page too large");
    const size_t need = lumina_canny_workspace_bytes(n, h, w);
    if (workspace_bytes < need) return set_error(LUMINA_E_NOMEM, "canny workspace too small: need %zu bytes", need);
    LUMINA_REQUIRE((((uintptr_t)d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    const size_t px = (size_t)n * h * w;
    uint8_t *map = (uint8_t *)d_workspace;
    int *labels = (int *)((uint8_t *)d_workspace + ((px + 255) & ~(size_t)255));
    dim3 grid(div_up(w, CN_TW), div_up(h, CN_TH), n);
    LUMINA_REQUIRE(grid.y <= 65535 && n <= 65535, "image too large for grid");
    if (c == 3) canny_nms_kernel<3><<<grid, 256, 0, st>>>(d_src, map, h, w, low, high);
    else canny_nms_kernel<1><<<grid, 256, 0, st>>>(d_src, map, h, w, low, high);
    LUMINA_KERNEL_CHECK("canny_nms_kernel");
    const long long hw = (long long)h * w;
    if ((hw & 3) == 0 && (((uintptr_t)d_edges) & 3) == 0) {
        const long long groups = (long long)px / 4;
        const unsigned gg = (unsigned)((groups + 255) / 256);
        canny_init4_kernel<<<gg, 256, 0, st>>>(map, labels, w, hw, groups);
        LUMINA_KERNEL_CHECK("canny_init4_kernel");
        canny_merge4_kernel<<<gg, 256, 0, st>>>(map, labels, h, w, hw, groups);
        LUMINA_KERNEL_CHECK("canny_merge4_kernel");
        canny_flatten4_kernel<<<gg, 256, 0, st>>>(map, labels, hw, groups);
        LUMINA_KERNEL_CHECK("canny_flatten4_kernel");
        canny_final4_kernel<<<gg, 256, 0, st>>>(map, labels, d_edges, hw, groups);
        LUMINA_KERNEL_CHECK("canny_final4_kernel");
        return LUMINA_OK;
    }
    CannyFG fg{map};
    const long long nseg = (long long)n * h * ((w + 31) / 32);
    LUMINA_REQUIRE(nseg < (1LL << 31), "batch too large");
    ccl_init_rows_kernel<CannyFG><<<(unsigned)((nseg * 32 + 255) / 256), 256, 0, st>>>(fg, labels, h, w, (int)nseg);
    LUMINA_KERNEL_CHECK("ccl_init_rows_kernel<canny>");
    const unsigned gpx = (unsigned)((px + 255) / 256);
    ccl_merge_kernel<CannyFG><<<gpx, 256, 0, st>>>(fg, labels, h, w, (long long)px);
    LUMINA_KERNEL_CHECK("ccl_merge_kernel<canny>");
    canny_flatten_kernel<<<gpx, 256, 0, st>>>(map, labels, hw, (long long)px);
    LUMINA_KERNEL_CHECK("canny_flatten_kernel");
    canny_final_kernel<<<gpx, 256, 0, st>>>(map, labels, d_edges, hw, (long long)px);
    LUMINA_KERNEL_CHECK("canny_final_kernel");
    return LUMINA_OK;
}
