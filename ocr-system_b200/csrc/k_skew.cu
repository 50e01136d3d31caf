// k_skew.cu -- fast skew-angle estimator (projection profiles), the flag-gated alternative to the exact deskew.
//
// Reference: backend/utils/image_preprocessing.py:372-460 (deskew).  The product path reproduces the reference's
// angle exactly (Canny + cv2.HoughLinesP + median of the folded segment angles: k_canny.cu, k_ppht.cu) and that
// exact replica of a sequential randomised algorithm is 80 % of the device step.  BASELINE.json's north_star also
// names "warp-shuffle histogram and projection-profile reductions for deskew angle search"; SURVEY 7.3 #1 asks to
// ship that as a second, tolerance-certified mode.  This is it -- NEVER the default: any angle difference changes
// every byte of the rotated raster, so this mode is certified on the angle only (tests: |angle - reference angle|
// and agreement on the reference's 0.5-degree gate over 1024 synthetic pages).
//
// Estimator: for a candidate angle a, edge pixel (x, y) falls into profile bin round(y cos a - x sin a): pixels of
// one text line (direction (cos a, sin a)) share a bin when a is the skew.  score(a) = sum of squared bin counts.
//   coarse pass: a = -45 .. 45 degrees in 0.5-degree steps;  fine pass: 0.02-degree steps around the coarse
//   maximum, then a parabola through the three best scores.
// One CTA per (page, group of 4 angles): the four profiles live in shared memory (shared-memory atomics), the edge
// map is read with 128-bit loads (it stays L2 resident: 0.65 MB per page), scores are reduced with warp shuffles.
#include "common.cuh"

namespace lumina {

constexpr int SK_ANGLES_PER_CTA = 4;
constexpr int SK_COARSE = 181;        // -45 .. 45 step 0.5
constexpr int SK_FINE = 61;           // +-0.6 degree in 0.02 steps around the coarse maximum
constexpr double SK_COARSE_STEP = 0.5, SK_FINE_STEP = 0.02;

__device__ __forceinline__ double sk_coarse_angle(int i) { return -45.0 + SK_COARSE_STEP * i; }

__device__ int sk_argmax(const unsigned long long *s, int n) {
    int b = 0;
    for (int i = 1; i < n; i++)
        if (s[i] > s[b]) b = i;
    return b;
}

// edge map -> one bit per pixel, rows padded to whole 32-bit words (one warp per word)
__global__ void __launch_bounds__(256) skew_bitmask_kernel(const uint8_t *__restrict__ edges, uint32_t *__restrict__ bits, int w,
                                                           int wpr, long long total_words) {
    const long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (word >= total_words) return;
    const int lane = threadIdx.x & 31;
    const long long row = word / wpr;                 // global row over the batch
    const int x = (int)(word - row * wpr) * 32 + lane;
    const bool e = x < w && edges[row * w + x] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, e);
    if (lane == 0) bits[word] = m;
}

// fine == 0: coarse angles; fine == 1: angles around the coarse maximum of `coarse_scores`.
// A warp walks tiles of 32 rows x 32 columns, one row per lane (one 32-bit word of the bitmask): the lanes of a
// warp then hit different profile bins (bin ~ row for small angles), so the shared-memory atomics do not collide,
// and the pixels of a lane's segment that fall into the same bin are counted in registers first.
__global__ void __launch_bounds__(256) skew_profile_kernel(const uint32_t *__restrict__ bits, int h, int w, int wpr, int fine,
                                                           const unsigned long long *__restrict__ coarse_scores,
                                                           unsigned long long *__restrict__ scores) {
    extern __shared__ unsigned int sk_prof[];     // [4][nbins]
    const int page = blockIdx.y;
    const int nang = fine ? SK_FINE : SK_COARSE;
    const int a0 = blockIdx.x * SK_ANGLES_PER_CTA;
    const int nbins = h + 2 * w + 3;               // y cos a - x sin a lies in [-w, h + w] for |a| <= 45 degrees
    const int off = w + 1;
    for (int i = threadIdx.x; i < SK_ANGLES_PER_CTA * nbins; i += 256) sk_prof[i] = 0u;
    double centre = 0.0;
    if (fine) centre = sk_coarse_angle(sk_argmax(coarse_scores + (size_t)page * SK_COARSE, SK_COARSE));
    float cs[SK_ANGLES_PER_CTA], sn[SK_ANGLES_PER_CTA];
#pragma unroll
    for (int k = 0; k < SK_ANGLES_PER_CTA; k++) {
        const int ai = min(a0 + k, nang - 1);
        const double deg = fine ? centre + SK_FINE_STEP * (ai - SK_FINE / 2) : sk_coarse_angle(ai);
        const double rad = deg * 3.14159265358979323846 / 180.0;
        cs[k] = (float)cos(rad);
        sn[k] = (float)sin(rad);
    }
    __syncthreads();
    const uint32_t *B = bits + (size_t)page * h * wpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles_y = (h + 31) >> 5, ntiles = tiles_y * wpr;
    // the coarse pass only has to find the right half-degree cell: it looks at every fourth 32-pixel COLUMN band
    // (whole rows of bins stay populated; cutting row bands instead would favour 0 degrees, where the cut is a bin edge)
    const int tx_step = fine ? 1 : 4;
    for (int t = warp; t < ntiles; t += 8) {
        const int ty = t / wpr, tx = t - ty * wpr;
        if (tx % tx_step) continue;
        const int y = ty * 32 + lane;
        uint32_t m = y < h ? __ldg(B + (size_t)y * wpr + tx) : 0u;
        if (!m) continue;
        const float fy = (float)y;
        int cur[SK_ANGLES_PER_CTA], cnt[SK_ANGLES_PER_CTA];
#pragma unroll
        for (int k = 0; k < SK_ANGLES_PER_CTA; k++) { cur[k] = -1; cnt[k] = 0; }
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float fx = (float)(tx * 32 + j);
#pragma unroll
            for (int k = 0; k < SK_ANGLES_PER_CTA; k++) {
                const int b = __float2int_rn(fy * cs[k] - fx * sn[k]) + off;
                if (b != cur[k]) {
                    if (cnt[k]) atomicAdd(&sk_prof[k * nbins + cur[k]], (unsigned int)cnt[k]);
                    cur[k] = b;
                    cnt[k] = 0;
                }
                cnt[k]++;
            }
        }
#pragma unroll
        for (int k = 0; k < SK_ANGLES_PER_CTA; k++)
            if (cnt[k]) atomicAdd(&sk_prof[k * nbins + cur[k]], (unsigned int)cnt[k]);
    }
    __syncthreads();
    __shared__ unsigned long long s_part[8];
    for (int k = 0; k < SK_ANGLES_PER_CTA; k++) {
        unsigned long long acc = 0;
        for (int i = threadIdx.x; i < nbins; i += 256) {
            const unsigned long long c = sk_prof[k * nbins + i];
            acc += c * c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0 && a0 + k < nang) {
            unsigned long long tt = 0;
            for (int i = 0; i < 8; i++) tt += s_part[i];
            scores[(size_t)page * nang + a0 + k] = tt;
        }
        __syncthreads();
    }
}

__global__ void skew_angle_kernel(const unsigned long long *__restrict__ coarse, const unsigned long long *__restrict__ fine_s,
                                  int n, double *__restrict__ angles) {
    const int page = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n) return;
    const double centre = sk_coarse_angle(sk_argmax(coarse + (size_t)page * SK_COARSE, SK_COARSE));
    const unsigned long long *f = fine_s + (size_t)page * SK_FINE;
    const int b = sk_argmax(f, SK_FINE);
    double a = centre + SK_FINE_STEP * (b - SK_FINE / 2);
    if (b > 0 && b < SK_FINE - 1) {   // parabola through the three scores around the maximum
        const double y0 = (double)f[b - 1], y1 = (double)f[b], y2 = (double)f[b + 1];
        const double den = y0 - 2.0 * y1 + y2;
        if (den < 0.0) a += SK_FINE_STEP * 0.5 * (y0 - y2) / den;
    }
    angles[page] = a;
}

}  // namespace lumina

using namespace lumina;

static size_t sk_scores_bytes(int n) { return ((size_t)n * (SK_COARSE + SK_FINE) * 8 + 255) & ~(size_t)255; }

LUMINA_API size_t lumina_skew_workspace_bytes_for(int n, int h, int w) {
    return n > 0 && h > 0 && w > 0 ? sk_scores_bytes(n) + (size_t)n * h * ((w + 31) / 32) * 4 + 256 : 0;
}

LUMINA_API int lumina_skew_estimate_fast(const uint8_t *d_edges, int n, int h, int w, double *d_angles, void *d_workspace,
                                         size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(d_edges && d_angles && d_workspace, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && n <= 65535, "bad batch");
    LUMINA_REQUIRE(workspace_bytes >= lumina_skew_workspace_bytes_for(n, h, w), "skew workspace too small");
    const size_t smem = (size_t)SK_ANGLES_PER_CTA * (h + 2 * w + 3) * 4;
    LUMINA_REQUIRE(smem <= 200 * 1024, "page too large for the shared-memory profiles");
    cudaStream_t st = as_stream(stream);
    unsigned long long *coarse = reinterpret_cast<unsigned long long *>(d_workspace);
    unsigned long long *fine = coarse + (size_t)n * SK_COARSE;
    uint32_t *bits = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(d_workspace) + sk_scores_bytes(n));
    const int wpr = (w + 31) / 32;
    const long long words = (long long)n * h * wpr;
    skew_bitmask_kernel<<<(unsigned)((words * 32 + 255) / 256), 256, 0, st>>>(d_edges, bits, w, wpr, words);
    LUMINA_KERNEL_CHECK("skew_bitmask_kernel");
    LUMINA_CUDA_TRY(cudaFuncSetAttribute(skew_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    skew_profile_kernel<<<dim3(div_up(SK_COARSE, SK_ANGLES_PER_CTA), n), 256, smem, st>>>(bits, h, w, wpr, 0, nullptr, coarse);
    LUMINA_KERNEL_CHECK("skew_profile_kernel");
    skew_profile_kernel<<<dim3(div_up(SK_FINE, SK_ANGLES_PER_CTA), n), 256, smem, st>>>(bits, h, w, wpr, 1, coarse, fine);
    LUMINA_KERNEL_CHECK("skew_profile_kernel");
    skew_angle_kernel<<<(n + 63) / 64, 64, 0, st>>>(coarse, fine, n, d_angles);
    LUMINA_KERNEL_CHECK("skew_angle_kernel");
    return LUMINA_OK;
}
