// k_synth.cu -- synthetic A4 text pages generated directly in HBM (bench / test
// workload; the reference ships no fixtures).  Device build of
// include/lumina_synth.h: integer arithmetic only, so the bytes are identical to
// the host build used by the CPU oracle.
#include "common.cuh"
#include "../../include/lumina_synth.h"

namespace lumina {

__global__ void __launch_bounds__(256) synth_pages_kernel(uint8_t *__restrict__ dst, int h, int w, uint64_t seed0) {
    const int page = blockIdx.z;
    __shared__ lsyn_page_t pg;
    if (threadIdx.x == 0) lsyn_page_init(&pg, h, w, seed0 + (uint64_t)page);
    __syncthreads();
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    uint8_t *o = dst + ((size_t)page * h * w + (size_t)y * w + x) * 3;
    o[0] = lsyn_pixel(&pg, x, y, 0);
    o[1] = lsyn_pixel(&pg, x, y, 1);
    o[2] = lsyn_pixel(&pg, x, y, 2);
}

__global__ void __launch_bounds__(256) synth_prob_kernel(float *__restrict__ dst, int h, int w, uint64_t seed0) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, m = blockIdx.z;
    if (x >= w) return;
    const uint64_t s64 = seed0 + (uint64_t)m;
    const uint32_t seed = (uint32_t)(s64 ^ (s64 >> 32)) * 2654435761u + 777u;
    dst[((size_t)m * h + y) * w + x] = lsyn_prob(w, seed, x, y);
}

// one CTA per (crop, step) row of C classes
__global__ void __launch_bounds__(256) synth_ctc_kernel(float *__restrict__ dst, int T, int C, uint64_t n0, uint32_t seed) {
    const size_t row = blockIdx.x;
    const uint32_t n = (uint32_t)(n0 + row / T);
    const int t = (int)(row % T);
    uint32_t win, tie;
    lsyn_ctc_step(seed, C, n, t, &win, &tie);
    float *o = dst + row * (size_t)C;
    for (int c = threadIdx.x; c < C; c += 256) o[c] = lsyn_ctc_value(seed, T, n, t, (uint32_t)c, win, tie);
}

}  // namespace lumina

using namespace lumina;

LUMINA_API int lumina_synth_prob_maps_f32(float *d_dst, int n, int h, int w, uint64_t seed0, void *stream) {
    LUMINA_REQUIRE(d_dst, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && h <= 65535 && n <= 65535, "bad geometry");
    synth_prob_kernel<<<dim3(div_up(w, 256), h, n), 256, 0, as_stream(stream)>>>(d_dst, h, w, seed0);
    LUMINA_KERNEL_CHECK("synth_prob_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_synth_ctc_f32(float *d_dst, int n, int t, int c, uint64_t crop0, uint32_t seed, void *stream) {
    LUMINA_REQUIRE(d_dst, "null pointer");
    LUMINA_REQUIRE(n > 0 && t > 0 && c > 1 && (long long)n * t < (1LL << 31), "bad geometry");
    synth_ctc_kernel<<<(unsigned)((size_t)n * t), 256, 0, as_stream(stream)>>>(d_dst, t, c, crop0, seed);
    LUMINA_KERNEL_CHECK("synth_ctc_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_synth_pages_u8(uint8_t *d_dst, int n, int h, int w, uint64_t seed0, void *stream) {
    LUMINA_REQUIRE(d_dst, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    dim3 grid(div_up(w, 64), div_up(h, 4), n);
    LUMINA_REQUIRE(grid.y <= 65535 && n <= 65535, "image too large for grid");
    synth_pages_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_dst, h, w, seed0);
    LUMINA_KERNEL_CHECK("synth_pages_kernel");
    return LUMINA_OK;
}
