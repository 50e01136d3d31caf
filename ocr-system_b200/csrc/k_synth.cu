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

}  // namespace lumina

using namespace lumina;

LUMINA_API int lumina_synth_pages_u8(uint8_t *d_dst, int n, int h, int w, uint64_t seed0, void *stream) {
    LUMINA_REQUIRE(d_dst, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    dim3 grid(div_up(w, 64), div_up(h, 4), n);
    LUMINA_REQUIRE(grid.y <= 65535 && n <= 65535, "image too large for grid");
    synth_pages_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_dst, h, w, seed0);
    LUMINA_KERNEL_CHECK("synth_pages_kernel");
    return LUMINA_OK;
}
