// k_ctc.cu -- [upstream PaddleOCR] CTCLabelDecode (greedy): argmax + max over
// classes per time step, collapse repeats, drop blank (0), confidence = float32
// mean of the kept max-probabilities (numpy pairwise order).
//
// Not in the reference tree.  HBM-bound: the [N,T,C]
// posterior tensor is read exactly once with 128-bit loads, one warp per (n,t)
// row, arg-max by warp shuffles with numpy's first-index tie rule (NaN counts as
// the maximum, first NaN wins); a second tiny kernel collapses each sequence.
#include "common.cuh"

namespace lumina {

__device__ __forceinline__ bool ctc_better(float v, int i, float bv, int bi) {
    // "v at index i beats (bv, bi)" under numpy argmax: NaN is maximal, ties -> lower index
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (vn) return i < bi;
    return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256) ctc_argmax_kernel(const float *__restrict__ probs, int rows, int c,
                                                         int *__restrict__ arg, float *__restrict__ val) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *p = probs + (size_t)row * c;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    // peel to 16-byte alignment
    const int mis = (int)(((uintptr_t)p >> 2) & 3);
    const int head = min(c, (4 - mis) & 3);
    if (lane < head) { bv = p[lane]; bi = lane; }
    const int nvec = (c - head) >> 2;
    const float4 *pv = reinterpret_cast<const float4 *>(p + head);
    // 8 independent 128-bit loads in flight per lane before the (dependent) compare chain runs
    int i = lane;
    for (; i + 7 * 32 < nvec; i += 8 * 32) {
        uint4 u[8];
#pragma unroll
        for (int q = 0; q < 8; q++) u[q] = ldg_stream_u4(pv + i + q * 32);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float f[4] = {__uint_as_float(u[q].x), __uint_as_float(u[q].y), __uint_as_float(u[q].z), __uint_as_float(u[q].w)};
            const int base = head + (i + q * 32) * 4;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (ctc_better(f[k], base + k, bv, bi)) { bv = f[k]; bi = base + k; }
        }
    }
    for (; i < nvec; i += 32) {
        const uint4 u = ldg_stream_u4(pv + i);
        const float f[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
        const int base = head + i * 4;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (ctc_better(f[k], base + k, bv, bi)) { bv = f[k]; bi = base + k; }
    }
    for (int i = head + nvec * 4 + lane; i < c; i += 32) {
        const float f = p[i];
        if (ctc_better(f, i, bv, bi)) { bv = f; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, bv, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ctc_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { arg[row] = bi; val[row] = bv; }
}

// numpy float32 add.reduce (pairwise): n < 8 sequential; else 8 accumulators over
// blocks, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), remainder sequential;
// recursive halving above 128 elements.
__device__ float np_pairwise_sum(const float *a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; i++) r = __fadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int q = 0; q < 8; q++) r[q] = a[q];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int q = 0; q < 8; q++) r[q] = __fadd_rn(r[q], a[i + q]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

__global__ void __launch_bounds__(128) ctc_collapse_kernel(const int *__restrict__ arg, float *__restrict__ val, int n, int t,
                                                           int *__restrict__ idx_out, int *__restrict__ pos_out,
                                                           int *__restrict__ len_out, float *__restrict__ conf_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int *a = arg + (size_t)b * t;
    float *v = val + (size_t)b * t;  // compacted in place (kept values move left)
    int *io = idx_out + (size_t)b * t, *po = pos_out + (size_t)b * t;
    int prev = -1, len = 0;
    for (int s = 0; s < t; s++) {
        const int cur = a[s];
        const bool keep = (s == 0 || cur != prev) && cur != 0;
        prev = cur;
        if (keep) { io[len] = cur; po[len] = s; v[len] = v[s]; len++; }
    }
    for (int s = len; s < t; s++) { io[s] = -1; po[s] = -1; }
    len_out[b] = len;
    conf_out[b] = len ? __fdiv_rn(np_pairwise_sum(v, len), (float)len) : 0.0f;
}

}  // namespace lumina

using namespace lumina;

LUMINA_API size_t lumina_ctc_workspace_bytes(int n, int t) { return (size_t)n * t * 8 + 256; }

LUMINA_API int lumina_ctc_greedy(const float *d_probs, int n, int t, int c, int32_t *d_idx, int32_t *d_pos, int32_t *d_len,
                                 float *d_conf, void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(d_probs && d_idx && d_pos && d_len && d_conf && d_workspace, "null pointer");
    LUMINA_REQUIRE(n > 0 && t > 0 && c > 0, "empty batch");
    LUMINA_REQUIRE((((uintptr_t)d_probs) & 3) == 0, "probs must be 4-byte aligned");
    const size_t need = lumina_ctc_workspace_bytes(n, t);
    if (workspace_bytes < need) return set_error(LUMINA_E_NOMEM, "ctc workspace too small: need %zu bytes", need);
    cudaStream_t st = as_stream(stream);
    const long long rows = (long long)n * t;
    LUMINA_REQUIRE(rows < (1LL << 31), "batch too large");
    int *arg = (int *)d_workspace;
    float *val = (float *)((uint8_t *)d_workspace + (((size_t)rows * 4 + 127) & ~(size_t)127));
    ctc_argmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(d_probs, (int)rows, c, arg, val);
    LUMINA_KERNEL_CHECK("ctc_argmax_kernel");
    ctc_collapse_kernel<<<(n + 127) / 128, 128, 0, st>>>(arg, val, n, t, d_idx, d_pos, d_len, d_conf);
    LUMINA_KERNEL_CHECK("ctc_collapse_kernel");
    return LUMINA_OK;
}
