// common.cuh -- shared host/device helpers for the lumina_b200 C-ABI library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/lumina_b200.h"

#define LUMINA_API extern "C" __attribute__((visibility("default")))

namespace lumina {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int set_error(int code, const char *fmt, ...);

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Count + check a kernel launch.  Used right after every <<<>>>.
#define LUMINA_KERNEL_CHECK(name)                                                                  \
    do {                                                                                           \
        ::lumina::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess)                                                                    \
            return ::lumina::set_error(LUMINA_E_CUDA, "%s launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

#define LUMINA_CUDA_TRY(expr)                                                                      \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return ::lumina::set_error(LUMINA_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__));   \
    } while (0)

#define LUMINA_REQUIRE(cond, msg)                                                                  \
    do {                                                                                           \
        if (!(cond)) return ::lumina::set_error(LUMINA_E_INVALID, "%s (%s)", msg, #cond);          \
    } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers -------------------------------------------------------
#ifdef __CUDACC__
// streaming 128-bit / 32-bit loads that do not pollute L1 (read-once data)
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return (w >> (8 * k)) & 0xffu; }
__device__ __forceinline__ uint32_t sat_u8(int v) { return (uint32_t)min(max(v, 0), 255); }
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return a | (b << 8) | (c << 16) | (d << 24);
}
#endif

}  // namespace lumina
