// ccl.cuh -- 8-connected union-find labelling on a page plane, shared by the
// Canny hysteresis (k_canny.cu) and the DB mask labelling (k_dbpost.cu).
//
// Labels are page-relative pixel indices (int32); a component's final label is
// the minimum raster index of its pixels (canonical form the parity tests use).
//   pass 1  ccl_init_rows : warp ballot over 32 consecutive pixels of a row; every
//           foreground pixel starts labelled with the first pixel of its run
//           inside the 32-pixel segment (no atomics, no pointer chasing).
//   pass 2  ccl_merge     : run heads union with the left segment and with the row
//           above (N, else NW/NE) through atomicMin on the roots.
//   pass 3  caller-specific flatten (find root, write final label).
#pragma once
#include "common.cuh"

namespace lumina {

__device__ __forceinline__ int ccl_find(const int *__restrict__ L, int a) {
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}

__device__ __forceinline__ void ccl_union(int *L, int a, int b) {
    bool done;
    do {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a < b) {
            int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else done = true;
    } while (!done);
}

// FG(page_ptr, idx) -> bool.  grid: (ceil(w/32) * h / warps_per_block ...) flattened.
// One warp per 32-pixel row segment.
template <typename FG>
__global__ void __launch_bounds__(256) ccl_init_rows_kernel(FG fg, int *__restrict__ labels, int h, int w, int nseg_total) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= nseg_total) return;
    const int segs_per_row = (w + 31) >> 5;
    const long long row_g = gw / segs_per_row;  // global row over the batch
    const int seg = (int)(gw - row_g * segs_per_row);
    const int page = (int)(row_g / h), y = (int)(row_g - (long long)page * h);
    const int x = seg * 32 + lane;
    const size_t pbase = (size_t)page * h * w;
    const bool in = x < w;
    const int idx = y * w + x;
    const bool f = in && fg(pbase, idx);
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (in) {
        int lab = -1;
        if (f) {
            // first lane of the run containing `lane`: highest zero bit below lane, +1
            const unsigned below = ~m & ((1u << lane) - 1u);
            const int start = below ? 32 - __clz(below) : 0;
            lab = idx - (lane - start);
        }
        labels[pbase + idx] = lab;
    }
}

template <typename FG>
__global__ void __launch_bounds__(256) ccl_merge_kernel(FG fg, int *__restrict__ labels, int h, int w, long long total_px) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_px) return;
    const long long hw = (long long)h * w;
    const int page = (int)(g / hw);
    const int idx = (int)(g - (long long)page * hw);
    const size_t pbase = (size_t)page * hw;
    int *L = labels + pbase;
    if (L[idx] < 0) return;
    const int y = idx / w, x = idx - y * w;
    // left neighbour: only needed at a 32-pixel segment boundary (runs inside a segment are pre-linked)
    if ((x & 31) == 0 && x > 0 && fg(pbase, idx - 1)) ccl_union(L, idx, idx - 1);
    if (y > 0) {
        const int up = idx - w;
        if (fg(pbase, up)) {
            // N present: only the run head (or a pixel whose left neighbour is background above) needs to link;
            // linking every pixel is correct, restrict to cut atomics: link when left pixel is bg or above-left is bg
            const bool left_fg = x > 0 && L[idx - 1] >= 0;
            const bool upleft_fg = x > 0 && fg(pbase, up - 1);
            if (!(left_fg && upleft_fg)) ccl_union(L, idx, up);
        } else {
            if (x > 0 && fg(pbase, up - 1)) ccl_union(L, idx, up - 1);
            if (x + 1 < w && fg(pbase, up + 1)) ccl_union(L, idx, up + 1);
        }
    }
}

}  // namespace lumina
