#include <cstdio>
#include <cuda_runtime.h>
#include "../../ocr-system_b200/csrc/db_geom.h"
__global__ void k(const float *bx, const float *by, int w, int h, int *out) {
    const int lane = threadIdx.x & 31;
    float bxs[4], bys[4];
    for (int i = 0; i < 4; i++) { bxs[i] = bx[i]; bys[i] = by[i]; }
    const float fminx = fminf(fminf(bxs[0], bxs[1]), fminf(bxs[2], bxs[3]));
    const float fmaxx = fmaxf(fmaxf(bxs[0], bxs[1]), fmaxf(bxs[2], bxs[3]));
    const float fminy = fminf(fminf(bys[0], bys[1]), fminf(bys[2], bys[3]));
    const float fmaxy = fmaxf(fmaxf(bys[0], bys[1]), fmaxf(bys[2], bys[3]));
    const int xmin = min(max((int)floorf(fminx), 0), w - 1), xmax = min(max((int)ceilf(fmaxx), 0), w - 1);
    const int ymin = min(max((int)floorf(fminy), 0), h - 1), ymax = min(max((int)ceilf(fmaxy), 0), h - 1);
    DbgPt q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { q[i].x = (int)(bxs[i] - (float)xmin); q[i].y = (int)(bys[i] - (float)ymin); }
    const int mh = ymax - ymin + 1, mw = xmax - xmin + 1;
    double sum = 0.0;
    int cnt = 0;
    for (int ry = 0; ry < mh; ry++) {
        int lo[5], hi[5];
        int c = dbg_row_cover(q, ry, lo, hi);
        c = dbg_merge(lo, hi, c);
        for (int i = 0; i < c; i++) {
            const int a = max(lo[i], 0), b = min(hi[i], mw - 1);
            for (int x = a + lane; x <= b; x += 32) { sum += 1.0; cnt++; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { out[0] = cnt; out[1] = (int)sum; out[2] = xmin; out[3] = ymin; out[4] = mw; out[5] = mh; for (int i = 0; i < 4; i++) { out[6+2*i] = q[i].x; out[7+2*i] = q[i].y; } }
}
int main() {
    float bx[4] = {443.2399597167969f, 465.6399841308594f, 467.2999572753906f, 444.8999328613281f};
    float by[4] = {557.6799926757812f, 554.4800415039062f, 566.1000366210938f, 569.2999877929688f};
    float *dx, *dy; int *d; cudaMalloc(&dx, 16); cudaMalloc(&dy, 16); cudaMalloc(&d, 64 * 4);
    cudaMemcpy(dx, bx, 16, cudaMemcpyHostToDevice); cudaMemcpy(dy, by, 16, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(dx, dy, 960, 960, d);
    int h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("cnt=%d sum=%d xmin=%d ymin=%d mw=%d mh=%d q=", h[0], h[1], h[2], h[3], h[4], h[5]);
    for (int i = 0; i < 4; i++) printf("(%d,%d) ", h[6+2*i], h[7+2*i]);
    printf("\n");
    return 0;
}
