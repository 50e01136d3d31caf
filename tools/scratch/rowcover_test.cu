#include <cstdio>
#include <cuda_runtime.h>
#include "../../ocr-system_b200/csrc/db_geom.h"
__global__ void k(DbgPt q0, DbgPt q1, DbgPt q2, DbgPt q3, int rows, int *out) {
    DbgPt q[4] = {q0, q1, q2, q3};
    for (int ry = 0; ry < rows; ry++) {
        int lo[5], hi[5];
        int c = dbg_row_cover(q, ry, lo, hi);
        int c0 = c;
        c = dbg_merge(lo, hi, c);
        out[ry * 12] = c0; out[ry * 12 + 1] = c;
        for (int i = 0; i < 5; i++) { out[ry * 12 + 2 + 2 * i] = i < c ? lo[i] : -99; out[ry * 12 + 3 + 2 * i] = i < c ? hi[i] : -99; }
    }
}
int main() {
    DbgPt q[4] = {{0, 6}, {43, 0}, {45, 12}, {1, 18}};
    int rows = 20;
    int *d; cudaMalloc(&d, rows * 12 * 4);
    k<<<1, 1>>>(q[0], q[1], q[2], q[3], rows, d);
    int h[20 * 12]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int ry = 0; ry < rows; ry++) {
        int lo[5], hi[5];
        int c = dbg_row_cover(q, ry, lo, hi); int c0 = c; c = dbg_merge(lo, hi, c);
        printf("row %2d host c0=%d c=%d:", ry, c0, c); for (int i = 0; i < c; i++) printf(" [%d,%d]", lo[i], hi[i]);
        printf("   dev c0=%d c=%d:", h[ry*12], h[ry*12+1]); for (int i = 0; i < h[ry*12+1]; i++) printf(" [%d,%d]", h[ry*12+2+2*i], h[ry*12+3+2*i]);
        printf("\n");
    }
    return 0;
}
