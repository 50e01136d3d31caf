// k_dbpost.cu -- [upstream PaddleOCR] DBPostProcess on the GPU:
// probability map -> mask -> connected components -> min-area quads -> mean score ->
// Clipper round-offset ("unclip") -> min-area quads -> boxes in source coordinates.
//
// Not in the reference tree.  cv2.findContours(RETR_LIST) is
// replaced by its component-level equivalent (verified with cv2):
//   outer border  <-> 8-connected foreground component
//   hole border   <-> 4-connected background component that does not touch the image
//                     border; its points are the foreground pixels 4-adjacent to it
//   contour order <-> descending raster position of the component's first pixel
// One union-find labelling pass handles both classes at once (ballot-initialised row
// runs, atomicMin unions).  Per candidate, one warp gathers the row extremes of the
// component (the only pixels a convex hull can use), then runs the geometry of
// db_geom.h: hull -> rotating calipers -> fillPoly-exact mean score (lanes stride the
// covered pixels) -> Clipper offset -> hull -> calipers -> scale/clip/round.
#include "ccl.cuh"
#include "db_geom.h"

namespace lumina {

struct DbLayout {
    size_t mask_off, labels_off, cand_off, bbox_off, ncand_off, pool_off, poolctr_off, accept_off, cmeta_off, tmpbox_off,
        tmpscore_off, total;
    size_t pool_pts_per_map;
};

static size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

static DbLayout db_layout(int n, int h, int w, int maxc) {
    DbLayout L;
    const size_t px = (size_t)h * w;
    size_t off = 0;
    L.mask_off = off; off = a256(off + (size_t)n * px);
    L.labels_off = off; off = a256(off + (size_t)n * px * 4);
    L.cand_off = off; off = a256(off + (size_t)n * maxc * 4);           // root position per slot (<0: hole, encoded)
    L.bbox_off = off; off = a256(off + (size_t)n * maxc * 4 * 4);      // xmin, ymin, xmax, ymax
    L.ncand_off = off; off = a256(off + (size_t)n * 2 * 4);            // total found, kept
    L.pool_pts_per_map = px + 4096;  // DbgPt (8 B): row extremes + hull of every candidate (4*rows+2 each)
    L.pool_off = off; off = a256(off + (size_t)n * L.pool_pts_per_map * 8);
    L.poolctr_off = off; off = a256(off + (size_t)n * 4);
    L.accept_off = off; off = a256(off + (size_t)n * maxc);
    L.cmeta_off = off; off = a256(off + (size_t)n * maxc * 2 * 4);
    L.tmpbox_off = off; off = a256(off + (size_t)n * maxc * 8 * 4);
    L.tmpscore_off = off; off = a256(off + (size_t)n * maxc * 4);
    L.total = off;
    return L;
}

// ---- 1. mask ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) db_mask_kernel(const float *__restrict__ pred, uint8_t *__restrict__ mask, size_t total,
                                                      float thresh) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= total) return;
    if (i + 4 <= total && ((((uintptr_t)pred) & 15) == 0)) {
        const uint4 u = ldg_stream_u4(pred + i);
        const uint32_t m = (__uint_as_float(u.x) > thresh ? 1u : 0u) | (__uint_as_float(u.y) > thresh ? 0x100u : 0u) |
                           (__uint_as_float(u.z) > thresh ? 0x10000u : 0u) | (__uint_as_float(u.w) > thresh ? 0x1000000u : 0u);
        *reinterpret_cast<uint32_t *>(mask + i) = m;
    } else {
        for (size_t k = i; k < total && k < i + 4; k++) mask[k] = pred[k] > thresh ? 1 : 0;
    }
}

// use_dilation=True (upstream: cv2.dilate(mask, [[1,1],[1,1]]), anchor (1,1), pixels outside the map ignored):
// mask(x, y) = OR of (pred > thresh) over x-1..x, y-1..y.  4 pixels per thread.
__global__ void __launch_bounds__(256) db_mask_dilate_kernel(const float *__restrict__ pred, uint8_t *__restrict__ mask, int n,
                                                             int h, int w, float thresh) {
    const int groups = (w + 3) >> 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * h * groups) return;
    const int g = (int)(gid % groups);
    const long long row = gid / groups;
    const int y = (int)(row % h);
    const float *cur = pred + (size_t)row * w, *up = cur - w;
    const int x0 = g * 4;
    bool left = false;   // column x0 - 1 (this row or the one above)
    if (x0 > 0) left = cur[x0 - 1] > thresh || (y > 0 && up[x0 - 1] > thresh);
    for (int i = 0; i < 4 && x0 + i < w; i++) {
        const bool here = cur[x0 + i] > thresh || (y > 0 && up[x0 + i] > thresh);
        mask[(size_t)row * w + x0 + i] = (left || here) ? 1 : 0;
        left = here;
    }
}

// ---- 2. two-class union-find labelling ----------------------------------------------
// class bits (fg = 1) of 4 consecutive pixels from one aligned 32-bit load of the byte mask
__device__ __forceinline__ uint32_t db_class4(const uint8_t *p) {
    const uint32_t wv = *reinterpret_cast<const uint32_t *>(p) & 0x01010101u;
    return (wv & 1u) | ((wv >> 7) & 2u) | ((wv >> 14) & 4u) | ((wv >> 21) & 8u);
}

// Two-class row runs: every pixel starts labelled with the first pixel of its run (same class) inside its aligned
// 32-pixel segment.  4 pixels per thread (one 32-bit mask load, one 128-bit label store); the 8 lanes of a segment
// assemble its 32 class bits with three xor-shuffles.  Requires w % 4 == 0 (else the scalar kernel below).
__global__ void __launch_bounds__(256) db_ccl_init4_kernel(const uint8_t *__restrict__ mask, int *__restrict__ labels, int h,
                                                           int w, long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int groups = (w + 31) >> 5 << 3;          // lanes per row, padded to whole segments (8 lanes each)
    const bool active = g < total_groups;
    const long long row = active ? g / groups : 0;
    const int q = active ? (int)(g - row * groups) : 0;   // 4-pixel group inside the row
    const int x0 = q * 4;
    const bool in = active && x0 < w;
    const size_t base = (size_t)row * w;                   // rows are contiguous over the batch
    uint32_t bits = in ? db_class4(mask + base + x0) : 0u;
    const int sub = q & 7;                                 // position of this lane inside its 32-pixel segment
    uint32_t seg = bits << (4 * sub);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        seg |= __shfl_xor_sync(0xffffffffu, seg, o);
    }
    if (!in) return;
    const int idx0 = (int)(((size_t)row % h) * w) + x0;    // page-relative index of the first pixel
    int out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int pos = 4 * sub + k;
        const uint32_t c = (seg >> pos) & 1u;
        const uint32_t same = c ? seg : ~seg;              // pixels of my class
        const uint32_t below = ~same & ((1u << pos) - 1u); // other-class pixels before me in the segment
        const int start = below ? 32 - __clz(below) : 0;
        out[k] = idx0 + k - (pos - start);
    }
    *reinterpret_cast<int4 *>(labels + base + x0) = make_int4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) db_ccl_init_kernel(const uint8_t *__restrict__ mask, int *__restrict__ labels, int h,
                                                          int w, long long nseg_total) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= nseg_total) return;
    const int segs_per_row = (w + 31) >> 5;
    const long long row_g = gw / segs_per_row;
    const int seg = (int)(gw - row_g * segs_per_row);
    const int page = (int)(row_g / h), y = (int)(row_g - (long long)page * h);
    const int x = seg * 32 + lane;
    const size_t pbase = (size_t)page * h * w;
    const bool in = x < w;
    const int idx = y * w + x;
    const int c = in ? (mask[pbase + idx] & 1) : 0;
    const unsigned m1 = __ballot_sync(0xffffffffu, in && c == 1);
    const unsigned m0 = __ballot_sync(0xffffffffu, in && c == 0);
    if (in) {
        const unsigned mine = c ? m1 : m0;
        const unsigned below = ~mine & ((1u << lane) - 1u);
        const int start = below ? 32 - __clz(below) : 0;
        labels[pbase + idx] = idx - (lane - start);
    }
}

// 4 pixels per thread: the classes of this row and the row above come from two word loads (+ the three bytes left /
// right of them); labels are only touched where a union is due (a few per cent of the pixels).
__global__ void __launch_bounds__(256) db_ccl_merge_kernel(const uint8_t *__restrict__ mask, int *__restrict__ labels, int h,
                                                           int w, long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const int groups = (w + 3) >> 2;
    const long long row = g / groups;
    const int x0 = (int)(g - row * groups) * 4;
    const int page = (int)(row / h), y = (int)(row - (long long)page * h);
    const size_t hw = (size_t)h * w;
    const uint8_t *M = mask + (size_t)page * hw;
    int *L = labels + (size_t)page * hw;
    const uint8_t *cur = M + (size_t)y * w;
    // bit k+1 = class of pixel x0+k; bit 0 = pixel x0-1; bit 5 = pixel x0+4 (row above only)
    uint32_t c = 0, u = 0;
    if ((w & 3) == 0) {
        c = db_class4(cur + x0) << 1;
        if (y > 0) u = db_class4(cur - w + x0) << 1;
    } else {
        for (int k = 0; k < 4 && x0 + k < w; k++) {
            c |= (uint32_t)(cur[x0 + k] & 1) << (k + 1);
            if (y > 0) u |= (uint32_t)(cur[x0 + k - w] & 1) << (k + 1);
        }
    }
    if (x0 > 0) {
        c |= cur[x0 - 1] & 1u;
        if (y > 0) u |= cur[x0 - 1 - w] & 1u;
    }
    if (y > 0 && x0 + 4 < w) u |= (uint32_t)(cur[x0 + 4 - w] & 1) << 5;
    const int nvalid = min(4, w - x0);
    for (int k = 0; k < nvalid; k++) {
        const int x = x0 + k, idx = y * w + x;
        const uint32_t cc = (c >> (k + 1)) & 1u;
        const bool left_same = x > 0 && ((c >> k) & 1u) == cc;
        if ((x & 31) == 0 && left_same) ccl_union(L, idx, idx - 1);
        if (y > 0) {
            const int up = idx - w;
            if (((u >> (k + 1)) & 1u) == cc) {
                const bool upleft_same = x > 0 && ((u >> k) & 1u) == cc;
                if (!(left_same && upleft_same)) ccl_union(L, idx, up);
            } else if (cc == 1u) {  // foreground is 8-connected
                if (x > 0 && ((u >> k) & 1u)) ccl_union(L, idx, up - 1);
                if (x + 1 < w && ((u >> (k + 2)) & 1u)) ccl_union(L, idx, up + 1);
            }
        }
    }
}

// ---- tile-local labelling (w % 4 == 0): a CTA labels a 32-row x 128-column tile entirely in shared memory (row runs, unions
// with shared-memory atomics, local flatten) and writes every pixel's label as the PAGE index of its tile-local root; the
// links that cross a tile border are made afterwards on those labels (db_ccl_border_kernel), then the usual flatten.  Local
// raster order equals page raster order inside a tile, so roots stay "minimum raster index" and the final labels are the
// same as the global union-find's.
constexpr int DBT_W = 128, DBT_H = 32;

__device__ __forceinline__ int dbt_find(const int *L, int a) {
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}
__device__ __forceinline__ void dbt_union(int *L, int a, int b) {
    bool done;
    do {
        a = dbt_find(L, a);
        b = dbt_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

__global__ void __launch_bounds__(256) db_ccl_tile_kernel(const uint8_t *__restrict__ mask, int *__restrict__ labels, int h, int w,
                                                          int tiles_x, int tiles_y) {
    __shared__ __align__(16) int sl[DBT_H * DBT_W];
    __shared__ uint32_t scls[DBT_H][DBT_W / 32], sval[DBT_H][DBT_W / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, page = blockIdx.x / (tiles_x * tiles_y);
    const int x0t = tx * DBT_W, y0t = ty * DBT_H;
    const size_t hw = (size_t)h * w;
    const uint8_t *M = mask + (size_t)page * hw;
    int *L = labels + (size_t)page * hw;
    const int lx0 = lane * 4, sub = lane & 7;
    // phase 1: class / validity bits of the tile, row runs inside 32-pixel segments
#pragma unroll
    for (int i = 0; i < DBT_H / 8; i++) {
        const int ly = i * 8 + warp, y = y0t + ly, x = x0t + lx0;
        uint32_t bits = 0, val = 0;
        if (y < h) {
            if (x + 4 <= w) { bits = db_class4(M + (size_t)y * w + x); val = 0xFu; }
            else for (int k = 0; k < 4; k++) if (x + k < w) { bits |= (uint32_t)(M[(size_t)y * w + x + k] & 1) << k; val |= 1u << k; }
        }
        uint32_t seg = bits << (4 * sub), vseg = val << (4 * sub);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            seg |= __shfl_xor_sync(0xffffffffu, seg, o);
            vseg |= __shfl_xor_sync(0xffffffffu, vseg, o);
        }
        int out[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int pos = 4 * sub + k;
            const uint32_t c = (seg >> pos) & 1u;
            const uint32_t same = (c ? seg : ~seg) & vseg;
            const uint32_t below = ~same & ((1u << pos) - 1u);
            const int start = below ? 32 - __clz(below) : 0;
            out[k] = ly * DBT_W + (lane >> 3) * 32 + start;
        }
        *reinterpret_cast<int4 *>(&sl[ly * DBT_W + lx0]) = make_int4(out[0], out[1], out[2], out[3]);
        if (sub == 0) { scls[ly][lane >> 3] = seg; sval[ly][lane >> 3] = vseg; }
    }
    __syncthreads();
    // phase 2: unions inside the tile (pixels outside the tile or the map count as absent).  Bit-parallel: a thread owns half a
    // 32-pixel class word of one tile row, derives the pixels that need a link (run heads under a same-class pixel,
    // foreground under background with a foreground diagonal, segment starts with a same-class left pixel) as bit masks and
    // walks only their set bits.
    {
        const int t = threadIdx.x, ly = t >> 3, wi = (t >> 1) & 3;
        const uint32_t half = (t & 1) ? 0xffff0000u : 0x0000ffffu;
        const uint32_t C = scls[ly][wi], V = sval[ly][wi];
        if (V & half) {
            const uint32_t cl = wi > 0 ? scls[ly][wi - 1] >> 31 : 0u;
            const uint32_t HL = (V << 1) | (wi > 0 ? 1u : 0u);              // the left neighbour is inside the tile
            const uint32_t left_same = HL & ~(C ^ ((C << 1) | cl)) & V;
            uint32_t mL = left_same & 1u & half, mN = 0, mNW = 0, mNE = 0;
            if (ly > 0) {
                const uint32_t U = scls[ly - 1][wi];
                const uint32_t ul = wi > 0 ? scls[ly - 1][wi - 1] >> 31 : 0u;
                const uint32_t ur = wi < 3 ? scls[ly - 1][wi + 1] & 1u : 0u, vr = wi < 3 ? sval[ly][wi + 1] & 1u : 0u;
                const uint32_t Ul = (U << 1) | ul, Ur = (U >> 1) | (ur << 31), HR = (V >> 1) | (vr << 31);
                const uint32_t up_same = ~(C ^ U) & V;
                const uint32_t upleft_same = HL & ~(C ^ Ul);
                mN = up_same & ~(left_same & upleft_same) & half;
                const uint32_t fg_no_up = C & ~U & V & half;
                mNW = fg_no_up & HL & Ul;
                mNE = fg_no_up & HR & Ur;
            }
            const int base = ly * DBT_W + wi * 32;
            if (mL) dbt_union(sl, base, base - 1);
            for (uint32_t m = mN; m; m &= m - 1) { const int idx = base + __ffs(m) - 1; dbt_union(sl, idx, idx - DBT_W); }
            for (uint32_t m = mNW; m; m &= m - 1) { const int idx = base + __ffs(m) - 1; dbt_union(sl, idx, idx - DBT_W - 1); }
            for (uint32_t m = mNE; m; m &= m - 1) { const int idx = base + __ffs(m) - 1; dbt_union(sl, idx, idx - DBT_W + 1); }
        }
    }
    __syncthreads();
    // phase 3: local roots as page indices
#pragma unroll 1
    for (int i = 0; i < DBT_H / 8; i++) {
        const int ly = i * 8 + warp, y = y0t + ly, x = x0t + lx0;
        if (y >= h || x >= w) continue;
        int out[4];
        int prev_lab = -1, prev_root = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int lab = sl[ly * DBT_W + lx0 + k];
            if (lab != prev_lab) { prev_lab = lab; prev_root = dbt_find(sl, lab); }
            out[k] = (y0t + (prev_root >> 7)) * w + x0t + (prev_root & (DBT_W - 1));
        }
        int *dst = L + (size_t)y * w + x;
        if (x + 4 <= w) *reinterpret_cast<int4 *>(dst) = make_int4(out[0], out[1], out[2], out[3]);
        else for (int k = 0; k < 4; k++) if (x + k < w) dst[k] = out[k];
    }
}

// links across tile borders: the first row of every tile row (N, or NW / NE for foreground under background) and the two
// pixel columns either side of every vertical tile border (left link, NW of the right pixel, NE of the left pixel)
__global__ void __launch_bounds__(256) db_ccl_border_kernel(const uint8_t *__restrict__ mask, int *__restrict__ labels, int h, int w,
                                                            int tiles_x, int tiles_y, long long n_rows_part, long long total) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const size_t hw = (size_t)h * w;
    if (g < n_rows_part) {
        const int x = (int)(g % w);
        const long long r = g / w;
        const int ty = (int)(r % (tiles_y - 1)) + 1, page = (int)(r / (tiles_y - 1));
        const int y = ty * DBT_H;
        const uint8_t *M = mask + (size_t)page * hw;
        int *L = labels + (size_t)page * hw;
        const int idx = y * w + x, up = idx - w;
        const uint32_t cc = M[idx] & 1u;
        if ((M[up] & 1u) == cc) {
            const bool left_same = x > 0 && (M[idx - 1] & 1u) == cc, upleft_same = x > 0 && (M[up - 1] & 1u) == cc;
            if (!(left_same && upleft_same)) ccl_union(L, idx, up);
        } else if (cc == 1u) {
            if (x > 0 && (M[up - 1] & 1u)) ccl_union(L, idx, up - 1);
            if (x + 1 < w && (M[up + 1] & 1u)) ccl_union(L, idx, up + 1);
        }
    } else {
        const long long q = g - n_rows_part;
        const int y = (int)(q % h);
        const long long r = q / h;
        const int tx = (int)(r % (tiles_x - 1)) + 1, page = (int)(r / (tiles_x - 1));
        const int x = tx * DBT_W;
        const uint8_t *M = mask + (size_t)page * hw;
        int *L = labels + (size_t)page * hw;
        const int idx = y * w + x;
        const uint32_t cc = M[idx] & 1u, cl = M[idx - 1] & 1u;
        if (cl == cc) ccl_union(L, idx, idx - 1);
        if (y > 0) {
            const uint32_t cu = M[idx - w] & 1u, cul = M[idx - w - 1] & 1u;
            if (cc == 1u && cu == 0u && cul == 1u) ccl_union(L, idx, idx - w - 1);      // NW of the right pixel
            if (cl == 1u && cul == 0u && cu == 1u) ccl_union(L, idx - 1, idx - w);      // NE of the left pixel
        }
    }
}

// flatten, 4 pixels per thread (128-bit label load / store; neighbours of a run share the root that was just found)
__global__ void __launch_bounds__(256) db_ccl_flatten4_kernel(uint8_t *__restrict__ mask, int *__restrict__ labels, int h, int w,
                                                              long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const int groups = w >> 2;
    const long long row = g / groups;
    const int x0 = (int)(g - row * groups) * 4;
    const int page = (int)(row / h), y = (int)(row - (long long)page * h);
    const size_t hw = (size_t)h * w;
    uint8_t *M = mask + (size_t)page * hw;
    int *L = labels + (size_t)page * hw;
    const int idx0 = y * w + x0;
    const int4 lv = *reinterpret_cast<const int4 *>(L + idx0);
    const int lab[4] = {lv.x, lv.y, lv.z, lv.w};
    const uint32_t cls = db_class4(M + idx0);
    int out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k > 0 && lab[k] == lab[k - 1]) { out[k] = out[k - 1]; continue; }
        int a = lab[k], p = a == idx0 + k ? a : L[a];
        while (p != a) { a = p; p = L[a]; }
        out[k] = a;
    }
    *reinterpret_cast<int4 *>(L + idx0) = make_int4(out[0], out[1], out[2], out[3]);
    if (y == 0 || y == h - 1) {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (!((cls >> k) & 1u)) M[out[k]] = 2;    // same value from every writer
    } else {
        if (x0 == 0 && !(cls & 1u)) M[out[0]] = 2;
        if (x0 + 4 == w && !((cls >> 3) & 1u)) M[out[3]] = 2;
    }
}

// flatten; background components touching the image border get bit 1 on their root's mask byte
__global__ void __launch_bounds__(256) db_ccl_flatten_kernel(uint8_t *__restrict__ mask, int *__restrict__ labels, int h, int w,
                                                             long long total_px) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_px) return;
    const long long hw = (long long)h * w;
    const int page = (int)(g / hw);
    const int idx = (int)(g - (long long)page * hw);
    uint8_t *M = mask + (size_t)page * hw;
    int *L = labels + (size_t)page * hw;
    const int root = ccl_find(L, idx);
    L[idx] = root;
    if ((M[idx] & 1) == 0) {
        const int y = idx / w, x = idx - y * w;
        if (x == 0 || y == 0 || x == w - 1 || y == h - 1) M[root] = 2;  // same value from every writer
    }
}

// ---- 3. candidates in findContours order ----------------------------------------------
// One CTA per map: raster-order compaction of component roots (outer: fg root; hole: root of
// a bg component that does not touch the border).  Discovery order == ascending root position;
// findContours returns the reverse, truncated to max_candidates.  Two passes: count, then place.
__global__ void __launch_bounds__(1024) db_candidates_kernel(const uint8_t *__restrict__ mask, int *__restrict__ labels,
                                                             int *__restrict__ cand, int *__restrict__ bbox,
                                                             int *__restrict__ ncand, int h, int w, int maxc) {
    const int page = blockIdx.x;
    const int px = h * w;
    const uint8_t *M = mask + (size_t)page * px;
    int *L = labels + (size_t)page * px;
    int *C = cand + (size_t)page * maxc;
    int *B = bbox + (size_t)page * maxc * 4;
    __shared__ int wsum[32];
    __shared__ int chunk_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < maxc; i += blockDim.x) {
        C[i] = 0x7fffffff;
        B[i * 4 + 0] = 0x7fffffff; B[i * 4 + 1] = 0x7fffffff; B[i * 4 + 2] = -1; B[i * 4 + 3] = -1;
    }
    int total = 0;
    for (int pass = 0; pass < 2; pass++) {
        int base = 0;
        for (int p0 = 0; p0 < px; p0 += 1024 * 4) {
            const int p = p0 + threadIdx.x * 4;
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int q = p + k;
                if (q < px) {
                    const int lab = L[q];
                    const uint8_t mv = M[q];
                    const bool root = pass == 0 ? (lab == q) : (lab == q || lab < 0);
                    if (root && ((mv & 1) || mv == 0)) bits |= 1u << k;  // fg root, or bg root not touching the border
                }
            }
            const int c = __popc(bits);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                const int v = wsum[lane];
                int s = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, s, o);
                    if (lane >= o) s += t;
                }
                wsum[lane] = s - v;
                if (lane == 31) chunk_total = s;
            }
            __syncthreads();
            if (pass == 1) {
                int i = base + wsum[warp] + inc - c;  // ascending discovery index
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (bits & (1u << k)) {
                        const int q = p + k;
                        const int slot = total - 1 - i;  // findContours order = reverse discovery
                        if (slot < maxc) {
                            C[slot] = (M[q] & 1) ? q : -1 - q;  // hole candidates are stored negated
                            L[q] = -2 - slot;                   // root now carries its slot
                        } else {
                            L[q] = -1;                          // dropped by max_candidates
                        }
                        i++;
                    }
            }
            base += chunk_total;
            __syncthreads();
        }
        if (pass == 0) total = base;
    }
    if (threadIdx.x == 0) { ncand[page * 2] = total; ncand[page * 2 + 1] = total < maxc ? total : maxc; }
}

__device__ __forceinline__ int db_slot_of(const int *L, int p) {
    const int lab = L[p];
    if (lab < 0) return -2 - lab;          // p is a root (slot, or -1 when dropped)
    const int r = L[lab];
    return r < 0 ? -2 - r : -1;
}

// ---- 4. bounding boxes of the candidates' point sets -----------------------------------
// Only outline pixels of a foreground component (a 4-neighbour is background or the map edge) and background pixels
// next to foreground (hole borders) can move a box.  The byte mask decides that first -- 4 pixels per thread, the rows
// above / below as one 32-bit load each -- and only those pixels chase their labels and issue atomics.
__device__ __forceinline__ void db_bbox_update(int *B, int slot, int x, int y) {
    atomicMin(&B[slot * 4 + 0], x); atomicMin(&B[slot * 4 + 1], y);
    atomicMax(&B[slot * 4 + 2], x); atomicMax(&B[slot * 4 + 3], y);
}

__global__ void __launch_bounds__(256) db_bbox_kernel(const uint8_t *__restrict__ mask, const int *__restrict__ labels,
                                                      int *__restrict__ bbox, int h, int w, int maxc, long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const int groups = (w + 3) >> 2;
    const long long row = g / groups;
    const int x0 = (int)(g - row * groups) * 4;
    const int page = (int)(row / h), y = (int)(row - (long long)page * h);
    const size_t hw = (size_t)h * w;
    const uint8_t *M = mask + (size_t)page * hw;
    const int *L = labels + (size_t)page * hw;
    int *B = bbox + (size_t)page * maxc * 4;
    const uint8_t *cur = M + (size_t)y * w;
    // fg bits of this row for x0-1 .. x0+4 (bit k+1 = pixel x0+k), and of the rows above / below for x0 .. x0+3
    uint32_t c = 0, up = 0, dn = 0;
    const bool vec = (w & 3) == 0;
    if (vec) {
        const uint32_t wc = *reinterpret_cast<const uint32_t *>(cur + x0) & 0x01010101u;
        c = ((wc & 1u) | ((wc >> 7) & 2u) | ((wc >> 14) & 4u) | ((wc >> 21) & 8u)) << 1;
        if (y > 0) {
            const uint32_t wu = *reinterpret_cast<const uint32_t *>(cur - w + x0) & 0x01010101u;
            up = (wu & 1u) | ((wu >> 7) & 2u) | ((wu >> 14) & 4u) | ((wu >> 21) & 8u);
        }
        if (y < h - 1) {
            const uint32_t wd = *reinterpret_cast<const uint32_t *>(cur + w + x0) & 0x01010101u;
            dn = (wd & 1u) | ((wd >> 7) & 2u) | ((wd >> 14) & 4u) | ((wd >> 21) & 8u);
        }
    } else {
        for (int k = 0; k < 4 && x0 + k < w; k++) {
            c |= (uint32_t)(cur[x0 + k] & 1) << (k + 1);
            if (y > 0) up |= (uint32_t)(cur[x0 + k - w] & 1) << k;
            if (y < h - 1) dn |= (uint32_t)(cur[x0 + k + w] & 1) << k;
        }
    }
    if (x0 > 0) c |= cur[x0 - 1] & 1u;
    if (x0 + 4 < w) c |= (uint32_t)(cur[x0 + 4] & 1) << 5;
    const int nvalid = min(4, w - x0);
    for (int k = 0; k < nvalid; k++) {
        const int x = x0 + k, idx = y * w + x;
        const bool fg = (c >> (k + 1)) & 1u;
        const bool l = (c >> k) & 1u, r = (c >> (k + 2)) & 1u, u = (up >> k) & 1u, d = (dn >> k) & 1u;
        if (fg) {
            const bool inner = x > 0 && x < w - 1 && y > 0 && y < h - 1 && l && r && u && d;
            if (inner) continue;
            const int slot = db_slot_of(L, idx);
            if (slot >= 0) db_bbox_update(B, slot, x, y);
        } else {
            // hole pixel: its 4-adjacent foreground pixels are the hole border's point set
            const bool fl = x > 0 && l, fr = x < w - 1 && r, fu = u, fd = d;   // up / dn are already 0 outside the map
            if (!(fl | fr | fu | fd)) continue;
            const int slot = db_slot_of(L, idx);
            if (slot < 0) continue;
            if (fl) db_bbox_update(B, slot, x - 1, y);
            if (fr) db_bbox_update(B, slot, x + 1, y);
            if (fu) db_bbox_update(B, slot, x, y - 1);
            if (fd) db_bbox_update(B, slot, x, y + 1);
        }
    }
}

// ---- 5. per-candidate geometry: one warp per candidate ----------------------------------
struct DbParams {
    const float *pred;
    const uint8_t *mask;
    const int *labels;
    const int *cand;
    const int *bbox;
    const int *ncand;
    DbgPt *pool;
    int *poolctr;
    uint8_t *accept;
    int *cmeta;          // [n][maxc][2] pool offset and point count of every candidate
    int *tmpbox;
    float *tmpscore;
    const int *src_hw;  // [n][2] device copy
    size_t pool_pts_per_map;
    int h, w, maxc, min_size;
    double box_thresh, unclip_ratio;
};

// merged coverage intervals of one mask row.  Deliberately out of line: inlined into the candidate
// kernel nvcc 12.9 produced wrong interval bounds for x-major edges (caught by the parity test).
__device__ __noinline__ int db_row_intervals(const DbgPt *q4, int ry, int mw, int mh, int *lo, int *hi) {
    DbgPt q[4] = {q4[0], q4[1], q4[2], q4[3]};
    int l[5] = {0, 0, 0, 0, 0}, h[5] = {-1, -1, -1, -1, -1};
    int c = dbg_row_cover(q, ry, mw, mh, l, h);
    c = dbg_merge(l, h, c);
    for (int i = 0; i < 5; i++) { lo[i] = l[i]; hi[i] = h[i]; }
    return c;
}

constexpr int DB_WARPS = 4;
constexpr int DB_OFFS_MAX = 384;  // Clipper offset vertices kept in shared memory per candidate

// The per-candidate geometry runs as four small kernels.  As one kernel (one warp per candidate, the serial
// routines on lane 0) its dominant stall was instruction fetch: thousands of warps, each at a different
// place of a large body.  Split, the serial routines run one candidate per THREAD (32 candidates share every
// instruction fetch) and the warp-parallel parts (pixel scans) keep one warp per candidate.
// accept[] carries the state between them: 2 = row extremes ready, 3 = first quad ready, 4 = score passed,
// 1 = accepted; 11 / 12 / 13 / 15 = reject reasons (too small, score, unclip, pool overflow); 0 = empty slot.

// ---- 5a. row extremes of the candidate's point set (one warp per candidate) -------------------------------
__global__ void __launch_bounds__(DB_WARPS * 32) db_cand_rows_kernel(const DbParams p) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int page = blockIdx.y;
    const int slot = blockIdx.x * DB_WARPS + wib;
    if (slot >= p.ncand[page * 2 + 1]) return;
    const int px = p.h * p.w;
    const uint8_t *M = p.mask + (size_t)page * px;
    const int *L = p.labels + (size_t)page * px;
    const int centry = p.cand[(size_t)page * p.maxc + slot];
    const bool hole = centry < 0;
    const int root = hole ? -1 - centry : centry;
    const int *B = p.bbox + ((size_t)page * p.maxc + slot) * 4;
    const int bx0 = B[0], by0 = B[1], bx1 = B[2], by1 = B[3];
    uint8_t *acc = p.accept + (size_t)page * p.maxc + slot;
    if (lane == 0) *acc = 0;
    if (bx1 < bx0 || by1 < by0) return;
    const int rows = by1 - by0 + 1;
    // pool space: 2*rows sorted points + (2*rows + 2) hull points
    int pbase = 0;
    if (lane == 0) pbase = atomicAdd(p.poolctr + page, 4 * rows + 2);
    pbase = __shfl_sync(0xffffffffu, pbase, 0);
    if ((size_t)pbase + 4 * rows + 2 > p.pool_pts_per_map) { if (lane == 0) *acc = 15; return; }  // pathological masks only (> px/4 candidate rows): dropped
    DbgPt *pts = p.pool + (size_t)page * p.pool_pts_per_map + pbase;
    int npts = 0;
    for (int y = by0; y <= by1; y++) {
        int rmin = 0x7fffffff, rmax = -1;
        for (int xb = bx0; xb <= bx1; xb += 32) {
            const int x = xb + lane;
            bool member = false;
            if (x <= bx1) {
                const int q = y * p.w + x;
                if (M[q] & 1) {
                    if (!hole) member = (q == root) || (L[q] == root);
                    else {
                        // foreground pixel 4-adjacent to a pixel of this hole
                        if (x > 0 && !(M[q - 1] & 1) && (q - 1 == root || L[q - 1] == root)) member = true;
                        else if (x + 1 < p.w && !(M[q + 1] & 1) && (q + 1 == root || L[q + 1] == root)) member = true;
                        else if (y > 0 && !(M[q - p.w] & 1) && (q - p.w == root || L[q - p.w] == root)) member = true;
                        else if (y + 1 < p.h && !(M[q + p.w] & 1) && (q + p.w == root || L[q + p.w] == root)) member = true;
                    }
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, member);
            if (b) {
                rmin = min(rmin, xb + __ffs(b) - 1);
                rmax = max(rmax, xb + 31 - __clz(b));
            }
        }
        if (rmax >= 0) {
            if (lane == 0) {
                pts[npts].x = rmin; pts[npts].y = y;
                pts[npts + 1].x = rmax; pts[npts + 1].y = y;
            }
            npts += 2;
        }
    }
    if (lane == 0) {
        int *meta = p.cmeta + ((size_t)page * p.maxc + slot) * 2;
        meta[0] = pbase; meta[1] = npts;
        *acc = 2;
    }
}

// ---- 5b. hull + first min-area quad (one THREAD per candidate) -------------------------------------------------
__global__ void __launch_bounds__(128) db_cand_rect_kernel(const DbParams p) {
    const int page = blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= p.ncand[page * 2 + 1]) return;
    uint8_t *acc = p.accept + (size_t)page * p.maxc + slot;
    if (*acc != 2) return;
    const int centry = p.cand[(size_t)page * p.maxc + slot];
    const bool hole = centry < 0;
    const int root = hole ? -1 - centry : centry;
    const int *meta = p.cmeta + ((size_t)page * p.maxc + slot) * 2;
    const int *B = p.bbox + ((size_t)page * p.maxc + slot) * 4;
    const int rows = B[3] - B[1] + 1;
    DbgPt *pts = p.pool + (size_t)page * p.pool_pts_per_map + meta[0];
    DbgPt *hull = pts + 2 * rows;
    const int hn = dbg_hull_sorted(pts, meta[1], hull);
    // note: roots carry the slot code in L; the root position itself is the (y,x)-smallest pixel
    const int hy = root / p.w, hx = root - hy * p.w;
    dbg_hull_rotate(hull, hn, hole ? 2 : 1, hx - 1, hy);
    const DbgRect r = dbg_min_area_rect(hull, hn);
    DbgPtF o[4];
    const float sside = dbg_mini_box(r, o);
    if (sside < (float)p.min_size) { *acc = 11; return; }
    float *qf = reinterpret_cast<float *>(p.tmpbox + ((size_t)page * p.maxc + slot) * 8);   // the quad parks in tmpbox
    for (int i = 0; i < 4; i++) { qf[i * 2] = o[i].x; qf[i * 2 + 1] = o[i].y; }
    *acc = 3;
}

// ---- 5c. box_score_fast: mean of pred over cv2.fillPoly(int32(box - (xmin, ymin))) (one warp per candidate) ---
__global__ void __launch_bounds__(DB_WARPS * 32) db_cand_score_kernel(const DbParams p) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int page = blockIdx.y;
    const int slot = blockIdx.x * DB_WARPS + wib;
    if (slot >= p.ncand[page * 2 + 1]) return;
    uint8_t *acc = p.accept + (size_t)page * p.maxc + slot;
    if (*acc != 3) return;
    const float *P = p.pred + (size_t)page * p.h * p.w;
    const float *qf = reinterpret_cast<const float *>(p.tmpbox + ((size_t)page * p.maxc + slot) * 8);
    float bxs[4], bys[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { bxs[i] = qf[i * 2]; bys[i] = qf[i * 2 + 1]; }
    const float fminx = fminf(fminf(bxs[0], bxs[1]), fminf(bxs[2], bxs[3]));
    const float fmaxx = fmaxf(fmaxf(bxs[0], bxs[1]), fmaxf(bxs[2], bxs[3]));
    const float fminy = fminf(fminf(bys[0], bys[1]), fminf(bys[2], bys[3]));
    const float fmaxy = fmaxf(fmaxf(bys[0], bys[1]), fmaxf(bys[2], bys[3]));
    const int xmin = min(max((int)floorf(fminx), 0), p.w - 1), xmax = min(max((int)ceilf(fmaxx), 0), p.w - 1);
    const int ymin = min(max((int)floorf(fminy), 0), p.h - 1), ymax = min(max((int)ceilf(fmaxy), 0), p.h - 1);
    DbgPt q[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { q[i].x = (int)(bxs[i] - (float)xmin); q[i].y = (int)(bys[i] - (float)ymin); }
    const int mh = ymax - ymin + 1, mw = xmax - xmin + 1;
    // a lane owns a mask row: its coverage intervals (serial integer geometry, now 32 rows at a time) and its sum
    double sum = 0.0;
    int cnt = 0;
    for (int ry = lane; ry < mh; ry += 32) {
        int lo[5], hi[5];
        const int c = db_row_intervals(q, ry, mw, mh, lo, hi);
        const float *prow = P + (size_t)(ymin + ry) * p.w + xmin;
        for (int i = 0; i < c; i++) {
            const int a = max(lo[i], 0), b = min(hi[i], mw - 1);
            for (int x = a; x <= b; x++) { sum += (double)__ldg(prow + x); cnt++; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const double score = cnt > 0 ? sum / (double)cnt : 0.0;
    if (lane == 0) {
        p.tmpscore[(size_t)page * p.maxc + slot] = (float)score;   // for a reject: the score that failed (diagnostics)
        *acc = p.box_thresh > score ? 12 : 4;
    }
}

// ---- 5c'. box_score_slow: mean of pred over cv2.fillPoly(contour) (one warp per candidate) ---------------------------
// The traced contour never exists here, and is not needed: the polygon through a border's pixels covers exactly
//   outer border of component C :  every pixel that cannot leave C's outline over non-C pixels (4-moves) = C, its holes
//                                  and whatever is nested in them;
//   hole border of hole H       :  the border's own pixels (foreground 4-adjacent to H), H, and whatever is nested in H
// (checked against cv2.findContours + cv2.fillPoly on 7 400 random contours with nesting, tests/test_db_geom.py).
// "Nested in" is read off the labels without any flood fill: the raster-first pixel of a component is its root; the
// pixel left of a foreground root belongs to the background component around that component, the pixel left of a
// hole's root to the foreground component the hole is a hole of -- the same walk findContours' hierarchy is built by.
__device__ __forceinline__ int db_root_pos(const int *L, int p) {
    const int lab = L[p];
    return lab < 0 ? p : lab;
}

// does the component of pixel p lie inside `target` (a root position; hole or outer as `target_hole` says)?
__device__ __forceinline__ bool db_enclosed_by(const uint8_t *M, const int *L, int w, int p, int target, bool target_hole) {
    bool fg = M[p] & 1;
    int X = db_root_pos(L, p);
    for (;;) {
        if (fg) {
            if (!target_hole && X == target) return true;
            if (X % w == 0) return false;       // the component's first pixel sits in column 0: nothing around it
            X = db_root_pos(L, X - 1);
            fg = false;
        } else {
            if (target_hole && X == target) return true;
            if (M[X] == 2) return false;        // background joined to the image border
            X = db_root_pos(L, X - 1);          // a hole does not touch the border: its first pixel has a left neighbour
            fg = true;
        }
    }
}

__global__ void __launch_bounds__(DB_WARPS * 32) db_cand_score_slow_kernel(const DbParams p) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int page = blockIdx.y;
    const int slot = blockIdx.x * DB_WARPS + wib;
    if (slot >= p.ncand[page * 2 + 1]) return;
    uint8_t *acc = p.accept + (size_t)page * p.maxc + slot;
    if (*acc != 3) return;
    const int px = p.h * p.w;
    const float *P = p.pred + (size_t)page * px;
    const uint8_t *M = p.mask + (size_t)page * px;
    const int *L = p.labels + (size_t)page * px;
    const int centry = p.cand[(size_t)page * p.maxc + slot];
    const bool hole = centry < 0;
    const int root = hole ? -1 - centry : centry;
    const int *B = p.bbox + ((size_t)page * p.maxc + slot) * 4;   // bounding box of the contour's points
    const int bx0 = B[0], by0 = B[1], bx1 = B[2], by1 = B[3];
    double sum = 0.0;
    int cnt = 0;
    for (int y = by0; y <= by1; y++) {
        for (int x = bx0 + lane; x <= bx1; x += 32) {
            const int q = y * p.w + x;
            bool member = false;
            if (hole && (M[q] & 1)) {
                // a pixel of the hole's border
                if (x > 0 && !(M[q - 1] & 1) && db_root_pos(L, q - 1) == root) member = true;
                else if (x + 1 < p.w && !(M[q + 1] & 1) && db_root_pos(L, q + 1) == root) member = true;
                else if (y > 0 && !(M[q - p.w] & 1) && db_root_pos(L, q - p.w) == root) member = true;
                else if (y + 1 < p.h && !(M[q + p.w] & 1) && db_root_pos(L, q + p.w) == root) member = true;
            }
            if (!member) member = db_enclosed_by(M, L, p.w, q, root, hole);
            if (member) { sum += (double)__ldg(P + q); cnt++; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const double score = cnt > 0 ? sum / (double)cnt : 0.0;
    if (lane == 0) {
        p.tmpscore[(size_t)page * p.maxc + slot] = (float)score;
        *acc = p.box_thresh > score ? 12 : 4;
    }
}

// ---- 5d. unclip + second min-area quad (one THREAD per candidate, offset polygon in shared memory) ------------
constexpr int DB_UNCLIP_THREADS = 32;
constexpr int DB_OFFS_SMALL = 64;   // offset polygons of ordinary text boxes have 12-40 vertices

// unclip of one candidate with the offset polygon in `offs` (room for max_pts) and its hull in `offs_hull`.
// Returns 1 accepted (box written), 0 rejected, -1 the polygon does not fit max_pts.
__device__ __forceinline__ int db_unclip_one(const DbParams &p, int page, int *tb, DbgPt *offs, DbgPt *offs_hull, int max_pts) {
    const float *qf = reinterpret_cast<const float *>(tb);
    DbgPtF b4[4];
    for (int i = 0; i < 4; i++) { b4[i].x = qf[i * 2]; b4[i].y = qf[i * 2 + 1]; }
    const double dist = dbg_unclip_distance(b4, p.unclip_ratio);
    if (!(dist >= 0)) return 0;
    const int m = dbg_clipper_offset(b4, dist, offs, max_pts);
    if (m < 0) return -1;
    if (m < 3) return 0;
    dbg_sort(offs, m);
    const int hn = dbg_hull_sorted(offs, m, offs_hull);
    const DbgRect r2 = dbg_min_area_rect(offs_hull, hn);
    DbgPtF o2[4];
    const float ss2 = dbg_mini_box(r2, o2);
    if (ss2 < (float)(p.min_size + 2)) return 0;
    const double dw = (double)p.src_hw[page * 2 + 1], dh = (double)p.src_hw[page * 2];
    int outq[8];
    for (int i = 0; i < 4; i++) {
        outq[i * 2] = dbg_scale_coord(o2[i].x, p.w, dw);
        outq[i * 2 + 1] = dbg_scale_coord(o2[i].y, p.h, dh);
    }
    for (int i = 0; i < 8; i++) tb[i] = outq[i];
    return 1;
}

// first pass: per-thread buffers for DB_OFFS_SMALL vertices in local memory (no shared memory: full occupancy);
// a candidate whose polygon is larger keeps state 4 for the second pass
__global__ void __launch_bounds__(128) db_cand_unclip_small_kernel(const DbParams p) {
    const int page = blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= p.ncand[page * 2 + 1]) return;
    uint8_t *acc = p.accept + (size_t)page * p.maxc + slot;
    if (*acc != 4) return;
    DbgPt offs[DB_OFFS_SMALL + 2], offs_hull[DB_OFFS_SMALL + 2];
    const int r = db_unclip_one(p, page, p.tmpbox + ((size_t)page * p.maxc + slot) * 8, offs, offs_hull, DB_OFFS_SMALL);
    if (r >= 0) *acc = r ? 1 : 13;
}

// second pass (rare): polygons of up to DB_OFFS_MAX vertices, buffers in shared memory
__global__ void __launch_bounds__(DB_UNCLIP_THREADS) db_cand_unclip_kernel(const DbParams p) {
    extern __shared__ __align__(16) unsigned char db_smem[];
    DbgPt *offs = reinterpret_cast<DbgPt *>(db_smem) + (size_t)threadIdx.x * 2 * (DB_OFFS_MAX + 2);
    DbgPt *offs_hull = offs + (DB_OFFS_MAX + 2);
    const int page = blockIdx.y;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= p.ncand[page * 2 + 1]) return;
    uint8_t *acc = p.accept + (size_t)page * p.maxc + slot;
    if (*acc != 4) return;
    const int r = db_unclip_one(p, page, p.tmpbox + ((size_t)page * p.maxc + slot) * 8, offs, offs_hull, DB_OFFS_MAX);
    *acc = r == 1 ? 1 : 13;
}

// ---- 6. ordered compaction of the accepted candidates ------------------------------------
__global__ void __launch_bounds__(1024) db_compact_kernel(const uint8_t *__restrict__ accept, const int *__restrict__ tmpbox,
                                                          const float *__restrict__ tmpscore, const int *__restrict__ ncand,
                                                          int *__restrict__ boxes, float *__restrict__ scores,
                                                          int *__restrict__ counts, int maxc) {
    const int page = blockIdx.x;
    const int kept = ncand[page * 2 + 1];
    __shared__ int wsum[32];
    __shared__ int chunk_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int s0 = 0; s0 < kept; s0 += 1024) {
        const int s = s0 + threadIdx.x;
        const int a = (s < kept && accept[(size_t)page * maxc + s] == 1) ? 1 : 0;  // other codes = reject reasons
        int inc = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int v = wsum[lane];
            int t = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            wsum[lane] = t - v;
            if (lane == 31) chunk_total = t;
        }
        __syncthreads();
        if (a) {
            const int dst = base + wsum[warp] + inc - 1;
            const int *tb = tmpbox + ((size_t)page * maxc + s) * 8;
            int *ob = boxes + ((size_t)page * maxc + dst) * 8;
#pragma unroll
            for (int i = 0; i < 8; i++) ob[i] = tb[i];
            scores[(size_t)page * maxc + dst] = tmpscore[(size_t)page * maxc + s];
        }
        base += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[page] = base;
}

// canonical labels for the stage-level parity test: label = min raster index + 1 for fg, 0 for bg
__global__ void __launch_bounds__(256) db_export_labels_kernel(const uint8_t *__restrict__ mask, const int *__restrict__ labels,
                                                               uint8_t *__restrict__ mask_out, int *__restrict__ labels_out,
                                                               long long total_px) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_px) return;
    const int fg = mask[g] & 1;
    mask_out[g] = (uint8_t)fg;
    labels_out[g] = fg ? labels[g] + 1 : 0;
}

static int db_label(const float *d_pred, int n, int h, int w, float thresh, uint8_t *mask, int *labels, cudaStream_t st,
                    bool dilate = false) {
    const size_t px = (size_t)n * h * w;
    if (dilate) {
        const long long groups = (long long)n * h * ((w + 3) / 4);
        db_mask_dilate_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(d_pred, mask, n, h, w, thresh);
        LUMINA_KERNEL_CHECK("db_mask_dilate_kernel");
    } else {
        db_mask_kernel<<<(unsigned)((px / 4 + 256) / 256), 256, 0, st>>>(d_pred, mask, px, thresh);
        LUMINA_KERNEL_CHECK("db_mask_kernel");
    }
    const long long nseg = (long long)n * h * ((w + 31) / 32);
    if ((w & 3) == 0 && !getenv("LUMINA_DB_GLOBAL_CCL")) {   // the global union-find stays as the A/B and fallback path
        const int tiles_x = (w + DBT_W - 1) / DBT_W, tiles_y = (h + DBT_H - 1) / DBT_H;
        const long long ctas = (long long)n * tiles_x * tiles_y;
        LUMINA_REQUIRE(ctas < (1LL << 31), "batch too large for grid");
        db_ccl_tile_kernel<<<(unsigned)ctas, 256, 0, st>>>(mask, labels, h, w, tiles_x, tiles_y);
        LUMINA_KERNEL_CHECK("db_ccl_tile_kernel");
        const long long rows_part = (long long)n * (tiles_y - 1) * w, cols_part = (long long)n * (tiles_x - 1) * h;
        if (rows_part + cols_part > 0) {
            db_ccl_border_kernel<<<(unsigned)((rows_part + cols_part + 255) / 256), 256, 0, st>>>(mask, labels, h, w, tiles_x, tiles_y,
                                                                                              rows_part, rows_part + cols_part);
            LUMINA_KERNEL_CHECK("db_ccl_border_kernel");
        }
        const long long fgroups = (long long)n * h * (w >> 2);
        db_ccl_flatten4_kernel<<<(unsigned)((fgroups + 255) / 256), 256, 0, st>>>(mask, labels, h, w, fgroups);
        LUMINA_KERNEL_CHECK("db_ccl_flatten4_kernel");
        return LUMINA_OK;
    }
    if ((w & 3) == 0) {
        const long long igroups = (long long)n * h * (((w + 31) >> 5) << 3);
        db_ccl_init4_kernel<<<(unsigned)((igroups + 255) / 256), 256, 0, st>>>(mask, labels, h, w, igroups);
        LUMINA_KERNEL_CHECK("db_ccl_init4_kernel");
    } else {
        db_ccl_init_kernel<<<(unsigned)((nseg * 32 + 255) / 256), 256, 0, st>>>(mask, labels, h, w, nseg);
        LUMINA_KERNEL_CHECK("db_ccl_init_kernel");
    }
    const unsigned gpx = (unsigned)((px + 255) / 256);
    const long long mgroups = (long long)n * h * ((w + 3) / 4);
    db_ccl_merge_kernel<<<(unsigned)((mgroups + 255) / 256), 256, 0, st>>>(mask, labels, h, w, mgroups);
    LUMINA_KERNEL_CHECK("db_ccl_merge_kernel");
    if ((w & 3) == 0) {
        const long long fgroups = (long long)n * h * (w >> 2);
        db_ccl_flatten4_kernel<<<(unsigned)((fgroups + 255) / 256), 256, 0, st>>>(mask, labels, h, w, fgroups);
        LUMINA_KERNEL_CHECK("db_ccl_flatten4_kernel");
    } else {
        db_ccl_flatten_kernel<<<gpx, 256, 0, st>>>(mask, labels, h, w, (long long)px);
        LUMINA_KERNEL_CHECK("db_ccl_flatten_kernel");
    }
    return LUMINA_OK;
}

}  // namespace lumina

using namespace lumina;

LUMINA_API size_t lumina_db_workspace_bytes(int n, int h, int w, int max_candidates) {
    if (n <= 0 || h <= 0 || w <= 0 || max_candidates <= 0) return 0;
    return db_layout(n, h, w, max_candidates).total + (size_t)n * 2 * 4 + 256;
}

LUMINA_API int lumina_db_postprocess(const float *d_pred, int n, int h, int w, float thresh, double box_thresh,
                                     double unclip_ratio, int max_candidates, int min_size, const int32_t *h_src_hw,
                                     int32_t *d_boxes, float *d_scores, int32_t *d_counts, void *d_workspace,
                                     size_t workspace_bytes, void *stream) {
    return lumina_db_postprocess_ex(d_pred, n, h, w, thresh, box_thresh, unclip_ratio, max_candidates, min_size, 0, h_src_hw,
                                    d_boxes, d_scores, d_counts, d_workspace, workspace_bytes, stream);
}

LUMINA_API int lumina_db_postprocess_ex(const float *d_pred, int n, int h, int w, float thresh, double box_thresh,
                                        double unclip_ratio, int max_candidates, int min_size, int flags,
                                        const int32_t *h_src_hw, int32_t *d_boxes, float *d_scores, int32_t *d_counts,
                                        void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE((flags & ~3) == 0, "unknown flag");
    LUMINA_REQUIRE(d_pred && h_src_hw && d_boxes && d_scores && d_counts && d_workspace, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && max_candidates > 0, "empty batch");
    LUMINA_REQUIRE((long long)h * w < (1LL << 30), "map too large");
    LUMINA_REQUIRE((((uintptr_t)d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
    const size_t need = lumina_db_workspace_bytes(n, h, w, max_candidates);
    if (workspace_bytes < need) return set_error(LUMINA_E_NOMEM, "db workspace too small: need %zu bytes", need);
    const DbLayout L = db_layout(n, h, w, max_candidates);
    cudaStream_t st = as_stream(stream);
    uint8_t *ws = (uint8_t *)d_workspace;
    uint8_t *mask = ws + L.mask_off;
    int *labels = (int *)(ws + L.labels_off);
    int *src_hw_dev = (int *)(ws + L.total);
    LUMINA_CUDA_TRY(cudaMemcpyAsync(src_hw_dev, h_src_hw, (size_t)n * 2 * 4, cudaMemcpyHostToDevice, st));
    LUMINA_CUDA_TRY(cudaMemsetAsync(ws + L.poolctr_off, 0, (size_t)n * 4, st));
    int rc = db_label(d_pred, n, h, w, thresh, mask, labels, st, (flags & 1) != 0);
    if (rc != LUMINA_OK) return rc;
    LUMINA_REQUIRE(n <= 65535, "batch too large for grid");
    db_candidates_kernel<<<n, 1024, 0, st>>>(mask, labels, (int *)(ws + L.cand_off), (int *)(ws + L.bbox_off),
                                             (int *)(ws + L.ncand_off), h, w, max_candidates);
    LUMINA_KERNEL_CHECK("db_candidates_kernel");
    const long long bgroups = (long long)n * h * ((w + 3) / 4);
    db_bbox_kernel<<<(unsigned)((bgroups + 255) / 256), 256, 0, st>>>(mask, labels, (int *)(ws + L.bbox_off), h, w, max_candidates,
                                                                       bgroups);
    LUMINA_KERNEL_CHECK("db_bbox_kernel");
    DbParams p;
    p.pred = d_pred; p.mask = mask; p.labels = labels;
    p.cand = (const int *)(ws + L.cand_off); p.bbox = (const int *)(ws + L.bbox_off); p.ncand = (const int *)(ws + L.ncand_off);
    p.pool = (DbgPt *)(ws + L.pool_off); p.poolctr = (int *)(ws + L.poolctr_off);
    p.accept = ws + L.accept_off; p.cmeta = (int *)(ws + L.cmeta_off); p.tmpbox = (int *)(ws + L.tmpbox_off); p.tmpscore = (float *)(ws + L.tmpscore_off);
    p.src_hw = src_hw_dev; p.pool_pts_per_map = L.pool_pts_per_map;
    p.h = h; p.w = w; p.maxc = max_candidates; p.min_size = min_size;
    p.box_thresh = box_thresh; p.unclip_ratio = unclip_ratio;
    db_cand_rows_kernel<<<dim3(div_up(max_candidates, DB_WARPS), n), DB_WARPS * 32, 0, st>>>(p);
    LUMINA_KERNEL_CHECK("db_cand_rows_kernel");
    db_cand_rect_kernel<<<dim3(div_up(max_candidates, 128), n), 128, 0, st>>>(p);
    LUMINA_KERNEL_CHECK("db_cand_rect_kernel");
    if (flags & 2) {
        db_cand_score_slow_kernel<<<dim3(div_up(max_candidates, DB_WARPS), n), DB_WARPS * 32, 0, st>>>(p);
        LUMINA_KERNEL_CHECK("db_cand_score_slow_kernel");
    } else {
        db_cand_score_kernel<<<dim3(div_up(max_candidates, DB_WARPS), n), DB_WARPS * 32, 0, st>>>(p);
        LUMINA_KERNEL_CHECK("db_cand_score_kernel");
    }
    {
        const size_t smem = (size_t)DB_UNCLIP_THREADS * 2 * (DB_OFFS_MAX + 2) * sizeof(DbgPt);
        LUMINA_CUDA_TRY(cudaFuncSetAttribute(db_cand_unclip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        db_cand_unclip_small_kernel<<<dim3(div_up(max_candidates, 128), n), 128, 0, st>>>(p);
        LUMINA_KERNEL_CHECK("db_cand_unclip_small_kernel");
        db_cand_unclip_kernel<<<dim3(div_up(max_candidates, DB_UNCLIP_THREADS), n), DB_UNCLIP_THREADS, smem, st>>>(p);
        LUMINA_KERNEL_CHECK("db_cand_unclip_kernel");
    }
    db_compact_kernel<<<n, 1024, 0, st>>>(p.accept, p.tmpbox, p.tmpscore, p.ncand, d_boxes, d_scores, d_counts, max_candidates);
    LUMINA_KERNEL_CHECK("db_compact_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_db_mask_ccl(const float *d_pred, int n, int h, int w, float thresh, uint8_t *d_mask, int32_t *d_labels,
                                  void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(d_pred && d_mask && d_labels && d_workspace, "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE((((uintptr_t)d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
    const size_t px = (size_t)n * h * w;
    const size_t need = a256(px) + px * 4;
    if (workspace_bytes < need) return set_error(LUMINA_E_NOMEM, "db label workspace too small: need %zu bytes", need);
    cudaStream_t st = as_stream(stream);
    uint8_t *mask = (uint8_t *)d_workspace;
    int *labels = (int *)((uint8_t *)d_workspace + a256(px));
    int rc = db_label(d_pred, n, h, w, thresh, mask, labels, st);
    if (rc != LUMINA_OK) return rc;
    db_export_labels_kernel<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(mask, labels, d_mask, d_labels, (long long)px);
    LUMINA_KERNEL_CHECK("db_export_labels_kernel");
    return LUMINA_OK;
}
