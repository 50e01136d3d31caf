// k_ppht.cu -- cv2.HoughLinesP (progressive probabilistic Hough transform),
// identical line list in identical order.
//
// Replaces OpenCV hough.cpp HoughLinesProbabilistic reached from
// backend/utils/image_preprocessing.py:402-407.  Arithmetic: SURVEY App. A8.
//
// The algorithm is a sequential randomised loop (cv::RNG seeded with (uint64)-1,
// swap-remove of a raster-ordered point list, vote / arg-max / line walk per
// point).  What is parallel is exploited without changing the result:
//   ppht_collect_kernel    : raster-order compaction of edge pixels (CTA-wide scan)
//   ppht_bitmask_kernel    : the edge map as a bitmask (one 32-pixel word per thread)
//   ppht_order_kernel      : the visiting order depends only on N and the seed, not on the votes, so the
//                            RNG + swap-remove permutation is produced ahead of time, 32 draws per step,
//                            in shared memory once the live list fits, with an exact in-order replay of
//                            the steps whose draws interact
//   ppht_page_order_kernel : pages are launched heaviest first (longest-processing-time-first)
//   ppht_cluster_lm_kernel : (k_ppht_cluster.cuh) accumulator rows distributed over the shared memory of
//                            a thread-block cluster, a private edge bitmask per CTA -- the product path
//   ppht_main_kernel       : last resort for pages whose theta rows do not fit 8 CTAs: one CTA per page,
//                            the 180 theta votes of a point spread over the lanes (L2 atomics on packed
//                            16-bit counters), batched with exact rank resolution
// This stage is latency-bound (a serial dependency chain per page), not bandwidth-bound.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace lumina {

struct PphtLayout {
    size_t acc_off, mask_off, nz_off, order_off, count_off, trig_off, step_off, stats_off, rholo_off, celloff_off, evbuf_off, pageorder_off, bits_off, total;
    size_t bits_stride;         // words per page of the shared edge bitmask (multiple of 4)
    size_t evbuf_words;
    size_t acc_words_per_page;  // uint32 words (2 counters per word)
    int numangle, numrho;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static PphtLayout ppht_layout(int n, int h, int w, double rho_d, double theta_d) {
    PphtLayout L;
    const float rho = (float)rho_d, theta = (float)theta_d;
    L.numangle = (int)lrint(M_PI / theta);
    L.numrho = (int)lrint(((w + h) * 2 + 1) / rho);
    const size_t px = (size_t)h * w;
    L.acc_words_per_page = ((size_t)L.numangle * L.numrho + 1) / 2;
    size_t off = 0;
    L.acc_off = off; off = align256(off + (size_t)n * L.acc_words_per_page * 4);
    L.mask_off = off; off = align256(off + (size_t)n * px);
    L.nz_off = off; off = align256(off + (size_t)n * px * 4);
    L.order_off = off; off = align256(off + (size_t)n * px * 4);
    L.count_off = off; off = align256(off + (size_t)n * 4);
    L.trig_off = off; off = align256(off + (size_t)L.numangle * 2 * 4);
    L.step_off = off; off = align256(off + (size_t)L.numangle * 3 * 4);
    L.stats_off = off; off = align256(off + (size_t)n * 8 * 4 + (size_t)n * 10 * 8);  // per page: N, votes, events, good, walk windows
    L.rholo_off = off; off = align256(off + (size_t)L.numangle * 4);
    L.celloff_off = off; off = align256(off + (size_t)L.numangle * 4);
    L.evbuf_words = ((size_t)h * w + 31) / 32;  // words of one private bitmask copy
    L.evbuf_off = off; off = align256(off + (size_t)n * 8 * (((L.evbuf_words + 3) & ~(size_t)3)) * 4);
    L.pageorder_off = off; off = align256(off + (size_t)n * 4);
    L.bits_stride = (L.evbuf_words + 3) & ~(size_t)3;
    L.bits_off = off; off = align256(off + (size_t)n * L.bits_stride * 4);
    L.total = off;
    return L;
}

// ---- A: raster-order compaction ------------------------------------------------
__global__ void __launch_bounds__(1024) ppht_collect_kernel(const uint8_t *__restrict__ edges, uint8_t *__restrict__ mask,
                                                            uint32_t *__restrict__ nz, int *__restrict__ count, int h, int w) {
    const int page = blockIdx.x;
    const int px = h * w;
    const uint8_t *e = edges + (size_t)page * px;
    uint8_t *m = mask + (size_t)page * px;
    uint32_t *out = nz + (size_t)page * px;
    __shared__ int wsum[32];
    __shared__ int chunk_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;  // uniform running count
    for (int p0 = 0; p0 < px; p0 += 1024 * 4) {
        const int p = p0 + threadIdx.x * 4;
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (p + k < px && e[p + k]) bits |= 1u << k;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (p + k < px) m[p + k] = (uint8_t)((bits >> k) & 1u);
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
            wsum[lane] = s - v;  // exclusive warp offsets
            if (lane == 31) chunk_total = s;
        }
        __syncthreads();
        int off = base + wsum[warp] + inc - c;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (bits & (1u << k)) {
                const int q = p + k;
                const int y = q / w, x = q - y * w;
                out[off++] = ((uint32_t)y << 16) | (uint32_t)x;
            }
        base += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[page] = base;
}

// ---- B: visiting order (cv::RNG MWC + swap-remove), data independent ------------
// One CTA per page; warp 0 runs the sequential process, 32 draws per step:
//   * the MWC chain of the next step is computed while the memory operations of this step are in flight;
//   * a step whose 32 draws do not interact (distinct indices, none inside the 32 tail slots) is one
//     parallel gather + scatter;
//   * a step with interacting draws is replayed in order on a 64-slot shared-memory image of the cells it
//     touches (one memory round trip, not one per draw);
//   * once the live list fits in shared memory (cap entries) it is moved there by the whole CTA and the
//     remaining steps never leave the SM.
constexpr int PORD_THREADS = 256;
__global__ void __launch_bounds__(PORD_THREADS) ppht_order_kernel(uint32_t *__restrict__ nz_all, uint32_t *__restrict__ order_all,
                                                                  const int *__restrict__ count, int px, int cap) {
    extern __shared__ uint32_t sA[];          // [cap] the live list once it fits
    __shared__ uint32_t V[64], O[32];
    __shared__ int slot_of[32];
    const int page = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *nz = nz_all + (size_t)page * px;
    uint32_t *order = order_all + (size_t)page * px;
    const int N = count[page];
    unsigned long long state = ~0ull;
    int cnt = N, i = 0;
    // the 32 draws of a step (lane k gets draw k); chained multiply-with-carry, fully unrolled for full steps
    auto draw = [&](int c) -> int {
        unsigned long long s = state;
        uint32_t my_r = 0;
        if (c >= 32) {
#pragma unroll
            for (int k = 0; k < 32; k++) {
                s = (unsigned long long)(uint32_t)s * 4164903690ull + (s >> 32);
                if (k == lane) my_r = (uint32_t)s;
            }
        } else {
            for (int k = 0; k < c; k++) {
                s = (unsigned long long)(uint32_t)s * 4164903690ull + (s >> 32);
                if (k == lane) my_r = (uint32_t)s;
            }
        }
        state = s;
        return lane < c ? (int)(my_r % (uint32_t)(c - lane)) : -1 - lane;
    };
    int idx = 0;
    if (warp == 0) {
        idx = cnt > 0 ? draw(cnt) : 0;
        // ---- phase A: the list is still larger than shared memory, cells live in L2 ----
        while (cnt > cap) {
            const int b = 32;  // cap >= 32, so every step here is full
            const unsigned same = __match_any_sync(0xffffffffu, idx);
            const int first = __ffs(same) - 1;
            const bool in_tail = idx >= cnt - b;
            const bool slow = __any_sync(0xffffffffu, first != lane || in_tail);
            const uint32_t pt = __ldcg(nz + idx), tl = __ldcg(nz + (cnt - 1 - lane));      // in flight ...
            const int idx_next = draw(cnt - b);                                            // ... while the next draws are computed
            if (!slow) {
                __stcg(nz + idx, tl);
                __stcg(order + i + lane, pt);
            } else {
                // replay in order on the touched cells: slots 0..31 = the tail cells, 32+k = first use of an index
                const int slot = in_tail ? cnt - 1 - idx : 32 + first;
                V[lane] = tl;
                if (!in_tail && first == lane) V[32 + lane] = pt;
                slot_of[lane] = slot;
                __syncwarp();
                if (lane == 0) {
                    for (int k = 0; k < b; k++) {
                        const int sk = slot_of[k];
                        const uint32_t o = V[sk];
                        V[sk] = V[k];
                        O[k] = o;
                    }
                }
                __syncwarp();
                __stcg(order + i + lane, O[lane]);
                if (!in_tail && first == lane) __stcg(nz + idx, V[32 + lane]);   // tail cells die with this step
            }
            __syncwarp();
            cnt -= b;
            i += b;
            idx = idx_next;
        }
    }
    __syncthreads();
    // ---- phase B: the remaining list lives in shared memory ----
    const int nB = N > cap ? N - ((N - cap + 31) / 32) * 32 : N;   // entries left when phase A ends (same on all threads)
    for (int t = tid; t < nB; t += PORD_THREADS) sA[t] = __ldcg(nz + t);
    __syncthreads();
    if (warp != 0) return;
    while (cnt > 0) {
        const int b = cnt < 32 ? cnt : 32;
        const bool act = lane < b;
        const unsigned same = __match_any_sync(0xffffffffu, idx);
        const int first = __ffs(same) - 1;
        const bool slow = __any_sync(0xffffffffu, act && (first != lane || idx >= cnt - b));
        uint32_t pt = 0, tl = 0;
        if (act) { pt = sA[idx]; tl = sA[cnt - 1 - lane]; }
        const int idx_next = cnt - b > 0 ? draw(cnt - b) : 0;
        if (!slow) {
            __syncwarp();
            if (act) { sA[idx] = tl; __stcg(order + i + lane, pt); }
        } else {
            slot_of[lane] = idx;
            __syncwarp();
            if (lane == 0) {
                for (int k = 0; k < b; k++) {
                    const int ik = slot_of[k];
                    const uint32_t o = sA[ik];
                    sA[ik] = sA[cnt - 1 - k];
                    O[k] = o;
                }
            }
            __syncwarp();
            if (act) __stcg(order + i + lane, O[lane]);
        }
        __syncwarp();
        cnt -= b;
        i += b;
        idx = idx_next;
    }
}

// ---- A2: edge bitmask (raster bit order): one 32-pixel word per thread ---------------------------
__global__ void __launch_bounds__(256) ppht_bitmask_kernel(const uint8_t *__restrict__ edges, uint32_t *__restrict__ bits,
                                                           int px, int stride_words) {
    const int page = blockIdx.y;
    const int wd = blockIdx.x * blockDim.x + threadIdx.x;
    if (wd >= stride_words) return;
    const uint8_t *e = edges + (size_t)page * px + (size_t)wd * 32;
    uint32_t b = 0;
    const int left = px - wd * 32;
    if (left >= 32 && ((uintptr_t)e & 15) == 0) {
        const uint4 lo = ldg_stream_u4(e), hi = ldg_stream_u4(e + 16);
        const uint32_t wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
#pragma unroll
            for (int q = 0; q < 4; q++) b |= (((wv[k] >> (8 * q)) & 0xffu) != 0u ? 1u : 0u) << (k * 4 + q);
        }
    } else {
        for (int k = 0; k < 32 && k < left; k++) b |= (e[k] != 0 ? 1u : 0u) << k;
    }
    bits[(size_t)page * stride_words + wd] = b;
}

// ---- B2: launch order of the pages: most edge points first (longest-processing-time-first keeps the
// last wave of clusters short); ties by page index.  O(n^2) compares, n is a batch of pages.
__global__ void __launch_bounds__(256) ppht_page_order_kernel(const int *__restrict__ count, int *__restrict__ page_order, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int ci = count[i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
        const int cj = __ldg(count + j);
        rank += (cj > ci) || (cj == ci && j < i);
    }
    page_order[rank] = i;
}

// ---- C: main loop -----------------------------------------------------------------
struct PphtParams {
    uint32_t *acc;         // packed 2 x 16-bit counters
    uint8_t *mask;
    const uint32_t *order;
    const int *count;
    const float *trig;     // [numangle][2] cos, sin (pre-divided by rho)
    const int *step;       // [numangle][3] xflag, dx0, dy0
    int32_t *lines;        // [n][max_lines][4]
    int32_t *nlines;       // [n]
    int32_t *stats;        // [n][8] diagnostics: N, votes, events, good lines, walk windows
    size_t acc_words_per_page;
    int h, w, numangle, numrho, threshold, line_length, line_gap, max_lines;
};

__device__ __forceinline__ int cvround_f(float v) { return __float2int_rn(v); }

// vote (+1) for point (x,y); returns packed key of this lane's best: (val << 16) | (65535 - n)
template <int NPER>
__device__ __forceinline__ uint32_t ppht_vote(uint32_t *acc, const float *tc, const float *ts, int lane, int numangle,
                                              int numrho, int x, int y) {
    const float fx = (float)x, fy = (float)y;
    uint32_t old[NPER];
    int half[NPER];
#pragma unroll
    for (int q = 0; q < NPER; q++) {
        const int n = lane + 32 * q;
        if (n < numangle) {
            const int r = cvround_f(__fadd_rn(__fmul_rn(fx, tc[q]), __fmul_rn(fy, ts[q]))) + (numrho - 1) / 2;
            const size_t cell = (size_t)n * numrho + r;
            half[q] = (int)(cell & 1);
            old[q] = atomicAdd(acc + (cell >> 1), half[q] ? 0x10000u : 1u);
        }
    }
    uint32_t best = 0;
#pragma unroll
    for (int q = 0; q < NPER; q++) {
        const int n = lane + 32 * q;
        if (n < numangle) {
            // counters are biased by PPHT_BIAS (OpenCV's int accumulator goes negative when a
            // not-yet-visited pixel of a good line is un-voted); the biased value is monotonic
            const uint32_t val = ((half[q] ? (old[q] >> 16) : (old[q] & 0xffffu)) + 1u) & 0xffffu;
            const uint32_t key = (val << 16) | (uint32_t)(65535 - n);
            best = max(best, key);
        }
    }
    return best;
}

template <int NPER>
__device__ __forceinline__ void ppht_unvote(uint32_t *acc, const float *tc, const float *ts, int lane, int numangle,
                                            int numrho, int x, int y) {
    const float fx = (float)x, fy = (float)y;
#pragma unroll
    for (int q = 0; q < NPER; q++) {
        const int n = lane + 32 * q;
        if (n < numangle) {
            const int r = cvround_f(__fadd_rn(__fmul_rn(fx, tc[q]), __fmul_rn(fy, ts[q]))) + (numrho - 1) / 2;
            const size_t cell = (size_t)n * numrho + r;
            atomicAdd(acc + (cell >> 1), (cell & 1) ? 0xffff0000u : 0xffffffffu);  // -1 in the half (result unused -> RED)
        }
    }
}

constexpr int PPHT_MAXWIN = 256;   // 32-position windows per direction (covers 8192-pixel walks)
constexpr int PPHT_BIAS = 0x4040;  // every accumulator byte is memset to 0x40
constexpr int PPHT_WARPS = 16;      // points of the visiting order in flight per page

// One line event (hough.cpp: walk both ways from the point along theta = max_n, decide good /
// not good, clear the mask along the segment, un-vote when good, emit the line).  Runs on ONE
// warp: 32 walk positions per step, ballots find the stop position, RED un-votes.
template <int NPER>
__device__ __forceinline__ void ppht_line_event(const PphtParams &p, uint32_t *acc, uint8_t *mask, int32_t *lines,
                                                const float *tc, const float *ts, uint32_t (*setbits)[PPHT_MAXWIN], int lane,
                                                int j, int i, int max_n, int &nl, int &n_win) {
    const int shift = 16;
    const int xflag = p.step[max_n * 3], dx0 = p.step[max_n * 3 + 1], dy0 = p.step[max_n * 3 + 2];
    int x0 = j, y0 = i;
    if (xflag) y0 = (y0 << shift) + (1 << (shift - 1));
    else x0 = (x0 << shift) + (1 << (shift - 1));
    int endk[2], ex[2], ey[2];
#pragma unroll
    for (int d = 0; d < 2; d++) {
        const int dx = d ? -dx0 : dx0, dy = d ? -dy0 : dy0;
        int gap = 0, ek = 0;  // position 0 is the (set) start point in walk 1
        int base = 0, win = 0;
        for (;; base += 32, win++) {
            n_win++;
            const int kp = base + lane;
            const int X = x0 + kp * dx, Y = y0 + kp * dy;
            const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
            const bool inb = j1 >= 0 && j1 < p.w && i1 >= 0 && i1 < p.h;
            const bool st = inb && __ldcg(mask + (size_t)i1 * p.w + j1) != 0;
            const unsigned bset = __ballot_sync(0xffffffffu, st);
            if (lane == 0 && win < PPHT_MAXWIN) setbits[d][win] = bset;
            // gap seen by this lane if it is unset: distance to the last set position
            const unsigned below = bset & ((2u << lane) - 1u);  // bits <= lane
            const int gk = below ? lane - (31 - __clz(below)) : gap + lane + 1;
            const bool brk = !inb || (!st && gk > p.line_gap);
            const unsigned bbrk = __ballot_sync(0xffffffffu, brk);
            if (bbrk) {
                const int fb = __ffs(bbrk) - 1;
                const unsigned sb = fb ? (bset & ((1u << fb) - 1u)) : 0u;
                if (sb) ek = base + 31 - __clz(sb);
                break;
            }
            if (bset) { ek = base + 31 - __clz(bset); gap = __clz(bset); }
            else gap += 32;
        }
        endk[d] = ek;
        const int X = x0 + ek * dx, Y = y0 + ek * dy;
        ex[d] = xflag ? X : (X >> shift);
        ey[d] = xflag ? (Y >> shift) : Y;
    }
    const bool good = abs(ex[1] - ex[0]) >= p.line_length || abs(ey[1] - ey[0]) >= p.line_length;
    __syncwarp();
#pragma unroll
    for (int d = 0; d < 2; d++) {
        const int dx = d ? -dx0 : dx0, dy = d ? -dy0 : dy0;
        for (int base = 0, win = 0; base <= endk[d]; base += 32, win++) {
            unsigned bset;
            if (win < PPHT_MAXWIN) bset = setbits[d][win];
            else {  // beyond the recorded windows (never for pages < 8192 px): re-read
                const int kp = base + lane;
                const int X = x0 + kp * dx, Y = y0 + kp * dy;
                const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                const bool inb = j1 >= 0 && j1 < p.w && i1 >= 0 && i1 < p.h;
                bset = __ballot_sync(0xffffffffu, inb && __ldcg(mask + (size_t)i1 * p.w + j1) != 0);
            }
            const int rem = endk[d] - base;  // positions base..endk
            if (rem < 31) bset &= (2u << rem) - 1u;
            if (d == 1 && base == 0) bset &= ~1u;  // start pixel already cleared by direction 0
            if (bset & (1u << lane)) {
                const int kp = base + lane;
                const int X = x0 + kp * dx, Y = y0 + kp * dy;
                const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                __stcg(mask + (size_t)i1 * p.w + j1, (uint8_t)0);
            }
            if (good) {
                unsigned bb = bset;
                while (bb) {
                    const int b = __ffs(bb) - 1;
                    bb &= bb - 1;
                    const int kp = base + b;
                    const int X = x0 + kp * dx, Y = y0 + kp * dy;
                    const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                    ppht_unvote<NPER>(acc, tc, ts, lane, p.numangle, p.numrho, j1, i1);
                }
            }
        }
    }
    if (good) {
        if (lane == 0 && nl < p.max_lines) {
            lines[nl * 4 + 0] = ex[0]; lines[nl * 4 + 1] = ey[0];
            lines[nl * 4 + 2] = ex[1]; lines[nl * 4 + 3] = ey[1];
        }
        nl++;
    }
}

// Batched main loop: one CTA per page, NW warps = NW consecutive points of the visiting order in
// flight.  Votes commute, so a batch is voted speculatively (every atomic of NW points in flight at
// once: one L2 round trip per batch instead of one per point).  The atomics return the cell values in
// an arbitrary order, but per cell the multiset of returned values equals the multiset of the values
// the sequential loop would see; hence
//   * if no returned value reaches the threshold, no point of the batch triggers a line: commit;
//   * otherwise the exact sequential value of (point k, theta) is  min(returned values of the
//     batch points hitting that cell) + #(batch points <= k hitting that cell), resolved through
//     shared memory.  The first point k* whose exact maximum reaches the threshold runs its line
//     event; points after k* take their votes back (RED -1) and are replayed in the next batch,
//     because the event changes the mask / accumulator they must see.
// The result is identical to the sequential algorithm (same lines, same order).
template <int NPER, int NW>
__global__ void __launch_bounds__(NW * 32) ppht_main_kernel(const PphtParams p) {
    constexpr int ORD_CHUNK = 1024;
    const int page = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int px = p.h * p.w;
    uint32_t *acc = p.acc + (size_t)page * p.acc_words_per_page;
    uint8_t *mask = p.mask + (size_t)page * px;
    const uint32_t *order = p.order + (size_t)page * px;
    int32_t *lines = p.lines + (size_t)page * p.max_lines * 4;
    const int N = p.count[page];
    __shared__ uint32_t setbits[2][PPHT_MAXWIN];
    __shared__ uint32_t ordbuf[ORD_CHUNK];
    __shared__ short rr[NW][NPER * 32];
    __shared__ unsigned short oo[NW][NPER * 32];
    __shared__ uint32_t wkey[NW];
    __shared__ uint32_t wspec[2][NW];  // double-buffered by batch parity: a warp is at most one batch ahead
    __shared__ int s_nl, s_stats[4];

    float tc[NPER], ts[NPER];
#pragma unroll
    for (int q = 0; q < NPER; q++) {
        const int n = lane + 32 * q;
        tc[q] = n < p.numangle ? p.trig[n * 2] : 0.f;
        ts[q] = n < p.numangle ? p.trig[n * 2 + 1] : 0.f;
    }
    if (threadIdx.x == 0) { s_nl = 0; s_stats[0] = s_stats[1] = s_stats[2] = s_stats[3] = 0; }
    int pos = 0, buf_lo = 0, buf_hi = 0;
    int n_votes = 0, n_events = 0, n_win = 0, n_batches = 0;
    __syncthreads();
    while (pos < N) {
        if (pos + NW > buf_hi && buf_hi < N) {  // refill the visiting-order window (uniform branch)
            __syncthreads();
            buf_lo = pos;
            buf_hi = min(N, pos + ORD_CHUNK);
            for (int t = threadIdx.x; t < buf_hi - buf_lo; t += NW * 32) ordbuf[t] = __ldcg(order + buf_lo + t);
            __syncthreads();
        }
        n_batches++;
        const int e = pos + w;
        const bool live = e < N;
        const uint32_t pt = live ? ordbuf[e - buf_lo] : 0u;
        const int j = (int)(pt & 0xffffu), i = (int)(pt >> 16);
        const bool m = live && __ldcg(mask + (size_t)i * p.w + j) != 0;
        // ---- speculative vote ----
        int rq[NPER];
        uint32_t oq[NPER];
        uint32_t spec = 0;
        if (m) {
            const float fx = (float)j, fy = (float)i;
            int half[NPER];
#pragma unroll
            for (int q = 0; q < NPER; q++) {
                const int n = lane + 32 * q;
                if (n < p.numangle) {
                    rq[q] = cvround_f(__fadd_rn(__fmul_rn(fx, tc[q]), __fmul_rn(fy, ts[q]))) + (p.numrho - 1) / 2;
                    const size_t cell = (size_t)n * p.numrho + rq[q];
                    half[q] = (int)(cell & 1);
                    oq[q] = atomicAdd(acc + (cell >> 1), half[q] ? 0x10000u : 1u);
                }
            }
#pragma unroll
            for (int q = 0; q < NPER; q++) {
                const int n = lane + 32 * q;
                if (n < p.numangle) {
                    oq[q] = half[q] ? (oq[q] >> 16) : (oq[q] & 0xffffu);
                    rr[w][n] = (short)rq[q];
                    oo[w][n] = (unsigned short)oq[q];
                    spec = max(spec, oq[q] + 1u);
                }
            }
            n_votes++;
        } else {
#pragma unroll
            for (int q = 0; q < NPER; q++) rr[w][lane + 32 * q] = (short)-1;
        }
        spec = __reduce_max_sync(0xffffffffu, spec);
        if (lane == 0) wspec[n_batches & 1][w] = spec;
        __syncthreads();  // (A) every vote of the batch has been performed; rr/oo/wspec visible
        {
            const uint32_t v = lane < NW ? wspec[n_batches & 1][lane] : 0u;
            const bool any_hit = __any_sync(0xffffffffu, (int)v - PPHT_BIAS >= p.threshold);
            if (!any_hit) {  // nothing in this batch can trigger: commit all votes (one barrier per batch)
                pos += NW;
                continue;
            }
        }
        // ---- exact sequential values of this warp's point ----
        uint32_t key = 0;
        if (m) {
#pragma unroll
            for (int q = 0; q < NPER; q++) {
                const int n = lane + 32 * q;
                if (n < p.numangle) {
                    uint32_t base = oq[q], rank = 1;
                    const short r = (short)rq[q];
                    for (int k = 0; k < NW; k++) {
                        if (k == w) continue;
                        if (rr[k][n] == r) {
                            base = min(base, (uint32_t)oo[k][n]);
                            rank += (k < w) ? 1u : 0u;
                        }
                    }
                    key = max(key, ((base + rank) << 16) | (uint32_t)(65535 - n));
                }
            }
        }
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0) wkey[w] = key;
        __syncthreads();  // (B) exact keys visible
        int ks;
        {
            const uint32_t v = lane < NW ? wkey[lane] : 0u;
            const unsigned hits = __ballot_sync(0xffffffffu, v != 0u && (int)(v >> 16) - PPHT_BIAS >= p.threshold);
            ks = hits ? __ffs(hits) - 1 : NW;  // the multiset bound is not tight per point: may be none
        }
        if (ks < NW) {
            if (w > ks && m) ppht_unvote<NPER>(acc, tc, ts, lane, p.numangle, p.numrho, j, i);  // replayed next batch
            if (w == ks) {
                int nl = s_nl;
                const int max_n = 65535 - (int)(wkey[ks] & 0xffffu);
                ppht_line_event<NPER>(p, acc, mask, lines, tc, ts, setbits, lane, j, i, max_n, nl, n_win);
                if (lane == 0) s_nl = nl;
                n_events++;
            }
            pos += ks + 1;
        } else {
            pos += NW;
        }
        __syncthreads();  // (C) event stores / rollbacks ordered before the next batch
    }
    // statistics (per-thread counters: votes per warp, events on the event warps)
    if (lane == 0) { atomicAdd(&s_stats[1], n_votes); atomicAdd(&s_stats[2], n_events); atomicAdd(&s_stats[3], n_win); }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.nlines[page] = s_nl;
        int32_t *st = p.stats + page * 8;
        st[0] = N; st[1] = s_stats[1]; st[2] = s_stats[2]; st[3] = s_nl; st[4] = s_stats[3]; st[5] = n_batches;
    }
}

}  // namespace lumina

#include "k_ppht_pipe.cuh"

using namespace lumina;

#ifdef LUMINA_PPHT_PROFILE
// diagnostics build only (tools/ppht_pipe_stats.py): where the per-page statistics live in the workspace
LUMINA_API size_t lumina_ppht_stats_offset(int n, int h, int w, double rho, double theta) {
    return ppht_layout(n, h, w, rho, theta).stats_off;
}
#endif

LUMINA_API size_t lumina_ppht_workspace_bytes(int n, int h, int w, double rho, double theta) {
    if (n <= 0 || h <= 0 || w <= 0 || !(rho > 0) || !(theta > 0)) return 0;
    return ppht_layout(n, h, w, rho, theta).total;
}

// phases: 1 = prepare (tables, point collection, bitmask, visiting order, page order; touches only the workspace),
//         2 = line extraction from a prepared workspace, 3 = both
static int ppht_run(const uint8_t *d_edges, int n, int h, int w, double rho_d, double theta_d, int threshold, int min_line_length,
                    int max_line_gap, int32_t *d_lines, int32_t *d_nlines, int max_lines, void *d_workspace,
                    size_t workspace_bytes, int phases, void *stream) {
    LUMINA_REQUIRE(d_edges && d_workspace && ((phases & 2) == 0 || (d_lines && d_nlines)), "null pointer");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0, "empty batch");
    LUMINA_REQUIRE(h < 65536 && w < 65536, "page too large (16-bit packed coordinates)");
    LUMINA_REQUIRE(rho_d > 0 && theta_d > 0, "rho/theta must be positive");
    const PphtLayout L = ppht_layout(n, h, w, rho_d, theta_d);
    if (workspace_bytes < L.total) return set_error(LUMINA_E_NOMEM, "ppht workspace too small: need %zu bytes", L.total);
    LUMINA_REQUIRE((((uintptr_t)d_workspace) & 255) == 0, "workspace must be 256-byte aligned");
    LUMINA_REQUIRE(L.numangle >= 1 && L.numangle <= 192, "numangle must be <= 192 (theta >= pi/192)");
    cudaStream_t st = as_stream(stream);
    uint8_t *ws = (uint8_t *)d_workspace;

    // host tables exactly as hough.cpp builds them (float theta, double cos/sin, float store)
    const float rho = (float)rho_d, theta = (float)theta_d;
    const float irho = 1 / rho;
    std::vector<float> trig((size_t)L.numangle * 2);
    std::vector<int> step((size_t)L.numangle * 3);
    for (int a = 0; a < L.numangle; a++) {
        trig[a * 2] = (float)(cos((double)a * theta) * irho);
        trig[a * 2 + 1] = (float)(sin((double)a * theta) * irho);
        const float fa = -trig[a * 2 + 1], fb = trig[a * 2];
        int xflag, dx0, dy0;
        if (fabsf(fa) > fabsf(fb)) {
            xflag = 1; dx0 = fa > 0 ? 1 : -1;
            dy0 = (int)lrintf(fb * (1 << 16) / fabsf(fa));
        } else {
            xflag = 0; dy0 = fb > 0 ? 1 : -1;
            dx0 = (int)lrintf(fa * (1 << 16) / fabsf(fb));
        }
        step[a * 3] = xflag; step[a * 3 + 1] = dx0; step[a * 3 + 2] = dy0;
    }
    const bool lpt_order = n > 1 && n <= 8192;
    if (phases & 1) {
    LUMINA_CUDA_TRY(cudaMemcpyAsync(ws + L.trig_off, trig.data(), trig.size() * 4, cudaMemcpyHostToDevice, st));
    LUMINA_CUDA_TRY(cudaMemcpyAsync(ws + L.step_off, step.data(), step.size() * 4, cudaMemcpyHostToDevice, st));
    // pageable-source async copies are staged before returning, so the vectors may die here

    ppht_collect_kernel<<<n, 1024, 0, st>>>(d_edges, ws + L.mask_off, (uint32_t *)(ws + L.nz_off), (int *)(ws + L.count_off), h, w);
    LUMINA_KERNEL_CHECK("ppht_collect_kernel");
    ppht_bitmask_kernel<<<dim3((unsigned)((L.bits_stride + 255) / 256), (unsigned)n), 256, 0, st>>>(
        d_edges, (uint32_t *)(ws + L.bits_off), h * w, (int)L.bits_stride);
    LUMINA_KERNEL_CHECK("ppht_bitmask_kernel");
    {
        int max_optin = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        LUMINA_CUDA_TRY(cudaFuncGetAttributes(&fa, ppht_order_kernel));
        long long cap = ((long long)max_optin - (long long)fa.sharedSizeBytes - 256) / 4;
        if (cap > (long long)h * w) cap = (long long)h * w;
        if (const char *e = getenv("LUMINA_PPHT_ORDER_CAP")) { const long long c = atoll(e); if (c > 0 && c < cap) cap = c; }   // experiment knob
        if (cap < 32) cap = 32;
        LUMINA_CUDA_TRY(cudaFuncSetAttribute(ppht_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(cap * 4)));
        ppht_order_kernel<<<n, PORD_THREADS, (size_t)cap * 4, st>>>((uint32_t *)(ws + L.nz_off), (uint32_t *)(ws + L.order_off),
                                                                    (const int *)(ws + L.count_off), h * w, (int)cap);
    }
    LUMINA_KERNEL_CHECK("ppht_order_kernel");
    if (lpt_order) {
        ppht_page_order_kernel<<<(n + 255) / 256, 256, 0, st>>>((const int *)(ws + L.count_off), (int *)(ws + L.pageorder_off), n);
        LUMINA_KERNEL_CHECK("ppht_page_order_kernel");
    }
    }
    if ((phases & 2) == 0) return LUMINA_OK;
    // ---- shared-memory (cluster) paths: per-theta rho range the page can reach ----
    std::vector<int> rho_lo(L.numangle), row_cells(L.numangle);
    {
        const int half = (L.numrho - 1) / 2;
        for (int a = 0; a < L.numangle; a++) {
            int lo = 1 << 30, hi = -(1 << 30);
            for (int c = 0; c < 4; c++) {
                const float fx = (float)((c & 1) ? w - 1 : 0), fy = (float)((c & 2) ? h - 1 : 0);
                const int r = (int)lrintf(fx * trig[a * 2] + fy * trig[a * 2 + 1]);
                lo = r < lo ? r : lo; hi = r > hi ? r : hi;
            }
            lo -= 1; hi += 1;  // rounding slack
            if (lo + half < 0) lo = -half;
            if (hi + half > L.numrho - 1) hi = L.numrho - 1 - half;
            rho_lo[a] = lo + half;
            row_cells[a] = hi - lo + 1;
        }
    }
    const char *force = getenv("LUMINA_PPHT");  // diagnostics: "l2" | "cluster" | "nopipe" force a slower variant
    const bool want_l2 = force && force[0] == 'l', want_cluster = force && force[0] == 'c', no_pipe = force && force[0] == 'n';
    // variant 0: accumulator slice + a private copy of the edge bitmask in each CTA's shared memory
    //            (software-pipelined kernel, or the one-phase-at-a-time kernel with LUMINA_PPHT=nopipe)
    // variant 1: accumulator slices in shared memory, the private bitmask copies in L2
    for (int variant = (want_cluster ? 1 : 0); variant < 2 && !want_l2; variant++) {
        const bool lm = variant == 0;
        const bool pipe = !no_pipe;
        int max_optin = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        const cudaError_t fe = pipe ? (lm ? cudaFuncGetAttributes(&fa, ppht_cluster_pipe_kernel<true>)
                                          : cudaFuncGetAttributes(&fa, ppht_cluster_pipe_kernel<false>))
                               : lm ? cudaFuncGetAttributes(&fa, ppht_cluster_lm_kernel<true>)
                                    : cudaFuncGetAttributes(&fa, ppht_cluster_lm_kernel<false>);
        if (fe != cudaSuccess) { cudaGetLastError(); continue; }
        const long long mask_bytes = lm ? (long long)L.bits_stride * 4 : 0;
        const long long budget = (long long)max_optin - (long long)fa.sharedSizeBytes - 1024 - mask_bytes;
        for (int c = 1; c <= 8 && budget > 0; c++) {
            const int T = (L.numangle + c - 1) / c;
            if (T > PCL_THREADS) continue;
            int worst = 0;
            std::vector<int> offs(L.numangle);
            for (int r0 = 0; r0 < L.numangle; r0 += T) {
                int sum = 0;
                for (int a = r0; a < L.numangle && a < r0 + T; a++) { offs[a] = sum; sum += row_cells[a]; }
                worst = sum > worst ? sum : worst;
            }
            if ((long long)worst * 2 + 16 > budget) continue;
            LUMINA_CUDA_TRY(cudaMemcpyAsync(ws + L.rholo_off, rho_lo.data(), rho_lo.size() * 4, cudaMemcpyHostToDevice, st));
            LUMINA_CUDA_TRY(cudaMemcpyAsync(ws + L.celloff_off, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice, st));
            PphtLmParams q;
            q.bits = (const uint32_t *)(ws + L.bits_off); q.bits_stride = (int)L.bits_stride;
            q.gbits = (uint32_t *)(ws + L.evbuf_off);
            q.order = (const uint32_t *)(ws + L.order_off); q.count = (const int *)(ws + L.count_off);
            q.trig = (const float *)(ws + L.trig_off); q.step = (const int *)(ws + L.step_off);
            q.rho_lo = (const int *)(ws + L.rholo_off); q.cell_off = (const int *)(ws + L.celloff_off);
            q.page_order = lpt_order ? (const int *)(ws + L.pageorder_off) : nullptr;
            q.lines = d_lines; q.nlines = d_nlines; q.stats = (int32_t *)(ws + L.stats_off);
            q.stats_ll = (long long *)(ws + L.stats_off + (((size_t)n * 8 * 4 + 7) & ~(size_t)7));
            q.h = h; q.w = w; q.numangle = L.numangle; q.numrho = L.numrho; q.theta_per_cta = T;
            q.slice_cells = worst; q.threshold = threshold; q.line_length = min_line_length;
            q.line_gap = max_line_gap; q.max_lines = max_lines;
            const size_t dyn = (((size_t)worst * 2 + 15) & ~(size_t)15) + (size_t)mask_bytes;
            if (pipe && lm) LUMINA_CUDA_TRY(cudaFuncSetAttribute(ppht_cluster_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            else if (pipe) LUMINA_CUDA_TRY(cudaFuncSetAttribute(ppht_cluster_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            else if (lm) LUMINA_CUDA_TRY(cudaFuncSetAttribute(ppht_cluster_lm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            else LUMINA_CUDA_TRY(cudaFuncSetAttribute(ppht_cluster_lm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(n * c));
            cfg.blockDim = dim3(pipe ? PPI_THREADS : PCL_THREADS);
            cfg.dynamicSmemBytes = dyn;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (pipe && lm) LUMINA_CUDA_TRY(cudaLaunchKernelEx(&cfg, ppht_cluster_pipe_kernel<true>, q));
            else if (pipe) LUMINA_CUDA_TRY(cudaLaunchKernelEx(&cfg, ppht_cluster_pipe_kernel<false>, q));
            else if (lm) LUMINA_CUDA_TRY(cudaLaunchKernelEx(&cfg, ppht_cluster_lm_kernel<true>, q));
            else LUMINA_CUDA_TRY(cudaLaunchKernelEx(&cfg, ppht_cluster_lm_kernel<false>, q));
            LUMINA_KERNEL_CHECK(pipe ? "ppht_cluster_pipe_kernel" : "ppht_cluster_lm_kernel");
            return LUMINA_OK;
        }
    }
    // ---- (3) fallback: accumulator in L2 (pages whose rows do not fit 8 CTAs of shared memory) ----
    LUMINA_CUDA_TRY(cudaMemsetAsync(ws + L.acc_off, 0x40, (size_t)n * L.acc_words_per_page * 4, st));
    PphtParams p;
    p.acc = (uint32_t *)(ws + L.acc_off); p.mask = ws + L.mask_off;
    p.order = (const uint32_t *)(ws + L.order_off); p.count = (const int *)(ws + L.count_off);
    p.trig = (const float *)(ws + L.trig_off); p.step = (const int *)(ws + L.step_off);
    p.lines = d_lines; p.nlines = d_nlines; p.stats = (int32_t *)(ws + L.stats_off); p.acc_words_per_page = L.acc_words_per_page;
    p.h = h; p.w = w; p.numangle = L.numangle; p.numrho = L.numrho;
    p.threshold = threshold; p.line_length = min_line_length; p.line_gap = max_line_gap; p.max_lines = max_lines;
    ppht_main_kernel<6, PPHT_WARPS><<<n, PPHT_WARPS * 32, 0, st>>>(p);
    LUMINA_KERNEL_CHECK("ppht_main_kernel");
    return LUMINA_OK;
}

LUMINA_API int lumina_ppht(const uint8_t *d_edges, int n, int h, int w, double rho_d, double theta_d, int threshold,
                           int min_line_length, int max_line_gap, int32_t *d_lines, int32_t *d_nlines, int max_lines,
                           void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(max_lines > 0, "empty batch");
    return ppht_run(d_edges, n, h, w, rho_d, theta_d, threshold, min_line_length, max_line_gap, d_lines, d_nlines, max_lines,
                    d_workspace, workspace_bytes, 3, stream);
}

// The same call in two halves, so that a pipeline can run the wide preparation kernels and the long cluster
// kernel on different streams: lumina_ppht_prepare + lumina_ppht_lines(same arguments) == lumina_ppht.
LUMINA_API int lumina_ppht_prepare(const uint8_t *d_edges, int n, int h, int w, double rho_d, double theta_d, void *d_workspace,
                                   size_t workspace_bytes, void *stream) {
    return ppht_run(d_edges, n, h, w, rho_d, theta_d, 0, 0, 0, nullptr, nullptr, 1, d_workspace, workspace_bytes, 1, stream);
}
LUMINA_API int lumina_ppht_lines(const uint8_t *d_edges, int n, int h, int w, double rho_d, double theta_d, int threshold,
                                 int min_line_length, int max_line_gap, int32_t *d_lines, int32_t *d_nlines, int max_lines,
                                 void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(max_lines > 0, "empty batch");
    return ppht_run(d_edges, n, h, w, rho_d, theta_d, threshold, min_line_length, max_line_gap, d_lines, d_nlines, max_lines,
                    d_workspace, workspace_bytes, 2, stream);
}

// (iv)+(v) image_preprocessing.py:414-428 on the host.  np.median of already folded per-line angles: the middle
// element, or the mean of the two middle elements, of the sorted values -- selection yields the same values as a sort.
static double median_of(double *a, int n) {
    std::nth_element(a, a + n / 2, a + n);
    const double hi = a[n / 2];
    if (n & 1) return hi;
    const double lo = *std::max_element(a, a + n / 2);
    return (lo + hi) / 2.0;
}
// glibc atan2 for the per-line angle.  NOTE: numpy's arctan2 is glibc's only where numpy does not dispatch to its
// bundled SIMD math (AVX-512 builds use SVML, whose result differs from glibc's in the last place for ~0.3 % of
// integer (dy, dx) pairs -- measured, DESIGN "deskew angle").  The Python host layer therefore computes the per-line
// angles with numpy itself (the reference's own expression) and calls lumina_deskew_decide_angles_host; this entry
// is for hosts without numpy.
LUMINA_API double lumina_median_angle_host(const int32_t *h_lines, int nlines) {
    if (!h_lines || nlines <= 0) return 0.0;
    std::vector<double> ang((size_t)nlines);
    for (int i = 0; i < nlines; i++) {
        const double dy = (double)(h_lines[i * 4 + 3] - h_lines[i * 4 + 1]);
        const double dx = (double)(h_lines[i * 4 + 2] - h_lines[i * 4 + 0]);
        double a = atan2(dy, dx) * (180.0 / M_PI);
        if (a < -45) a += 90;
        else if (a > 45) a -= 90;
        ang[i] = a;
    }
    return median_of(ang.data(), nlines);
}

// A small persistent pool for the per-page decisions (no thread creation on the path the GPU waits on).  run() is
// serialised; the workers live for the life of the process.
namespace {
class DecidePool {
  public:
    DecidePool() {
        unsigned hw = std::thread::hardware_concurrency();
        int nt = hw > 1 ? (int)hw - 1 : 1;
        if (nt > 7) nt = 7;                         // + the calling thread
        for (int t = 0; t < nt; t++) std::thread([this] { worker(); }).detach();
    }
    void run(int n, const std::function<void(int)> &fn) {
        std::lock_guard<std::mutex> serial(run_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; next_.store(0); done_ = 0; gen_++;
        }
        cv_.notify_all();
        int mine = 0;
        for (int i; (i = next_.fetch_add(1)) < n;) { fn(i); mine++; }
        std::unique_lock<std::mutex> lk(mu_);
        done_ += mine;
        cv_done_.wait(lk, [&] { return done_ >= n_ && active_ == 0; });   // no worker still holds this call's fn
        fn_ = nullptr;
    }

  private:
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *fn;
            int n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                fn = fn_; n = n_;
                if (fn) active_++;
            }
            if (!fn) continue;
            int mine = 0;
            for (int i; (i = next_.fetch_add(1)) < n;) { (*fn)(i); mine++; }
            {
                std::lock_guard<std::mutex> lk(mu_);
                done_ += mine;
                active_--;
                cv_done_.notify_all();
            }
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, done_ = 0, active_ = 0;
    unsigned long long gen_ = 0;
};
DecidePool &decide_pool() {
    static DecidePool *p = new DecidePool();   // never destroyed: detached workers may outlive static destructors
    return *p;
}
}  // namespace

// The per-page decision of deskew (image_preprocessing.py:409-444) for a whole batch in one host call:
// no lines -> (0.0, keep); |median| < 0.5 -> (median, keep); |median| > 45 -> (0.0, keep); else rotate by
// getRotationMatrix2D((w//2, h//2), median, 1.0).
static void deskew_gate(double a, int has_lines, int h, int w, double *angle, double *m6, uint8_t *apply) {
    *angle = 0.0;
    *apply = 0;
    for (int k = 0; k < 6; k++) m6[k] = 0.0;
    if (!has_lines) return;
    if (fabs(a) < 0.5) { *angle = a; return; }
    if (fabs(a) > 45) return;
    *angle = a;
    *apply = 1;
    lumina_rotation_matrix_host((double)(w / 2), (double)(h / 2), a, 1.0, m6);
}

LUMINA_API void lumina_deskew_decide_host(const int32_t *h_lines, const int32_t *h_nlines, int n, int lines_stride,
                                          int h, int w, double *h_angles, double *h_m6, uint8_t *h_apply) {
    auto one = [=](int i) {
        const int nl = h_nlines[i] < lines_stride ? h_nlines[i] : lines_stride;
        const double a = nl > 0 ? lumina_median_angle_host(h_lines + (size_t)i * lines_stride * 4, nl) : 0.0;
        deskew_gate(a, nl > 0, h, w, h_angles + i, h_m6 + (size_t)i * 6, h_apply + i);
    };
    // ~20 us per page (libm atan2 per line + selection); the GPU waits for this, so a batch is spread over a few
    // host threads (pages are independent; same libm, same result)
    if (n < 16) {
        for (int i = 0; i < n; i++) one(i);
        return;
    }
    decide_pool().run(n, one);
}

// The same decision from per-line angles the caller computed (degrees, already folded to +-45): the Python host layer
// evaluates np.degrees(np.arctan2(dy, dx)) with numpy -- the reference's own expression, :421 -- so that the angle is
// the reference's on every host, whatever math library its numpy dispatches to.  h_line_angles [n][angle_stride] is
// not modified.
LUMINA_API void lumina_deskew_decide_angles_host(const double *h_line_angles, const int32_t *h_nlines, int n, int angle_stride,
                                                 int h, int w, double *h_angles, double *h_m6, uint8_t *h_apply) {
    auto one = [=](int i) {
        const int nl = h_nlines[i] < angle_stride ? h_nlines[i] : angle_stride;
        double a = 0.0;
        if (nl > 0) {
            thread_local std::vector<double> tmp;
            tmp.assign(h_line_angles + (size_t)i * angle_stride, h_line_angles + (size_t)i * angle_stride + nl);
            a = median_of(tmp.data(), nl);
        }
        deskew_gate(a, nl > 0, h, w, h_angles + i, h_m6 + (size_t)i * 6, h_apply + i);
    };
    if (n < 16) {
        for (int i = 0; i < n; i++) one(i);
        return;
    }
    decide_pool().run(n, one);
}
