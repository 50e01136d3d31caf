// k_ppht_cluster.cuh -- HoughLinesP main loop with the accumulator in (distributed) shared memory.
//
// The L2-atomic kernel in k_ppht.cu is bound by the ~180 scattered L2 accesses every vote costs
// (measured: ~7 cycles per returning atomic per SM).  Here a page is owned by a thread-block CLUSTER
// of CS CTAs: CTA c keeps the accumulator rows of theta in [c*T, (c+1)*T) in its own shared memory
// (16-bit biased counters, only the rho range the page can reach: ~0.64*(w+h) cells per theta), so a
// vote is a shared-memory read-modify-write by the thread that owns the theta row: no atomics, no L2.
//   * 32 points of the (precomputed) visiting order per batch; the theta-thread applies them in
//     order (groups of 4, duplicates resolved in registers), so every per-(point,theta) value is the
//     exact sequential value;
//   * per point the row warps reduce max/first-theta (redux.sync), CTAs exchange their 32 keys through
//     DSMEM (st.shared::cluster) and one cluster barrier, then all CTAs take the same decision;
//   * the first point that reaches the threshold runs its line event, replayed by every CTA (the walk
//     is deterministic given the mask); a line that is not "good" only clears mask pixels, so the scan
//     continues inside the batch; a good line (un-votes) or a cleared batch point makes the later
//     points take their votes back and replay.
// Identical result to cv2.HoughLinesP (same lines, same order).
#pragma once
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace lumina {

constexpr int PCL_THREADS = 192;   // >= theta rows per CTA
constexpr int PCL_B = 32;          // points per batch
constexpr int PCL_ORD = 1024;      // visiting-order window in shared memory
constexpr int PCL_MKWIN = 256;     // mask values refreshed per L2 trip
constexpr int PCL_BIAS = 0x4000;
constexpr int PCL_EVMAX = 512;     // un-vote pixel list entries per chunk
#define PCL_TICK(i) do { const long long t__ = clock64(); tph[i] += t__ - tc0; tc0 = t__; } while (0)

// ---------------------------------------------------------------------------------------------
// When a CTA can also hold the page's edge mask as a bitmask (h*w/8 bytes) next
// to its accumulator slice, every CTA of the cluster keeps its own copy and replays each line event
// locally (the walk is deterministic given the mask), so events need no cluster barrier, no L2 round
// trips and no published pixel list.  The only cross-CTA traffic is 32 keys per batch over DSMEM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int pcl_rho(float fx, float fy, float c, float s, int half_rho) {
    return __float2int_rn(__fadd_rn(__fmul_rn(fx, c), __fmul_rn(fy, s))) + half_rho;
}

struct PphtLmParams {
    const uint8_t *edges;    // [n][h][w] the Canny output (nonzero = edge)
    uint32_t *gbits;         // [n][8][(h*w+31)/32] per-CTA private edge bitmasks in L2 (LM=false variant only)
    const uint32_t *order;
    const int *count;
    const float *trig;
    const int *step;
    const int *rho_lo;
    const int *cell_off;
    int32_t *lines, *nlines, *stats;
    long long *stats_ll;
    int h, w, numangle, numrho, theta_per_cta, slice_cells;
    int threshold, line_length, line_gap, max_lines;
};

// apply `sign` (+1 vote / -1 take back) for up to 4 points at once on this thread's theta row.
// Loads are issued together; duplicates inside the group are resolved in registers so the values
// equal the sequential ones.  Returns the post-update values in v[].
__device__ __forceinline__ void pcl_group_update(unsigned short *row, const int r[4], int cnt, int sign, int v[4]) {
    int a[4];
#pragma unroll
    for (int g = 0; g < 4; g++) a[g] = g < cnt ? (int)row[r[g]] : 0;
    v[0] = a[0] + sign;
    v[1] = a[1] + sign * (1 + (r[1] == r[0]));
    v[2] = a[2] + sign * (1 + (r[2] == r[0]) + (r[2] == r[1]));
    v[3] = a[3] + sign * (1 + (r[3] == r[0]) + (r[3] == r[1]) + (r[3] == r[2]));
#pragma unroll
    for (int g = 0; g < 4; g++)
        if (g < cnt) row[r[g]] = (unsigned short)v[g];  // in order: a later duplicate carries the larger count
}

// Every CTA of the cluster owns a private copy of the page's edge bitmask and replays each line event on
// it, so events never need cluster-wide synchronisation.
// LM = true : the copy lives in shared memory next to the accumulator slice (fastest per page, but at
//             678x960 it needs 3 CTAs per page: 49 pages resident on 148 SMs).
// LM = false: the copy lives in L2 (81 KB per CTA at 678x960, 10 MB per 64-page batch); mask reads cost an
//             L2 round trip, but large pages (bitmask > shared memory) are covered and the slices of a
//             678x960 page fit 2 CTAs.
template <bool LM>
__global__ void __launch_bounds__(PCL_THREADS) ppht_cluster_lm_kernel(const PphtLmParams p) {
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks();
    const int rank = (int)cl.block_rank();
    const int page = blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px = p.h * p.w;
    const uint8_t *edges = p.edges + (size_t)page * px;
    const int nwords_all = (px + 31) >> 5;
    uint32_t *gbits = LM ? nullptr : p.gbits + ((size_t)page * 8 + rank) * nwords_all;
    const uint32_t *order = p.order + (size_t)page * px;
    int32_t *lines = p.lines + (size_t)page * p.max_lines * 4;
    const int N = p.count[page];
    const int half_rho = (p.numrho - 1) / 2;

    extern __shared__ __align__(16) unsigned char dynsm[];
    unsigned short *acc = reinterpret_cast<unsigned short *>(dynsm);                              // [slice_cells]
    uint32_t *mbits = reinterpret_cast<uint32_t *>(dynsm + (((size_t)p.slice_cells * 2 + 15) & ~(size_t)15));  // [(px+31)/32]
    auto mask_set = [&](int bidx) -> bool {
        if (LM) return (mbits[bidx >> 5] >> (bidx & 31)) & 1u;
        return (__ldcg(gbits + (bidx >> 5)) >> (bidx & 31)) & 1u;
    };
    auto mask_clear = [&](int bidx) {
        if (LM) atomicAnd(&mbits[bidx >> 5], ~(1u << (bidx & 31)));
        else atomicAnd(gbits + (bidx >> 5), ~(1u << (bidx & 31)));  // result unused: RED to L2
    };
    __shared__ uint32_t ordbuf[PCL_ORD];
    __shared__ uint32_t wkeys[PCL_THREADS / 32][PCL_B];   // [row warp][compact slot of the live point]
    __shared__ float2 lpt[PCL_THREADS / 32][PCL_B];      // per-warp compact table of the batch's live points (x, y)
    __shared__ uint32_t keys[2][8][PCL_B];
    __shared__ uint32_t setbits[2][PPHT_MAXWIN];
    __shared__ uint32_t evpx[PCL_EVMAX];
    __shared__ int ev_n, ev_end[2], ev_ex[2], ev_ey[2], ev_done[2];
    __shared__ float s_cos[PCL_THREADS], s_sin[PCL_THREADS];  // row tables for the all-thread un-vote
    __shared__ int s_rlo[PCL_THREADS], s_coff[PCL_THREADS];
    __shared__ int s_step[192 * 3];  // per-theta walk direction table (xflag, dx0, dy0)

    const int th0 = rank * p.theta_per_cta;
    const int nth = max(0, min(p.theta_per_cta, p.numangle - th0));
    const int row_warps = (nth + 31) >> 5;
    const bool has_row = tid < nth;
    const int theta = th0 + tid;
    float cth = 0.f, sth = 0.f;
    int rlo = 0;
    unsigned short *row = acc;
    if (has_row) { cth = p.trig[theta * 2]; sth = p.trig[theta * 2 + 1]; rlo = p.rho_lo[theta]; row = acc + p.cell_off[theta]; }
    s_cos[tid] = cth; s_sin[tid] = sth; s_rlo[tid] = rlo; s_coff[tid] = has_row ? p.cell_off[theta] : 0;
    for (int i = tid; i < p.slice_cells; i += PCL_THREADS) acc[i] = (unsigned short)PCL_BIAS;
    for (int i = tid; i < p.numangle * 3; i += PCL_THREADS) s_step[i] = p.step[i];
    // bitmask of the edge map (raster bit order), built from global memory once
    const int nwords = (px + 31) >> 5;
    for (int wd = warp; wd < nwords; wd += PCL_THREADS / 32) {
        const int q = wd * 32 + lane;
        const bool e = q < px && __ldg(edges + q) != 0;
        const unsigned b = __ballot_sync(0xffffffffu, e);
        if (lane == 0) { if (LM) mbits[wd] = b; else __stcg(gbits + wd, b); }
    }

    int pos = 0, buf_lo = 0, buf_hi = 0;
    int nl = 0, n_votes = 0, n_events = 0, n_batches = 0;
    long long tph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = clock64();
    __syncthreads();
    cl.sync();
    PCL_TICK(9);

    while (pos < N) {
        if (pos + PCL_B > buf_hi && buf_hi < N) {
            __syncthreads();
            buf_lo = pos;
            buf_hi = min(N, pos + PCL_ORD);
            for (int t = tid; t < buf_hi - buf_lo; t += PCL_THREADS) ordbuf[t] = __ldcg(order + buf_lo + t);
            __syncthreads();
        }
        PCL_TICK(0);
        const int par = n_batches & 1;
        n_batches++;
        const int nb = min(PCL_B, N - pos);
        // lane k of every warp looks at point k of the batch
        const uint32_t mypt = lane < nb ? ordbuf[pos + lane - buf_lo] : 0u;
        const int myx = (int)(mypt & 0xffffu), myy = (int)(mypt >> 16);
        const int mybit = myy * p.w + myx;
        const bool mylive = lane < nb && mask_set(mybit);
        const unsigned livebits = __ballot_sync(0xffffffffu, mylive);
        const float myfx = (float)myx, myfy = (float)myy;
        // ---- votes: groups of 4 live points, all rows of this warp in lock step ----
        const int nlive = __popc(livebits);
        const int myslot = __popc(livebits & ((1u << lane) - 1u));
        if (warp < row_warps) {
            if (mylive) lpt[warp][myslot] = make_float2(myfx, myfy);
            __syncwarp();
            for (int j0 = 0; j0 < nlive; j0 += 4) {
                const int cnt = min(4, nlive - j0);
                int r[4], v[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    r[g] = -1 - g;
                    if (g < cnt) {
                        const float2 q = lpt[warp][j0 + g];
                        r[g] = pcl_rho(q.x, q.y, cth, sth, half_rho) - rlo;
                    }
                }
                if (has_row) pcl_group_update(row, r, cnt, +1, v);
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    if (g < cnt) {
                        uint32_t key = has_row ? (((uint32_t)v[g] << 16) | (uint32_t)(65535 - theta)) : 0u;
                        key = __reduce_max_sync(0xffffffffu, key);
                        if (lane == 0) wkeys[warp][j0 + g] = key;
                    }
                }
            }
        }
        __syncthreads();
        PCL_TICK(1);
        // ---- combine the row warps, publish this CTA's 32 keys to every CTA of the cluster ----
        if (warp == 0) {
            uint32_t key = 0;
            if ((livebits >> lane) & 1u)
                for (int wv = 0; wv < row_warps; wv++) key = max(key, wkeys[wv][myslot]);
            for (int c = 0; c < CS; c++) *cl.map_shared_rank(&keys[par][rank][lane], c) = key;
        }
        PCL_TICK(2);
        cl.sync();
        PCL_TICK(3);
        // every lane k holds the cluster-wide key (exact sequential value, first theta) of point k
        uint32_t g = 0;
        if (lane < nb)
            for (int c = 0; c < CS; c++) g = max(g, keys[par][c][lane]);
        const bool reaches = g != 0u && (int)(g >> 16) - PCL_BIAS >= p.threshold;
        int scan_lo = 0;
        bool restart = false;
        // A line that turns out "not good" only clears mask pixels: the votes of the later batch points and
        // therefore their exact values stay valid, so the scan continues inside the same batch.  Only a good
        // line (un-votes) or a cleared batch point forces the later points to be taken back and replayed.
        for (;;) {
            const unsigned hits = __ballot_sync(0xffffffffu, reaches && lane >= scan_lo) & livebits;
            if (!hits) break;
            const int ks = __ffs(hits) - 1;
            const uint32_t gkey = __shfl_sync(0xffffffffu, g, ks);
            n_events++;
            const int ej = __shfl_sync(0xffffffffu, myx, ks), ei = __shfl_sync(0xffffffffu, myy, ks);
            const int max_n = 65535 - (int)(gkey & 0xffffu);
            const int shift = 16;
            const int xflag = s_step[max_n * 3], dx0 = s_step[max_n * 3 + 1], dy0 = s_step[max_n * 3 + 2];
            int x0 = ej, y0 = ei;
            if (xflag) y0 = (y0 << shift) + (1 << (shift - 1));
            else x0 = (x0 << shift) + (1 << (shift - 1));
            // ---- walk 1: warp d follows direction d on this CTA's mask copy ----
            if (warp < 2) {
                const int d = warp;
                const int dx = d ? -dx0 : dx0, dy = d ? -dy0 : dy0;
                int gap = 0, ek = 0, base = 0, win = 0;
                for (;; base += 32, win++) {
                    const int kp = base + lane;
                    const int X = x0 + kp * dx, Y = y0 + kp * dy;
                    const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                    const bool inb = j1 >= 0 && j1 < p.w && i1 >= 0 && i1 < p.h;
                    const int bidx = i1 * p.w + j1;
                    const bool st = inb && mask_set(bidx);
                    const unsigned bset = __ballot_sync(0xffffffffu, st);
                    if (lane == 0 && win < PPHT_MAXWIN) setbits[d][win] = bset;
                    const unsigned below = bset & ((2u << lane) - 1u);
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
                if (lane == 0) {
                    const int X = x0 + ek * dx, Y = y0 + ek * dy;
                    ev_end[d] = ek;
                    ev_ex[d] = xflag ? X : (X >> shift);
                    ev_ey[d] = xflag ? (Y >> shift) : Y;
                }
            }
            __syncthreads();
            const bool good = abs(ev_ex[1] - ev_ex[0]) >= p.line_length || abs(ev_ey[1] - ev_ey[0]) >= p.line_length;
            if (!good) {
                if (warp < 2) {  // clear the segment (start pixel included by direction 0)
                    const int d = warp;
                    const int dx = d ? -dx0 : dx0, dy = d ? -dy0 : dy0;
                    const int end_d = ev_end[d];
                    for (int base = 0, win = 0; base <= end_d; base += 32, win++) {
                        unsigned bset = setbits[d][win < PPHT_MAXWIN ? win : PPHT_MAXWIN - 1];
                        const int rem = end_d - base;
                        if (rem < 31) bset &= (2u << rem) - 1u;
                        if (bset & (1u << lane)) {
                            const int kp = base + lane;
                            const int X = x0 + kp * dx, Y = y0 + kp * dy;
                            const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                            const int bidx = i1 * p.w + j1;
                            mask_clear(bidx);
                        }
                    }
                }
            } else {
                // good line: warps 0/1 clear their direction and list the set pixels; then ALL threads
                // un-vote (pixel, row) pairs with shared-memory atomics on the packed 16-bit counters
                // (subtraction commutes; rows keep their owner for everything order-sensitive).
                int done[2] = {0, 0};  // windows already consumed per direction (chunked when the list is full)
                for (;;) {
                    if (tid == 0) ev_n = 0;
                    __syncthreads();
                    if (warp < 2) {
                        const int d = warp;
                        const int dx = d ? -dx0 : dx0, dy = d ? -dy0 : dy0;
                        const int end_d = ev_end[d];
                        int win = done[d];
                        for (int base = win * 32; base <= end_d; base += 32, win++) {
                            unsigned bset = setbits[d][win < PPHT_MAXWIN ? win : PPHT_MAXWIN - 1];
                            const int rem = end_d - base;
                            if (rem < 31) bset &= (2u << rem) - 1u;
                            if (d == 1 && base == 0) bset &= ~1u;  // the start pixel belongs to direction 0
                            int slot0 = 0;
                            if (lane == 0) slot0 = atomicAdd(&ev_n, __popc(bset));
                            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                            if (slot0 + 32 > PCL_EVMAX) {  // list full: give the slots back, continue in the next chunk
                                if (lane == 0) atomicSub(&ev_n, __popc(bset));
                                break;
                            }
                            if (bset & (1u << lane)) {
                                const int kp = base + lane;
                                const int X = x0 + kp * dx, Y = y0 + kp * dy;
                                const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                                const int bidx = i1 * p.w + j1;
                                mask_clear(bidx);
                                evpx[slot0 + __popc(bset & ((1u << lane) - 1u))] = ((uint32_t)i1 << 16) | (uint32_t)j1;
                            }
                        }
                        if (lane == 0) ev_done[d] = win;
                    }
                    __syncthreads();
                    const int npx = ev_n;
                    done[0] = ev_done[0]; done[1] = ev_done[1];
                    // (pixel, row) pairs spread over every thread of the CTA
                    if (nth > 0) {
                        uint32_t *acc32 = reinterpret_cast<uint32_t *>(acc);
                        for (int pr = tid; pr < npx * nth; pr += PCL_THREADS) {
                            const int q = pr / nth, t = pr - q * nth;
                            const uint32_t pt = evpx[q];
                            const int cell = s_coff[t] + pcl_rho((float)(pt & 0xffffu), (float)(pt >> 16), s_cos[t], s_sin[t], half_rho) - s_rlo[t];
                            atomicSub(&acc32[cell >> 1], (cell & 1) ? 0x10000u : 1u);
                        }
                    }
                    const bool more = done[0] * 32 <= ev_end[0] || done[1] * 32 <= ev_end[1];
                    __syncthreads();
                    if (!more) break;
                }
                if (rank == 0 && tid == 0 && nl < p.max_lines) {
                    lines[nl * 4 + 0] = ev_ex[0]; lines[nl * 4 + 1] = ev_ey[0];
                    lines[nl * 4 + 2] = ev_ex[1]; lines[nl * 4 + 3] = ev_ey[1];
                }
                nl++;
            }
            __syncthreads();  // mask clears visible; ev_* / setbits may be reused
            // did the cleared segment take one of the later batch points?
            const bool stilllive = lane < nb && mask_set(mybit);
            const unsigned later = ks < 31 ? ~((2u << ks) - 1u) : 0u;
            const unsigned nowlive = __ballot_sync(0xffffffffu, stilllive);
            if (good || ((livebits ^ nowlive) & later)) {
                // ---- later live points of the batch take their votes back and are replayed ----
                if (warp < row_warps) {
                    const int first = __popc(livebits & ((2u << ks) - 1u));  // compact slot of the first later point
                    for (int j0 = first; j0 < nlive; j0 += 4) {
                        const int cnt = min(4, nlive - j0);
                        int r[4], v[4];
#pragma unroll
                        for (int gq = 0; gq < 4; gq++) {
                            r[gq] = -1 - gq;
                            if (gq < cnt) {
                                const float2 q = lpt[warp][j0 + gq];
                                r[gq] = pcl_rho(q.x, q.y, cth, sth, half_rho) - rlo;
                            }
                        }
                        if (has_row) pcl_group_update(row, r, cnt, -1, v);
                    }
                }
                n_votes += __popc(livebits & ((2u << ks) - 1u));
                pos += ks + 1;
                restart = true;
                __syncthreads();
                break;
            }
            scan_lo = ks + 1;
        }
        PCL_TICK(5);
        if (!restart) {
            n_votes += __popc(livebits);
            pos += nb;
        }
    }
    if (rank == 0 && tid == 0) {
        p.nlines[page] = nl;
        int32_t *st = p.stats + page * 8;
        st[0] = N; st[1] = n_votes; st[2] = n_events; st[3] = nl; st[4] = 1; st[5] = n_batches; st[6] = CS;
        for (int i = 0; i < 10; i++) p.stats_ll[page * 10 + i] = tph[i];
    }
}

}  // namespace lumina
