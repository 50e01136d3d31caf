// k_ppht_cluster.cuh -- HoughLinesP main loop with the accumulator in (distributed) shared memory.
//
// The L2-atomic kernel in k_ppht.cu is bound by the ~180 scattered L2 accesses every vote costs
// (measured: ~7 cycles per returning atomic per SM).  Here a page is owned by a thread-block CLUSTER
// of CS CTAs: CTA c keeps the accumulator rows of theta in [c*T, (c+1)*T) in its own shared memory
// (16-bit biased counters, only the rho range the page can reach: ~0.64*(w+h) cells per theta), so a
// vote is a shared-memory read-modify-write by the thread that owns the theta row: no atomics, no L2.
//   * 32 points of the (precomputed) visiting order per batch; the theta-thread applies them in
//     order (groups of 4, duplicates resolved in registers), so every per-(point,theta) value is the
//     exact sequential value; a row that reaches the threshold posts value<<16 | first theta with a
//     shared-memory atomicMax (rows below the threshold post nothing);
//   * warp 0 sends the CTA's 32 keys to every CTA of the cluster with st.async stores that complete
//     the receiver's transaction mbarrier and waits on its own: no cluster barrier, no fence;
//   * warp 0 alone runs the line events that are not "good" (such a line only clears mask pixels, so
//     the scan continues inside the batch); every CTA replays them on its private mask copy; the CTA
//     joins for a good line (un-vote by all threads) or when a later batch point was cleared (those
//     points take their votes back and the batch restarts behind the event).
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
#ifdef LUMINA_PPHT_PROFILE
#define PCL_TICK(i) do { const long long t__ = clock64(); tph[i] += t__ - tc0; tc0 = t__; } while (0)
#else
#define PCL_TICK(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// When a CTA can also hold the page's edge mask as a bitmask (h*w/8 bytes) next
// to its accumulator slice, every CTA of the cluster keeps its own copy and replays each line event
// locally (the walk is deterministic given the mask), so events need no cluster barrier, no L2 round
// trips and no published pixel list.  The only cross-CTA traffic is 32 keys per batch over DSMEM.
// ---------------------------------------------------------------------------------------------
// cvRound(x*cos + y*sin) + half_rho.  The float -> int conversion (round half to even) is done by adding
// 1.5 * 2^23: exact for |v| < 2^22 (rho never exceeds 2 * 65535), and it stays off the quarter-rate
// conversion pipe.
__device__ __forceinline__ int pcl_round(float v) {
    return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}
__device__ __forceinline__ int pcl_rho(float fx, float fy, float c, float s, int half_rho) {
    return pcl_round(__fadd_rn(__fmul_rn(fx, c), __fmul_rn(fy, s))) + half_rho;
}

// ---- cluster key exchange: remote shared-memory stores that complete a transaction barrier -------
__device__ __forceinline__ uint32_t pcl_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pcl_mapa(uint32_t cta_addr, int cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void pcl_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pcl_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pcl_st_async(uint32_t cluster_addr, uint32_t value, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(cluster_addr), "r"(value), "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void pcl_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PCL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PCL_DONE;\n"
        "bra PCL_WAIT;\n"
        "PCL_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

struct PphtLmParams {
    const uint32_t *bits;    // [n][bits_stride] edge bitmask (raster bit order), built by ppht_bitmask_kernel
    int bits_stride;         // words per page (multiple of 4)
    uint32_t *gbits;         // [n][8][(h*w+31)/32] per-CTA private edge bitmasks in L2 (LM=false variant only)
    const uint32_t *order;
    const int *count;
    const float *trig;
    const int *step;
    const int *rho_lo;
    const int *cell_off;
    const int *page_order;   // [n] pages in launch order (heaviest first), or null
    int32_t *lines, *nlines, *stats;
    long long *stats_ll;
    int h, w, numangle, numrho, theta_per_cta, slice_cells;
    int threshold, line_length, line_gap, max_lines;
};

// apply `sign` (+1 vote / -1 take back) for up to 4 points at once on this thread's theta row.  Entries
// with r < 0 are inactive (and pairwise distinct, so they never look like duplicates).  Loads are issued
// together; duplicates inside the group are resolved in registers so the values equal the sequential
// ones.  Returns the post-update values in v[].
__device__ __forceinline__ void pcl_group_update(unsigned short *row, const int r[4], int sign, int v[4]) {
    int a[4];
#pragma unroll
    for (int g = 0; g < 4; g++) a[g] = r[g] >= 0 ? (int)row[r[g]] : 0;
    v[0] = a[0] + sign;
    v[1] = a[1] + sign * (1 + (r[1] == r[0]));
    v[2] = a[2] + sign * (1 + (r[2] == r[0]) + (r[2] == r[1]));
    v[3] = a[3] + sign * (1 + (r[3] == r[0]) + (r[3] == r[1]) + (r[3] == r[2]));
#pragma unroll
    for (int g = 0; g < 4; g++)
        if (r[g] >= 0) row[r[g]] = (unsigned short)v[g];  // in order: a later duplicate carries the final count
}

// Every CTA of the cluster owns a private copy of the page's edge bitmask and replays each line event on
// it, so events never need cluster-wide synchronisation.
// LM = true : the copy lives in shared memory next to the accumulator slice (fastest per page, but at
//             678x960 it needs 3 CTAs per page: 49 pages resident on 148 SMs).
// LM = false: the copy lives in L2 (81 KB per CTA at 678x960, 10 MB per 64-page batch); mask reads cost an
//             L2 round trip, but large pages (bitmask > shared memory) are covered and the slices of a
//             678x960 page fit 2 CTAs.
// Per batch of 32 points of the visiting order:
//   votes   : the row warps compute the 32 rho values of their theta row up front (independent), then
//             apply them in order, 4 at a time; a row that reaches the threshold posts its key
//             (value << 16 | first theta) with a shared-memory atomicMax -- rows below the threshold
//             (almost all) post nothing, so there is no per-point reduction;
//   exchange: warp 0 sends the CTA's 32 keys to every CTA of the cluster (DSMEM), one cluster barrier;
//   events  : warp 0 alone walks, tests and clears the lines that are not "good" (both directions
//             interleaved in one warp, no CTA barrier); the CTA joins only for a good line (un-vote by
//             all threads) or when later points of the batch have to take their votes back.
template <bool LM>
__global__ void __launch_bounds__(PCL_THREADS) ppht_cluster_lm_kernel(const PphtLmParams p) {
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks();
    const int rank = (int)cl.block_rank();
    const int page = p.page_order ? p.page_order[blockIdx.x / CS] : (int)(blockIdx.x / CS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px = p.h * p.w;
    const uint32_t *bits = p.bits + (size_t)page * p.bits_stride;
    uint32_t *gbits = LM ? nullptr : p.gbits + ((size_t)page * 8 + rank) * p.bits_stride;
    const uint32_t *order = p.order + (size_t)page * px;
    int32_t *lines = p.lines + (size_t)page * p.max_lines * 4;
    const int N = p.count[page];
    const int half_rho = (p.numrho - 1) / 2;
    const int thr_b = p.threshold + PCL_BIAS;
    // unset run length that stops a walk, and the shift schedule that detects such runs in a 32-bit window
    // (doubling steps 1, 2, 4, ... then the remainder; six 5-bit fields, unused ones 0)
    const int gap_m = p.line_gap + 1;
    const bool gap_small = gap_m <= 32;
    uint32_t gap_sh = 0;
    if (gap_small) {
        int len = 1, i = 0;
        for (; 2 * len <= gap_m; len *= 2) gap_sh |= (uint32_t)len << (5 * i++);
        if (len < gap_m) gap_sh |= (uint32_t)(gap_m - len) << (5 * i);
    }

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
    __shared__ uint32_t hkey[2][PCL_B];                  // [parity][compact slot]: best key of the rows that reached the threshold
    __shared__ float2 lpt[PCL_THREADS / 32][PCL_B];      // per-warp compact table of the batch's live points (x, y)
    __shared__ uint32_t keys[2][8][PCL_B];
    __shared__ uint32_t setbits[2][PPHT_MAXWIN];
    __shared__ uint32_t evpx[PCL_EVMAX];
    __shared__ int ev_n, ev_end[2], ev_ex[2], ev_ey[2], ev_done[2];
    __shared__ int s_status, s_ks, s_maxn;
    __shared__ __align__(8) unsigned long long xbar[2];  // key-exchange transaction barriers, one per batch parity
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
    if (tid < 2 * PCL_B) hkey[tid / PCL_B][tid % PCL_B] = 0u;
    if (tid == 0) {
        pcl_mbar_init(pcl_smem_u32(&xbar[0]), 1);
        pcl_mbar_init(pcl_smem_u32(&xbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < p.slice_cells; i += PCL_THREADS) acc[i] = (unsigned short)PCL_BIAS;
    for (int i = tid; i < p.numangle * 3; i += PCL_THREADS) s_step[i] = p.step[i];
    // private copy of the page's edge bitmask (128-bit loads; bits_stride is a multiple of 4 words)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(bits);
        uint4 *dst = LM ? reinterpret_cast<uint4 *>(mbits) : reinterpret_cast<uint4 *>(gbits);
        const int nq = (((px + 31) >> 5) + 3) >> 2;
        for (int i = tid; i < nq; i += PCL_THREADS) {
            const uint4 v = __ldg(src + i);
            if (LM) dst[i] = v; else __stcg(dst + i, v);
        }
    }
    // un-vote mapping: thread = (row ut, pixel phase us); the row constants stay in registers
    const int nthpad = row_warps * 32;
    const int ugroups = nthpad > 0 ? PCL_THREADS / nthpad : 0;
    const int ut = nthpad > 0 ? tid % nthpad : 0, us = nthpad > 0 ? tid / nthpad : 0;
    const bool uact = us < ugroups && ut < nth;
    float ucos = 0.f, usin = 0.f;
    int ubase = 0;
    if (uact) {
        ucos = p.trig[(th0 + ut) * 2]; usin = p.trig[(th0 + ut) * 2 + 1];
        ubase = p.cell_off[th0 + ut] - p.rho_lo[th0 + ut] + half_rho;
    }

    int pos = 0, buf_lo = 0, buf_hi = 0;
    int nl = 0, n_votes = 0, n_events = 0, n_batches = 0, n_fast = 0;
#ifdef LUMINA_PPHT_PROFILE
    long long tph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = clock64();
#endif
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
        bool mylive = false;
        if (warp < row_warps || warp == 0) mylive = lane < nb && mask_set(mybit);
        const unsigned livebits = __ballot_sync(0xffffffffu, mylive);
        const int nlive = __popc(livebits);
        const int myslot = __popc(livebits & ((1u << lane) - 1u));
        // ---- votes: all rho values first (independent), then groups of 4 in order ----
        int rr[PCL_B];
        if (warp < row_warps) {
            if (mylive) lpt[warp][myslot] = make_float2((float)myx, (float)myy);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < PCL_B; j++) {  // branch-free: entries past nlive hold stale coordinates, masked below
                const float2 q = lpt[warp][j];
                const int r = pcl_rho(q.x, q.y, cth, sth, half_rho) - rlo;
                rr[j] = (j < nlive && has_row) ? r : -1 - (j & 3);
            }
            PCL_TICK(6);
#pragma unroll
            for (int j0 = 0; j0 < PCL_B; j0 += 4) {
                if (j0 < nlive) {
                    int v[4];
                    pcl_group_update(row, &rr[j0], +1, v);  // inactive entries come back far below the threshold
                    const bool h0 = v[0] >= thr_b, h1 = v[1] >= thr_b, h2 = v[2] >= thr_b, h3 = v[3] >= thr_b;
                    if (h0 | h1 | h2 | h3) {
                        const uint32_t kth = (uint32_t)(65535 - theta);
                        if (h0) atomicMax(&hkey[par][j0 + 0], ((uint32_t)v[0] << 16) | kth);
                        if (h1) atomicMax(&hkey[par][j0 + 1], ((uint32_t)v[1] << 16) | kth);
                        if (h2) atomicMax(&hkey[par][j0 + 2], ((uint32_t)v[2] << 16) | kth);
                        if (h3) atomicMax(&hkey[par][j0 + 3], ((uint32_t)v[3] << 16) | kth);
                    }
                }
            }
            PCL_TICK(7);
        }
        __syncthreads();
        PCL_TICK(1);
        // ---- exchange: this CTA's 32 keys go to every CTA of the cluster as asynchronous remote stores that
        // complete the receiver's transaction barrier; only warp 0 waits (no cluster-wide barrier, no fence) ----
        if (warp == 0) {
            const uint32_t key = ((livebits >> lane) & 1u) ? hkey[par][myslot] : 0u;
            __syncwarp();
            hkey[par][lane] = 0u;  // next used two batches from now
            const uint32_t bar = pcl_smem_u32(&xbar[par]);
            if (lane == 0) pcl_mbar_expect_tx(bar, (uint32_t)(CS * PCL_B * 4));
            const uint32_t slot = pcl_smem_u32(&keys[par][rank][lane]);
            for (int c = 0; c < CS; c++) pcl_st_async(pcl_mapa(slot, c), key, pcl_mapa(bar, c));
            PCL_TICK(2);
            pcl_mbar_wait(bar, (uint32_t)(((n_batches - 1) >> 1) & 1));
        }
        PCL_TICK(3);
        // ---- events: warp 0 handles every line that is not good on its own ----
        // A line that turns out "not good" only clears mask pixels: the votes of the later batch points and
        // therefore their exact values stay valid, so the scan continues inside the same batch.  Only a good
        // line (un-votes) or a cleared batch point forces the later points to be taken back and replayed.
        if (warp == 0) {
            uint32_t g = 0;
            if (lane < nb)
                for (int c = 0; c < CS; c++) g = max(g, keys[par][c][lane]);
            const bool reaches = g != 0u && (int)(g >> 16) >= thr_b;
            // every lane prepares the walk of its own point (as if it were the event) in unified 16.16
            // coordinates: X = xs + k * dxs, Y = ys + k * dys, pixel = (X >> 16, Y >> 16); unsigned
            // arithmetic, so a coordinate that leaves the page on the low side wraps above w / h.
            uint32_t my_xs = 0, my_ys = 0, my_dxs = 0, my_dys = 0;
            if (reaches) {
                const int mn = 65535 - (int)(g & 0xffffu);
                const int xf = s_step[mn * 3], d0 = s_step[mn * 3 + 1], d1 = s_step[mn * 3 + 2];
                my_xs = ((uint32_t)myx << 16) + (xf ? 0u : 0x8000u);
                my_ys = ((uint32_t)myy << 16) + (xf ? 0x8000u : 0u);
                my_dxs = xf ? ((uint32_t)d0 << 16) : (uint32_t)d0;
                my_dys = xf ? (uint32_t)d1 : ((uint32_t)d1 << 16);
            }
            int status = 0, ks = 0, max_n = 0;
            unsigned hits = __ballot_sync(0xffffffffu, reaches) & livebits;
            for (; hits; hits &= ~((2u << ks) - 1u)) {
                ks = __ffs(hits) - 1;
                n_events++;
                // ---- walk: both directions together, 32 positions per step; everything that decides is
                // warp-uniform bit arithmetic on the ballots ----
                const uint32_t xs = __shfl_sync(0xffffffffu, my_xs, ks), ys = __shfl_sync(0xffffffffu, my_ys, ks);
                const uint32_t dxs = __shfl_sync(0xffffffffu, my_dxs, ks), dys = __shfl_sync(0xffffffffu, my_dys, ks);
                // window 0 (positions 0..31 of both directions) is straight-line code: most events end here
                int bia0, bib0, eka, ekb, carrya, carryb;
                unsigned Wa0, Wb0;
                bool fina, finb;
                {
                    const uint32_t ja = (xs + lane * dxs) >> 16, ia = (ys + lane * dys) >> 16;
                    const uint32_t jb = (xs - lane * dxs) >> 16, ib = (ys - lane * dys) >> 16;
                    const bool ina = ja < (uint32_t)p.w && ia < (uint32_t)p.h, inb = jb < (uint32_t)p.w && ib < (uint32_t)p.h;
                    bia0 = (int)(ia * p.w + ja); bib0 = (int)(ib * p.w + jb);
                    const bool sa = ina && mask_set(bia0), sb = inb && mask_set(bib0);
                    const unsigned Ba = __ballot_sync(0xffffffffu, sa), Bb = __ballot_sync(0xffffffffu, sb);
                    const unsigned Oa = __ballot_sync(0xffffffffu, !ina), Ob = __ballot_sync(0xffffffffu, !inb);
                    // warp-uniform: bit k of R = the gap_m positions ending at k are all unset (position 0 is set)
                    unsigned Ra = gap_small ? ~Ba : 0u, Rb = gap_small ? ~Bb : 0u;
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        const int sh = (gap_sh >> (5 * i)) & 31;
                        Ra &= Ra << sh; Rb &= Rb << sh;
                    }
                    const unsigned brka = Ra | Oa, brkb = Rb | Ob;
                    fina = brka != 0u; finb = brkb != 0u;
                    Wa0 = fina ? Ba & ((1u << (__ffs(brka) - 1)) - 1u) : Ba;   // bit 0 (the start pixel) is always set
                    Wb0 = finb ? Bb & ((1u << (__ffs(brkb) - 1)) - 1u) : Bb;
                    eka = 31 - __clz(Wa0); ekb = 31 - __clz(Wb0);
                    carrya = __clz(Wa0); carryb = __clz(Wb0);
                }
                const bool one_window = fina && finb;
                if (!one_window) {
                    // ---- longer walks: further windows, same warp-uniform logic plus the run carried over ----
                    if (lane == 0) { setbits[0][0] = Wa0; setbits[1][0] = Wb0; }
                    for (int base = 32, win = 1;; base += 32, win++) {
                        const uint32_t k = (uint32_t)(base + lane);
                        const uint32_t ja = (xs + k * dxs) >> 16, ia = (ys + k * dys) >> 16;
                        const uint32_t jb = (xs - k * dxs) >> 16, ib = (ys - k * dys) >> 16;
                        const bool ina = fina || (ja < (uint32_t)p.w && ia < (uint32_t)p.h);
                        const bool inb = finb || (jb < (uint32_t)p.w && ib < (uint32_t)p.h);
                        const bool sa = !fina && ina && mask_set((int)(ia * p.w + ja));
                        const bool sb = !finb && inb && mask_set((int)(ib * p.w + jb));
                        const unsigned Ba = __ballot_sync(0xffffffffu, sa), Bb = __ballot_sync(0xffffffffu, sb);
                        const unsigned Oa = __ballot_sync(0xffffffffu, !ina), Ob = __ballot_sync(0xffffffffu, !inb);
                        unsigned Ra = gap_small ? ~Ba : 0u, Rb = gap_small ? ~Bb : 0u;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            const int sh = (gap_sh >> (5 * i)) & 31;
                            Ra &= Ra << sh; Rb &= Rb << sh;
                        }
                        if (!fina) {
                            // a run that started in the previous window completes at position gap_m - carry - 1
                            const int need = gap_m - carrya, z = Ba ? __ffs(Ba) - 1 : 32;
                            unsigned brk = Ra | Oa;
                            if (need <= 32 && z >= need) brk |= 1u << (need - 1);
                            unsigned Wa = Ba;
                            if (brk) { Wa = Ba & ((1u << (__ffs(brk) - 1)) - 1u); fina = true; }
                            if (Wa) { eka = base + 31 - __clz(Wa); carrya = __clz(Wa); }
                            else carrya += 32;
                            if (lane == 0 && win < PPHT_MAXWIN) setbits[0][win] = Wa;
                        }
                        if (!finb) {
                            const int need = gap_m - carryb, z = Bb ? __ffs(Bb) - 1 : 32;
                            unsigned brk = Rb | Ob;
                            if (need <= 32 && z >= need) brk |= 1u << (need - 1);
                            unsigned Wb = Bb;
                            if (brk) { Wb = Bb & ((1u << (__ffs(brk) - 1)) - 1u); finb = true; }
                            if (Wb) { ekb = base + 31 - __clz(Wb); carryb = __clz(Wb); }
                            else carryb += 32;
                            if (lane == 0 && win < PPHT_MAXWIN) setbits[1][win] = Wb;
                        }
                        if (fina && finb) break;
                    }
                }
                if (one_window) n_fast++;
                const int exa = (int)((xs + (uint32_t)eka * dxs) >> 16), eya = (int)((ys + (uint32_t)eka * dys) >> 16);
                const int exb = (int)((xs - (uint32_t)ekb * dxs) >> 16), eyb = (int)((ys - (uint32_t)ekb * dys) >> 16);
                if (abs(exb - exa) >= p.line_length || abs(eyb - eya) >= p.line_length) {
                    if (lane == 0) {
                        ev_end[0] = eka; ev_ex[0] = exa; ev_ey[0] = eya;
                        ev_end[1] = ekb; ev_ex[1] = exb; ev_ey[1] = eyb;
                        if (one_window) { setbits[0][0] = Wa0; setbits[1][0] = Wb0; }
                    }
                    max_n = 65535 - (int)(__shfl_sync(0xffffffffu, g, ks) & 0xffffu);
                    status = 2;
                    break;
                }
                // ---- not good: clear the segment (start pixel by direction a); window 0 from registers ----
                if ((Wa0 >> lane) & 1u) mask_clear(bia0);
                if (((Wb0 >> lane) & 1u) && lane != 0) mask_clear(bib0);
                if (eka >= 32 || ekb >= 32) {
                    __syncwarp();
                    for (int base = 32, win = 1; base <= eka; base += 32, win++) {
                        const unsigned bs = setbits[0][win < PPHT_MAXWIN ? win : PPHT_MAXWIN - 1];
                        if ((bs >> lane) & 1u) {
                            const uint32_t k = (uint32_t)(base + lane);
                            mask_clear((int)(((ys + k * dys) >> 16) * p.w + ((xs + k * dxs) >> 16)));
                        }
                    }
                    for (int base = 32, win = 1; base <= ekb; base += 32, win++) {
                        const unsigned bs = setbits[1][win < PPHT_MAXWIN ? win : PPHT_MAXWIN - 1];
                        if ((bs >> lane) & 1u) {
                            const uint32_t k = (uint32_t)(base + lane);
                            mask_clear((int)(((ys - k * dys) >> 16) * p.w + ((xs - k * dxs) >> 16)));
                        }
                    }
                }
                __syncwarp();
                // did the cleared segment take one of the later batch points?
                const bool stilllive = lane < nb && mask_set(mybit);
                const unsigned later = ks < 31 ? ~((2u << ks) - 1u) : 0u;
                const unsigned nowlive = __ballot_sync(0xffffffffu, stilllive);
                if ((livebits ^ nowlive) & later) { status = 1; break; }
            }
            if (lane == 0) { s_status = status; s_ks = ks; s_maxn = max_n; }
        }
        __syncthreads();
        PCL_TICK(4);
        const int status = s_status;
        if (status != 0) {
            const int ks = s_ks;
            if (status == 2) {
                // good line: warps 0/1 clear their direction and list the set pixels; then ALL threads
                // un-vote (pixel, row) pairs with shared-memory atomics on the packed 16-bit counters
                // (subtraction commutes; rows keep their owner for everything order-sensitive).
                const int max_n = s_maxn;
                const int shift = 16;
                const int xflag = s_step[max_n * 3], dx0 = s_step[max_n * 3 + 1], dy0 = s_step[max_n * 3 + 2];
                const uint32_t ept = ordbuf[pos + ks - buf_lo];
                int x0 = (int)(ept & 0xffffu), y0 = (int)(ept >> 16);
                if (xflag) y0 = (y0 << shift) + (1 << (shift - 1));
                else x0 = (x0 << shift) + (1 << (shift - 1));
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
                    // (pixel, row) pairs: thread (ut, us) takes pixels us, us + ugroups, ... on row ut
                    if (uact) {
                        uint32_t *acc32 = reinterpret_cast<uint32_t *>(acc);
                        for (int q = us; q < npx; q += ugroups) {
                            const uint32_t pt = evpx[q];
                            const int cell = ubase + pcl_round(__fadd_rn(__fmul_rn((float)(pt & 0xffffu), ucos), __fmul_rn((float)(pt >> 16), usin)));
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
            // ---- later live points of the batch take their votes back and are replayed ----
            if (warp < row_warps && has_row) {
                const int first = __popc(livebits & ((2u << ks) - 1u));  // compact slot of the first later point
#pragma unroll
                for (int j0 = 0; j0 < PCL_B; j0 += 4) {
                    if (j0 + 4 > first && j0 < nlive) {
                        int r4[4], v[4];
#pragma unroll
                        for (int g = 0; g < 4; g++) r4[g] = (j0 + g >= first) ? rr[j0 + g] : -1 - g;
                        pcl_group_update(row, r4, -1, v);
                    }
                }
            }
            n_votes += __popc(livebits & ((2u << ks) - 1u));
            pos += ks + 1;
            __syncthreads();
        } else {
            n_votes += __popc(livebits);
            pos += nb;
        }
        PCL_TICK(5);
    }
    cl.sync();  // no CTA leaves while a peer could still address its shared memory
    if (rank == 0 && tid == 0) {
        p.nlines[page] = nl;
        int32_t *st = p.stats + page * 8;
        st[0] = N; st[1] = n_votes; st[2] = n_events; st[3] = nl; st[4] = 1; st[5] = n_batches; st[6] = CS; st[7] = n_fast;
#ifdef LUMINA_PPHT_PROFILE
        for (int i = 0; i < 10; i++) p.stats_ll[page * 10 + i] = tph[i];
#endif
    }
}

}  // namespace lumina
