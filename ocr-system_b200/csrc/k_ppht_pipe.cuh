// k_ppht_pipe.cuh -- two-stage software pipeline of the cluster HoughLinesP kernel (k_ppht_cluster.cuh).
//
// In ppht_cluster_lm_kernel a batch goes through votes (row warps), key exchange and line events (warp 0) one
// after the other, each phase on one or two warps.  Here the phases of CONSECUTIVE batches overlap:
//   warp 0 ("control")   exchanges the keys of batch b and runs its line events, then takes the liveness
//                        snapshot of the batch after next;
//   warps 1..R ("rows")  meanwhile vote batch b+1 speculatively from the snapshot taken one step earlier.
// One CTA barrier per step.  The speculation holds when batch b ends without a restart and its events did
// not clear a point of batch b+1 (checked by comparing the snapshot with the mask after the events); then
// b+1's exact sequential values were observed, because a line that is not "good" never changes the
// accumulator.  Otherwise the speculative votes (and, for a restart inside b, the votes of b's later points)
// are taken back, a good line is un-voted by all threads, and the pipeline refills behind the event.
// Every CTA of the cluster runs the same deterministic schedule, so the key exchange stays one transaction
// barrier per judged batch.  Same result as the unpipelined kernel and as cv2.HoughLinesP.
#pragma once
#include "k_ppht_cluster.cuh"

namespace lumina {

constexpr int PPI_THREADS = 224;  // warp 0 control + up to 6 row warps (192 theta rows)

template <bool LM>
__global__ void __launch_bounds__(PPI_THREADS) ppht_cluster_pipe_kernel(const PphtLmParams p) {
    cg::cluster_group cl = cg::this_cluster();
    const int CS = (int)cl.num_blocks();
    const int rank = (int)cl.block_rank();
    const int page = p.page_order ? p.page_order[blockIdx.x / CS] : (int)(blockIdx.x / CS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px = p.h * p.w;
    const uint32_t *bits = p.bits + (size_t)page * p.bits_stride;
    uint32_t *gbits = LM ? nullptr : p.gbits + ((size_t)page * 8 + rank) * p.bits_stride;   // private mask copy in L2
    const uint32_t *order = p.order + (size_t)page * px;
    int32_t *lines = p.lines + (size_t)page * p.max_lines * 4;
    const int N = p.count[page];
    const int half_rho = (p.numrho - 1) / 2;
    const int thr_b = p.threshold + PCL_BIAS;
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
    __shared__ uint32_t hkey[2][PCL_B];       // [batch parity][compact slot]
    __shared__ float2 lpt[2][PCL_B];          // [batch parity] compact table of the live points (x, y)
    __shared__ uint32_t s_live[2];            // [batch parity] liveness snapshot the votes were made from
    __shared__ uint32_t keys[2][8][PCL_B];    // [exchange parity][cta][point]
    __shared__ uint32_t setbits[2][PPHT_MAXWIN];
    __shared__ uint32_t evpx[PCL_EVMAX];
    __shared__ int ev_n, ev_end[2], ev_ex[2], ev_ey[2], ev_done[2];
    __shared__ int s_status, s_ks, s_maxn, s_nxt_ok;
    __shared__ __align__(8) unsigned long long xbar[2];
    __shared__ int s_step[192 * 3];

    const int th0 = rank * p.theta_per_cta;
    const int nth = max(0, min(p.theta_per_cta, p.numangle - th0));
    const int row_warps = (nth + 31) >> 5;
    const int rt = tid - 32;                       // row index of a row-warp thread
    const bool is_row_warp = warp >= 1 && warp <= row_warps;
    const bool has_row = rt >= 0 && rt < nth;
    const int theta = th0 + rt;
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
    for (int i = tid; i < p.slice_cells; i += PPI_THREADS) acc[i] = (unsigned short)PCL_BIAS;
    for (int i = tid; i < p.numangle * 3; i += PPI_THREADS) s_step[i] = p.step[i];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(bits);
        uint4 *dst = LM ? reinterpret_cast<uint4 *>(mbits) : reinterpret_cast<uint4 *>(gbits);
        const int nq = (((px + 31) >> 5) + 3) >> 2;
        for (int i = tid; i < nq; i += PPI_THREADS) {
            const uint4 v = __ldg(src + i);
            if (LM) dst[i] = v; else __stcg(dst + i, v);
        }
    }
    // un-vote mapping: thread = (row ut, pixel phase us)
    const int nthpad = row_warps * 32;
    const int ugroups = nthpad > 0 ? PPI_THREADS / nthpad : 0;
    const int ut = nthpad > 0 ? tid % nthpad : 0, us = nthpad > 0 ? tid / nthpad : 0;
    const bool uact = us < ugroups && ut < nth;
    float ucos = 0.f, usin = 0.f;
    int ubase = 0;
    if (uact) {
        ucos = p.trig[(th0 + ut) * 2]; usin = p.trig[(th0 + ut) * 2 + 1];
        ubase = p.cell_off[th0 + ut] - p.rho_lo[th0 + ut] + half_rho;
    }

    int buf_lo = 0, buf_hi = 0;
    int nl = 0, n_events = 0, n_batches = 0, n_exch = 0, n_flush = 0;
#ifdef LUMINA_PPHT_PROFILE
    long long tph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // control warp: 0 exchange wait, 1 events, 2 snapshot, 3 barrier wait, 4 flush, 9 total
    long long tc0 = clock64();
    const long long t_begin = tc0;
#endif
    __syncthreads();
    cl.sync();

    // liveness snapshot + compact point table of the batch at `pos` (warp 0 only)
    auto prepare = [&](int pos, int q) {
        const int nb = max(0, min(PCL_B, N - pos));
        const uint32_t pt = lane < nb ? ordbuf[pos + lane - buf_lo] : 0u;
        const int x = (int)(pt & 0xffffu), y = (int)(pt >> 16);
        const bool live = lane < nb && mask_set(y * p.w + x);
        const unsigned lb = __ballot_sync(0xffffffffu, live);
        if (live) lpt[q][__popc(lb & ((1u << lane) - 1u))] = make_float2((float)x, (float)y);
        if (lane == 0) s_live[q] = lb;
    };
    // make the order window cover [from, from + 3 batches) (all threads)
    auto refill = [&](int from) {
        if (from + 3 * PCL_B > buf_hi && buf_hi < N) {
            __syncthreads();
            buf_lo = from;
            buf_hi = min(N, from + PCL_ORD);
            for (int t = tid; t < buf_hi - buf_lo; t += PPI_THREADS) ordbuf[t] = __ldcg(order + buf_lo + t);
            __syncthreads();
        }
    };

    // pipeline state (uniform): batch `cur` (voted, to be judged) and `nxt` (to be voted this step)
    int pos_cur = 0, nb_cur = 0, qc = 0;
    bool have_cur = false;
    int pos_nxt = 0, qn = 0;
    int rr_cur[PCL_B], rr_nxt[PCL_B];
#pragma unroll
    for (int j = 0; j < PCL_B; j++) { rr_cur[j] = -1 - (j & 3); rr_nxt[j] = -1 - (j & 3); }
    refill(0);
    if (warp == 0) prepare(0, 0);
    __syncthreads();

    while (have_cur || pos_nxt < N) {
        refill(have_cur ? pos_cur : pos_nxt);
        const int nb_nxt = max(0, min(PCL_B, N - pos_nxt));
        n_batches++;
        if (warp == 0) {
            int status = 0, ks = 0, max_n = 0;
            if (have_cur) {
                // ---- exchange the keys of `cur` ----
                const int xp = n_exch & 1;
                const unsigned livebits = s_live[qc];
                const int myslot = __popc(livebits & ((1u << lane) - 1u));
                const uint32_t key = ((livebits >> lane) & 1u) ? hkey[qc][myslot] : 0u;
                __syncwarp();
                hkey[qc][lane] = 0u;
                const uint32_t bar = pcl_smem_u32(&xbar[xp]);
                if (lane == 0) pcl_mbar_expect_tx(bar, (uint32_t)(CS * PCL_B * 4));
                const uint32_t slot = pcl_smem_u32(&keys[xp][rank][lane]);
                for (int c = 0; c < CS; c++) pcl_st_async(pcl_mapa(slot, c), key, pcl_mapa(bar, c));
                PCL_TICK(5);
                pcl_mbar_wait(bar, (uint32_t)((n_exch >> 1) & 1));
                PCL_TICK(0);
                // ---- line events of `cur` (same code path as ppht_cluster_lm_kernel) ----
                const int nb = nb_cur;
                const uint32_t mypt = lane < nb ? ordbuf[pos_cur + lane - buf_lo] : 0u;
                const int myx = (int)(mypt & 0xffffu), myy = (int)(mypt >> 16);
                const int mybit = myy * p.w + myx;
                uint32_t g = 0;
                if (lane < nb)
                    for (int c = 0; c < CS; c++) g = max(g, keys[xp][c][lane]);
                const bool reaches = g != 0u && (int)(g >> 16) >= thr_b;
                uint32_t my_xs = 0, my_ys = 0, my_dxs = 0, my_dys = 0;
                if (reaches) {
                    const int mn = 65535 - (int)(g & 0xffffu);
                    const int xf = s_step[mn * 3], d0 = s_step[mn * 3 + 1], d1 = s_step[mn * 3 + 2];
                    my_xs = ((uint32_t)myx << 16) + (xf ? 0u : 0x8000u);
                    my_ys = ((uint32_t)myy << 16) + (xf ? 0x8000u : 0u);
                    my_dxs = xf ? ((uint32_t)d0 << 16) : (uint32_t)d0;
                    my_dys = xf ? (uint32_t)d1 : ((uint32_t)d1 << 16);
                }
                unsigned hits = __ballot_sync(0xffffffffu, reaches) & livebits;
                for (; hits; hits &= ~((2u << ks) - 1u)) {
                    ks = __ffs(hits) - 1;
                    n_events++;
                    const uint32_t xs = __shfl_sync(0xffffffffu, my_xs, ks), ys = __shfl_sync(0xffffffffu, my_ys, ks);
                    const uint32_t dxs = __shfl_sync(0xffffffffu, my_dxs, ks), dys = __shfl_sync(0xffffffffu, my_dys, ks);
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
                        unsigned Ra = gap_small ? ~Ba : 0u, Rb = gap_small ? ~Bb : 0u;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            const int sh = (gap_sh >> (5 * i)) & 31;
                            Ra &= Ra << sh; Rb &= Rb << sh;
                        }
                        const unsigned brka = Ra | Oa, brkb = Rb | Ob;
                        fina = brka != 0u; finb = brkb != 0u;
                        Wa0 = fina ? Ba & ((1u << (__ffs(brka) - 1)) - 1u) : Ba;
                        Wb0 = finb ? Bb & ((1u << (__ffs(brkb) - 1)) - 1u) : Bb;
                        eka = 31 - __clz(Wa0); ekb = 31 - __clz(Wb0);
                        carrya = __clz(Wa0); carryb = __clz(Wb0);
                    }
                    const bool one_window = fina && finb;
                    if (!one_window) {
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
                    const bool stilllive = lane < nb && mask_set(mybit);
                    const unsigned later = ks < 31 ? ~((2u << ks) - 1u) : 0u;
                    const unsigned nowlive = __ballot_sync(0xffffffffu, stilllive);
                    if ((livebits ^ nowlive) & later) { status = 1; break; }
                }
                n_exch++;
                PCL_TICK(1);
            }
            int nxt_ok = 1;
            if (status == 0) {
                __syncwarp();
                // did the events of `cur` clear a point the speculative votes of `nxt` were made for?
                const uint32_t pt = lane < nb_nxt ? ordbuf[pos_nxt + lane - buf_lo] : 0u;
                const bool live = lane < nb_nxt && mask_set((int)(pt >> 16) * p.w + (int)(pt & 0xffffu));
                nxt_ok = __ballot_sync(0xffffffffu, live) == s_live[qn];
                if (nxt_ok) prepare(pos_nxt + nb_nxt, qn ^ 1);   // the batch after `nxt` (parity of `cur`, now free)
            }
            if (lane == 0) { s_status = status; s_ks = ks; s_maxn = max_n; s_nxt_ok = nxt_ok; }
            PCL_TICK(2);
        } else if (is_row_warp && nb_nxt > 0) {
            // ---- speculative votes of `nxt` from its snapshot ----
            const unsigned livebits = s_live[qn];
            const int nlive = __popc(livebits);
#pragma unroll
            for (int j = 0; j < PCL_B; j++) {
                const float2 q = lpt[qn][j];
                const int r = pcl_rho(q.x, q.y, cth, sth, half_rho) - rlo;
                rr_nxt[j] = (j < nlive && has_row) ? r : -1 - (j & 3);
            }
#pragma unroll
            for (int j0 = 0; j0 < PCL_B; j0 += 4) {
                if (j0 < nlive) {
                    int v[4];
                    pcl_group_update(row, &rr_nxt[j0], +1, v);
                    const bool h0 = v[0] >= thr_b, h1 = v[1] >= thr_b, h2 = v[2] >= thr_b, h3 = v[3] >= thr_b;
                    if (h0 | h1 | h2 | h3) {
                        const uint32_t kth = (uint32_t)(65535 - theta);
                        if (h0) atomicMax(&hkey[qn][j0 + 0], ((uint32_t)v[0] << 16) | kth);
                        if (h1) atomicMax(&hkey[qn][j0 + 1], ((uint32_t)v[1] << 16) | kth);
                        if (h2) atomicMax(&hkey[qn][j0 + 2], ((uint32_t)v[2] << 16) | kth);
                        if (h3) atomicMax(&hkey[qn][j0 + 3], ((uint32_t)v[3] << 16) | kth);
                    }
                }
            }
        }
        __syncthreads();
        if (warp == 0) { PCL_TICK(3); }
        const int status = s_status, nxt_ok = s_nxt_ok;
        if (status == 0 && nxt_ok) {
            // commit: `nxt` becomes the batch to judge, the prepared batch becomes `nxt`
            if (is_row_warp) {   // only the row warps keep rho tables; the control warp goes straight to the next exchange
#pragma unroll
                for (int j = 0; j < PCL_B; j++) rr_cur[j] = rr_nxt[j];
            }
            pos_cur = pos_nxt; nb_cur = nb_nxt; qc = qn;
            have_cur = nb_nxt > 0;
            pos_nxt += nb_nxt; qn ^= 1;
            continue;
        }
        // ---- the speculation failed: take everything back that depends on it ----
        n_flush++;
        const int ks = s_ks;
        if (status == 2) {
            // good line: warps 0/1 clear their direction and list the set pixels; then ALL threads un-vote
            const int max_n = s_maxn;
            const int shift = 16;
            const int xflag = s_step[max_n * 3], dx0 = s_step[max_n * 3 + 1], dy0 = s_step[max_n * 3 + 2];
            const uint32_t ept = ordbuf[pos_cur + ks - buf_lo];
            int x0 = (int)(ept & 0xffffu), y0 = (int)(ept >> 16);
            if (xflag) y0 = (y0 << shift) + (1 << (shift - 1));
            else x0 = (x0 << shift) + (1 << (shift - 1));
            int done[2] = {0, 0};
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
                        if (d == 1 && base == 0) bset &= ~1u;
                        int slot0 = 0;
                        if (lane == 0) slot0 = atomicAdd(&ev_n, __popc(bset));
                        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                        if (slot0 + 32 > PCL_EVMAX) {
                            if (lane == 0) atomicSub(&ev_n, __popc(bset));
                            break;
                        }
                        if (bset & (1u << lane)) {
                            const int kp = base + lane;
                            const int X = x0 + kp * dx, Y = y0 + kp * dy;
                            const int j1 = xflag ? X : (X >> shift), i1 = xflag ? (Y >> shift) : Y;
                            mask_clear(i1 * p.w + j1);
                            evpx[slot0 + __popc(bset & ((1u << lane) - 1u))] = ((uint32_t)i1 << 16) | (uint32_t)j1;
                        }
                    }
                    if (lane == 0) ev_done[d] = win;
                }
                __syncthreads();
                const int npx = ev_n;
                done[0] = ev_done[0]; done[1] = ev_done[1];
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
        if (is_row_warp && has_row) {
            // the speculative votes of `nxt` ...
            if (nb_nxt > 0) {
                const int nlive = __popc(s_live[qn]);
#pragma unroll
                for (int j0 = 0; j0 < PCL_B; j0 += 4) {
                    if (j0 < nlive) {
                        int v[4];
                        pcl_group_update(row, &rr_nxt[j0], -1, v);
                    }
                }
            }
            // ... and, for a restart inside `cur`, the votes of its points behind the event
            if (status != 0) {
                const unsigned lb = s_live[qc];
                const int nlive = __popc(lb);
                const int first = __popc(lb & ((2u << ks) - 1u));
#pragma unroll
                for (int j0 = 0; j0 < PCL_B; j0 += 4) {
                    if (j0 + 4 > first && j0 < nlive) {
                        int r4[4], v[4];
#pragma unroll
                        for (int g = 0; g < 4; g++) r4[g] = (j0 + g >= first) ? rr_cur[j0 + g] : -1 - g;
                        pcl_group_update(row, r4, -1, v);
                    }
                }
            }
        }
        if (tid < PCL_B) hkey[qn][tid] = 0u;      // keys posted by the discarded votes
        if (status != 0) pos_nxt = pos_cur + ks + 1;   // else: `cur` is done, `nxt` is voted again from a fresh snapshot
        have_cur = false;
        __syncthreads();
        refill(pos_nxt);
        if (warp == 0) prepare(pos_nxt, qn);
        __syncthreads();
        if (warp == 0) { PCL_TICK(4); }
    }
    cl.sync();
    if (rank == 0 && tid == 0) {
        p.nlines[page] = nl;
        int32_t *st = p.stats + page * 8;
        st[0] = N; st[1] = n_flush; st[2] = n_events; st[3] = nl; st[4] = 2; st[5] = n_batches; st[6] = CS; st[7] = n_exch;
#ifdef LUMINA_PPHT_PROFILE
        tph[9] = clock64() - t_begin;
        for (int i = 0; i < 10; i++) p.stats_ll[page * 10 + i] = tph[i];
#endif
    }
}

}  // namespace lumina
