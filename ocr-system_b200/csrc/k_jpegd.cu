// k_jpegd.cu -- baseline-JPEG page ingest on the device ("next" row 3 of SURVEY 8f: page ingest).
//
// Reference: backend/utils/image_preprocessing.py:57-75 (load_image / load_image_bytes: Image.open + mode rule)
// and backend/services/ocr_service.py:494-496,716-718 (the engine is handed files / bytes).  Today those bytes
// are decoded by Pillow -> libjpeg-turbo on one host core (65-90 ms per A4 page) and the 26 MB raster crosses
// PCIe; here the *file* crosses PCIe (1-2 MB) and is decoded in HBM to the same raster, byte for byte:
//   jpegd_unstuff_*_kernel scan bytes -> big-endian word stream without 0xFF00 stuffing / RSTn markers (count, scatter)
//   jpegd_entropy_kernel   self-synchronising parallel Huffman decode (jpegd_core.h), one launch: sub-sequence
//                          decode from guessed states, in-CTA propagation of exit states, CTA-to-CTA chain with
//                          provisional / final hand-over, block-index scan, coefficient write
//   jpegd_dc_kernel        DC prediction = segmented prefix sum of the differences (reset per restart interval)
//   jpegd_idct_kernel      dequantise + jidctint.c islow IDCT + level shift / clamp -> component planes
//   jpegd_colour_kernel    jdsample.c fancy upsampling (h2v2 / h2v1 / none) + jdcolor.c YCbCr->RGB -> NHWC
// Files outside the subset (progressive, CMYK, 4:4:0, multi-scan) are reported as LUMINA_E_UNSUPPORTED by the
// host-side probe and stay on the host codec, as in the reference.
#include <vector>

#include "common.cuh"
#include "jpegd_core.h"

namespace lumina {

#ifndef LUMINA_JD_SWL
#define LUMINA_JD_SWL 7
#endif
constexpr uint32_t kJdSwl = LUMINA_JD_SWL;   // log2(32-bit words per sub-sequence)
constexpr int kJdSubBits = 32 << kJdSwl;     // sub-sequence length (bits of unstuffed stream)
#ifndef LUMINA_JD_CHUNK
#define LUMINA_JD_CHUNK 256
#endif
constexpr int kJdChunk = LUMINA_JD_CHUNK;        // sub-sequences (threads) per CTA of the entropy kernel
constexpr int kUnstuffThreads = 1024;

__constant__ uint8_t c_jd_zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ---- workspace carve-up (host) -------------------------------------------------------------------------------
struct JdLayout {
    size_t blob, pages, stream, stream_bits, n_rst, rst, tiles, chain, chain_base, ticket, coef, dc, planes, total;
    size_t stream_words, blob_bytes, nblk, plane_bytes, dc_stride;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static JdLayout jd_layout(int n, int h, int w, int channels, int hs, int vs, size_t total_file_bytes) {
    JdLayout L;
    int mcux = (w + 8 * hs - 1) / (8 * hs), mcuy = (h + 8 * vs - 1) / (8 * vs);
    int bpm = channels == 1 ? 1 : hs * vs + 2;
    L.nblk = (size_t)mcux * mcuy * bpm;
    L.plane_bytes = L.nblk * 64;
    L.blob_bytes = align_up(total_file_bytes + 64, 256);
    // every page's stream region: its stuffed length rounded up + 16 bytes of zero padding, in words
    L.stream_words = total_file_bytes / 4 + (size_t)n * 16 + 64;
    size_t max_cta = (total_file_bytes * 8) / ((size_t)kJdSubBits * kJdChunk) + 2 * (size_t)n + 8;
    size_t max_rst = (size_t)n * (size_t)mcux * mcuy + 8;
    size_t max_tiles = total_file_bytes / (kUnstuffThreads * 16) + 2 * (size_t)n + 8;
    {
        size_t nmcu = (size_t)mcux * mcuy, cnt0 = channels == 1 ? 1 : (size_t)hs * vs;
        L.dc_stride = (nmcu * cnt0 + 7) / 8 * 8 + (channels == 1 ? 0 : 2 * ((nmcu + 7) / 8 * 8));
    }
    size_t o = 0;
    L.blob = o, o += L.blob_bytes;
    L.pages = o, o += align_up((size_t)n * sizeof(JdPage), 256);
    L.stream = o, o += align_up(L.stream_words * 4, 256);
    L.stream_bits = o, o += align_up((size_t)n * 4, 256);
    L.n_rst = o, o += align_up((size_t)n * 4, 256);
    L.rst = o, o += align_up(max_rst * 4, 256);
    L.tiles = o, o += align_up(max_tiles * sizeof(uint64_t), 256);
    L.chain = o, o += align_up(max_cta * 8, 256);
    L.chain_base = o, o += align_up(max_cta * 4, 256);
    L.ticket = o, o += 256;
    L.coef = o, o += align_up((size_t)n * L.nblk * 128, 256);
    L.dc = o, o += align_up((size_t)n * L.dc_stride * 2, 256);
    L.planes = o, o += align_up((size_t)n * L.plane_bytes, 256);
    L.total = o;
    return L;
}

// ---- 1. unstuff ----------------------------------------------------------------------------------------------
// A byte is dropped when it is the 0x00 behind a 0xFF, a fill 0xFF, or part of an RSTn marker; the first other
// marker ends the scan.  Two launches over 16 KB tiles (16 bytes per thread): count (kept bytes, restart markers,
// terminating marker per tile), then scatter at the tile's offset (sum over the page's earlier tiles).
struct JdTile {
    uint32_t counts; /* kept bytes | restart markers << 16 */
    uint32_t term;   /* position of the terminating marker in this tile (virtual stream coordinates), ~0 if none */
};

struct JdClass {
    uint32_t keep, rmask, term, wv[4];
};

__device__ __forceinline__ JdClass jd_classify(const uint8_t *src, uint32_t pos, uint32_t lead, uint32_t vlen) {
    JdClass c;
    c.keep = c.rmask = 0;
    c.term = 0xFFFFFFFFu;
    c.wv[0] = c.wv[1] = c.wv[2] = c.wv[3] = 0;
    if (pos >= vlen) return c;
    const uint4 v = *reinterpret_cast<const uint4 *>(src + pos);
    c.wv[0] = v.x, c.wv[1] = v.y, c.wv[2] = v.z, c.wv[3] = v.w;
    uint32_t prev = (pos > lead) ? src[pos - 1] : 0u;
    const uint32_t next = (pos + 16 < vlen) ? src[pos + 16] : 0xD9u;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t b = (c.wv[j >> 2] >> (8 * (j & 3))) & 255u;
        uint32_t nx = j < 15 ? (c.wv[(j + 1) >> 2] >> (8 * ((j + 1) & 3))) & 255u : next;
        if (pos + j + 1 >= vlen) nx = 0xD9u;
        const bool valid = pos + j >= lead && pos + j < vlen;
        if (!valid) b = 0u;
        bool k;
        if (b == 0xFFu) {
            k = nx == 0u;
            const bool isr = nx >= 0xD0u && nx <= 0xD7u;
            if (valid && isr) c.rmask |= 1u << j;
            if (valid && !k && !isr && nx != 0xFFu && c.term == 0xFFFFFFFFu) c.term = pos + j;
        } else {
            k = prev != 0xFFu;
        }
        if (valid && k) c.keep |= 1u << j;
        prev = b;
    }
    return c;
}

// tile -> page (tiles are page-major; pages are few)
__device__ __forceinline__ int jd_page_of_tile(const JdPage *pages, int n_pages, uint32_t tile) {
    int lo = 0, hi = n_pages - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pages[mid].h.tile_off <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// block-wide exclusive scan of a packed (low 16 | high 16) counter; *total = sum over the CTA
__device__ __forceinline__ uint32_t jd_block_excl_scan(uint32_t mine, uint32_t *s_warp, uint32_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = s_warp[lane], ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti += u;
        }
        s_warp[lane] = ti - t;
        if (lane == 31) s_warp[32] = ti;
    }
    __syncthreads();
    *total = s_warp[32];
    return incl - mine + s_warp[wid];
}

__global__ void __launch_bounds__(kUnstuffThreads) jpegd_unstuff_count_kernel(const uint8_t *__restrict__ blob,
                                                                              const JdPage *__restrict__ pages,
                                                                              int n_pages, JdTile *__restrict__ tiles) {
    const int page = jd_page_of_tile(pages, n_pages, blockIdx.x);
    const JdPageHdr &pg = pages[page].h;
    const uint32_t abase = pg.scan_off & ~15u, lead = pg.scan_off - abase, vlen = lead + pg.scan_len;
    const uint32_t pos = (blockIdx.x - pg.tile_off) * (kUnstuffThreads * 16) + threadIdx.x * 16;
    __shared__ uint32_t s_term, s_warp[33];
    if (threadIdx.x == 0) s_term = 0xFFFFFFFFu;
    __syncthreads();
    JdClass c = jd_classify(blob + abase, pos, lead, vlen);
    if (c.term != 0xFFFFFFFFu) atomicMin(&s_term, c.term);
    __syncthreads();
    const uint32_t tpos = s_term;
    if (tpos != 0xFFFFFFFFu && pos + 16 > tpos) {
        const uint32_t m = tpos <= pos ? 0u : (1u << (tpos - pos)) - 1u;
        c.keep &= m, c.rmask &= m;
    }
    uint32_t total;
    jd_block_excl_scan((uint32_t)__popc(c.keep) | ((uint32_t)__popc(c.rmask) << 16), s_warp, &total);
    if (threadIdx.x == 0) {
        tiles[blockIdx.x].counts = total;
        tiles[blockIdx.x].term = tpos;
    }
}

__global__ void __launch_bounds__(kUnstuffThreads) jpegd_unstuff_scatter_kernel(
    const uint8_t *__restrict__ blob, const JdPage *__restrict__ pages, int n_pages, const JdTile *__restrict__ tiles,
    uint32_t *__restrict__ stream, uint32_t *__restrict__ stream_bits, uint32_t *__restrict__ rst,
    uint32_t *__restrict__ n_rst) {
    const int page = jd_page_of_tile(pages, n_pages, blockIdx.x);
    const JdPageHdr &pg = pages[page].h;
    const uint32_t abase = pg.scan_off & ~15u, lead = pg.scan_off - abase, vlen = lead + pg.scan_len;
    const uint32_t my_tile = blockIdx.x - pg.tile_off;
    const uint32_t pos = my_tile * (kUnstuffThreads * 16) + threadIdx.x * 16;
    __shared__ uint32_t s_warp[33], s_base[2], s_dead;
    // offset of this tile = sum over the page's earlier tiles (a few hundred at most); dead if one of them ended the scan
    {
        uint32_t kept = 0, nr = 0, dead = 0;
        for (uint32_t t = threadIdx.x; t < my_tile; t += kUnstuffThreads) {
            const JdTile ti = tiles[pg.tile_off + t];
            kept += ti.counts & 0xFFFFu, nr += ti.counts >> 16;
            dead |= ti.term != 0xFFFFFFFFu;
        }
        if (threadIdx.x == 0) s_base[0] = s_base[1] = s_dead = 0;
        __syncthreads();
        if (kept | nr | dead) {
            atomicAdd(&s_base[0], kept);
            atomicAdd(&s_base[1], nr);
            if (dead) s_dead = 1;
        }
        __syncthreads();
    }
    if (s_dead) return;
    const uint32_t base_bytes = s_base[0], base_rst = s_base[1];
    const uint32_t tpos = tiles[blockIdx.x].term;
    JdClass c = jd_classify(blob + abase, pos, lead, vlen);
    if (tpos != 0xFFFFFFFFu && pos + 16 > tpos) {
        const uint32_t m = tpos <= pos ? 0u : (1u << (tpos - pos)) - 1u;
        c.keep &= m, c.rmask &= m;
    }
    uint32_t total;
    const uint32_t excl = jd_block_excl_scan((uint32_t)__popc(c.keep) | ((uint32_t)__popc(c.rmask) << 16), s_warp, &total);
    uint8_t *dst = reinterpret_cast<uint8_t *>(stream + pg.stream_word_off);
    uint32_t *rpos = rst + pg.rst_off;
    const uint32_t nrmax = pg.n_rst_max;
    uint32_t ob = base_bytes + (excl & 0xFFFFu), orst = base_rst + (excl >> 16);
    if (c.keep == 0xFFFFu && (ob & 3u) == 0u) {   // the common case: 16 kept bytes landing on a word boundary
        uint4 o;
        o.x = __byte_perm(c.wv[0], 0, 0x0123), o.y = __byte_perm(c.wv[1], 0, 0x0123);
        o.z = __byte_perm(c.wv[2], 0, 0x0123), o.w = __byte_perm(c.wv[3], 0, 0x0123);
        uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
        const uint32_t wi = ob >> 2;
        d32[jd_word_index(wi, kJdSwl)] = o.x, d32[jd_word_index(wi + 1, kJdSwl)] = o.y;
        d32[jd_word_index(wi + 2, kJdSwl)] = o.z, d32[jd_word_index(wi + 3, kJdSwl)] = o.w;
    } else if (c.keep | c.rmask) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (c.keep & (1u << j)) {
                dst[(jd_word_index(ob >> 2, kJdSwl) << 2) | ((ob & 3u) ^ 3u)] = (uint8_t)((c.wv[j >> 2] >> (8 * (j & 3))) & 255u);
                ob++;
            } else if (c.rmask & (1u << j)) {
                if (orst < nrmax) rpos[orst] = ob * 8u;
                orst++;
            }
        }
    }
    // the tile that ends the scan (or the page's last tile) publishes the totals
    if (threadIdx.x == 0 && (tpos != 0xFFFFFFFFu || my_tile + 1 == pg.n_tiles)) {
        const uint32_t bytes = base_bytes + (total & 0xFFFFu), nr = base_rst + (total >> 16);
        stream_bits[page] = bytes * 8u;
        n_rst[page] = nr < nrmax ? nr : nrmax;
    }
}

// ---- 2. entropy decode ---------------------------------------------------------------------------------------
// chain word: p << 32 | sk << 8 | status  (status 1 = provisional exit state of the CTA, 2 = final)
__device__ __forceinline__ unsigned long long jd_pack(JdState s, uint32_t status) {
    return ((unsigned long long)s.p << 32) | ((unsigned long long)(s.sk & 0xFFFFu) << 8) | status;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool jd_same(JdState a, JdState b) { return a.p == b.p && a.sk == b.sk; }

struct JdScanElem {
    int32_t has_abs, val;
};
__device__ __forceinline__ JdScanElem jd_combine(JdScanElem a, JdScanElem b) {
    JdScanElem r;
    r.has_abs = a.has_abs | b.has_abs;
    r.val = b.has_abs ? b.val : a.val + b.val;
    return r;
}

__global__ void __launch_bounds__(kJdChunk) jpegd_entropy_kernel(const JdPage *__restrict__ pages, int n_pages,
                                                                 const uint32_t *__restrict__ stream,
                                                                 const uint32_t *__restrict__ stream_bits,
                                                                 const uint32_t *__restrict__ rst,
                                                                 const uint32_t *__restrict__ n_rst,
                                                                 unsigned long long *chain, int32_t *chain_base,
                                                                 uint32_t *ticket, int16_t *__restrict__ coef,
                                                                 int16_t *__restrict__ dcdiff, int32_t nblk_total,
                                                                 uint32_t dc_stride, int32_t *__restrict__ status) {
    __shared__ JdHuff s_tab[JD_MAX_TABLES];
    __shared__ JdPageHdr s_hdr;
    __shared__ uint8_t s_zz[64];
    __shared__ JdState s_E[2][kJdChunk];
    __shared__ int32_t s_N[kJdChunk], s_B[kJdChunk];
    __shared__ uint8_t s_chg[2][kJdChunk];
    __shared__ JdScanElem s_scan[2][kJdChunk];
    __shared__ uint32_t s_ticket;
    __shared__ unsigned long long s_pred;
    __shared__ int32_t s_base_in;
    __shared__ int s_flags;
    const int t = threadIdx.x;
    if (t == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t cta = s_ticket;
    // page of this CTA: pages are few, CTAs page-major
    int page = 0;
    {
        int lo = 0, hi = n_pages - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (pages[mid].h.cta_off <= cta) lo = mid; else hi = mid - 1;
        }
        page = lo;
    }
    const JdPage &gp = pages[page];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&gp.h);
        uint32_t *d = reinterpret_cast<uint32_t *>(&s_hdr);
        for (int i = t; i < (int)(sizeof(JdPageHdr) / 4); i += kJdChunk) d[i] = src[i];
        const int ntab = gp.h.ntab;
        const uint32_t *ts = reinterpret_cast<const uint32_t *>(gp.tab);
        uint32_t *td = reinterpret_cast<uint32_t *>(s_tab);
        for (int i = t; i < ntab * (int)(sizeof(JdHuff) / 4); i += kJdChunk) td[i] = ts[i];
        if (t < 64) s_zz[t] = c_jd_zz[t];
    }
    __syncthreads();
    const uint32_t chunk = cta - s_hdr.cta_off;
    const uint32_t *words = stream + s_hdr.stream_word_off;
    const uint32_t total_bits = stream_bits[page];
    const uint32_t *prst = rst + s_hdr.rst_off;
    const int nr = s_hdr.restart_interval ? (int)n_rst[page] : 0;
    const uint32_t sub = chunk * kJdChunk + t;
    const uint32_t end_bit = (sub + 1) * kJdSubBits;
    int16_t *pcoef = coef + (size_t)page * nblk_total * 64;
    int16_t *pdc = dcdiff + (size_t)page * dc_stride;

    // round 0: guessed entry (exact for the first sub-sequence of the page)
    JdState used = {sub * (uint32_t)kJdSubBits, 0u};
    {
        JdSubResult r = jd_decode_sub<false>(s_hdr, s_tab, s_zz, words, kJdSwl, total_bits, prst, nr, used, end_bit, 0,
                                             nblk_total, nullptr, nullptr);
        s_E[0][t] = r.exit;
        s_N[t] = r.nblocks;
        s_B[t] = r.abs_base;
        s_chg[0][t] = 1;
    }
    __syncthreads();
    int cur = 0;
    // in-CTA propagation: re-decode from the left neighbour's exit state until nothing changes
    auto propagate = [&]() {
        for (;;) {
            JdState e = s_E[cur][t];
            int mine = 0;
            if (t > 0 && s_chg[cur][t - 1]) {
                JdState en = s_E[cur][t - 1];
                if (!jd_same(en, used)) {
                    used = en;
                    JdSubResult r = jd_decode_sub<false>(s_hdr, s_tab, s_zz, words, kJdSwl, total_bits, prst, nr, en, end_bit, 0,
                                                         nblk_total, nullptr, nullptr);
                    s_N[t] = r.nblocks;
                    s_B[t] = r.abs_base;
                    if (!jd_same(r.exit, e)) {
                        e = r.exit;
                        mine = 1;
                    }
                }
            }
            s_E[cur ^ 1][t] = e;
            s_chg[cur ^ 1][t] = (uint8_t)mine;
            cur ^= 1;
            if (!__syncthreads_or(mine)) break;
        }
    };
    // exclusive scan of (restart base, block count) over the chunk; s_scan[.][kJdChunk-1] inclusive = chunk total
    JdScanElem my_excl, chunk_total;
    auto scan_blocks = [&]() {
        JdScanElem v;
        v.has_abs = s_B[t] >= 0;
        v.val = s_B[t] >= 0 ? s_B[t] + s_N[t] : s_N[t];
        int b = 0;
        s_scan[0][t] = v;
        __syncthreads();
        for (int d = 1; d < kJdChunk; d <<= 1) {
            JdScanElem x = s_scan[b][t];
            if (t >= d) x = jd_combine(s_scan[b][t - d], x);
            s_scan[b ^ 1][t] = x;
            b ^= 1;
            __syncthreads();
        }
        chunk_total = s_scan[b][kJdChunk - 1];
        if (t > 0) my_excl = s_scan[b][t - 1];
        else my_excl.has_abs = 0, my_excl.val = 0;
        __syncthreads();
    };
    propagate();
    scan_blocks();

    int32_t base_in = 0;
    if (chunk == 0) {
        if (t == 0) {
            chain_base[cta] = chunk_total.val;  // has_abs or not, base_in = 0
            __threadfence();
            atomicExch(&chain[cta], jd_pack(s_E[cur][kJdChunk - 1], 2u));
        }
    } else {
        unsigned long long last_pub = 0, last_seen = 0;
        if (t == 0) {
            last_pub = jd_pack(s_E[cur][kJdChunk - 1], 1u);
            atomicExch(&chain[cta], last_pub);
        }
        for (;;) {
            if (t == 0) {
                unsigned long long w;
                do {
                    w = ld_volatile_u64(&chain[cta - 1]);
                } while ((w & 3u) == 0u || w == last_seen);
                last_seen = w;
                if ((w & 3u) == 2u) {
                    __threadfence();
                    s_base_in = *reinterpret_cast<volatile int32_t *>(&chain_base[cta - 1]);
                }
                s_pred = w;
            }
            __syncthreads();
            const unsigned long long w = s_pred;
            JdState x;
            x.p = (uint32_t)(w >> 32);
            x.sk = (uint32_t)(w >> 8) & 0xFFFFu;
            int flags = 0;  // bit 0: thread 0's exit state changed, bit 1: thread 0 decoded again (its block count may differ)
            if (t == 0 && !jd_same(x, used)) {
                used = x;
                JdSubResult r = jd_decode_sub<false>(s_hdr, s_tab, s_zz, words, kJdSwl, total_bits, prst, nr, x, end_bit, 0,
                                                     nblk_total, nullptr, nullptr);
                s_N[0] = r.nblocks;
                s_B[0] = r.abs_base;
                flags = 2;
                if (!jd_same(r.exit, s_E[cur][0])) {
                    s_E[cur][0] = r.exit;
                    flags = 3;
                }
                s_chg[cur][0] = (uint8_t)(flags & 1);
            }
            if (t == 0) s_flags = flags;
            __syncthreads();
            flags = s_flags;
            if (flags & 1) propagate();
            if (flags) scan_blocks();
            if ((w & 3u) == 2u) {
                base_in = s_base_in;
                if (t == 0) {
                    chain_base[cta] = chunk_total.has_abs ? chunk_total.val : base_in + chunk_total.val;
                    __threadfence();
                    atomicExch(&chain[cta], jd_pack(s_E[cur][kJdChunk - 1], 2u));
                }
                break;
            }
            if (t == 0) {
                unsigned long long pub = jd_pack(s_E[cur][kJdChunk - 1], 1u);
                if (pub != last_pub) {
                    last_pub = pub;
                    atomicExch(&chain[cta], pub);
                }
            }
        }
    }
    // write pass: every entry state and block index is final now
    {
        JdState e = t == 0 ? used : s_E[cur][t - 1];
        int32_t blk0 = my_excl.has_abs ? my_excl.val : base_in + my_excl.val;
        jd_decode_sub<true>(s_hdr, s_tab, s_zz, words, kJdSwl, total_bits, prst, nr, e, end_bit, blk0, nblk_total, pcoef, pdc);
    }
    if (t == 0 && chunk + 1 == s_hdr.n_cta) {
        int32_t total = chunk_total.has_abs ? chunk_total.val : base_in + chunk_total.val;
        status[page] = total == nblk_total ? 0 : 1;
    }
}

// ---- 3. DC prediction ----------------------------------------------------------------------------------------
// grid (ncomp, pages).  Component c's differences are contiguous (dc_off[c], MCU-major, cnt_c blocks per MCU);
// the running sum restarts at every MCU whose index is a multiple of the restart interval.  8 elements per
// thread and step (one 128-bit load), CTA-wide segmented scan, carry from step to step.
__global__ void __launch_bounds__(1024) jpegd_dc_kernel(const JdPage *__restrict__ pages, int16_t *__restrict__ dc,
                                                       uint32_t dc_stride) {
    const JdPageHdr &pg = pages[blockIdx.y].h;
    const int c = blockIdx.x;
    const uint32_t cnt = c == 0 ? (pg.ncomp == 1 ? 1u : (uint32_t)pg.hs * pg.vs) : 1u;
    const uint32_t n = (uint32_t)pg.mcux * pg.mcuy * cnt;
    const uint32_t seg = pg.restart_interval ? pg.restart_interval * cnt : 0xFFFFFFFFu;  // elements per restart interval
    int16_t *d = dc + (size_t)blockIdx.y * dc_stride + pg.dc_off[c];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    __shared__ int s_v[32], s_f[32], s_carry;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 8192) {
        const uint32_t e0 = base + t * 8;
        int x[8];
        int f = 0, v = 0;
        if (e0 < n) {
            const uint4 q = *reinterpret_cast<const uint4 *>(d + e0);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int val = (int16_t)((w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                if (e0 + j >= n) val = 0;
                if ((e0 + j) % seg == 0) v = 0, f = 1;
                v += val;
                x[j] = v;   // running sum inside the thread (relative to its start or the last reset)
            }
        }
        // segmented inclusive scan of (f, v) over the CTA
        int sv = v, sf = f;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            int pv = __shfl_up_sync(0xffffffffu, sv, k), pf = __shfl_up_sync(0xffffffffu, sf, k);
            if (lane >= k) {
                if (!sf) sv += pv;
                sf |= pf;
            }
        }
        if (lane == 31) s_v[wid] = sv, s_f[wid] = sf;
        __syncthreads();
        const int carry = s_carry;
        if (wid == 0) {
            int wv = s_v[lane], wf = s_f[lane];
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                int pv = __shfl_up_sync(0xffffffffu, wv, k), pf = __shfl_up_sync(0xffffffffu, wf, k);
                if (lane >= k) {
                    if (!wf) wv += pv;
                    wf |= pf;
                }
            }
            s_v[lane] = wv, s_f[lane] = wf;
        }
        __syncthreads();
        // exclusive prefix of this thread: lanes before it in the warp, warps before it, the carry of earlier steps
        int ev = __shfl_up_sync(0xffffffffu, sv, 1), ef = __shfl_up_sync(0xffffffffu, sf, 1);
        if (lane == 0) ev = 0, ef = 0;
        if (wid > 0 && !ef) ev += s_v[wid - 1], ef |= s_f[wid - 1];
        if (!ef) ev += carry;
        if (e0 < n) {
            uint32_t o[4] = {0, 0, 0, 0};
            bool past_reset = false;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if ((e0 + j) % seg == 0) past_reset = true;
                const int val = past_reset ? x[j] : x[j] + ev;
                o[j >> 1] |= ((uint32_t)val & 0xFFFFu) << (16 * (j & 1));
            }
            *reinterpret_cast<uint4 *>(d + e0) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        if (t == 1023) s_carry = s_f[31] ? s_v[31] : s_v[31] + carry;
        __syncthreads();
    }
}

// ---- 4. IDCT -------------------------------------------------------------------------------------------------
// One thread per 8x8 block, threads in plane raster order (neighbouring threads write neighbouring 8-byte runs).
__global__ void __launch_bounds__(128) jpegd_idct_kernel(const JdPage *__restrict__ pages,
                                                         const int16_t *__restrict__ coef,
                                                         const int16_t *__restrict__ dc, uint8_t *__restrict__ planes,
                                                         int32_t nblk_total, uint32_t dc_stride) {
    const int page = blockIdx.y;
    const JdPage &gp = pages[page];
    __shared__ uint16_t s_q[3][64];
    for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) s_q[i / 64][i % 64] = gp.qt[i / 64][i % 64];
    __syncthreads();
    const JdPageHdr &pg = gp.h;
    const int hs = pg.ncomp == 1 ? 1 : pg.hs, vs = pg.ncomp == 1 ? 1 : pg.vs;
    const int wb0 = pg.mcux * hs, hb0 = pg.mcuy * vs, wb1 = pg.mcux, hb1 = pg.mcuy;
    const int n0 = wb0 * hb0, n1 = wb1 * hb1;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nblk_total) return;
    int c, bx, by, wb;
    size_t plane_off;
    if (idx < n0) c = 0, wb = wb0, plane_off = 0;
    else if (idx < n0 + n1) c = 1, idx -= n0, wb = wb1, plane_off = (size_t)n0 * 64;
    else c = 2, idx -= n0 + n1, wb = wb1, plane_off = (size_t)(n0 + n1) * 64;
    by = idx / wb, bx = idx - by * wb;
    int blk;
    uint32_t dci;
    if (c == 0) {
        const int m = (by / vs) * pg.mcux + bx / hs, j = (by % vs) * hs + bx % hs;
        blk = m * pg.bpm + j;
        dci = pg.dc_off[0] + (uint32_t)(m * hs * vs + j);
    } else {
        const int m = by * pg.mcux + bx;
        blk = m * pg.bpm + hs * vs + c - 1;
        dci = pg.dc_off[c] + (uint32_t)m;
    }
    const int16_t dcv = dc[(size_t)page * dc_stride + dci];
    const uint4 *src = reinterpret_cast<const uint4 *>(coef + ((size_t)page * nblk_total + blk) * 64);
    const uint16_t *q = s_q[c];
    int32_t ws[64];
    {
        int32_t in[64];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            uint4 v = src[r];
            uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                int16_t cv = (int16_t)((wv[j >> 1] >> (16 * (j & 1))) & 0xFFFFu);
                if (r == 0 && j == 0) cv = dcv;
                in[r * 8 + j] = (int32_t)(int16_t)(cv * (int32_t)q[r * 8 + j]);
            }
        }
#pragma unroll
        for (int col = 0; col < 8; col++) {
            int32_t o[8];
            jd_idct_1d(in[col], in[8 + col], in[16 + col], in[24 + col], in[32 + col], in[40 + col], in[48 + col],
                       in[56 + col], 11, o);
#pragma unroll
            for (int r = 0; r < 8; r++) ws[r * 8 + col] = o[r];
        }
    }
    uint8_t *dst = planes + (size_t)page * nblk_total * 64 + plane_off + ((size_t)by * 8 * wb + bx) * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int32_t o[8];
        jd_idct_1d(ws[r * 8], ws[r * 8 + 1], ws[r * 8 + 2], ws[r * 8 + 3], ws[r * 8 + 4], ws[r * 8 + 5], ws[r * 8 + 6],
                   ws[r * 8 + 7], 18, o);
        uint2 out;
        out.x = pack4(jd_clamp_u8(o[0] + 128), jd_clamp_u8(o[1] + 128), jd_clamp_u8(o[2] + 128), jd_clamp_u8(o[3] + 128));
        out.y = pack4(jd_clamp_u8(o[4] + 128), jd_clamp_u8(o[5] + 128), jd_clamp_u8(o[6] + 128), jd_clamp_u8(o[7] + 128));
        *reinterpret_cast<uint2 *>(dst + (size_t)r * wb * 8) = out;
    }
}

// ---- 5. upsample + colour ------------------------------------------------------------------------------------
// One thread per 16 output pixels of a row.
__global__ void __launch_bounds__(256) jpegd_colour_kernel(const JdPage *__restrict__ pages,
                                                           const uint8_t *__restrict__ planes, uint8_t *__restrict__ out,
                                                           int n, int h, int w, int channels, int32_t nblk_total) {
    const int groups = (w + 15) / 16;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * h * groups) return;
    const int g = (int)(gid % groups);
    const int y = (int)((gid / groups) % h);
    const int page = (int)(gid / ((long long)groups * h));
    const JdPageHdr &pg = pages[page].h;
    const uint8_t *pl = planes + (size_t)page * nblk_total * 64;
    const int x0 = g * 16, nx = min(16, w - x0);
    if (channels == 1) {
        const int pitch = pg.mcux * 8;
        const uint8_t *s = pl + (size_t)y * pitch + x0;
        uint8_t *o = out + ((size_t)page * h + y) * w + x0;
        for (int i = 0; i < nx; i++) o[i] = s[i];
        return;
    }
    const int hs = pg.hs, vs = pg.vs;
    const int pitch0 = pg.mcux * 8 * hs, rows0 = pg.mcuy * 8 * vs, pitch1 = pg.mcux * 8, rows1 = pg.mcuy * 8;
    const uint8_t *py = pl + (size_t)y * pitch0 + x0;
    const uint8_t *pcb = pl + (size_t)pitch0 * rows0, *pcr = pcb + (size_t)pitch1 * rows1;
    const int mode = hs == 1 ? 0 : (vs == 1 ? 1 : 2);
    const int dw = (w + hs - 1) / hs, dh = (h + vs - 1) / vs;
    uint8_t *o = out + (((size_t)page * h + y) * w + x0) * 3;
    if (mode == 2 && nx == 16 && dw > 2 && (w & 15) == 0) {
        // fast path: 16 luma bytes, 8 (+2) chroma samples from two rows, three 128-bit stores
        const uint4 yv = *reinterpret_cast<const uint4 *>(py);
        const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
        int r0 = min(y >> 1, dh - 1);
        int r1 = (y & 1) ? (y >> 1) + 1 : (y >> 1) - 1;
        r1 = max(0, min(r1, dh - 1));
        const int i0 = x0 >> 1;
        int cs[2][10];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint8_t *p = c ? pcr : pcb;
            const uint8_t *a = p + (size_t)r0 * pitch1 + i0, *b = p + (size_t)r1 * pitch1 + i0;
            const uint2 av = *reinterpret_cast<const uint2 *>(a), bv = *reinterpret_cast<const uint2 *>(b);
            const uint32_t aw[2] = {av.x, av.y}, bw[2] = {bv.x, bv.y};
#pragma unroll
            for (int j = 0; j < 8; j++)
                cs[c][j + 1] = (int)((aw[j >> 2] >> (8 * (j & 3))) & 255u) * 3 + (int)((bw[j >> 2] >> (8 * (j & 3))) & 255u);
            cs[c][0] = i0 > 0 ? a[-1] * 3 + b[-1] : 0;
            cs[c][9] = i0 + 8 < dw ? a[8] * 3 + b[8] : 0;
        }
        uint32_t ob[12];
#pragma unroll
        for (int j = 0; j < 12; j++) ob[j] = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int i = j >> 1, gi = i0 + i;
            int cv[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int cur = cs[c][i + 1];
                if (j & 1) cv[c] = gi == dw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + cs[c][i + 2] + 7) >> 4;
                else cv[c] = gi == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + cs[c][i] + 8) >> 4;
            }
            uint32_t r, gg, b;
            jd_ycc_to_rgb((int)((yw[j >> 2] >> (8 * (j & 3))) & 255u), cv[0], cv[1], r, gg, b);
            ob[(3 * j) >> 2] |= r << (8 * ((3 * j) & 3));
            ob[(3 * j + 1) >> 2] |= gg << (8 * ((3 * j + 1) & 3));
            ob[(3 * j + 2) >> 2] |= b << (8 * ((3 * j + 2) & 3));
        }
        uint4 *o4 = reinterpret_cast<uint4 *>(o);
        stg_stream_u4(o4, make_uint4(ob[0], ob[1], ob[2], ob[3]));
        stg_stream_u4(o4 + 1, make_uint4(ob[4], ob[5], ob[6], ob[7]));
        stg_stream_u4(o4 + 2, make_uint4(ob[8], ob[9], ob[10], ob[11]));
        return;
    }
    for (int i = 0; i < nx; i++) {
        const int x = x0 + i;
        const int cb = jd_upsample_at(pcb, pitch1, dw, dh, mode, x, y), cr = jd_upsample_at(pcr, pitch1, dw, dh, mode, x, y);
        uint32_t r, gg, b;
        jd_ycc_to_rgb(py[i], cb, cr, r, gg, b);
        o[3 * i] = (uint8_t)r, o[3 * i + 1] = (uint8_t)gg, o[3 * i + 2] = (uint8_t)b;
    }
}

}  // namespace lumina

using namespace lumina;

// ---- C-ABI ---------------------------------------------------------------------------------------------------
LUMINA_API int lumina_jpeg_probe(const uint8_t *h_file, size_t len, lumina_jpeg_info *info) {
    if (!h_file || !info) return set_error(LUMINA_E_INVALID, "lumina_jpeg_probe: null argument");
    JdInfo ji;
    JdPage *pg = new JdPage;
    int rc = jd_parse(h_file, len, &ji, pg);
    delete pg;
    if (rc == -4) return set_error(LUMINA_E_UNSUPPORTED, "JPEG outside the device decoder's subset (host codec decodes it)");
    if (rc) return set_error(LUMINA_E_INVALID, "malformed JPEG");
    info->width = ji.width, info->height = ji.height, info->channels = ji.ncomp, info->hs = ji.hs, info->vs = ji.vs;
    return LUMINA_OK;
}

LUMINA_API size_t lumina_jpeg_decode_workspace_bytes(int n, int h, int w, int channels, int hs, int vs,
                                                     size_t total_file_bytes) {
    if (n <= 0 || h <= 0 || w <= 0) return 0;
    return jd_layout(n, h, w, channels, channels == 1 ? 1 : hs, channels == 1 ? 1 : vs, total_file_bytes).total;
}

LUMINA_API size_t lumina_jpeg_decode_stage_bytes(int n) { return n > 0 ? (size_t)n * sizeof(JdPage) : 0; }

LUMINA_API int lumina_jpeg_decode_batch(const uint8_t *h_blob, const int64_t *h_offsets, int n, int h, int w,
                                        int channels, uint8_t *d_out, int32_t *d_status, void *h_stage,
                                        void *d_workspace, size_t workspace_bytes, void *stream) {
    LUMINA_REQUIRE(h_blob && h_offsets && d_out && d_status && h_stage && d_workspace, "null argument");
    LUMINA_REQUIRE(n > 0 && h > 0 && w > 0 && (channels == 1 || channels == 3), "bad batch shape");
    cudaStream_t st = as_stream(stream);
    JdPage *hp = reinterpret_cast<JdPage *>(h_stage);
    const size_t total_bytes = (size_t)(h_offsets[n] - h_offsets[0]);
    LUMINA_REQUIRE(total_bytes < (1ull << 31), "batch blob too large (2 GiB limit)");
    int hs = 1, vs = 1;
    uint32_t word_off = 0, cta_off = 0, rst_off = 0, tile_off = 0;
    for (int i = 0; i < n; i++) {
        JdInfo ji;
        const size_t off = (size_t)(h_offsets[i] - h_offsets[0]), len = (size_t)(h_offsets[i + 1] - h_offsets[i]);
        int rc = jd_parse(h_blob + h_offsets[i], len, &ji, &hp[i]);
        if (rc == -4) return set_error(LUMINA_E_UNSUPPORTED, "page %d: JPEG outside the device decoder's subset", i);
        if (rc) return set_error(LUMINA_E_INVALID, "page %d: malformed JPEG", i);
        if (ji.width != w || ji.height != h || ji.ncomp != channels)
            return set_error(LUMINA_E_INVALID, "page %d is %dx%dx%d, the batch is %dx%dx%d", i, ji.width, ji.height,
                             ji.ncomp, w, h, channels);
        if (i == 0) hs = ji.hs, vs = ji.vs;
        if (ji.hs != hs || ji.vs != vs) return set_error(LUMINA_E_INVALID, "page %d: chroma sampling differs from page 0", i);
        JdPageHdr &ph = hp[i].h;
        ph.scan_off += (uint32_t)off;
        ph.stream_word_off = word_off;
        word_off += (ph.scan_len + 3) / 4 + 8;
        ph.cta_off = cta_off;
        ph.n_cta = (uint32_t)(((size_t)ph.scan_len * 8 + (size_t)kJdSubBits * kJdChunk - 1) / ((size_t)kJdSubBits * kJdChunk));
        if (ph.n_cta == 0) ph.n_cta = 1;
        cta_off += ph.n_cta;
        ph.rst_off = rst_off;
        const uint32_t nmcu = (uint32_t)ph.mcux * ph.mcuy;
        ph.n_rst_max = ph.restart_interval ? (nmcu - 1) / ph.restart_interval : 0;
        rst_off += ph.n_rst_max;
        ph.tile_off = tile_off;
        ph.n_tiles = ((ph.scan_off & 15u) + ph.scan_len + kUnstuffThreads * 16 - 1) / (kUnstuffThreads * 16);
        tile_off += ph.n_tiles;
    }
    const JdLayout L = jd_layout(n, h, w, channels, hs, vs, total_bytes);
    if (workspace_bytes < L.total)
        return set_error(LUMINA_E_NOMEM, "jpeg decode workspace too small: %zu < %zu", workspace_bytes, L.total);
    LUMINA_REQUIRE((size_t)word_off <= L.stream_words, "internal: stream region overflow");
    uint8_t *ws = reinterpret_cast<uint8_t *>(d_workspace);
    uint8_t *d_blob = ws + L.blob;
    JdPage *d_pages = reinterpret_cast<JdPage *>(ws + L.pages);
    uint32_t *d_stream = reinterpret_cast<uint32_t *>(ws + L.stream);
    uint32_t *d_bits = reinterpret_cast<uint32_t *>(ws + L.stream_bits);
    uint32_t *d_nrst = reinterpret_cast<uint32_t *>(ws + L.n_rst);
    uint32_t *d_rst = reinterpret_cast<uint32_t *>(ws + L.rst);
    JdTile *d_tiles = reinterpret_cast<JdTile *>(ws + L.tiles);
    unsigned long long *d_chain = reinterpret_cast<unsigned long long *>(ws + L.chain);
    int32_t *d_chain_base = reinterpret_cast<int32_t *>(ws + L.chain_base);
    uint32_t *d_ticket = reinterpret_cast<uint32_t *>(ws + L.ticket);
    int16_t *d_coef = reinterpret_cast<int16_t *>(ws + L.coef);
    int16_t *d_dc = reinterpret_cast<int16_t *>(ws + L.dc);
    uint8_t *d_planes = ws + L.planes;
    const int32_t nblk = (int32_t)L.nblk;

    LUMINA_CUDA_TRY(cudaMemcpyAsync(d_blob, h_blob + h_offsets[0], total_bytes, cudaMemcpyHostToDevice, st));
    LUMINA_CUDA_TRY(cudaMemcpyAsync(d_pages, hp, (size_t)n * sizeof(JdPage), cudaMemcpyHostToDevice, st));
    LUMINA_CUDA_TRY(cudaMemsetAsync(d_blob + total_bytes, 0, L.blob_bytes - total_bytes, st));
    // stream (zero tail bits), chain words, ticket, coefficients and DC differences start from zero
    LUMINA_CUDA_TRY(cudaMemsetAsync(ws + L.stream, 0, (size_t)word_off * 4, st));
    LUMINA_CUDA_TRY(cudaMemsetAsync(ws + L.chain, 0, L.coef - L.chain, st));
    LUMINA_CUDA_TRY(cudaMemsetAsync(ws + L.coef, 0, L.planes - L.coef, st));

    jpegd_unstuff_count_kernel<<<tile_off, kUnstuffThreads, 0, st>>>(d_blob, d_pages, n, d_tiles);
    LUMINA_KERNEL_CHECK("jpegd_unstuff_count_kernel");
    jpegd_unstuff_scatter_kernel<<<tile_off, kUnstuffThreads, 0, st>>>(d_blob, d_pages, n, d_tiles, d_stream, d_bits, d_rst, d_nrst);
    LUMINA_KERNEL_CHECK("jpegd_unstuff_scatter_kernel");
    jpegd_entropy_kernel<<<cta_off, kJdChunk, 0, st>>>(d_pages, n, d_stream, d_bits, d_rst, d_nrst, d_chain, d_chain_base,
                                                      d_ticket, d_coef, d_dc, nblk, (uint32_t)L.dc_stride, d_status);
    LUMINA_KERNEL_CHECK("jpegd_entropy_kernel");
    jpegd_dc_kernel<<<dim3(channels, n), 1024, 0, st>>>(d_pages, d_dc, (uint32_t)L.dc_stride);
    LUMINA_KERNEL_CHECK("jpegd_dc_kernel");
    jpegd_idct_kernel<<<dim3(div_up(nblk, 128), n), 128, 0, st>>>(d_pages, d_coef, d_dc, d_planes, nblk, (uint32_t)L.dc_stride);
    LUMINA_KERNEL_CHECK("jpegd_idct_kernel");
    const long long groups = (long long)n * h * ((w + 15) / 16);
    jpegd_colour_kernel<<<(unsigned)div_up(groups, 256), 256, 0, st>>>(d_pages, d_planes, d_out, n, h, w, channels, nblk);
    LUMINA_KERNEL_CHECK("jpegd_colour_kernel");
    return LUMINA_OK;
}
