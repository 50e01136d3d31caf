// Host build of ocr-system_b200/csrc/jpegd_core.h for the CPU unit tests (tests/test_jpegd_core.py): the same
// parser, sub-sequence decoder, IDCT, upsampling and colour conversion the CUDA kernels run, driven by a
// sequential simulation of the kernels' schedule (one loop iteration = one thread), checked against Pillow
// without a GPU.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../ocr-system_b200/csrc/jpegd_core.h"

extern "C" {
// stats[0] = sync rounds used, stats[1] = sub-sequences, stats[2] = re-decodes in rounds >= 1,
// stats[3] = sub-sequences whose exit state was verified in the write pass (must equal stats[1])
int jdh_decode(const uint8_t *f, size_t len, int sub_bits, uint8_t *out, int *whc, long long *stats) {
    JdInfo info;
    JdPage *pg = new JdPage;
    int rc = jd_parse(f, len, &info, pg);
    if (rc) { delete pg; return rc; }
    whc[0] = info.width; whc[1] = info.height; whc[2] = info.ncomp;
    if (!out) { delete pg; return 0; }
    // ---- unstuff (jpegd_unstuff_kernel) ----
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> rst;
    const uint8_t *s = f + pg->h.scan_off;
    for (size_t i = 0; i < pg->h.scan_len; i++) {
        uint8_t c = s[i];
        if (c != 0xFF) { bytes.push_back(c); continue; }
        uint8_t nx = i + 1 < pg->h.scan_len ? s[i + 1] : 0xD9;
        if (nx == 0x00) { bytes.push_back(0xFF); i++; }
        else if (nx >= 0xD0 && nx <= 0xD7) { rst.push_back((uint32_t)bytes.size() * 8); i++; }
        else if (nx == 0xFF) { }
        else break;
    }
    uint32_t total_bits = (uint32_t)bytes.size() * 8;
    uint32_t swl = 0;
    while ((32u << swl) < (uint32_t)sub_bits) swl++;
    const size_t chunk_words = (size_t)1 << (swl + JD_CHUNK_LOG2);
    std::vector<uint32_t> words(((bytes.size() / 4 + 4) / chunk_words + 1) * chunk_words, 0);
    for (size_t i = 0; i < bytes.size(); i++) words[jd_word_index((uint32_t)(i >> 2), swl)] |= (uint32_t)bytes[i] << (24 - 8 * (i & 3));
    int n_rst = (int)rst.size();
    if (!pg->h.restart_interval) n_rst = 0;
    // ---- sync rounds ----
    uint32_t S = (uint32_t)sub_bits;
    int n_sub = (int)((total_bits + S - 1) / S);
    if (n_sub == 0) n_sub = 1;
    std::vector<JdState> E(n_sub), E2(n_sub);
    std::vector<int32_t> N(n_sub), B(n_sub);
    std::vector<uint8_t> chg(n_sub, 1), chg2(n_sub, 0);
    int nblk_total = pg->h.mcux * pg->h.mcuy * pg->h.bpm;
    for (int i = 0; i < n_sub; i++) {
        JdState e = {(uint32_t)i * S, 0};
        JdSubResult r = jd_decode_sub<false>(pg->h, pg->tab, kJdZigzag, words.data(), swl, total_bits, rst.data(), n_rst, e,
                                             (i + 1) * S, 0, nblk_total, nullptr, nullptr);
        E[i] = r.exit; N[i] = r.nblocks; B[i] = r.abs_base;
    }
    long long rounds = 0, redec = 0;
    for (;;) {
        bool any = false;
        rounds++;
        for (int i = 0; i < n_sub; i++) {
            E2[i] = E[i]; chg2[i] = 0;
            if (i == 0 || !chg[i - 1]) continue;
            JdState e = E[i - 1];
            if (rounds == 1 && e.p == (uint32_t)i * S && e.sk == 0) continue;
            redec++;
            JdSubResult r = jd_decode_sub<false>(pg->h, pg->tab, kJdZigzag, words.data(), swl, total_bits, rst.data(), n_rst, e,
                                                 (i + 1) * S, 0, nblk_total, nullptr, nullptr);
            N[i] = r.nblocks; B[i] = r.abs_base;
            if (r.exit.p != E[i].p || r.exit.sk != E[i].sk) { E2[i] = r.exit; chg2[i] = 1; any = true; }
        }
        E.swap(E2); chg.swap(chg2);
        if (!any) break;
        if (rounds > n_sub + 2) { delete pg; return -9; }
    }
    stats[0] = rounds; stats[1] = n_sub; stats[2] = redec;
    // ---- block index scan ----
    std::vector<int32_t> base(n_sub);
    int32_t run = 0;
    for (int i = 0; i < n_sub; i++) { base[i] = run; run = B[i] >= 0 ? B[i] + N[i] : run + N[i]; }
    // ---- write pass ----
    std::vector<int16_t> coef((size_t)nblk_total * 64, 0), dc(pg->h.dc_off[pg->h.ncomp], 0);
    long long verified = 0;
    for (int i = 0; i < n_sub; i++) {
        JdState e = i ? E[i - 1] : JdState{0, 0};
        JdSubResult r = jd_decode_sub<true>(pg->h, pg->tab, kJdZigzag, words.data(), swl, total_bits, rst.data(), n_rst, e,
                                            (i + 1) * S, base[i], nblk_total, coef.data(), dc.data());
        verified += (r.exit.p == E[i].p && r.exit.sk == E[i].sk);
    }
    stats[3] = verified;
    // ---- DC prediction (jpegd_dc_kernel): prefix sum per component, reset every restart interval ----
    {
        int pred[3] = {0, 0, 0};
        int nmcu = pg->h.mcux * pg->h.mcuy;
        for (int m = 0; m < nmcu; m++) {
            if (pg->h.restart_interval && m % pg->h.restart_interval == 0) pred[0] = pred[1] = pred[2] = 0;
            for (int sl = 0; sl < pg->h.bpm; sl++) {
                int c = pg->h.slot_comp[sl];
                int16_t &d = dc[pg->h.slot_dcbase[sl] + (size_t)m * pg->h.slot_cnt[sl]];
                pred[c] += d;
                d = (int16_t)pred[c];
            }
        }
    }
    // ---- IDCT into planes ----
    int W = info.width, H = info.height, nc = info.ncomp, hs = info.hs, vs = info.vs;
    int pw[3], ph[3];
    std::vector<uint8_t> plane[3];
    for (int c = 0; c < nc; c++) {
        pw[c] = pg->h.mcux * 8 * (c == 0 ? hs : 1);
        ph[c] = pg->h.mcuy * 8 * (c == 0 ? vs : 1);
        plane[c].resize((size_t)pw[c] * ph[c]);
    }
    int nmcu = pg->h.mcux * pg->h.mcuy;
    for (int m = 0; m < nmcu; m++)
        for (int sl = 0; sl < pg->h.bpm; sl++) {
            int c = pg->h.slot_comp[sl], blk = m * pg->h.bpm + sl;
            int sub = c == 0 ? sl : 0;
            int bx = (m % pg->h.mcux) * (c == 0 ? hs : 1) + (c == 0 ? sub % hs : 0);
            int by = (m / pg->h.mcux) * (c == 0 ? vs : 1) + (c == 0 ? sub / hs : 0);
            int32_t ws[64], in[64];
            for (int i = 0; i < 64; i++)
                in[i] = (int16_t)((i ? coef[(size_t)blk * 64 + i] : dc[pg->h.slot_dcbase[sl] + (size_t)m * pg->h.slot_cnt[sl]]) * pg->qt[c][i]);
            for (int col = 0; col < 8; col++) {
                int32_t o[8];
                jd_idct_1d(in[col], in[8 + col], in[16 + col], in[24 + col], in[32 + col], in[40 + col], in[48 + col],
                           in[56 + col], 11, o);
                for (int r = 0; r < 8; r++) ws[r * 8 + col] = o[r];
            }
            for (int r = 0; r < 8; r++) {
                int32_t o[8];
                const int32_t *w = ws + r * 8;
                jd_idct_1d(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], 18, o);
                uint8_t *dst = plane[c].data() + (size_t)(by * 8 + r) * pw[c] + bx * 8;
                for (int x = 0; x < 8; x++) dst[x] = (uint8_t)jd_clamp_u8(o[x] + 128);
            }
        }
    // ---- upsample + colour ----
    if (nc == 1) {
        for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, plane[0].data() + (size_t)y * pw[0], W);
    } else {
        int mode = hs == 1 ? 0 : (vs == 1 ? 1 : 2);
        int dw = (W + hs - 1) / hs, dh = (H + vs - 1) / vs;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int Y = plane[0][(size_t)y * pw[0] + x];
                int cb = jd_upsample_at(plane[1].data(), pw[1], dw, dh, mode, x, y);
                int cr = jd_upsample_at(plane[2].data(), pw[2], dw, dh, mode, x, y);
                uint32_t r, g, b;
                jd_ycc_to_rgb(Y, cb, cr, r, g, b);
                uint8_t *o = out + ((size_t)y * W + x) * 3;
                o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b;
            }
    }
    delete pg;
    return 0;
}
}
