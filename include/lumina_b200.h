/*
 * lumina_b200.h -- C-ABI of the B200-native page-image hot path.
 *
 * Drop-in boundary for GothiProCoder/OCR-System ("Lumina OCR").  Every entry
 * point replaces the native routine that one reference call site reaches
 * (reference paths are relative to backend/utils/image_preprocessing.py unless
 * stated).  The reference is Python, so the reference-side binding is a ctypes
 * stub; nothing here mentions torch.
 *
 * Conventions
 *   - All pointers named d_* are DEVICE pointers valid on the current CUDA
 *     device; h_* are HOST pointers.  The caller owns every buffer; kernels
 *     never allocate.  Scratch is passed in and sized by *_workspace_bytes().
 *   - Page batches are NHWC uint8, tightly packed: [n][h][w][c], c in {1,3}.
 *     Planes (gray / masks / edges) are [n][h][w].
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls
 *     only enqueue work unless documented as synchronising.
 *   - Return value: 0 on success, negative LUMINA_E_* on failure.  Nothing
 *     throws, nothing calls exit().  lumina_last_error_string() is
 *     thread-local.
 *   - Re-entrant per (device, stream); the only global state is an init-once
 *     constant table cache (bicubic weights, Gaussian taps).
 */
#ifndef LUMINA_B200_H
#define LUMINA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUMINA_OK 0
#define LUMINA_E_INVALID (-1) /* bad argument                     */
#define LUMINA_E_CUDA (-2)    /* CUDA runtime error (see string)  */
#define LUMINA_E_NOMEM (-3)   /* workspace too small              */
#define LUMINA_E_UNSUPPORTED (-4)

int lumina_abi_version(void);
const char *lumina_last_error_string(void);
/* number of kernels this library has launched since load (bench gpu_launches) */
uint64_t lumina_launch_count(void);

/* ---- a2  auto_orient :171-173 (PIL ImageOps.exif_transpose) ------------- */
/* orientation 1..8 (EXIF); dst is [n][w][h][c] for 5..8 else [n][h][w][c]. */
int lumina_exif_transpose_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                             int orientation, void *stream);

/* ---- a3  resize_if_needed :81-110 (PIL Image.resize LANCZOS) ------------ */
/* Target size with the reference's int() truncation (:97-105). */
void lumina_target_size(int width, int height, int max_dim, int *out_w, int *out_h);
typedef struct lumina_resize_plan lumina_resize_plan; /* device coefficient tables */
/* Synchronising (allocates + uploads the 22-bit fixed-point tables once). */
int lumina_resize_plan_create(int in_h, int in_w, int out_h, int out_w, lumina_resize_plan **plan);
void lumina_resize_plan_destroy(lumina_resize_plan *plan);
size_t lumina_resize_workspace_bytes(const lumina_resize_plan *plan, int n, int c);
/* Horizontal pass then vertical pass on a uint8 intermediate (Resample.c order). */
int lumina_resize_lanczos_u8(const lumina_resize_plan *plan, const uint8_t *d_src, uint8_t *d_dst, int n,
                             int c, void *d_workspace, size_t workspace_bytes, void *stream);

/* Image.resize on mode "P" / "1" objects: Pillow forces NEAREST there (Geometry.c ImagingScaleAffine).  The table
 * of source indices per output coordinate is built on the host exactly as Pillow accumulates it (double, repeated
 * addition); planes are [n][h][w] single-byte pixels (palette indices, or 0/255 for mode "1"). */
void lumina_nearest_table_host(int in_size, int out_size, int32_t *h_tab);
int lumina_resize_nearest_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int in_h, int in_w, int out_h, int out_w,
                             const int32_t *d_xtab, const int32_t *d_ytab, void *stream);

/* ---- a4  convert_to_grayscale :167-169 ; deskew's cv gray :394-396 ------ */
int lumina_rgb2gray_pil_u8(const uint8_t *d_rgb, uint8_t *d_gray, size_t npx, void *stream);
int lumina_rgb2gray_cv_u8(const uint8_t *d_rgb, uint8_t *d_gray, size_t npx, void *stream);

/* ---- a5  enhance_contrast :132-144 (ImageStat mean + Blend.c) ----------- */
/* mean = int(sum(L)/count + 0.5) == ImageStat's sum(i*hist[i])/count, so the
 * exact integer sum of PIL-L is reduced instead of a histogram.
 * d_sum_scratch: [n] uint64 (zeroed by the call); d_mean: [n] int32. */
int lumina_contrast_mean_u8(const uint8_t *d_src, int n, int h, int w, int c, uint64_t *d_sum_scratch,
                            int32_t *d_mean, void *stream);
int lumina_contrast_apply_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                             const int32_t *d_mean, float factor, void *stream);

/* ---- a6  enhance_sharpness :146-158 (Filter.c SMOOTH 3x3 + Blend.c) ----- */
int lumina_sharpness_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, float factor,
                        void *stream);
/* contrast-apply fused into the sharpness stencil (one read, one write):
 * dst = sharpness(contrast(src, mean, contrast_factor), sharp_factor). */
int lumina_contrast_sharpness_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                                 const int32_t *d_mean, float contrast_factor, float sharp_factor,
                                 void *stream);

/* ---- a7  denoise :160-165 (RankFilter.c median 3x3, edge replicate) ----- */
int lumina_median3_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, void *stream);

/* ---- a8  binarize :175-185 (convert L, point >thr -> mode "1") ---------- */
/* c==3: fused PIL gray.  dst [n][h][w] in {0,255}. */
int lumina_binarize_u8(const uint8_t *d_src, uint8_t *d_dst, size_t npx, int c, int threshold, void *stream);
/* Ingest helper (load_image / pdf_to_images, image_preprocessing.py:57-75, :248-295): Pillow keeps mode "RGB"
 * images as 4 bytes per pixel (R, G, B, pad); npx such pixels -> tightly packed R, G, B. */
int lumina_rgbx_to_rgb_u8(const uint8_t *d_src, uint8_t *d_dst, size_t npx, void *stream);

/* ---- north_star extras: Otsu and Sauvola binarisation (not called by the reference) ---------- */
/* Otsu: per-page threshold as cv2.threshold(gray, 0, 255, THRESH_BINARY | THRESH_OTSU) computes it (bit-equal);
 * d_thresh [n] int32 receives it, d_dst [n][h][w] {0,255} the mask (may be NULL); d_hist_scratch: n * 256 uint32. */
int lumina_otsu_u8(const uint8_t *d_gray, uint8_t *d_dst, int n, int h, int w, int32_t *d_thresh,
                   uint32_t *d_hist_scratch, void *stream);
/* Sauvola: dst = 255 where x > m * (1 + k * (s / R - 1)), m / s = mean / standard deviation of the
 * window x window neighbourhood clipped to the page (exact integer sums, float64 formula); window odd, 3..49. */
int lumina_sauvola_u8(const uint8_t *d_gray, uint8_t *d_dst, int n, int h, int w, int window, double k, double R,
                      void *stream);

/* ---- a9  adaptive_binarize :462-494 (cv2.adaptiveThreshold GAUSSIAN 11,C) */
/* c==3: fused PIL gray.  dst [n][h][w] in {0,255}.  OpenCV's plain float path (cv2.setUseOptimized(False)). */
int lumina_adaptive_gauss11_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, int cval,
                               void *stream);
/* The float Gaussian behind adaptiveThreshold is the one dispatch-dependent step of the reference's OpenCV calls:
 * LUMINA_CV_PLAIN = every product and sum rounded separately (setUseOptimized(False), CPUs without AVX2);
 * LUMINA_CV_AVX2  = OpenCV's default dispatch on x86 hosts with AVX2 + FMA3 -- what the reference runs unless told
 * otherwise: fused multiply-add in the vector loops, not in the scalar remainders (row pass fused for
 * x < w - w % 4, column pass for x < w - w % 8).  Each is bit-equal to cv2 in that mode. */
#define LUMINA_CV_PLAIN 0
#define LUMINA_CV_AVX2 1
int lumina_adaptive_gauss11_ex_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c, int cval,
                                  int cv_dispatch, void *stream);

/* ---- a10 deskew :372-460 ------------------------------------------------ */
/* (i)+(ii) cv gray (c==3) + cv2.Canny(low, high, aperture 3, L1): edges [n][h][w] {0,255} */
size_t lumina_canny_workspace_bytes(int n, int h, int w);
int lumina_canny_u8(const uint8_t *d_src, uint8_t *d_edges, int n, int h, int w, int c, int low, int high,
                    void *d_workspace, size_t workspace_bytes, void *stream);
/* (iii) cv2.HoughLinesP (progressive probabilistic Hough, cv::RNG stream).
 * d_lines [n][max_lines][4] int32 (x0,y0,x1,y1) in OpenCV's emission order;
 * d_nlines [n] int32 (may exceed max_lines: truncated output). */
size_t lumina_ppht_workspace_bytes(int n, int h, int w, double rho, double theta);
int lumina_ppht(const uint8_t *d_edges, int n, int h, int w, double rho, double theta, int threshold,
                int min_line_length, int max_line_gap, int32_t *d_lines, int32_t *d_nlines, int max_lines,
                void *d_workspace, size_t workspace_bytes, void *stream);
/* The same call in two halves for stream pipelines: prepare = point collection, edge bitmask, visiting order
 * (wide, short kernels; touches only the workspace); lines = the long cluster kernel on a prepared workspace.
 * lumina_ppht_prepare + lumina_ppht_lines with the same arguments == lumina_ppht. */
int lumina_ppht_prepare(const uint8_t *d_edges, int n, int h, int w, double rho, double theta, void *d_workspace,
                        size_t workspace_bytes, void *stream);
int lumina_ppht_lines(const uint8_t *d_edges, int n, int h, int w, double rho, double theta, int threshold,
                      int min_line_length, int max_line_gap, int32_t *d_lines, int32_t *d_nlines, int max_lines,
                      void *d_workspace, size_t workspace_bytes, void *stream);
/* the first `lines` (<= max_lines) segments of every page of d_lines [n][max_lines][4] -> h_lines [n][lines][4]
 * (pinned host memory), one strided copy on `stream` */
int lumina_copy_lines_to_host(const int32_t *d_lines, int n, int max_lines, int lines, int32_t *h_lines, void *stream);
/* (iv)+(v) host: per-line degrees(arctan2) folded to +-45, np.median, with glibc's atan2.  nlines==0 -> 0.0.
 * numpy's arctan2 is glibc's only where numpy does not dispatch to its bundled SIMD math (AVX-512 builds: the last
 * place differs for ~0.3 % of (dy, dx) pairs); a host that has numpy computes the per-line angles with it and calls
 * lumina_deskew_decide_angles_host (what the Python layer of this repo does). */
double lumina_median_angle_host(const int32_t *h_lines, int nlines);
/* (iv)-(vi) for a batch in one call: per page the reference's gating (:409-439) and, when the page
 * is to be rotated, getRotationMatrix2D((w//2, h//2), angle, 1).  h_lines [n][lines_stride][4]. */
void lumina_deskew_decide_host(const int32_t *h_lines, const int32_t *h_nlines, int n, int lines_stride,
                               int h, int w, double *h_angles, double *h_m6, uint8_t *h_apply);
/* (iv)-(vi) from per-line angles the caller computed with the reference's own expression
 * np.degrees(np.arctan2(y2 - y1, x2 - x1)) and folded to +-45 (:421-426): median (:428), gating, rotation matrix.
 * h_line_angles [n][angle_stride] f64, the first min(h_nlines[i], angle_stride) of a page are used. */
void lumina_deskew_decide_angles_host(const double *h_line_angles, const int32_t *h_nlines, int n, int angle_stride,
                                      int h, int w, double *h_angles, double *h_m6, uint8_t *h_apply);
/* cv::invertAffineTransform's arithmetic as cv::warpAffine applies it to a forward 2x3 matrix (double) */
void lumina_invert_affine_host(const double *h_m6, double *h_minv6);
/* cv2.getRotationMatrix2D (double, 2x3 row-major) */
void lumina_rotation_matrix_host(double cx, double cy, double angle_deg, double scale, double *h_m6);
/* (vii) cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE), same size.
 * h_m6: [n][6] forward matrices on the HOST (copied as kernel arguments in
 * chunks); pages with h_apply[i]==0 are copied unchanged (|angle|<0.5 etc). */
int lumina_warp_affine_cubic_u8(const uint8_t *d_src, uint8_t *d_dst, int n, int h, int w, int c,
                                const double *h_m6, const uint8_t *h_apply, void *stream);

/* ---- flag-gated alternative to a10's angle: projection-profile skew estimate (north_star) ---- */
/* NOT what the reference computes (:402-428 HoughLinesP + median): a tolerance-certified estimate of the same
 * angle (degrees, sign as the reference: positive = text descends to the right) from the Canny edge maps
 * [n][h][w]; the caller applies the reference's gates (:433-439) and rotation.  d_angles [n] f64. */
size_t lumina_skew_workspace_bytes_for(int n, int h, int w);
int lumina_skew_estimate_fast(const uint8_t *d_edges, int n, int h, int w, double *d_angles, void *d_workspace,
                              size_t workspace_bytes, void *stream);

/* ---- a15 [upstream PaddleOCR] DetResizeForTest + NormalizeImage + ToCHW - */
void lumina_det_target_size(int h, int w, int limit_side_len, int *out_h, int *out_w);
/* upstream's other limit types (resize_image_type0): limit_type 0 = "max" (the entry above), 1 = "min" (enlarge
 * when the shorter side is below the limit), 2 = "resize_long".  Host only. */
int lumina_det_target_size_ex(int h, int w, int limit_side_len, int limit_type, int *out_h, int *out_w);
/* src [n][h][w][3] u8 -> dst [n][3][oh][ow] f32; cv2.resize INTER_LINEAR
 * (11-bit fixed point) then (x*scale - mean[c]) / std[c]. */
int lumina_det_resize_normalize(const uint8_t *d_src, float *d_dst, int n, int h, int w, int oh, int ow,
                                const float *h_mean3, const float *h_std3, float scale, void *stream);

/* ---- a17 [upstream PaddleOCR] CTCLabelDecode ----------------------------- */
/* probs [n][t][c] f32.  d_idx/d_pos [n][t] int32 kept class ids / time steps
 * (-1 padded), d_len [n], d_conf [n] (mean of kept max-probs, 0 if none). */
size_t lumina_ctc_workspace_bytes(int n, int t);
int lumina_ctc_greedy(const float *d_probs, int n, int t, int c, int32_t *d_idx, int32_t *d_pos,
                      int32_t *d_len, float *d_conf, void *d_workspace, size_t workspace_bytes,
                      void *stream);

/* ---- a16 [upstream PaddleOCR] DBPostProcess ------------------------------ */
/* pred [n][h][w] f32 (channel 0 of the DB head).  Output per map: up to
 * max_candidates quads d_boxes [n][max_candidates][4][2] int32 (scaled to
 * (src_w, src_h) = h_src_hw[i][1], h_src_hw[i][0], clipped, rounded), d_scores
 * [n][max_candidates] f32, d_counts [n] int32; box order = OpenCV findContours
 * (RETR_LIST) order of the surviving candidates.  thresh is compared in float32
 * (numpy: pred > thresh), box_thresh / unclip_ratio in float64 like upstream. */
size_t lumina_db_workspace_bytes(int n, int h, int w, int max_candidates);
int lumina_db_postprocess(const float *d_pred, int n, int h, int w, float thresh, double box_thresh,
                          double unclip_ratio, int max_candidates, int min_size, const int32_t *h_src_hw,
                          int32_t *d_boxes, float *d_scores, int32_t *d_counts, void *d_workspace,
                          size_t workspace_bytes, void *stream);
/* flags bit 0: use_dilation=True (the mask is dilated by upstream's 2x2 kernel before the contours are taken;
 * scores still come from the undilated probabilities).
 * flags bit 1: score_mode="slow" (upstream box_score_slow: the mean of the probabilities over
 * cv2.fillPoly(contour), i.e. over the component, its holes and everything nested in them, instead of over the
 * first min-area quad). */
int lumina_db_postprocess_ex(const float *d_pred, int n, int h, int w, float thresh, double box_thresh,
                             double unclip_ratio, int max_candidates, int min_size, int flags,
                             const int32_t *h_src_hw, int32_t *d_boxes, float *d_scores, int32_t *d_counts,
                             void *d_workspace, size_t workspace_bytes, void *stream);
/* Stage outputs for parity tests: binary mask {0,1} + 8-connected labels
 * (label = min raster index of the component + 1, 0 = background).
 * workspace >= align256(n*h*w) + 4*n*h*w bytes. */
int lumina_db_mask_ccl(const float *d_pred, int n, int h, int w, float thresh, uint8_t *d_mask,
                       int32_t *d_labels, void *d_workspace, size_t workspace_bytes, void *stream);

/* ---- "next" row: reading order / line merge ---------------- */
/* backend/utils/ocr_postprocessor.py:101-182 (group_into_lines + sort_and_merge_lines) for a batch of
 * pages.  Page p owns boxes [d_offsets[p], d_offsets[p+1]) of d_boxes [total][4][2] f64 (x, y quads in
 * detector order; Python floats) and d_conf [total] f64.  Outputs, all indexed from d_offsets[p]:
 *   d_order[k]   = index inside the page of the k-th block in reading order,
 *   d_line_of[k] = 0-based line of that block (lines top to bottom, blocks left to right),
 *   d_nlines[p], d_line_conf[l] = mean confidence, d_line_y[l] = mean y_center of line l (f64).
 * Stable sorts and float64 sums exactly as the reference's Python (sum() as in CPython >= 3.12).
 * one_line != 0: every page is ONE line whose input order breaks x ties (sort_and_merge_lines on lines that
 * were grouped elsewhere); the ratio is then unused.  A negative or NaN ratio keeps the reference's meaning
 * (the tolerance test never holds: one line per block).  max_boxes_per_page >= the largest page, <= 4096. */
int lumina_reading_order(const double *d_boxes, const double *d_conf, const int32_t *d_offsets, int n_pages,
                         int max_boxes_per_page, double y_tolerance_ratio, int one_line, int32_t *d_order,
                         int32_t *d_line_of, int32_t *d_nlines, double *d_line_conf, double *d_line_y,
                         void *stream);

/* ---- "next" row: JPEG encode for compress_for_azure ----------- */
/* image_preprocessing.py:496-557 (compress_for_azure: image.save(format='JPEG', quality=q, optimize=True) in
 * a quality loop) and :331-347 (image_to_bytes).  Baseline JPEG of n equally sized RGB pages [n][h][w][3]:
 * YCbCr 4:2:0, libjpeg's integer (islow) DCT, Annex K quantisers scaled by `quality` (1..100), standard or
 * per-image optimal Huffman tables -- the same byte stream Pillow / libjpeg-turbo writes for
 * image.save(format='JPEG', quality=quality, optimize=bool(flags & 1)).
 * flags bit 1: reuse the DCT coefficients the previous call left in this workspace (same pages, next quality
 * of the loop).  h_sizes[i] (host) receives the file size of page i; the file itself is written to
 * h_out + i * out_stride when it fits in out_stride bytes (the reference's "size <= target" test; a page that
 * does not fit is simply not copied).  Synchronous with respect to `stream`. */
size_t lumina_jpeg_workspace_bytes(int n, int h, int w);
int lumina_jpeg_encode_rgb(const uint8_t *d_rgb, int n, int h, int w, int quality, int flags, uint8_t *h_out,
                           size_t out_stride, int64_t *h_sizes, void *d_workspace, size_t workspace_bytes,
                           void *stream);

/* ---- "next" row: page ingest -- baseline JPEG decode in HBM ---- */
/* image_preprocessing.py:57-75 (load_image / load_image_bytes: Image.open(...)) and services/ocr_service.py:494-496,
 * 716-718 (the engine is handed files / bytes).  Replaces Pillow's JpegDecode.c -> libjpeg-turbo (Huffman decode,
 * islow IDCT, fancy upsampling, YCbCr->RGB): the raster equals np.asarray(Image.open(file)) byte for byte, mode
 * "RGB" (channels 3) or "L" (channels 1) as Pillow reports it.  Subset: 8-bit baseline / extended-sequential
 * Huffman, one interleaved scan, grayscale or YCbCr with 4:4:4 / 4:2:2 / 4:2:0 sampling, restart intervals.
 * Anything else (progressive, arithmetic, CMYK, RGB-tagged, 4:4:0, multi-scan) returns LUMINA_E_UNSUPPORTED from
 * the probe / the batch call and stays on the host codec, which is what the reference uses for every file. */
typedef struct lumina_jpeg_info {
    int32_t width, height, channels; /* channels: 1 (mode "L") or 3 (mode "RGB") */
    int32_t hs, vs;                  /* luma sampling factors: 1x1 = 4:4:4, 2x1 = 4:2:2, 2x2 = 4:2:0 */
} lumina_jpeg_info;
/* Host-only header parse.  LUMINA_OK / LUMINA_E_UNSUPPORTED / LUMINA_E_INVALID (malformed). */
int lumina_jpeg_probe(const uint8_t *h_file, size_t len, lumina_jpeg_info *info);
/* Device scratch for n files of one geometry whose sizes add up to total_file_bytes; host staging for the
 * per-page tables (pinned memory lets the call stay asynchronous). */
size_t lumina_jpeg_decode_workspace_bytes(int n, int h, int w, int channels, int hs, int vs, size_t total_file_bytes);
size_t lumina_jpeg_decode_stage_bytes(int n);
/* Decodes n files (file i = h_blob[h_offsets[i] .. h_offsets[i+1]), all w x h x channels with the same sampling)
 * into d_out [n][h][w][channels].  The call parses the headers on the host, enqueues the upload of the files and
 * tables and five kernels on `stream`, and returns; h_blob and h_stage must stay untouched until the stream has
 * passed this call.  d_status [n] int32 (device): 0 = page decoded, 1 = the entropy-coded data did not contain
 * exactly the blocks the header announces (truncated / corrupt file: decode that page on the host). */
int lumina_jpeg_decode_batch(const uint8_t *h_blob, const int64_t *h_offsets, int n, int h, int w, int channels,
                             uint8_t *d_out, int32_t *d_status, void *h_stage, void *d_workspace,
                             size_t workspace_bytes, void *stream);

/* ---- synthetic workloads (bench/test inputs generated in HBM) ------------ */
/* A4-like text page, seeded by page index; identical bytes to the host
 * generator in include/lumina_synth.h compiled for the CPU. */
int lumina_synth_pages_u8(uint8_t *d_dst, int n, int h, int w, uint64_t seed0, void *stream);
/* DB probability maps [n][h][w] f32 (map m seeded by seed0 + m: ~one text box per 60x30 cell, soft borders,
 * holes) and CTC posteriors [n][t][c] f32 (crop i seeded by crop0 + i: winner 0.9 per step, planted blanks,
 * repeats and exact ties) -- BASELINE.json configs[2..4]; same floats as the host build of lumina_synth.h. */
int lumina_synth_prob_maps_f32(float *d_dst, int n, int h, int w, uint64_t seed0, void *stream);
int lumina_synth_ctc_f32(float *d_dst, int n, int t, int c, uint64_t crop0, uint32_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LUMINA_B200_H */
