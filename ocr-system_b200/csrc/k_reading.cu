// k_reading.cu -- reading order / line grouping of detected text boxes, batched over pages.
//
// Reference: backend/utils/ocr_postprocessor.py
//   group_into_lines      :101-143  stable sort by y_center, tolerance = mean(height) * ratio, sequential
//                                   grouping against the running mean y of the open line
//   sort_and_merge_lines  :146-182  stable sort of every line by x_left, mean confidence / mean y per line,
//                                   stable sort of the lines by mean y
// One CTA per page.  The two sorts are bitonic sorts in shared memory on (major, double value, tie index)
// keys -- the tie index makes them reproduce Python's stable sorts; the grouping itself is a sequential
// recurrence (every decision depends on the running mean of the open line) and is walked by one thread, in
// float64, in the reference's order of operations.  Lines are contiguous runs of the y-sorted list, so their
// mean y is non-decreasing and the reference's final sort of the lines is the identity (ties keep creation
// order); the kernel therefore emits the lines in creation order.
// Sums follow CPython's sum() for floats (bltinmodule.c, 3.12+): left to right with Neumaier compensation,
// the correction added once at the end.  With float32-valued coordinates (what a detector produces) every
// partial sum is exact anyway, so older interpreters' plain sums give the same bits.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace lumina {

constexpr int RO_THREADS = 256;
constexpr int RO_MAX_BOXES = 4096;

struct RoKey {
    double val;
    int major;
    int idx;
};

// CPython >= 3.12 builtin sum() over floats: running sum f and compensation c; value = f + c (when c != 0)
struct PySum {
    double f = 0.0, c = 0.0;
    __device__ __forceinline__ void add(double x) {
        const double t = f + x;
        if (fabs(f) >= fabs(x)) c += (f - t) + x;
        else c += (x - t) + f;
        f = t;
    }
    __device__ __forceinline__ double value() const { return (c != 0.0 && isfinite(c)) ? f + c : f; }
};

__device__ __forceinline__ bool ro_less(const RoKey &a, const RoKey &b) {
    if (a.major != b.major) return a.major < b.major;
    if (a.val != b.val) return a.val < b.val;
    return a.idx < b.idx;
}

// in-place bitonic sort of keys[0..npad) (npad a power of two), all threads of the CTA
__device__ void ro_bitonic(RoKey *keys, int npad) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += RO_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const RoKey a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if (up ? ro_less(b, a) : ro_less(a, b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(RO_THREADS) reading_order_kernel(const double *__restrict__ boxes, const double *__restrict__ conf,
                                                                   const int32_t *__restrict__ offsets, double ratio, int one_line_mode,
                                                                   int32_t *__restrict__ order,
                                                                   int32_t *__restrict__ line_of, int32_t *__restrict__ nlines,
                                                                   double *__restrict__ line_conf, double *__restrict__ line_y, int npad) {
    extern __shared__ __align__(16) unsigned char ro_smem[];
    const int page = blockIdx.x, tid = threadIdx.x;
    const int o0 = offsets[page], n = offsets[page + 1] - o0;
    RoKey *keys = reinterpret_cast<RoKey *>(ro_smem);                 // [npad]
    double *yc = reinterpret_cast<double *>(keys + npad);              // [npad] y_center by original index
    double *xl = yc + npad;                                            // [npad] x_left by original index
    int *lstart = reinterpret_cast<int *>(xl + npad);                  // [npad + 1] first sorted position of every line
    __shared__ int s_nlines;
    if (n <= 0) {
        if (tid == 0) nlines[page] = 0;
        return;
    }
    const double *bx = boxes + (size_t)o0 * 8;
    const bool one_line = one_line_mode != 0;  // pre-grouped input: the page is one line, input order breaks x ties
    // (a negative or NaN ratio is NOT that mode: like the reference, abs(dy) <= tol is then never true and
    //  every block opens its own line)
    // TextBlock.y_center / x_left (ocr_postprocessor.py:26-35)
    for (int i = tid; i < npad; i += RO_THREADS) {
        RoKey k;
        if (i < n) {
            const double *q = bx + (size_t)i * 8;
            yc[i] = (q[1] + q[5]) / 2;
            double m = q[0];   // Python min(): keeps the first of equal values, a later one only if strictly smaller
            if (q[2] < m) m = q[2];
            if (q[4] < m) m = q[4];
            if (q[6] < m) m = q[6];
            xl[i] = m;
            k.val = one_line ? 0.0 : yc[i]; k.major = 0; k.idx = i;
        } else {
            k.val = 0.0; k.major = INT_MAX; k.idx = i;
        }
        keys[i] = k;
    }
    __syncthreads();
    ro_bitonic(keys, npad);  // sorted(blocks, key=y_center), stable
    // ---- sequential grouping (ocr_postprocessor.py:118-141) ----
    if (tid == 0) {
        PySum hs;
        for (int k = 0; k < n; k++) {
            const double *q = bx + (size_t)keys[k].idx * 8;
            hs.add(fabs(q[5] - q[1]));   // TextBlock.height
        }
        const double tol = one_line ? (double)INFINITY : (hs.value() / (double)n) * ratio;
        int nl = 0, len = 1;
        PySum cur;
        cur.add(yc[keys[0].idx]);
        double cur_y = yc[keys[0].idx];
        lstart[0] = 0;
        keys[0].major = 0;
        for (int k = 1; k < n; k++) {
            const double y = yc[keys[k].idx];
            if (fabs(y - cur_y) <= tol) {
                cur.add(y); len++;
                cur_y = cur.value() / (double)len;
            } else {
                nl++;
                lstart[nl] = k;
                cur = PySum(); cur.add(y); len = 1; cur_y = y;
            }
            keys[k].major = nl;
        }
        nl++;
        lstart[nl] = n;
        s_nlines = nl;
    }
    __syncthreads();
    const int nl = s_nlines;
    // ---- sorted(line, key=x_left), stable with respect to the y-sorted order ----
    for (int k = tid; k < n; k += RO_THREADS) {
        RoKey q = keys[k];
        q.val = xl[q.idx];
        q.idx = (q.idx & 0xffff) | (k << 16);   // tie: position in the y-sorted list; original index rides along
        keys[k] = q;
    }
    __syncthreads();
    ro_bitonic(keys, npad);
    for (int k = tid; k < n; k += RO_THREADS) {
        order[o0 + k] = keys[k].idx & 0xffff;
        line_of[o0 + k] = keys[k].major;
    }
    // ---- per line: mean confidence and mean y in x-sorted order (ocr_postprocessor.py:170-171) ----
    for (int l = tid; l < nl; l += RO_THREADS) {
        PySum cs, ys;
        const int a = lstart[l], b = lstart[l + 1];
        for (int k = a; k < b; k++) {
            const int oi = keys[k].idx & 0xffff;
            cs.add(conf[o0 + oi]);
            ys.add(yc[oi]);
        }
        line_conf[o0 + l] = cs.value() / (double)(b - a);
        line_y[o0 + l] = ys.value() / (double)(b - a);
    }
    // ---- merged.sort(key=lambda m: m.y_position), stable (ocr_postprocessor.py:181) ----
    // Lines are runs of the y-sorted list, so their means ascend -- except in the last place: with a zero tolerance two
    // lines may be one ulp apart and the rounded mean of the first may come out above the second's.  Found by the
    // randomised sweep (tools/sweep_paddle_vs_oracle.py, 1000 jittered boxes, ratio 0).  Rare: one thread repairs it.
    __syncthreads();
    if (tid == 0) {
        bool sorted = true;
        for (int l = 1; l < nl && sorted; l++) sorted = !(line_y[o0 + l] < line_y[o0 + l - 1]);
        if (!sorted) {
            int *perm = lstart + npad + 1;   // [npad]
            for (int l = 0; l < nl; l++) { perm[l] = l; yc[l] = line_y[o0 + l]; xl[l] = line_conf[o0 + l]; }   // yc / xl are free now
            for (int i = 1; i < nl; i++) {   // stable insertion sort: a line moves only past strictly larger means
                const int pl = perm[i];
                const double v = yc[pl];
                int j = i - 1;
                while (j >= 0 && yc[perm[j]] > v) { perm[j + 1] = perm[j]; j--; }
                perm[j + 1] = pl;
            }
            int pos = 0;
            for (int j = 0; j < nl; j++) {
                const int l = perm[j];
                for (int k = lstart[l]; k < lstart[l + 1]; k++, pos++) {
                    order[o0 + pos] = keys[k].idx & 0xffff;
                    line_of[o0 + pos] = j;
                }
                line_conf[o0 + j] = xl[l];
                line_y[o0 + j] = yc[l];
            }
        }
        nlines[page] = nl;
    }
}

}  // namespace lumina

using namespace lumina;

// ocr_postprocessor.py:101-182 for a batch of pages (see include/lumina_b200.h)
LUMINA_API int lumina_reading_order(const double *d_boxes, const double *d_conf, const int32_t *d_offsets, int n_pages,
                                    int max_boxes_per_page, double y_tolerance_ratio, int one_line, int32_t *d_order, int32_t *d_line_of,
                                    int32_t *d_nlines, double *d_line_conf, double *d_line_y, void *stream) {
    LUMINA_REQUIRE(d_boxes && d_conf && d_offsets && d_order && d_line_of && d_nlines && d_line_conf && d_line_y, "null pointer");
    LUMINA_REQUIRE(n_pages > 0, "empty batch");
    LUMINA_REQUIRE(max_boxes_per_page >= 0 && max_boxes_per_page <= RO_MAX_BOXES, "more than 4096 boxes on a page");
    int npad = 32;
    while (npad < max_boxes_per_page) npad <<= 1;
    const size_t smem = (size_t)npad * (sizeof(RoKey) + 16) + (size_t)(npad + 1) * 4 + (size_t)npad * 4;   // keys, yc, xl, lstart, perm
    if (smem > 48 * 1024)
        LUMINA_CUDA_TRY(cudaFuncSetAttribute(reading_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    reading_order_kernel<<<n_pages, RO_THREADS, smem, as_stream(stream)>>>(d_boxes, d_conf, d_offsets, y_tolerance_ratio, one_line, d_order,
                                                                          d_line_of, d_nlines, d_line_conf, d_line_y, npad);
    LUMINA_KERNEL_CHECK("reading_order_kernel");
    return LUMINA_OK;
}
