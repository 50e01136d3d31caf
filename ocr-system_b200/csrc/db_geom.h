// db_geom.h -- per-candidate geometry of DBPostProcess, host+device.
//
// [upstream PaddleOCR ppocr/postprocess/db_postprocess.py; not in the reference tree]
// Everything a candidate needs after labelling, as plain functions that compile for
// the GPU (k_dbpost.cu) and for the host (tests/geom_host.cpp checks them against
// cv2.minAreaRect / boxPoints / fillPoly+mean on the CPU):
//   dbg_hull            convex hull (monotone chain) of integer pixel centres
//   dbg_min_area_rect   rotating calipers in float32 (as cv::minAreaRect evaluates it)
//   dbg_box_points      cv::RotatedRect::points
//   dbg_mini_box        get_mini_boxes ordering + short side
//   dbg_row_cover       cv2.fillPoly coverage of a row: scan-line interior span
//                       (16.16 fixed point edges) + 8-connected Bresenham boundary
//                       runs in closed form; box_score_fast = mean over that cover
//   dbg_clipper_offset  Clipper 6.4.2 ClipperOffset(jtRound, etClosedPolygon)
//   dbg_scale_box       round(x / W * dest_w) clipped, int32
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define DBG_HD __host__ __device__ __forceinline__
// the large routines stay out of line on the device: the candidate kernel runs them on one lane of many warps that
// are all at different places, and its stalls are instruction-cache misses (ncu: no_instruction) -- one copy each
#define DBG_HD_BIG __host__ __device__ __noinline__
#else
#define DBG_HD static inline
#define DBG_HD_BIG static inline
#endif

#ifndef DBG_PI
#define DBG_PI 3.1415926535897932384626433832795
#endif

struct DbgPt { int x, y; };
struct DbgPtF { float x, y; };
struct DbgRect { float cx, cy, w, h, angle; };

DBG_HD long long dbg_cross(const DbgPt &o, const DbgPt &a, const DbgPt &b) {
    return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// in-place heap sort by (y, x)
DBG_HD bool dbg_less(const DbgPt &a, const DbgPt &b) { return a.y < b.y || (a.y == b.y && a.x < b.x); }
DBG_HD_BIG void dbg_sort(DbgPt *p, int n) {
    for (int start = n / 2 - 1; start >= 0; start--) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && dbg_less(p[child], p[child + 1])) child++;
            if (!dbg_less(p[root], p[child])) break;
            DbgPt t = p[root]; p[root] = p[child]; p[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; end--) {
        DbgPt t = p[0]; p[0] = p[end]; p[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && dbg_less(p[child], p[child + 1])) child++;
            if (!dbg_less(p[root], p[child])) break;
            DbgPt u = p[root]; p[root] = p[child]; p[child] = u;
            root = child;
        }
    }
}

// Monotone chain on points already sorted by (y, x) (duplicates allowed).  `hull` must hold n+1
// points.  Collinear points are dropped.  Returns the hull size (1, 2 or >= 3).
DBG_HD_BIG int dbg_hull_sorted(const DbgPt *p, int n, DbgPt *hull) {
    if (n <= 0) return 0;
    int k = 0;
    for (int i = 0; i < n; i++) {
        if (i > 0 && p[i].x == p[i - 1].x && p[i].y == p[i - 1].y) continue;
        while (k >= 2 && dbg_cross(hull[k - 2], hull[k - 1], p[i]) <= 0) k--;
        hull[k++] = p[i];
    }
    if (k == 1) return 1;
    const int lower = k + 1;
    for (int i = n - 2; i >= 0; i--) {
        if (p[i].x == p[i + 1].x && p[i].y == p[i + 1].y) continue;
        while (k >= lower && dbg_cross(hull[k - 2], hull[k - 1], p[i]) <= 0) k--;
        hull[k++] = p[i];
    }
    return k - 1;  // last point == first point
}

// cv::minAreaRect(contour) runs its calipers over cv::convexHull(contour), whose start vertex follows
// the contour's point order (it shifts the hull so the original indices are monotone).  With equal-area
// candidates (`area <= minarea`, last one wins) the start decides which rectangle is returned, so the
// hull built from the pixel set is rotated to the same start:
//   mode 1 (outer border, traced counter-clockwise on screen from its (y,x)-smallest pixel): that pixel LAST
//   mode 2 (hole border, traced clockwise from the pixel left of the hole's first pixel): that pixel first
//          when it is a hull vertex, otherwise the (y,x)-smallest vertex (already first)
DBG_HD void dbg_hull_rotate(DbgPt *h, int n, int mode, int sx, int sy) {
    if (n < 3) return;
    int shift = 0;
    if (mode == 1) shift = 1;
    else if (mode == 2) {
        for (int i = 0; i < n; i++)
            if (h[i].x == sx && h[i].y == sy) { shift = i; break; }
    }
    if (shift == 0) return;
    // rotate left by `shift` with three reversals
    for (int a = 0, b = shift - 1; a < b; a++, b--) { DbgPt t = h[a]; h[a] = h[b]; h[b] = t; }
    for (int a = shift, b = n - 1; a < b; a++, b--) { DbgPt t = h[a]; h[a] = h[b]; h[b] = t; }
    for (int a = 0, b = n - 1; a < b; a++, b--) { DbgPt t = h[a]; h[a] = h[b]; h[b] = t; }
}

// cv::minAreaRect on a convex hull (float32 evaluation as rotcalipers.cpp).
// cv2 (4.13) reports a rotated rectangle in the frame whose angle lies in [-90, 0): the angle (degrees, evaluated in
// double) is turned by quarter turns into that range, width and height changing places with every turn, and is rounded
// to float once at the end.  (Measured: bit-equal to cv2.minAreaRect on 2*10^5 random hulls, segments and points.)
DBG_HD void dbg_cv_frame(DbgRect &box, double adeg) {
    while (adeg >= 0.0) { adeg -= 90.0; const float t = box.w; box.w = box.h; box.h = t; }
    while (adeg < -90.0) { adeg += 90.0; const float t = box.w; box.w = box.h; box.h = t; }
    box.angle = (float)adeg;
}

DBG_HD_BIG DbgRect dbg_min_area_rect(const DbgPt *hp, int n) {
    DbgRect box;
    box.cx = box.cy = box.w = box.h = box.angle = 0.f;
    if (n <= 0) return box;
    if (n == 1) { box.cx = (float)hp[0].x; box.cy = (float)hp[0].y; dbg_cv_frame(box, 0.0); return box; }
    if (n == 2) {
        box.cx = ((float)hp[0].x + (float)hp[1].x) * 0.5f;
        box.cy = ((float)hp[0].y + (float)hp[1].y) * 0.5f;
        const double dx = (double)((float)hp[1].x - (float)hp[0].x), dy = (double)((float)hp[1].y - (float)hp[0].y);
        box.w = (float)sqrt(dx * dx + dy * dy);
        box.h = 0.f;
        dbg_cv_frame(box, atan2(dy, dx) * 180 / DBG_PI);
        return box;
    }
    // orientation of the hull
    float orientation = 0.f;
    {
        double ax = (double)(hp[0].x - hp[n - 1].x), ay = (double)(hp[0].y - hp[n - 1].y);
        for (int i = 0; i < n; i++) {
            const DbgPt &a = hp[i], &b = hp[(i + 1) % n];
            const double bx = (double)(b.x - a.x), by = (double)(b.y - a.y);
            const double conv = ax * by - ay * bx;
            if (conv != 0) { orientation = conv > 0 ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    int left = 0, bottom = 0, right = 0, top = 0;
    {
        float lx = (float)hp[0].x, rx = lx, ty = (float)hp[0].y, by = ty;
        for (int i = 0; i < n; i++) {
            const float x = (float)hp[i].x, y = (float)hp[i].y;
            if (x < lx) { lx = x; left = i; }
            if (x > rx) { rx = x; right = i; }
            if (y > ty) { ty = y; top = i; }
            if (y < by) { by = y; bottom = i; }
        }
    }
    int seq[4] = {bottom, right, top, left};
    float base_a = orientation, base_b = 0.f;
    float minarea = 3.402823466e+38f;
    int best_left = 0, best_bottom = 0;
    float best_a = 0.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
#define DBG_VX(i) ((float)(hp[((i) + 1) % n].x - hp[(i)].x))
#define DBG_VY(i) ((float)(hp[((i) + 1) % n].y - hp[(i)].y))
#define DBG_INVLEN(i) ((float)(1. / sqrt((double)DBG_VX(i) * (double)DBG_VX(i) + (double)DBG_VY(i) * (double)DBG_VY(i))))
    for (int k = 0; k < n; k++) {
        // which calipers side meets its polygon edge first: the four edges are turned into the frame of side 0
        // (side 1 by -90, side 2 by 180, side 3 by +90 degrees) and compared by the sign of a cross product
        // (exact on integer edge vectors; cv2 >= 4.5.2 -- older versions compared float cosines)
        float rvx[4], rvy[4];
        rvx[0] = DBG_VX(seq[0]);  rvy[0] = DBG_VY(seq[0]);
        rvx[1] = DBG_VY(seq[1]);  rvy[1] = -DBG_VX(seq[1]);
        rvx[2] = -DBG_VX(seq[2]); rvy[2] = -DBG_VY(seq[2]);
        rvx[3] = -DBG_VY(seq[3]); rvy[3] = DBG_VX(seq[3]);
        int main_element = 0;
        for (int i = 1; i < 4; i++) {
            // first vector to the right (clockwise) of the second
            if (rvy[i] * rvx[main_element] + (-rvx[i]) * rvy[main_element] < 0.f) main_element = i;
        }
        {
            const int pindex = seq[main_element];
            const float il = DBG_INVLEN(pindex);
            const float lead_x = DBG_VX(pindex) * il, lead_y = DBG_VY(pindex) * il;
            switch (main_element) {
                case 0: base_a = lead_x; base_b = lead_y; break;
                case 1: base_a = lead_y; base_b = -lead_x; break;
                case 2: base_a = -lead_x; base_b = -lead_y; break;
                default: base_a = -lead_y; base_b = lead_x; break;
            }
        }
        seq[main_element] += 1;
        if (seq[main_element] == n) seq[main_element] = 0;
        float dx = (float)hp[seq[1]].x - (float)hp[seq[3]].x, dy = (float)hp[seq[1]].y - (float)hp[seq[3]].y;
        const float width = dx * base_a + dy * base_b;
        dx = (float)hp[seq[2]].x - (float)hp[seq[0]].x;
        dy = (float)hp[seq[2]].y - (float)hp[seq[0]].y;
        const float height = -dx * base_b + dy * base_a;
        const float area = width * height;
        if (area <= minarea) {
            minarea = area;
            best_left = seq[3]; best_bottom = seq[0];
            best_a = base_a; best_b = base_b; best_w = width; best_h = height;
        }
    }
#undef DBG_VX
#undef DBG_VY
#undef DBG_INVLEN
    const float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
    const float C1 = A1 * (float)hp[best_left].x + (float)hp[best_left].y * B1;
    const float C2 = A2 * (float)hp[best_bottom].x + (float)hp[best_bottom].y * B2;
    const float idet = 1.f / (A1 * B2 - A2 * B1);
    const float px = (C1 * B2 - C2 * B1) * idet, py = (A1 * C2 - A2 * C1) * idet;
    const float o1x = A1 * best_w, o1y = B1 * best_w, o2x = A2 * best_h, o2y = B2 * best_h;
    box.cx = px + (o1x + o2x) * 0.5f;
    box.cy = py + (o1y + o2y) * 0.5f;
    box.w = (float)sqrt((double)o1x * o1x + (double)o1y * o1y);
    box.h = (float)sqrt((double)o2x * o2x + (double)o2y * o2y);
    dbg_cv_frame(box, atan2((double)o1y, (double)o1x) * 180 / DBG_PI);
    return box;
}

DBG_HD void dbg_box_points(const DbgRect &r, DbgPtF pt[4]) {
    const double ang = (double)r.angle * DBG_PI / 180.;
    const float b = (float)cos(ang) * 0.5f, a = (float)sin(ang) * 0.5f;
    pt[0].x = r.cx - a * r.h - b * r.w;
    pt[0].y = r.cy + b * r.h - a * r.w;
    pt[1].x = r.cx + a * r.h - b * r.w;
    pt[1].y = r.cy - b * r.h - a * r.w;
    pt[2].x = 2 * r.cx - pt[0].x;
    pt[2].y = 2 * r.cy - pt[0].y;
    pt[3].x = 2 * r.cx - pt[1].x;
    pt[3].y = 2 * r.cy - pt[1].y;
}

// get_mini_boxes: stable sort by x, then pick (tl, tr, br, bl).  Returns min(w, h).
DBG_HD float dbg_mini_box(const DbgRect &r, DbgPtF out[4]) {
    DbgPtF p[4];
    dbg_box_points(r, p);
    for (int i = 1; i < 4; i++) {  // stable insertion sort on x
        DbgPtF t = p[i];
        int j = i - 1;
        while (j >= 0 && p[j].x > t.x) { p[j + 1] = p[j]; j--; }
        p[j + 1] = t;
    }
    int i1, i2, i3, i4;
    if (p[1].y > p[0].y) { i1 = 0; i4 = 1; } else { i1 = 1; i4 = 0; }
    if (p[3].y > p[2].y) { i2 = 2; i3 = 3; } else { i2 = 3; i3 = 2; }
    out[0] = p[i1]; out[1] = p[i2]; out[2] = p[i3]; out[3] = p[i4];
    return r.w < r.h ? r.w : r.h;
}

// ---- cv2.fillPoly coverage of one row -------------------------------------------------
// Up to 5 closed intervals [lo, hi] per row: the scan-line interior span and the Bresenham
// run of each of the 4 boundary edges.  Returns the number of intervals written (unsorted,
// possibly overlapping); callers merge them.
DBG_HD long long dbg_floordiv(long long a, long long b) {  // b > 0
    long long q = a / b;
    if ((a % b != 0) && (a < 0)) q--;
    return q;
}
DBG_HD long long dbg_ceildiv(long long a, long long b) { return -dbg_floordiv(-a, b); }

// cv::clipLine(Size(mw, mh), p1, p2): Cohen-Sutherland on the end points, the moved coordinate truncated towards
// zero from a double product.  cv::Line runs its Bresenham on the CLIPPED segment, so an edge that leaves the mask
// paints other pixels than the same edge cut pixel by pixel.  Returns false when nothing of the line is inside.
DBG_HD bool dbg_clip_line(int mw, int mh, DbgPt &a, DbgPt &b) {
    const long long right = mw - 1, bottom = mh - 1;
    if (mw <= 0 || mh <= 0) return false;
    long long x1 = a.x, y1 = a.y, x2 = b.x, y2 = b.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long t;
        if (c1 & 12) {
            t = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(t - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = t;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            t = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(t - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = t;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                t = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(t - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = t;
                c1 = 0;
            }
            if (c2) {
                t = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(t - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = t;
                c2 = 0;
            }
        }
    }
    a.x = (int)x1; a.y = (int)y1; b.x = (int)x2; b.y = (int)y2;
    return (c1 | c2) == 0;
}

// (mw, mh) = size of the mask the quad is painted into (cv2.fillPoly clips to it); mw <= 0: unbounded.
DBG_HD_BIG int dbg_row_cover(const DbgPt q[4], int y, int mw, int mh, int lo[5], int hi[5]) {
    int cnt = 0;
    // boundary edges: cv::Line -> clipLine, then LineIterator(8-connected, left_to_right)
    for (int e = 0; e < 4; e++) {
        DbgPt p1 = q[(e + 3) & 3], p2 = q[e];
        if (mw > 0 && ((unsigned)p1.x >= (unsigned)mw || (unsigned)p2.x >= (unsigned)mw || (unsigned)p1.y >= (unsigned)mh ||
                       (unsigned)p2.y >= (unsigned)mh)) {
            if (!dbg_clip_line(mw, mh, p1, p2)) continue;
        }
        if (p2.x < p1.x) { DbgPt t = p1; p1 = p2; p2 = t; }  // left_to_right
        const int dx = p2.x - p1.x;
        int dy = p2.y - p1.y;
        const int sy = dy < 0 ? -1 : 1;
        dy = dy < 0 ? -dy : dy;
        const int ylo = p1.y < p2.y ? p1.y : p2.y, yhi = p1.y < p2.y ? p2.y : p1.y;
        if (y < ylo || y > yhi) continue;
        if (dy > dx) {
            // y-major: one pixel per row; k steps along y, minor moves in +x
            const long long k = (long long)(y - p1.y) * sy;
            const long long m = dx == 0 ? 0 : dbg_ceildiv(2LL * dx * k - dy, 2LL * dy);
            const int x = p1.x + (int)(m < 0 ? 0 : m);
            lo[cnt] = x; hi[cnt] = x; cnt++;
        } else {
            // x-major: run of pixels on this row; r = minor steps taken
            if (dy == 0) { lo[cnt] = p1.x; hi[cnt] = p2.x; cnt++; continue; }
            const long long r = (long long)(y - p1.y) * sy;
            // ceil((2 dy k - dx) / (2 dx)) == r  <=>  (2 dx (r-1) + dx) / (2 dy) < k <= (2 dx r + dx) / (2 dy)
            long long kmin = dbg_floordiv(2LL * dx * (r - 1) + dx, 2LL * dy) + 1;
            long long kmax = dbg_floordiv(2LL * dx * r + dx, 2LL * dy);
            if (kmin < 0) kmin = 0;
            if (kmax > dx) kmax = dx;
            if (kmin > kmax) continue;
            lo[cnt] = p1.x + (int)kmin; hi[cnt] = p1.x + (int)kmax; cnt++;
        }
    }
    // interior span: FillEdgeCollection, edges active on [y0, y1).  An edge with an end point outside the mask takes
    // its slope from the end points cv::clipLine leaves (when they still differ in y), extrapolated back to the
    // edge's first row (cv2 >= 4.5.x CollectPolyEdges: "use clipped endpoints to create a more accurate PolyEdge").
    // Measured against cv2 4.13: exact for quads whose bounding box is cut by the image border by a few pixels
    // (box_score_fast near the page edge, 20 000 cases); quads lying mostly outside the mask still differ in ~2 %.
    long long xs[4];
    int na = 0;
    for (int e = 0; e < 4; e++) {
        const DbgPt p0 = q[(e + 3) & 3], p1 = q[e];
        if (p0.y == p1.y) continue;
        long long c0x = (long long)p0.x << 16, c1x = (long long)p1.x << 16;
        int c0y = p0.y, c1y = p1.y;
        if (mw > 0 && ((unsigned)p0.x >= (unsigned)mw || (unsigned)p1.x >= (unsigned)mw || (unsigned)p0.y >= (unsigned)mh ||
                       (unsigned)p1.y >= (unsigned)mh)) {
            DbgPt t0 = p0, t1 = p1;
            dbg_clip_line(mw, mh, t0, t1);
            if (t0.y != t1.y) { c0x = (long long)t0.x << 16; c1x = (long long)t1.x << 16; c0y = t0.y; c1y = t1.y; }
        }
        const long long ddx = (c1x - c0x) / (c1y - c0y);  // C truncation, as OpenCV
        int y0, y1;
        long long xstart;
        if (p0.y < p1.y) { y0 = p0.y; y1 = p1.y; xstart = c0x + (long long)(p0.y - c0y) * ddx; }
        else { y0 = p1.y; y1 = p0.y; xstart = c1x + (long long)(p1.y - c1y) * ddx; }
        if (y < y0 || y >= y1) continue;
        xs[na++] = xstart + (long long)(y - y0) * ddx;
    }
    if (na >= 2) {
        // sort the (<= 4) crossings; fill between pairs (0,1), (2,3)
        for (int i = 1; i < na; i++) {
            long long t = xs[i];
            int j = i - 1;
            while (j >= 0 && xs[j] > t) { xs[j + 1] = xs[j]; j--; }
            xs[j + 1] = t;
        }
        for (int i = 0; i + 1 < na && cnt < 5; i += 2) {
            const int a = (int)((xs[i] + 0xFFFF) >> 16), b = (int)(xs[i + 1] >> 16);
            if (a <= b) { lo[cnt] = a; hi[cnt] = b; cnt++; }
        }
    }
    return cnt;
}

// merge intervals in place (sorted by lo); returns count
DBG_HD int dbg_merge(int *lo, int *hi, int n) {
    for (int i = 1; i < n; i++) {
        int l = lo[i], h = hi[i], j = i - 1;
        while (j >= 0 && lo[j] > l) { lo[j + 1] = lo[j]; hi[j + 1] = hi[j]; j--; }
        lo[j + 1] = l; hi[j + 1] = h;
    }
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (m > 0 && lo[i] <= hi[m - 1] + 1) { if (hi[i] > hi[m - 1]) hi[m - 1] = hi[i]; }
        else { lo[m] = lo[i]; hi[m] = hi[i]; m++; }
    }
    return m;
}

// ---- Clipper 6.4.2 ClipperOffset(jtRound, etClosedPolygon).Execute(delta), before the union ----
DBG_HD long long dbg_cround(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

// in: 4 float points (truncated to integers like pyclipper); out: up to max_out int points.
// Returns the number of points, 0 if the path degenerates, -1 on overflow.
DBG_HD_BIG int dbg_clipper_offset(const DbgPtF in[4], double delta, DbgPt *out, int max_out) {
    DbgPt c[4];
    int n = 0;
    {
        DbgPt p[4];
        for (int i = 0; i < 4; i++) { p[i].x = (int)in[i].x; p[i].y = (int)in[i].y; }
        int hi = 3;
        while (hi > 0 && p[0].x == p[hi].x && p[0].y == p[hi].y) hi--;
        c[n++] = p[0];
        for (int i = 1; i <= hi; i++)
            if (c[n - 1].x != p[i].x || c[n - 1].y != p[i].y) c[n++] = p[i];
    }
    if (n < 3) return 0;
    {
        double a = 0.0;
        for (int i = 0, j = n - 1; i < n; j = i++) a += ((double)c[j].x + c[i].x) * ((double)c[j].y - c[i].y);
        if (!(-a * 0.5 >= 0)) {
            for (int i = 0; i < n / 2; i++) { DbgPt t = c[i]; c[i] = c[n - 1 - i]; c[n - 1 - i] = t; }
        }
    }
    const double ad = fabs(delta);
    if (ad < 1e-20) {
        for (int i = 0; i < n; i++) out[i] = c[i];
        return n;
    }
    double y = 0.25;
    if (0.25 > ad * 0.25) y = ad * 0.25;
    double steps = DBG_PI / acos(1 - y / ad);
    if (steps > ad * DBG_PI) steps = ad * DBG_PI;
    double m_sin = sin(2 * DBG_PI / steps);
    const double m_cos = cos(2 * DBG_PI / steps);
    const double steps_per_rad = steps / (2 * DBG_PI);
    if (delta < 0.0) m_sin = -m_sin;
    double nx[4], ny[4];
    for (int i = 0; i < n; i++) {
        const DbgPt p1 = c[i], p2 = c[(i + 1) % n];
        double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
        if (dx == 0 && dy == 0) { nx[i] = ny[i] = 0; continue; }
        const double f = 1.0 / sqrt(dx * dx + dy * dy);
        dx *= f; dy *= f;
        nx[i] = dy; ny[i] = -dx;
    }
    int m = 0;
#define DBG_PUSH(X, Y) do { if (m >= max_out) return -1; out[m].x = (int)(X); out[m].y = (int)(Y); m++; } while (0)
    int k = n - 1;
    for (int j = 0; j < n; j++) {
        double sin_a = nx[k] * ny[j] - nx[j] * ny[k];
        const double sx = (double)c[j].x, sy = (double)c[j].y;
        if (fabs(sin_a * delta) < 1.0) {
            const double cos_a = nx[k] * nx[j] + ny[j] * ny[k];
            if (cos_a > 0) {
                DBG_PUSH(dbg_cround(sx + nx[k] * delta), dbg_cround(sy + ny[k] * delta));
                continue;  // clipper returns before k = j
            }
        } else if (sin_a > 1.0) sin_a = 1.0;
        else if (sin_a < -1.0) sin_a = -1.0;
        if (sin_a * delta < 0) {
            DBG_PUSH(dbg_cround(sx + nx[k] * delta), dbg_cround(sy + ny[k] * delta));
            DBG_PUSH(c[j].x, c[j].y);
            DBG_PUSH(dbg_cround(sx + nx[j] * delta), dbg_cround(sy + ny[j] * delta));
        } else {
            const double ang = atan2(sin_a, nx[k] * nx[j] + ny[k] * ny[j]);
            long long st = dbg_cround(steps_per_rad * fabs(ang));
            if (st < 1) st = 1;
            double X = nx[k], Y = ny[k];
            for (long long i = 0; i < st; i++) {
                DBG_PUSH(dbg_cround(sx + X * delta), dbg_cround(sy + Y * delta));
                const double X2 = X;
                X = X * m_cos - m_sin * Y;
                Y = X2 * m_sin + Y * m_cos;
            }
            DBG_PUSH(dbg_cround(sx + nx[j] * delta), dbg_cround(sy + ny[j] * delta));
        }
        k = j;
    }
#undef DBG_PUSH
    return m;
}

// shapely area * ratio / length on the float quad (float64)
DBG_HD double dbg_unclip_distance(const DbgPtF b[4], double ratio) {
    double area2 = 0.0, len = 0.0;
    for (int i = 0; i < 4; i++) {
        const double x0 = b[i].x, y0 = b[i].y, x1 = b[(i + 1) & 3].x, y1 = b[(i + 1) & 3].y;
        area2 += x0 * y1 - y0 * x1;
        len += sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
    }
    if (len == 0) return -1.0;
    return 0.5 * fabs(area2) * ratio / len;
}

// box[:,0] = clip(round(box[:,0] / width * dest_w), 0, dest_w)  (float32 division, float64 scale, half-even)
DBG_HD int dbg_scale_coord(float v, int size, double dest) {
    const float q = v / (float)size;
    double r = rint((double)q * dest);
    if (r < 0) r = 0;
    if (r > dest) r = dest;
    return (int)(float)r;
}
