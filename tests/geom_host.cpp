// Host build of ocr-system_b200/csrc/db_geom.h for the CPU unit tests (tests/test_db_geom.py):
// the same functions the CUDA kernels call, checked against cv2 without a GPU.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../ocr-system_b200/csrc/db_geom.h"

extern "C" {
// pts: n x (x,y) ints (any order).  out: cx, cy, w, h, angle.  Returns hull size.
// start_mode 0: hull as built (starts at the (y,x)-smallest point, clockwise on screen)
//            1: outer contour convention of cv::convexHull on a findContours border (first point last)
//            2: hole convention: start at (sx, sy) when it is a hull vertex
int geom_min_area_rect_ex(const int *pts, int n, int start_mode, int sx, int sy, float *out5) {
    std::vector<DbgPt> p(n), hull(n + 2);
    for (int i = 0; i < n; i++) { p[i].x = pts[2 * i]; p[i].y = pts[2 * i + 1]; }
    dbg_sort(p.data(), n);
    int hn = dbg_hull_sorted(p.data(), n, hull.data());
    dbg_hull_rotate(hull.data(), hn, start_mode, sx, sy);
    DbgRect r = dbg_min_area_rect(hull.data(), hn);
    out5[0] = r.cx; out5[1] = r.cy; out5[2] = r.w; out5[3] = r.h; out5[4] = r.angle;
    return hn;
}
// calipers on a hull given as-is (cv::convexHull's vertex order): isolates dbg_min_area_rect from the hull construction
void geom_min_area_rect_hull(const int *hull_pts, int n, float *out5) {
    std::vector<DbgPt> h(n);
    for (int i = 0; i < n; i++) { h[i].x = hull_pts[2 * i]; h[i].y = hull_pts[2 * i + 1]; }
    DbgRect r = dbg_min_area_rect(h.data(), n);
    out5[0] = r.cx; out5[1] = r.cy; out5[2] = r.w; out5[3] = r.h; out5[4] = r.angle;
}
int geom_min_area_rect(const int *pts, int n, float *out5) {
    std::vector<DbgPt> p(n), hull(n + 2);
    for (int i = 0; i < n; i++) { p[i].x = pts[2 * i]; p[i].y = pts[2 * i + 1]; }
    dbg_sort(p.data(), n);
    int hn = dbg_hull_sorted(p.data(), n, hull.data());
    DbgRect r = dbg_min_area_rect(hull.data(), hn);
    out5[0] = r.cx; out5[1] = r.cy; out5[2] = r.w; out5[3] = r.h; out5[4] = r.angle;
    return hn;
}
void geom_box_points(const float *rect5, float *out8) {
    DbgRect r = {rect5[0], rect5[1], rect5[2], rect5[3], rect5[4]};
    DbgPtF o[4];
    dbg_box_points(r, o);
    for (int i = 0; i < 4; i++) { out8[2 * i] = o[i].x; out8[2 * i + 1] = o[i].y; }
}
float geom_mini_box(const float *rect5, float *out8) {
    DbgRect r = {rect5[0], rect5[1], rect5[2], rect5[3], rect5[4]};
    DbgPtF o[4];
    float s = dbg_mini_box(r, o);
    for (int i = 0; i < 4; i++) { out8[2 * i] = o[i].x; out8[2 * i + 1] = o[i].y; }
    return s;
}
// rasterise the quad like cv2.fillPoly(mask, [quad], 1) (pixels clipped to the mask)
void geom_fill_quad(const int *quad8, int h, int w, unsigned char *mask) {
    DbgPt q[4];
    for (int i = 0; i < 4; i++) { q[i].x = quad8[2 * i]; q[i].y = quad8[2 * i + 1]; }
    for (int y = 0; y < h; y++) {
        int lo[5], hi[5];
        int c = dbg_row_cover(q, y, w, h, lo, hi);
        c = dbg_merge(lo, hi, c);
        for (int i = 0; i < c; i++)
            for (int x = (lo[i] < 0 ? 0 : lo[i]); x <= hi[i] && x < w; x++) mask[(size_t)y * w + x] = 1;
    }
}
int geom_clipper_offset(const float *in8, double delta, int *out, int max_out) {
    DbgPtF in[4];
    for (int i = 0; i < 4; i++) { in[i].x = in8[2 * i]; in[i].y = in8[2 * i + 1]; }
    std::vector<DbgPt> o(max_out);
    int m = dbg_clipper_offset(in, delta, o.data(), max_out);
    for (int i = 0; i < m; i++) { out[2 * i] = o[i].x; out[2 * i + 1] = o[i].y; }
    return m;
}
double geom_unclip_distance(const float *in8, double ratio) {
    DbgPtF in[4];
    for (int i = 0; i < 4; i++) { in[i].x = in8[2 * i]; in[i].y = in8[2 * i + 1]; }
    return dbg_unclip_distance(in, ratio);
}
int geom_scale_coord(float v, int size, double dest) { return dbg_scale_coord(v, size, dest); }

// The whole per-candidate sequence of db_candidate_geometry_kernel on the host:
// pixel set (any order) -> row extremes -> hull -> rect -> score -> unclip -> rect -> scaled box.
// Returns 1 accepted, 0 rejected (reason in *why: 1 sside, 2 score, 3 unclip, 4 sside2).
int geom_candidate(const int *pts, int n, int hole, int sx, int sy, const float *pred, int h, int w, double box_thresh,
                   double unclip_ratio, int min_size, double dest_w, double dest_h, int *out8, double *score_out,
                   float *box8_out, int *why) {
    // row extremes
    int ymin = 1 << 30, ymax = -1;
    for (int i = 0; i < n; i++) { if (pts[2*i+1] < ymin) ymin = pts[2*i+1]; if (pts[2*i+1] > ymax) ymax = pts[2*i+1]; }
    int rows = ymax - ymin + 1;
    std::vector<int> rmin(rows, 1 << 30), rmax(rows, -1);
    for (int i = 0; i < n; i++) { int r = pts[2*i+1] - ymin, x = pts[2*i]; if (x < rmin[r]) rmin[r] = x; if (x > rmax[r]) rmax[r] = x; }
    std::vector<DbgPt> p, hull(2 * rows + 4);
    for (int r = 0; r < rows; r++) if (rmax[r] >= 0) { p.push_back({rmin[r], ymin + r}); p.push_back({rmax[r], ymin + r}); }
    int hn = dbg_hull_sorted(p.data(), (int)p.size(), hull.data());
    dbg_hull_rotate(hull.data(), hn, hole ? 2 : 1, sx, sy);
    DbgRect r = dbg_min_area_rect(hull.data(), hn);
    DbgPtF o[4];
    float sside = dbg_mini_box(r, o);
    for (int i = 0; i < 4; i++) { box8_out[2*i] = o[i].x; box8_out[2*i+1] = o[i].y; }
    *why = 0;
    if (sside < (float)min_size) { *why = 1; return 0; }
    float fminx = o[0].x, fmaxx = o[0].x, fminy = o[0].y, fmaxy = o[0].y;
    for (int i = 1; i < 4; i++) { fminx = fminf(fminx, o[i].x); fmaxx = fmaxf(fmaxx, o[i].x); fminy = fminf(fminy, o[i].y); fmaxy = fmaxf(fmaxy, o[i].y); }
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    int xmin = clampi((int)floorf(fminx), 0, w - 1), xmax = clampi((int)ceilf(fmaxx), 0, w - 1);
    int yminb = clampi((int)floorf(fminy), 0, h - 1), ymaxb = clampi((int)ceilf(fmaxy), 0, h - 1);
    DbgPt q[4];
    for (int i = 0; i < 4; i++) { q[i].x = (int)(o[i].x - (float)xmin); q[i].y = (int)(o[i].y - (float)yminb); }
    int mh = ymaxb - yminb + 1, mw = xmax - xmin + 1;
    double sum = 0; long cnt = 0;
    for (int ry = 0; ry < mh; ry++) {
        int lo[5], hi[5];
        int c = dbg_merge(lo, hi, dbg_row_cover(q, ry, mw, mh, lo, hi));
        for (int i = 0; i < c; i++) {
            int a = lo[i] < 0 ? 0 : lo[i], b = hi[i] > mw - 1 ? mw - 1 : hi[i];
            for (int x = a; x <= b; x++) { sum += pred[(size_t)(yminb + ry) * w + xmin + x]; cnt++; }
        }
    }
    double score = cnt ? sum / cnt : 0.0;
    *score_out = score;
    if (box_thresh > score) { *why = 2; return 0; }
    double dist = dbg_unclip_distance(o, unclip_ratio);
    std::vector<DbgPt> offs(1024), oh(1026);
    int m = dist >= 0 ? dbg_clipper_offset(o, dist, offs.data(), 1024) : 0;
    if (m < 3) { *why = 3; return 0; }
    dbg_sort(offs.data(), m);
    int hn2 = dbg_hull_sorted(offs.data(), m, oh.data());
    DbgRect r2 = dbg_min_area_rect(oh.data(), hn2);
    DbgPtF o2[4];
    float ss2 = dbg_mini_box(r2, o2);
    if (ss2 < (float)(min_size + 2)) { *why = 4; return 0; }
    for (int i = 0; i < 4; i++) { out8[2*i] = dbg_scale_coord(o2[i].x, w, dest_w); out8[2*i+1] = dbg_scale_coord(o2[i].y, h, dest_h); }
    return 1;
}
}
