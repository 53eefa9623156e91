// hostsim.cpp — g++ build of the host/device core headers (deepemia_b200/csrc/core/*.cuh).
// TEST INFRASTRUCTURE ONLY: lets the CPU-only CI (`pytest -m "not gpu"`) check the exact algorithms the CUDA
// kernels execute against OpenCV / torch / the oracle without a GPU.  Nothing in deepemia_b200/ loads this.
#include <vector>
#include <cstring>
#include <cstdlib>
#include "../../deepemia_b200/csrc/core/emia_common.cuh"
#include "../../deepemia_b200/csrc/core/emia_contour.cuh"
#include "../../deepemia_b200/csrc/core/emia_hull.cuh"
#include "../../deepemia_b200/csrc/core/emia_ellipse.cuh"
#include "../../deepemia_b200/csrc/core/emia_measure.cuh"
#include "../../deepemia_b200/csrc/core/emia_paste.cuh"
#include "../../deepemia_b200/csrc/core/emia_moments.cuh"
#include "../../deepemia_b200/csrc/core/emia_scalebar.cuh"

static void pack_bits(const uint8_t* mask, int H, int W, std::vector<uint32_t>& bits, int& ww) {
    ww = (W + 31) / 32;
    bits.assign((size_t)H * ww, 0u);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (mask[(size_t)y * W + x]) bits[(size_t)y * ww + (x >> 5)] |= 1u << (x & 31);
}

static double g_last_perim = 0.0;
extern "C" {

// returns number of contours (discovery order) or -1 on overflow
int sim_find_contours(const uint8_t* mask, int H, int W, uint32_t* pts, int cap_pts, int* cstart, int cap_c,
                      int* n_pts) {
    std::vector<uint32_t> bits; int ww;
    pack_bits(mask, H, W, bits, ww);
    std::vector<uint32_t> mk((size_t)H * ww), ng((size_t)H * ww);
    EmiaBitView v{bits.data(), ww, H, ww, 0, 0};
    EmiaContourOut o{};
    o.pts = pts; o.cap_pts = cap_pts; o.cstart = cstart; o.cap_contours = cap_c; o.store = 1; o.track = 1; o.diag = nullptr;
    emia_find_external_contours(v, mk.data(), ng.data(), o);
    *n_pts = o.n_pts;
    g_last_perim = o.perim_last;
    return o.overflow ? -1 : o.n_contours;
}
// running arcLength of the last discovered contour of the previous sim_find_contours call
double sim_last_perimeter() { return g_last_perim; }
double sim_contour_area(const uint32_t* pts, int n) { return emia_contour_area(pts, n); }
double sim_arc_length(const uint32_t* pts, int n) { return emia_arc_length_closed(pts, n); }


int sim_convex_hull(const uint32_t* pts, int n, int clockwise, int* hull) {
    std::vector<uint64_t> keys(n + 1); std::vector<int> stack(n + 3), tmp(n + 1);
    return emia_convex_hull(pts, n, clockwise, keys.data(), stack.data(), hull, tmp.data());
}
// rect[5] = cx,cy,w,h,angle ; box[8] corners
int sim_min_area_rect(const uint32_t* pts, int n, int clockwise, float* rect, float* box) {
    std::vector<uint64_t> keys(n + 1); std::vector<int> stack(n + 3), tmp(n + 1), hull(n + 1);
    int nh = emia_convex_hull(pts, n, clockwise, keys.data(), stack.data(), hull.data(), tmp.data());
    std::vector<uint32_t> hq(nh + 2);
    for (int i = 0; i < nh; ++i) hq[i] = pts[hull[i]];
    EmiaRotRect r = emia_min_area_rect_from_hull(hq.data(), nh);
    if (nh > 2) emia_rotating_calipers(hq.data(), nh, rect + 5);
    rect[0] = r.cx; rect[1] = r.cy; rect[2] = r.w; rect[3] = r.h; rect[4] = r.angle;
    emia_box_points(r, box);
    return nh;
}

// out[5] = cx, cy, w, h, angle
int sim_fit_ellipse(const uint32_t* pts, int n, float* out) {
    EmiaEllipse e = emia_fit_ellipse(pts, n);
    out[0] = e.cx; out[1] = e.cy; out[2] = e.w; out[3] = e.h; out[4] = e.angle;
    return e.ok;
}

// rec[16]
void sim_measure_contour(const uint32_t* pts, int n, double um_pix, double* rec) {
    std::vector<uint64_t> scratch(emia_measure_scratch_bytes(n) / 8 + 2);
    emia_measure_contour(pts, n, um_pix, scratch.data(), rec);
}

// paste one instance into an H x W uint8 frame (zero-initialised by caller). returns valid flag; region[4]=rx0,ry0,rx1,ry1
int sim_paste(const float* prob, const float* box, float sx, float sy, int H, int W, uint8_t* out, int* region) {
    EmiaPasteBox b = emia_paste_prepare(box[0], box[1], box[2], box[3], sx, sy, W, H);
    region[0] = b.rx0; region[1] = b.ry0; region[2] = b.rx1; region[3] = b.ry1;
    if (!b.valid) return 0;
    for (int y = b.ry0; y < b.ry1; ++y) {
        EmiaAxisTap ay = emia_paste_axis(y, b.y0, b.y1);
        for (int x = b.rx0; x < b.rx1; ++x) {
            EmiaAxisTap ax = emia_paste_axis(x, b.x0, b.x1);
            out[(size_t)y * W + x] = emia_paste_sample(prob, ax, ay) >= 0.5f;
        }
    }
    return 1;
}
// cv2.moments of a 0/1 byte mask: out[24] (see emia_moments.cuh)
void sim_moments(const uint8_t* mask, int H, int W, double* out) {
    std::vector<uint32_t> bits; int ww;
    pack_bits(mask, H, W, bits, ww);
    long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int y = 0; y < H; ++y)
        for (int c = 0; c < ww; ++c) emia_word_moments(bits[(size_t)y * ww + c], c * 32, y, acc);
    emia_complete_moments(acc, out);
}
// ---- row f3: scale-bar line detection (emia_scalebar.cuh) -------------------------------------------------------------
void sim_bgr2gray(const uint8_t* bgr, int n, uint8_t* gray) {
    for (int i = 0; i < n; ++i) gray[i] = emia_bgr2gray(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}
// cv2.Canny(gray, low, high): edges 0 / 255
void sim_canny(const uint8_t* gray, int H, int W, int low, int high, uint8_t* edges) {
    std::vector<uint8_t> map((size_t)H * W);
    std::vector<int> stack;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int c = emia_canny_classify(gray, H, W, W, x, y, low, high);
            map[(size_t)y * W + x] = (uint8_t)c;
            if (c == 2) stack.push_back(y * W + x);
        }
    while (!stack.empty()) {
        int p = stack.back(); stack.pop_back();
        int y = p / W, x = p % W;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int xx = x + dx, yy = y + dy;
                if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
                if (map[(size_t)yy * W + xx] == 0) { map[(size_t)yy * W + xx] = 2; stack.push_back(yy * W + xx); }
            }
    }
    for (size_t i = 0; i < map.size(); ++i) edges[i] = map[i] == 2 ? 255 : 0;
}
// cv2.HoughLinesP(edges, rho=1/irho, theta, threshold, minLineLength, maxLineGap): serial orchestration of the core pieces
// (the kernel runs the same pieces warp-cooperatively).  trig: 2*numangle floats.  returns the number of lines (x1,y1,x2,y2).
int sim_hough_lines_p(const uint8_t* edges, int H, int W, const float* trig, int numangle, int numrho, int threshold, int line_len,
                      int line_gap, int* lines, int max_lines) {
    std::vector<int> accum((size_t)numangle * numrho, 0);
    std::vector<uint8_t> mask((size_t)H * W);
    std::vector<int> nz;
    for (int i = 0; i < H * W; ++i) { mask[i] = edges[i] != 0; if (edges[i]) nz.push_back(i); }
    uint64_t rng = (uint64_t)-1;
    int nl = 0;
    for (int count = (int)nz.size(); count > 0; --count) {
        int idx = emia_cv_rng_uniform0(rng, count);
        int p = nz[idx];
        nz[idx] = nz[count - 1];
        int i = p / W, j = p % W;
        if (!mask[p]) continue;
        int max_val = threshold - 1, max_n = 0;
        for (int n = 0; n < numangle; ++n) {
            int r = emia_hough_rho_bin(j, i, trig[2 * n], trig[2 * n + 1], numrho);
            int val = ++accum[(size_t)n * numrho + r];
            if (max_val < val) { max_val = val; max_n = n; }
        }
        if (max_val < threshold) continue;
        EmiaHoughWalk w = emia_hough_walk_setup(j, i, trig[2 * max_n], trig[2 * max_n + 1]);
        int end[2][2] = {{j, i}, {j, i}};
        int steps[2] = {0, 0};
        for (int k = 0; k < 2; ++k) {
            int gap = 0;
            for (int t = 0;; ++t) {
                int j1, i1;
                emia_hough_walk_at(w, k, t, j1, i1);
                if (j1 < 0 || j1 >= W || i1 < 0 || i1 >= H) break;
                if (mask[(size_t)i1 * W + j1]) { gap = 0; end[k][0] = j1; end[k][1] = i1; steps[k] = t; }
                else if (++gap > line_gap) break;
            }
        }
        int adx = end[1][0] - end[0][0], ady = end[1][1] - end[0][1];
        bool good = (adx < 0 ? -adx : adx) >= line_len || (ady < 0 ? -ady : ady) >= line_len;
        for (int k = 0; k < 2; ++k)
            for (int t = 0; t <= steps[k]; ++t) {
                int j1, i1;
                emia_hough_walk_at(w, k, t, j1, i1);
                if (mask[(size_t)i1 * W + j1]) {
                    if (good)
                        for (int n = 0; n < numangle; ++n)
                            accum[(size_t)n * numrho + emia_hough_rho_bin(j1, i1, trig[2 * n], trig[2 * n + 1], numrho)]--;
                    mask[(size_t)i1 * W + j1] = 0;
                }
            }
        if (good) {
            if (nl < max_lines) { lines[4 * nl] = end[0][0]; lines[4 * nl + 1] = end[0][1]; lines[4 * nl + 2] = end[1][0]; lines[4 * nl + 3] = end[1][1]; }
            ++nl;
        }
    }
    return nl;
}
// cv2.line(mask, p1, p2, 255, 2): mask (zeroed by the caller) gets 255 on the covered pixels
void sim_thick_line(uint8_t* mask, int H, int W, int x1, int y1, int x2, int y2) {
    emia_cv_thick_line2(W, H, x1, y1, x2, y2, [&](int x, int y) { mask[(size_t)y * W + x] = 255; });
}
}  // extern "C"
