// emia_scalebar_kernels.cuh — row f3: the line-detection part of detect_scale_bar (src/utils/scalebar_ocr.py:72-373) for a BATCH
// of micrographs: BGR2GRAY of the ROI -> Canny -> probabilistic Hough transform -> mean grey level under every line.
// The arithmetic is core/emia_scalebar.cuh (bit-identical to OpenCV's, checked on the CPU by tests/hostsim).  The reference
// does this once (twice, in fact: inference.py:756,763) per image on a ROI of a few thousand pixels, strictly serially; here one
// image is one CTA / one warp, so the batch is what fills the machine.
#pragma once
#include "core/emia_scalebar.cuh"

// ---- grey ROI -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scalebar_gray(const uint8_t* __restrict__ images, int B, int H, int W, int channels, int rx, int ry,
                                                       int rw, int rh, uint8_t* __restrict__ gray) {
    int64_t total = (int64_t)B * rh * rw;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(t / ((int64_t)rh * rw));
        int r = (int)(t - (int64_t)b * rh * rw);
        int y = r / rw, x = r - y * rw;
        const uint8_t* p = images + (((int64_t)b * H + (ry + y)) * W + (rx + x)) * channels;
        gray[t] = channels == 3 ? emia_bgr2gray(p[0], p[1], p[2]) : p[0];
    }
}

// ---- Canny: classification (one thread per pixel), then hysteresis to a fixed point (one CTA per image) -------------------------
__global__ void __launch_bounds__(256) k_canny_classify(const uint8_t* __restrict__ gray, int B, int H, int W, int low, int high,
                                                        uint8_t* __restrict__ map) {
    int64_t total = (int64_t)B * H * W;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int b = (int)(t / ((int64_t)H * W));
        int r = (int)(t - (int64_t)b * H * W);
        int y = r / W, x = r - y * W;
        map[t] = (uint8_t)emia_canny_classify(gray + (int64_t)b * H * W, H, W, W, x, y, low, high);
    }
}

// weak candidates (0) 8-connected to a strong pixel (2) become strong; the fixed point does not depend on the visiting order
// (OpenCV grows the same set from a stack).  Rows are swept alternately downwards and upwards so that a chain running along the
// sweep direction is absorbed in one pass.
__global__ void __launch_bounds__(1024) k_canny_hysteresis(uint8_t* __restrict__ map_all, int H, int W) {
    uint8_t* map = map_all + (int64_t)blockIdx.x * H * W;
    const int n = H * W;
    for (int pass = 0;; ++pass) {
        int changed = 0;
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            int tt = (pass & 1) ? n - 1 - t : t;
            if (map[tt] != 0) continue;
            int y = tt / W, x = tt - y * W;
            bool s = false;
            for (int dy = -1; dy <= 1 && !s; ++dy) {
                int yy = y + dy;
                if ((unsigned)yy >= (unsigned)H) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    int xx = x + dx;
                    if ((unsigned)xx >= (unsigned)W) continue;
                    if (map[yy * W + xx] == 2) { s = true; break; }
                }
            }
            if (s) { map[tt] = 2; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int t = threadIdx.x; t < n; t += blockDim.x) map[t] = map[t] == 2 ? 255 : 0;
}

extern "C" int emia_scalebar_edges(const uint8_t* images, int32_t B, int32_t H, int32_t W, int32_t channels, int32_t rx, int32_t ry,
                                   int32_t rw, int32_t rh, int32_t low, int32_t high, uint8_t* gray, uint8_t* edges, void* stream) {
    if (B < 0 || H <= 0 || W <= 0 || (channels != 1 && channels != 3) || rx < 0 || ry < 0 || rw <= 0 || rh <= 0 || rx + rw > W ||
        ry + rh > H || (int64_t)rw * rh > (int64_t)1 << 30)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_scalebar_edges: %s", "bad argument");
    if (B == 0) return EMIA_OK;
    if (!images || !gray || !edges) return emia_fail(EMIA_ERR_BAD_ARG, "emia_scalebar_edges: %s", "null pointer");
    if (low > high) { int t = low; low = high; high = t; }
    int64_t total = (int64_t)B * rw * rh;
    unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    k_scalebar_gray<<<blocks, 256, 0, (cudaStream_t)stream>>>(images, B, H, W, channels, rx, ry, rw, rh, gray);
    k_canny_classify<<<blocks, 256, 0, (cudaStream_t)stream>>>(gray, B, rh, rw, low, high, edges);
    k_canny_hysteresis<<<(unsigned)B, 1024, 0, (cudaStream_t)stream>>>(edges, rh, rw);
    return emia_check_launch("emia_scalebar_edges launch: %s");
}

// ---- cv2.HoughLinesP: one warp per image ------------------------------------------------------------------------------------
// The algorithm is sequential by construction (random visiting order from cv::RNG, every accepted line removes its points and
// their votes before the next point is drawn), so one image is one warp: the lanes share the votes of a point (numangle / 32
// accumulator cells each), the line walk (32 steps per probe) and the un-voting of a line's points.
extern "C" size_t emia_hough_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t numangle, int32_t numrho) {
    if (B < 0 || H <= 0 || W <= 0 || numangle <= 0 || numrho <= 0) return 0;
    size_t per = (size_t)numangle * numrho * 4 + (size_t)H * W * 4 + (((size_t)H * W + 15) & ~(size_t)15);
    return per * (size_t)B + 256;
}

// Point state of one image.  FAST: the mask as a bit plane and the point list (up to HOUGH_NZ_SMEM entries) in shared memory — a
// skipped point (already removed by an earlier line) then costs a few dozen cycles instead of two dependent L2 round trips;
// otherwise (regions above HOUGH_FAST_PIXELS) bytes / ints in the global workspace.
#define HOUGH_FAST_PIXELS (128 * 1024)
#define HOUGH_NZ_SMEM 6144

template <bool FAST>
struct HoughMask {
    uint32_t* bits;     // FAST
    uint8_t* bytes;     // !FAST
    __device__ __forceinline__ bool get(int t) const { return FAST ? ((bits[t >> 5] >> (t & 31)) & 1u) != 0 : bytes[t] != 0; }
    __device__ __forceinline__ void clear(int t) const {
        if (FAST) atomicAnd(&bits[t >> 5], ~(1u << (t & 31)));
        else bytes[t] = 0;
    }
};

#define HOUGH_THREADS 128
template <bool FAST>
__global__ void __launch_bounds__(HOUGH_THREADS) k_hough_lines_p(const uint8_t* __restrict__ edges_all, int H, int W, const float* __restrict__ trig,
                                                                 int numangle, int numrho, int threshold, int line_len, int line_gap, int max_lines,
                                                                 int32_t* __restrict__ lines_all, int32_t* __restrict__ n_lines, uint8_t* ws) {
    extern __shared__ uint32_t hsm[];
    __shared__ int s_cnt[HOUGH_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.x;
    const int n = H * W;
    const size_t mask_bytes = ((size_t)n + 15) & ~(size_t)15;
    const size_t per = (size_t)numangle * numrho * 4 + (size_t)n * 4 + mask_bytes;
    int* accum = (int*)(ws + per * b);          // cleared by the caller-side memset of emia_hough_lines_p
    int* nz = accum + (size_t)numangle * numrho;
    HoughMask<FAST> mask;
    mask.bytes = (uint8_t*)(nz + n);
    mask.bits = hsm;
    const int mask_words = (n + 31) >> 5;
    const uint8_t* edges = edges_all + (size_t)b * n;
    int32_t* lines = lines_all + (size_t)b * max_lines * 4;
    // stage 1: the non-zero points in row-major order.  Each of the four warps compacts a contiguous quarter of the pixels into
    // the start of that quarter's own range of the list; the pieces are then moved together (into shared memory when they fit).
    const int chunk = ((mask_words + HOUGH_THREADS / 32 - 1) / (HOUGH_THREADS / 32)) * 32;      // pixels per warp, a multiple of 32
    const int c0 = min(n, warp * chunk), c1 = min(n, c0 + chunk);
    int my_count = 0;
    for (int base = c0; base < c1; base += 32) {
        int t = base + lane;
        bool on = t < c1 && edges[t] != 0;
        unsigned bal = __ballot_sync(0xffffffffu, on);
        if (FAST) { if (lane == 0) hsm[base >> 5] = bal; }
        else if (t < c1) mask.bytes[t] = on ? 1 : 0;
        if (on) nz[c0 + my_count + __popc(bal & ((1u << lane) - 1u))] = t;
        my_count += __popc(bal);
    }
    if (lane == 0) s_cnt[warp] = my_count;
    __syncthreads();
    int count = 0, my_off = 0;
    for (int w = 0; w < HOUGH_THREADS / 32; ++w) { if (w == warp) my_off = count; count += s_cnt[w]; }
    if (FAST && count <= HOUGH_NZ_SMEM) {
        int* nzs = (int*)(hsm + mask_words);
        for (int t = lane; t < my_count; t += 32) nzs[my_off + t] = nz[c0 + t];
        nz = nzs;
        __syncthreads();
        if (warp != 0) return;
    } else {
        // (dense frames) close the gaps in place, piece after piece
        __syncthreads();
        if (warp != 0) return;
        int dst = s_cnt[0];
        for (int w = 1; w < HOUGH_THREADS / 32; ++w) {
            const int src = min(n, w * chunk), cnt = s_cnt[w];
            for (int t0 = 0; t0 < cnt; t0 += 32) {
                const int t = t0 + lane;
                const int v = t < cnt ? nz[src + t] : 0;
                __syncwarp();
                if (t < cnt) nz[dst + t] = v;
                __syncwarp();
            }
            dst += cnt;
        }
    }
    __syncwarp();
    // the angles of this lane (slot s: angle s * 32 + lane) for tables of up to 256 angles
    float tc[8], ts[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        int a = s * 32 + lane;
        tc[s] = a < numangle ? trig[2 * a] : 0.f;
        ts[s] = a < numangle ? trig[2 * a + 1] : 0.f;
    }
    const bool small_table = numangle <= 256;
    uint64_t rng = (uint64_t)-1;
    int nl = 0;
    // rho bin of angle slot s (tables of up to 256 angles) / of angle a
    auto cell = [&](int s, int j, int i) -> int* {
        return &accum[(size_t)(s * 32 + lane) * numrho + emia_hough_rho_bin(j, i, tc[s], ts[s], numrho)];
    };
    // (largest count, first angle reaching it) over the warp
    auto reduce_best = [&](int& best_val, int& best_n) {
        for (int o = 16; o > 0; o >>= 1) {
            int ov = __shfl_xor_sync(0xffffffffu, best_val, o), on_ = __shfl_xor_sync(0xffffffffu, best_n, o);
            if (ov > best_val || (ov == best_val && on_ < best_n)) { best_val = ov; best_n = on_; }
        }
    };
    // votes of ONE point.  A cell (angle a, rho bin) is only ever touched by lane a % 32 of this warp, so no atomic is needed: the
    // lane loads its (up to 8) cells back to back, adds and stores.  (Atomics WITH a returned value were the first version: 75 % of
    // the kernel's stall samples sat on them — about one returned atomic per 35 cycles per warp however many were in flight, while
    // the fire-and-forget reductions of the un-voting cost 30x less per operation.)  ld / st .cg: the cells live in L2.
    auto vote_point = [&](int j, int i, int& best_val, int& best_n) {
        best_val = threshold - 1; best_n = 0x7fffffff;
        if (small_table) {
            int* c[8];
            int vals[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) c[s] = cell(s, j, i);
#pragma unroll
            for (int s = 0; s < 8; ++s) vals[s] = (s * 32 + lane) < numangle ? __ldcg(c[s]) + 1 : -1;
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if ((s * 32 + lane) < numangle) __stcg(c[s], vals[s]);
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if (vals[s] > best_val) { best_val = vals[s]; best_n = s * 32 + lane; }
        } else {
            for (int a = lane; a < numangle; a += 32) {
                int* c = &accum[(size_t)a * numrho + emia_hough_rho_bin(j, i, trig[2 * a], trig[2 * a + 1], numrho)];
                const int val = __ldcg(c) + 1;
                __stcg(c, val);
                if (val > best_val) { best_val = val; best_n = a; }
            }
        }
        reduce_best(best_val, best_n);
    };
    // un-voting: fire-and-forget reductions (nothing waits for them; a later load of the same cell by the same lane is ordered
    // behind them) — a line of 70 points is 70 x 180 decrements, which as load / store pairs were one L2 round trip per point
    auto unvote_point = [&](int j, int i) {
        if (small_table) {
#pragma unroll
            for (int s = 0; s < 8; ++s)
                if ((s * 32 + lane) < numangle) atomicSub(cell(s, j, i), 1);
        } else {
            for (int a = lane; a < numangle; a += 32)
                atomicSub(&accum[(size_t)a * numrho + emia_hough_rho_bin(j, i, trig[2 * a], trig[2 * a + 1], numrho)], 1);
        }
    };
    // the most voted line through (j, i): walk both ways, then remove its points (and, for an accepted line, their votes)
    auto take_line = [&](int j, int i, int best_n) {
        EmiaHoughWalk w = emia_hough_walk_setup(j, i, trig[2 * best_n], trig[2 * best_n + 1]);
        int end_x[2], end_y[2], steps[2];
        for (int k = 0; k < 2; ++k) {
            int gap = 0, last = 0;
            bool stop = false;
            for (int base = 0; !stop; base += 32) {
                int j1, i1;
                emia_hough_walk_at(w, k, base + lane, j1, i1);
                bool inb = j1 >= 0 && j1 < W && i1 >= 0 && i1 < H;
                bool on = inb && mask.get(i1 * W + j1);
                unsigned b_in = __ballot_sync(0xffffffffu, inb), b_on = __ballot_sync(0xffffffffu, on);
                // serial semantics over the 32 probes (every lane computes the same thing from the ballots)
                for (int q = 0; q < 32; ++q) {
                    if (!((b_in >> q) & 1u)) { stop = true; break; }
                    if ((b_on >> q) & 1u) { gap = 0; last = base + q; }
                    else if (++gap > line_gap) { stop = true; break; }
                }
            }
            steps[k] = last;
            emia_hough_walk_at(w, k, last, end_x[k], end_y[k]);
        }
        int adx = end_x[1] - end_x[0], ady = end_y[1] - end_y[0];
        const bool good = (adx < 0 ? -adx : adx) >= line_len || (ady < 0 ? -ady : ady) >= line_len;
        for (int k = 0; k < 2; ++k) {
            for (int base = 0; base <= steps[k]; base += 32) {
                int t = base + lane, j1 = 0, i1 = 0;
                bool on = false;
                if (t <= steps[k]) {
                    emia_hough_walk_at(w, k, t, j1, i1);
                    on = mask.get(i1 * W + j1);
                    if (on) mask.clear(i1 * W + j1);
                }
                unsigned b_on = __ballot_sync(0xffffffffu, on);
                if (good) {
                    while (b_on) {
                        int q = __ffs((int)b_on) - 1;
                        b_on &= b_on - 1;
                        unvote_point(__shfl_sync(0xffffffffu, j1, q), __shfl_sync(0xffffffffu, i1, q));
                    }
                }
            }
            __syncwarp();
        }
        if (good) {
            if (lane == 0 && nl < max_lines) {
                lines[4 * nl] = end_x[0]; lines[4 * nl + 1] = end_y[0]; lines[4 * nl + 2] = end_x[1]; lines[4 * nl + 3] = end_y[1];
            }
            ++nl;
        }
    };
    // next point of OpenCV's random visiting order (the order depends on nothing but the point count)
    auto draw = [&]() -> int {
        int idx = emia_cv_rng_uniform0(rng, count);
        int p = nz[idx];
        __syncwarp();
        if (lane == 0) nz[idx] = nz[count - 1];
        __syncwarp();
        --count;
        return p;
    };
    if (small_table) {
        // A WINDOW of KW points votes at once: their cells are loaded back to back (one L2 round trip for the window instead of one
        // per point).  This is exact: the counts a point sees do not depend on LATER points, and an earlier point of the window
        // that hits the same cell is forwarded in registers (a cell belongs to one lane); only an accepted line makes the order
        // matter — then the speculative votes of the later points are taken back, the line is removed and the rest of the window is
        // redone one by one, as OpenCV does.
        constexpr int KW = 4;
        while (count > 0) {
            int wj[KW], wi[KW], nw = 0;
#pragma unroll
            for (int k = 0; k < KW; ++k) { wj[k] = 0; wi[k] = 0; }
            while (nw < KW && count > 0) {
                const int p = draw();
                if (!mask.get(p)) continue;
                const int i = p / W, j = p - i * W;
#pragma unroll
                for (int k = 0; k < KW; ++k)
                    if (k == nw) { wj[k] = j; wi[k] = i; }
                ++nw;
            }
            int idx[KW][8], vals[KW][8];
#pragma unroll
            for (int k = 0; k < KW; ++k)
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    idx[k][s] = (k < nw && (s * 32 + lane) < numangle)
                                    ? (s * 32 + lane) * numrho + emia_hough_rho_bin(wj[k], wi[k], tc[s], ts[s], numrho) : -1;
#pragma unroll
            for (int k = 0; k < KW; ++k)
#pragma unroll
                for (int s = 0; s < 8; ++s) vals[k][s] = idx[k][s] >= 0 ? __ldcg(accum + idx[k][s]) : -2;
#pragma unroll
            for (int s = 0; s < 8; ++s)
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    int v = vals[k][s];
#pragma unroll
                    for (int q = 0; q < KW; ++q)
                        if (q < k && idx[q][s] == idx[k][s]) v = vals[q][s];        // the latest earlier vote on the same cell
                    vals[k][s] = v + 1;
                    if (idx[k][s] >= 0) __stcg(accum + idx[k][s], vals[k][s]);        // same address: program order, the last one stays
                }
            bool redo = false;
#pragma unroll
            for (int k = 0; k < KW; ++k) {
                if (k >= nw) continue;
                if (redo) {
                    if (!mask.get(wi[k] * W + wj[k])) continue;
                    int bv, bn;
                    vote_point(wj[k], wi[k], bv, bn);
                    if (bv >= threshold) take_line(wj[k], wi[k], bn);
                    continue;
                }
                int bv = threshold - 1, bn = 0x7fffffff;
#pragma unroll
                for (int s = 0; s < 8; ++s)
                    if (idx[k][s] >= 0 && vals[k][s] > bv) { bv = vals[k][s]; bn = s * 32 + lane; }
                reduce_best(bv, bn);
                if (bv < threshold) continue;
#pragma unroll
                for (int q = 0; q < KW; ++q)
                    if (q > k && q < nw) unvote_point(wj[q], wi[q]);
                take_line(wj[k], wi[k], bn);
                redo = true;
            }
        }
    } else {
        while (count > 0) {
            const int p = draw();
            if (!mask.get(p)) continue;
            const int i = p / W, j = p - i * W;
            int bv, bn;
            vote_point(j, i, bv, bn);
            if (bv >= threshold) take_line(j, i, bn);
        }
    }
    if (lane == 0) n_lines[b] = nl;
}

extern "C" int emia_hough_lines_p(const uint8_t* edges, int32_t B, int32_t H, int32_t W, const float* trig, int32_t numangle,
                                  int32_t numrho, int32_t threshold, int32_t min_line_length, int32_t max_line_gap, int32_t max_lines,
                                  int32_t* lines, int32_t* n_lines, void* workspace, size_t workspace_bytes, void* stream) {
    if (B < 0 || H <= 0 || W <= 0 || numangle <= 0 || numrho <= 0 || max_lines <= 0 || (int64_t)H * W > (int64_t)1 << 30)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_hough_lines_p: %s", "bad argument");
    if (B == 0) return EMIA_OK;
    if (!edges || !trig || !lines || !n_lines || !workspace) return emia_fail(EMIA_ERR_BAD_ARG, "emia_hough_lines_p: %s", "null pointer");
    if (workspace_bytes < emia_hough_workspace_bytes(B, H, W, numangle, numrho))
        return emia_fail(EMIA_ERR_WORKSPACE, "emia_hough_lines_p: %s", "workspace too small");
    uint8_t* ws = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t per = (size_t)numangle * numrho * 4 + (size_t)H * W * 4 + (((size_t)H * W + 15) & ~(size_t)15);
    if (cudaMemset2DAsync(ws, per, 0, (size_t)numangle * numrho * 4, (size_t)B, (cudaStream_t)stream) != cudaSuccess)
        return emia_fail(EMIA_ERR_LAUNCH, "emia_hough_lines_p: %s", "clearing the accumulators failed");
    if ((int64_t)H * W <= HOUGH_FAST_PIXELS) {
        size_t smem = ((size_t)(H * W + 31) / 32 + HOUGH_NZ_SMEM) * 4;
        static const char* pad_env = getenv("EMIA_HOUGH_SMEM_PAD_KB");     // experiment knob: fewer frames in flight per SM
        if (pad_env) { smem += (size_t)atoi(pad_env) * 1024; emia_need_dyn_smem((const void*)k_hough_lines_p<true>, smem); }
        k_hough_lines_p<true><<<(unsigned)B, HOUGH_THREADS, smem, (cudaStream_t)stream>>>(edges, H, W, trig, numangle, numrho, threshold,
                                                                                          min_line_length, max_line_gap, max_lines, lines, n_lines, ws);
    } else {
        k_hough_lines_p<false><<<(unsigned)B, HOUGH_THREADS, 0, (cudaStream_t)stream>>>(edges, H, W, trig, numangle, numrho, threshold,
                                                                                        min_line_length, max_line_gap, max_lines, lines, n_lines, ws);
    }
    return emia_check_launch("emia_hough_lines_p launch: %s");
}

// ---- mean grey level under cv2.line(mask, p1, p2, 255, 2): one warp per (line, image), the mask as a bit plane in shared memory ---
__global__ void __launch_bounds__(32) k_line_mean(const uint8_t* __restrict__ gray_all, int H, int W, const int32_t* __restrict__ lines_all,
                                                  const int32_t* __restrict__ n_lines, int max_lines, int64_t* __restrict__ out) {
    extern __shared__ uint32_t plane[];
    const int b = blockIdx.y, lane = threadIdx.x;
    int nl = n_lines[b];
    nl = nl < max_lines ? nl : max_lines;
    const int words = (H * W + 31) / 32;
    // a frame has ~10 lines and room for hundreds: the grid covers 32 line slots per frame and loops over the rest
    for (int l = blockIdx.x; l < nl; l += gridDim.x) {
    const int32_t* ln = lines_all + ((size_t)b * max_lines + l) * 4;
    // only the rows the thick line can touch (its half width and its end caps reach one pixel: two rows of slack) are cleared / read
    const int ry0 = max(0, min(ln[1], ln[3]) - 2), ry1 = min(H - 1, max(ln[1], ln[3]) + 2);
    const int wlo = max(0, (ry0 * W) >> 5), whi = min(words, ((ry1 + 1) * W + 31) >> 5);
    __syncwarp();
    for (int t = wlo + lane; t < whi; t += 32) plane[t] = 0;
    __syncwarp();
    if (lane == 0) {
        uint32_t* pl = plane;
        emia_cv_thick_line2(W, H, ln[0], ln[1], ln[2], ln[3], [pl, W](int x, int y) { int t = y * W + x; pl[t >> 5] |= 1u << (t & 31); });
    }
    __syncwarp();
    const uint8_t* gray = gray_all + (size_t)b * H * W;
    long long sum = 0, cnt = 0;
    for (int t = wlo + lane; t < whi; t += 32) {
        uint32_t v = plane[t];
        cnt += __popc(v);
        while (v) {
            int q = __ffs((int)v) - 1;
            v &= v - 1;
            sum += gray[t * 32 + q];
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
        out[((size_t)b * max_lines + l) * 2] = sum;
        out[((size_t)b * max_lines + l) * 2 + 1] = cnt;
    }
    }
}

extern "C" int emia_line_mean(const uint8_t* gray, int32_t B, int32_t H, int32_t W, const int32_t* lines, const int32_t* n_lines,
                              int32_t max_lines, int64_t* sum_count, void* stream) {
    if (B < 0 || H <= 0 || W <= 0 || max_lines <= 0 || max_lines > 65535 || B > 65535)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_line_mean: %s", "bad argument");
    if (B == 0) return EMIA_OK;
    if (!gray || !lines || !n_lines || !sum_count) return emia_fail(EMIA_ERR_BAD_ARG, "emia_line_mean: %s", "null pointer");
    size_t smem = (((size_t)H * W + 31) / 32) * 4;
    if (smem > 200 * 1024) return emia_fail(EMIA_ERR_BAD_ARG, "emia_line_mean: %s", "region above 1.6 Mpixel (the bit plane lives in shared memory)");
    emia_need_dyn_smem((const void*)k_line_mean, smem);
    k_line_mean<<<dim3((unsigned)(max_lines < 32 ? max_lines : 32), (unsigned)B), 32, smem, (cudaStream_t)stream>>>(gray, H, W, lines, n_lines, max_lines, sum_count);
    return emia_check_launch("emia_line_mean launch: %s");
}
