// emia_tile_kernels.cuh — K3: nearest-neighbour resize of bit-packed instances + placement into a larger frame, the edge test of
// the tile pipeline, and K6: the RLE wire format + raw moments (part of emia_kernels.cu).
//
// Replaces (reference):
//   * tile back-projection  src/functions/inference.py:2399-2420 : cv2.resize(mask u8, (tile_w, tile_h), INTER_NEAREST) of a mask
//     predicted on the UPSCALED tile, is_edge_mask (:2522-2549), then placement of the tile-sized mask at (x_offset, y_offset)
//     of a zero (h, w) frame with clipping at the image border;
//   * scale back-projection src/functions/inference.py:2044-2054 : cv2.resize(mask, (w, h), INTER_NEAREST) (offsets 0);
//   * rle_encoding          src/utils/mask_utils.py:17-35 (column-major, 1-indexed (start, length) pairs);
//   * cv2.moments(mask)     src/functions/inference.py:1101-1104 (centroid of the label text).
// cv2.resize INTER_NEAREST: src = min(floor(dst * (1 / (dst_size / src_size))), src_size - 1), evaluated in double
// (SURVEY Appendix B.4).  The reference allocates one full (h, w) array per instance (64 MiB at 8192^2); here the result
// is again a word-aligned bit crop, so a 20 000-instance micrograph needs a few MB.
#pragma once

struct EmiaNN {
    double inv;     // 1 / (dst / src)
    double scale;   // dst / src
    int src, dst;
};
__device__ __forceinline__ EmiaNN emia_nn_make(int src, int dst) {
    EmiaNN m;
    m.scale = (double)dst / (double)src;
    m.inv = 1.0 / m.scale;
    m.src = src; m.dst = dst;
    return m;
}
__device__ __forceinline__ int emia_nn_map(const EmiaNN& m, int d) {
    const int s = (int)floor((double)d * m.inv);
    return s < m.src - 1 ? s : m.src - 1;
}
// destination indices [lo, hi] whose source index falls into [s_lo, s_hi] (the map is monotone); lo > hi: none
__device__ __forceinline__ void emia_nn_range(const EmiaNN& m, int s_lo, int s_hi, int* lo, int* hi) {
    int d0 = (int)floor((double)s_lo * m.scale) - 2;
    if (d0 < 0) d0 = 0;
    while (d0 > 0 && emia_nn_map(m, d0) >= s_lo) --d0;      // the estimate may already be inside the range
    while (d0 < m.dst && emia_nn_map(m, d0) < s_lo) ++d0;
    int d1 = (int)ceil((double)(s_hi + 1) * m.scale) + 2;
    if (d1 > m.dst - 1) d1 = m.dst - 1;
    while (d1 < m.dst - 1 && emia_nn_map(m, d1) <= s_hi) ++d1;
    while (d1 >= 0 && emia_nn_map(m, d1) > s_hi) --d1;
    *lo = d0; *hi = d1;
}

struct EmiaPlaceGeom {
    int dy_lo, dy_hi, dx_lo, dx_hi;   // unclipped extent of the resized mask in TILE coordinates (lo > hi: empty)
    int gy0, gy1, gx0, gx1;           // clipped extent in destination-frame coordinates, half-open (gy0 >= gy1: empty)
};
__device__ __forceinline__ EmiaPlaceGeom emia_place_geom(const int32_t* sb, const EmiaNN& my, const EmiaNN& mx, int ox, int oy,
                                                         int Hd, int Wd) {
    EmiaPlaceGeom g;
    g.dy_lo = g.dx_lo = 0; g.dy_hi = g.dx_hi = -1; g.gy0 = g.gy1 = g.gx0 = g.gx1 = 0;
    if (sb[0] < 0) return g;
    emia_nn_range(my, sb[0], sb[2], &g.dy_lo, &g.dy_hi);
    emia_nn_range(mx, sb[1], sb[3], &g.dx_lo, &g.dx_hi);
    if (g.dy_lo > g.dy_hi || g.dx_lo > g.dx_hi) { g.dy_lo = g.dx_lo = 0; g.dy_hi = g.dx_hi = -1; return g; }
    // global_mask[y_off : min(y_off + th, h), x_off : min(x_off + tw, w)] = downscaled[: y_end - y_off, : x_end - x_off]
    g.gy0 = oy + g.dy_lo; g.gy1 = min(oy + g.dy_hi + 1, min(oy + my.dst, Hd));
    g.gx0 = ox + g.dx_lo; g.gx1 = min(ox + g.dx_hi + 1, min(ox + mx.dst, Wd));
    if (g.gy0 >= g.gy1 || g.gx0 >= g.gx1) { g.gy0 = g.gy1 = g.gx0 = g.gx1 = 0; }
    return g;
}

__global__ void k_resize_place_plan(const int32_t* __restrict__ src_bbox, int64_t n, int Hs, int Ws, int th, int tw,
                                    const int32_t* __restrict__ off_xy, const int32_t* __restrict__ alive, int Hd, int Wd,
                                    emia_inst_meta* __restrict__ dst_meta, int64_t* __restrict__ dst_crop_words) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const EmiaNN my = emia_nn_make(Hs, th), mx = emia_nn_make(Ws, tw);
    const int ox = off_xy ? off_xy[2 * i] : 0, oy = off_xy ? off_xy[2 * i + 1] : 0;
    const EmiaPlaceGeom g = emia_place_geom(src_bbox + 4 * i, my, mx, ox, oy, Hd, Wd);
    emia_inst_meta m;
    m.valid = 1; m.reserved = 0;
    if (g.gy0 < g.gy1 && (!alive || alive[i])) {
        m.ry0 = g.gy0; m.ch = g.gy1 - g.gy0;
        m.rx0 = g.gx0; m.rx1 = g.gx1;
        m.wc0 = g.gx0 >> 5; m.cw = ((g.gx1 - 1) >> 5) - m.wc0 + 1;
    } else {
        m.ry0 = m.ch = m.rx0 = m.rx1 = m.wc0 = m.cw = 0;
    }
    dst_meta[i] = m;
    dst_crop_words[i] = (int64_t)m.ch * m.cw;
}

// one warp per instance
__global__ void __launch_bounds__(128) k_resize_nearest_place(
    const uint32_t* __restrict__ src_crops, const emia_inst_meta* __restrict__ src_meta, const int64_t* __restrict__ src_crop_off,
    const int32_t* __restrict__ src_bbox, int64_t n, int Hs, int Ws, int th, int tw, const int32_t* __restrict__ off_xy, int Hd,
    int Wd, int edge_width, int tile_size, const emia_inst_meta* __restrict__ dst_meta, const int64_t* __restrict__ dst_crop_off,
    uint32_t* __restrict__ dst_crops, int32_t* __restrict__ dst_bbox, int32_t* __restrict__ dst_area, int32_t* __restrict__ edge_flag,
    const int32_t* __restrict__ alive) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    if (alive && !alive[i]) {                 // not a member of any list any more: no pixels, flagged as an edge mask
        if (lane == 0) {
            dst_area[i] = 0;
            ((int4*)dst_bbox)[i] = make_int4(-1, -1, -1, -1);
            if (edge_flag) edge_flag[i] = 1;
        }
        return;
    }
    const EmiaNN my = emia_nn_make(Hs, th), mx = emia_nn_make(Ws, tw);
    const int ox = off_xy ? off_xy[2 * i] : 0, oy = off_xy ? off_xy[2 * i + 1] : 0;
    const EmiaPlaceGeom g = emia_place_geom(src_bbox + 4 * i, my, mx, ox, oy, Hd, Wd);
    const emia_inst_meta sm = src_meta[i];
    const emia_inst_meta dm = dst_meta[i];
    const uint32_t* sc = src_crops + src_crop_off[i];
    uint32_t* dc = dst_crops + dst_crop_off[i];
    int a = 0, ymin = 0x7fffffff, xmin = 0x7fffffff, ymax = -1, xmax = -1;        // placed (clipped) mask, frame coordinates
    int tymin = 0x7fffffff, txmin = 0x7fffffff, tymax = -1, txmax = -1;           // unclipped mask, tile coordinates
    if (g.dy_lo <= g.dy_hi) {
        const int w0 = (ox + g.dx_lo) >> 5, w1 = (ox + g.dx_hi) >> 5;             // destination word columns (unclipped)
        // word column by word column: the source column of a destination pixel depends on the column only, so its (double
        // precision) index map is evaluated once per lane and word column, not once per pixel
        for (int w = w0; w <= w1; ++w) {
            const int gx = w * 32 + lane;
            const int dx = gx - ox;
            const bool in_x = dx >= g.dx_lo && dx <= g.dx_hi;
            int c = -1, sh = 0;
            if (in_x) {
                const int sx = emia_nn_map(mx, dx);
                c = (sx >> 5) - sm.wc0;
                sh = sx & 31;
                if ((unsigned)c >= (unsigned)sm.cw) c = -1;
            }
            const bool stored_col = w >= dm.wc0 && w < dm.wc0 + dm.cw;
            uint32_t colmask = 0xffffffffu;
            const int lim = g.gx1 - w * 32;                                       // bits >= lim are outside the frame / tile
            if (lim < 32) colmask = (lim <= 0) ? 0u : ((1u << lim) - 1u);
            for (int dy = g.dy_lo; dy <= g.dy_hi; ++dy) {
                const int sy = emia_nn_map(my, dy) - sm.ry0;
                const int gy = oy + dy;
                bool bit = false;
                if (c >= 0 && (unsigned)sy < (unsigned)sm.ch) bit = (sc[(size_t)sy * sm.cw + c] >> sh) & 1u;
                const uint32_t word = __ballot_sync(0xffffffffu, bit);
                if (word) {
                    tymin = min(tymin, dy); tymax = max(tymax, dy);
                    txmin = min(txmin, w * 32 + (__ffs((int)word) - 1) - ox);
                    txmax = max(txmax, w * 32 + (31 - __clz((int)word)) - ox);
                }
                // stored part: inside the clipped extent
                if (stored_col && gy >= g.gy0 && gy < g.gy1) {
                    const uint32_t keep = word & colmask;
                    if (lane == 0) dc[(size_t)(gy - dm.ry0) * dm.cw + (w - dm.wc0)] = keep;
                    if (keep) {
                        a += __popc(keep);
                        ymin = min(ymin, gy); ymax = max(ymax, gy);
                        xmin = min(xmin, w * 32 + (__ffs((int)keep) - 1));
                        xmax = max(xmax, w * 32 + (31 - __clz((int)keep)));
                    }
                }
            }
        }
    }
    if (lane == 0) {
        dst_area[i] = a;
        ((int4*)dst_bbox)[i] = a > 0 ? make_int4(ymin, xmin, ymax, xmax) : make_int4(-1, -1, -1, -1);
        if (edge_flag) {
            int e = 1;                                                             // empty mask => edge (inference.py:2537-2538)
            if (tymax >= 0)
                e = (tymin < edge_width || tymax > tile_size - edge_width || txmin < edge_width || txmax > tile_size - edge_width) ? 1 : 0;
            edge_flag[i] = e;
        }
    }
}

extern "C" int emia_resize_place_plan(const int32_t* src_bbox, int64_t n, int Hs, int Ws, int th, int tw, const int32_t* off_xy,
                                      const int32_t* alive, int Hd, int Wd, emia_inst_meta* dst_meta, int64_t* dst_crop_words,
                                      void* stream) {
    if (n < 0 || Hs <= 0 || Ws <= 0 || th <= 0 || tw <= 0 || Hd <= 0 || Wd <= 0)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_resize_place_plan: %s", "bad shape");
    if (n == 0) return EMIA_OK;
    if (!src_bbox || !dst_meta || !dst_crop_words) return emia_fail(EMIA_ERR_BAD_ARG, "emia_resize_place_plan: %s", "null pointer");
    k_resize_place_plan<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(src_bbox, n, Hs, Ws, th, tw, off_xy, alive, Hd, Wd,
                                                                                    dst_meta, dst_crop_words);
    return emia_check_launch("emia_resize_place_plan launch: %s");
}
extern "C" int emia_resize_nearest_place(const uint32_t* src_crops, const emia_inst_meta* src_meta, const int64_t* src_crop_off,
                                         const int32_t* src_bbox, int64_t n, int Hs, int Ws, int th, int tw, const int32_t* off_xy,
                                         int Hd, int Wd, int edge_width, int tile_size, const emia_inst_meta* dst_meta,
                                         const int64_t* dst_crop_off, uint32_t* dst_crops, int32_t* dst_bbox, int32_t* dst_area,
                                         int32_t* edge_flag, const int32_t* alive, void* stream) {
    if (n < 0 || Hs <= 0 || Ws <= 0 || th <= 0 || tw <= 0 || Hd <= 0 || Wd <= 0)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_resize_nearest_place: %s", "bad shape");
    if (n == 0) return EMIA_OK;
    if (!src_crops || !src_meta || !src_crop_off || !src_bbox || !dst_meta || !dst_crop_off || !dst_crops || !dst_bbox || !dst_area)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_resize_nearest_place: %s", "null pointer");
    k_resize_nearest_place<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        src_crops, src_meta, src_crop_off, src_bbox, n, Hs, Ws, th, tw, off_xy, Hd, Wd, edge_width, tile_size, dst_meta, dst_crop_off,
        dst_crops, dst_bbox, dst_area, edge_flag, alive);
    return emia_check_launch("emia_resize_nearest_place launch: %s");
}

// list members whose flag equals keep_value, list order kept (edge filter of the tile pipeline, inference.py:2405-2407)
__global__ void k_group_filter_flag(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                    const int32_t* __restrict__ in_idx, const int32_t* __restrict__ flag, int keep_value,
                                    int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    const int base = cap_off[g], len = in_len[g];
    int run = 0;
    for (int k0 = 0; k0 < len; k0 += 32) {
        const int k = k0 + lane;
        const int inst = (k < len) ? in_idx[base + k] : 0;
        const int keep = (k < len) && flag[inst] == keep_value;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) out_idx[base + run + __popc(b & ((1u << lane) - 1u))] = inst;
        run += __popc(b);
    }
    if (lane == 0) out_len[g] = run;
}
extern "C" int emia_group_filter_flag(const int32_t* cap_off, int32_t G, const int32_t* in_len, const int32_t* in_idx,
                                      const int32_t* flag, int32_t keep_value, int32_t* out_len, int32_t* out_idx, void* stream) {
    if (G < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_group_filter_flag: %s", "bad G");
    if (G == 0) return EMIA_OK;
    if (!cap_off || !in_len || !in_idx || !flag || !out_len || !out_idx)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_group_filter_flag: %s", "null pointer");
    k_group_filter_flag<<<(unsigned)(((size_t)G * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(cap_off, G, in_len, in_idx, flag,
                                                                                                   keep_value, out_len, out_idx);
    return emia_check_launch("emia_group_filter_flag launch: %s");
}

// ---- K6: RLE wire format -----------------------------------------------------------------------------------------------
// rle_encoding(x): dots = where(x.T.flatten() == 1); runs of consecutive dots -> (start + 1, length), i.e. column-major runs;
// a run continues from the bottom of column c into the top of column c + 1 (consecutive flat indices).
// Pass 1 counts the runs of every instance, pass 2 writes (start, length) pairs; one warp per instance, lanes own columns.
// A run STARTS at flat index f = x * H + y when pixel (y, x) is set and the pixel at flat index f - 1 is not.
template <bool kStore>
__global__ void __launch_bounds__(128) k_rle(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                             const int64_t* __restrict__ crop_off, const int32_t* __restrict__ bbox, int64_t n, int H,
                                             int64_t* __restrict__ n_runs, const int64_t* __restrict__ run_off,
                                             int64_t* __restrict__ runs) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const int4 bb = ((const int4*)bbox)[i];
    if (bb.x < 0) { if (!kStore && lane == 0) n_runs[i] = 0; return; }
    const uint32_t* crop = crops + crop_off[i];
    auto px = [&](int y, int x) -> int {
        const int r = y - m.ry0, c = (x >> 5) - m.wc0;
        if ((unsigned)r >= (unsigned)m.ch || (unsigned)c >= (unsigned)m.cw) return 0;
        return (crop[(size_t)r * m.cw + c] >> (x & 31)) & 1u;
    };
    // columns x_min .. x_max, rows y_min .. y_max; the only cross-column continuation: (H-1, x-1) set and (0, x) set
    int64_t base = kStore ? run_off[i] : 0;
    int total = 0;
    for (int x0 = bb.y; x0 <= bb.w; x0 += 32) {
        const int x = x0 + lane;
        int cnt = 0;
        if (x <= bb.w) {
            int prev = (bb.x == 0 && x > 0) ? px(H - 1, x - 1) : 0;      // pixel at flat index f - 1 of the column's first row
            for (int y = bb.x; y <= bb.z; ++y) {
                const int cur = px(y, x);
                cnt += (cur && !prev);
                prev = cur;
            }
        }
        // exclusive scan of the counts over the lanes -> where this column's runs go
        int inc = cnt;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (kStore && x <= bb.w && cnt) {
            int64_t w = base + total + inc - cnt;
            // (start, length): the length may extend into following columns (only when the run reaches row H - 1)
            int prev = (bb.x == 0 && x > 0) ? px(H - 1, x - 1) : 0;
            for (int y = bb.x; y <= bb.z; ++y) {
                const int cur = px(y, x);
                if (cur && !prev) {
                    int64_t len = 0;
                    int yy = y, xx = x;
                    while (px(yy, xx)) {
                        ++len;
                        if (++yy == H) { yy = 0; ++xx; }
                    }
                    runs[2 * w] = (int64_t)x * H + y + 1;
                    runs[2 * w + 1] = len;
                    ++w;
                }
                prev = cur;
            }
        }
        total += tot;
    }
    if (!kStore && lane == 0) n_runs[i] = total;
}

extern "C" int emia_rle_count(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, const int32_t* bbox,
                              int64_t n, int H, int64_t* n_runs, void* stream) {
    if (n < 0 || H <= 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_rle_count: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !n_runs) return emia_fail(EMIA_ERR_BAD_ARG, "emia_rle_count: %s", "null pointer");
    k_rle<false><<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, bbox, n, H, n_runs, nullptr, nullptr);
    return emia_check_launch("emia_rle_count launch: %s");
}
extern "C" int emia_rle_encode(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, const int32_t* bbox,
                               int64_t n, int H, const int64_t* run_off, int64_t* runs, void* stream) {
    if (n < 0 || H <= 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_rle_encode: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !run_off || !runs) return emia_fail(EMIA_ERR_BAD_ARG, "emia_rle_encode: %s", "null pointer");
    k_rle<true><<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, bbox, n, H, nullptr, run_off, runs);
    return emia_check_launch("emia_rle_encode launch: %s");
}

// ---- raw image moments of a 0/1 mask up to order 1 (cv2.moments(mask)["m00" | "m10" | "m01"]): exact integer sums ------------
__global__ void __launch_bounds__(128) k_moments01(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                                   const int64_t* __restrict__ crop_off, int64_t n, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const uint32_t* crop = crops + crop_off[i];
    long long m00 = 0, m10 = 0, m01 = 0;
    for (int k = lane; k < m.ch * m.cw; k += 32) {
        uint32_t w = crop[k];
        if (!w) continue;
        const int r = k / m.cw, c = k - r * m.cw;
        const int y = m.ry0 + r, xb = (m.wc0 + c) * 32;
        const int pc = __popc(w);
        m00 += pc; m01 += (long long)pc * y;
        while (w) { const int b = __ffs((int)w) - 1; m10 += xb + b; w &= w - 1; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        m00 += __shfl_xor_sync(0xffffffffu, m00, o); m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    if (lane == 0) { out[3 * i] = m00; out[3 * i + 1] = m10; out[3 * i + 2] = m01; }
}
extern "C" int emia_moments01(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n, int64_t* out,
                              void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_moments01: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !out) return emia_fail(EMIA_ERR_BAD_ARG, "emia_moments01: %s", "null pointer");
    k_moments01<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, n, out);
    return emia_check_launch("emia_moments01 launch: %s");
}

// ---- all image moments up to order 3 (cv2.moments: raw m_pq exact, central mu_pq and normalised nu_pq as OpenCV computes them) ----
__global__ void __launch_bounds__(128) k_moments(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                                 const int64_t* __restrict__ crop_off, int64_t n, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const uint32_t* crop = crops + crop_off[i];
    long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < m.ch * m.cw; k += 32) {
        const int r = k / m.cw, c = k - r * m.cw;
        emia_word_moments(crop[k], (m.wc0 + c) * 32, m.ry0 + r, acc);
    }
    for (int q = 0; q < 10; ++q)
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    if (lane == 0) emia_complete_moments(acc, out + i * EMIA_MOMENT_FIELDS);
}
extern "C" int emia_moments(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n, double* out,
                            void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_moments: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !out) return emia_fail(EMIA_ERR_BAD_ARG, "emia_moments: %s", "null pointer");
    k_moments<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, n, out);
    return emia_check_launch("emia_moments launch: %s");
}

// ---- colour sums of the image pixels under every instance: out[i] = (sum B, sum G, sum R, pixel count) as exact integers ----------
// (mean colour -> rgb_to_wavelength, src/utils/measurements.py:32-111; the README's "Wavelength_nm" that no reference code writes, Q9)
__global__ void __launch_bounds__(128) k_color_sums(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                                    const int64_t* __restrict__ crop_off, int64_t n, const uint8_t* __restrict__ image,
                                                    int H, int W, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const uint32_t* crop = crops + crop_off[i];
    long long sb = 0, sg = 0, sr = 0, cnt = 0;
    for (int k = lane; k < m.ch * m.cw; k += 32) {
        uint32_t w = crop[k];
        const int r = k / m.cw, c = k - r * m.cw;
        const int y = m.ry0 + r;
        while (w) {
            const int b = __ffs((int)w) - 1;
            w &= w - 1;
            const int x = (m.wc0 + c) * 32 + b;
            if (x >= W || y >= H) continue;
            const uint8_t* px = image + ((size_t)y * W + x) * 3;
            sb += px[0]; sg += px[1]; sr += px[2]; cnt += 1;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        sb += __shfl_xor_sync(0xffffffffu, sb, o); sg += __shfl_xor_sync(0xffffffffu, sg, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { out[4 * i] = sb; out[4 * i + 1] = sg; out[4 * i + 2] = sr; out[4 * i + 3] = cnt; }
}
extern "C" int emia_color_sums(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n,
                               const uint8_t* image_bgr, int H, int W, int64_t* out, void* stream) {
    if (n < 0 || H <= 0 || W <= 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_color_sums: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !image_bgr || !out) return emia_fail(EMIA_ERR_BAD_ARG, "emia_color_sums: %s", "null pointer");
    k_color_sums<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, n, image_bgr, H, W, out);
    return emia_check_launch("emia_color_sums launch: %s");
}

// ---- visualisation overlay (src/functions/inference.py:1080-1145) ---------------------------------------------------------------
// Per mask, IN LIST ORDER (later masks blend over earlier ones): colored_mask = color where mask; vis = cv2.addWeighted(vis, 1.0,
// colored_mask, 0.5, 0) — i.e. every mask pixel becomes saturate(rint(v + 0.5 * color)) (round half to even), everything else is
// unchanged — then cv2.drawContours(vis, contours, -1, color, 1): with CHAIN_APPROX_SIMPLE contours every segment is an axial or
// diagonal run, so the drawn pixels are exactly the border-following chain, which is replayed here from the stored vertices.
// One CTA walks the masks in order (the order is the semantics); its threads split the pixels / segments of the current mask.
__global__ void __launch_bounds__(1024) k_overlay(uint8_t* __restrict__ image, int H, int W, const uint32_t* __restrict__ crops,
                                                  const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off,
                                                  const int32_t* __restrict__ order, int n, const int32_t* __restrict__ classes,
                                                  const uint8_t* __restrict__ colors, int ncolors, const uint32_t* __restrict__ pts,
                                                  const int64_t* __restrict__ pt_off, const int32_t* __restrict__ cstart, int cstart_stride,
                                                  const int64_t* __restrict__ inst_cont_off, const int64_t* __restrict__ n_contours) {
    for (int k = 0; k < n; ++k) {
        const int i = order ? order[k] : k;
        const emia_inst_meta m = meta[i];
        const int cls = classes[i];
        const uint8_t* col = colors + 3 * (((cls % ncolors) + ncolors) % ncolors);
        const float cb = 0.5f * col[0], cg = 0.5f * col[1], cr = 0.5f * col[2];
        const uint32_t* crop = crops + crop_off[i];
        const int npx = m.ch * m.cw * 32;
        for (int p = threadIdx.x; p < npx; p += blockDim.x) {
            const int wi = p >> 5, b = p & 31;
            if (!((crop[wi] >> b) & 1u)) continue;
            const int r = wi / m.cw, c = wi - r * m.cw;
            const int y = m.ry0 + r, x = (m.wc0 + c) * 32 + b;
            if (x >= W || y >= H) continue;
            uint8_t* px = image + ((size_t)y * W + x) * 3;
            px[0] = (uint8_t)min(255.0f, rintf((float)px[0] + cb));
            px[1] = (uint8_t)min(255.0f, rintf((float)px[1] + cg));
            px[2] = (uint8_t)min(255.0f, rintf((float)px[2] + cr));
        }
        __syncthreads();
        const int nc = (int)n_contours[i];
        const int32_t* cs = cstart + (cstart_stride ? (size_t)i * cstart_stride : (size_t)(inst_cont_off[i] + i));
        const uint32_t* pv = pts + pt_off[i];
        for (int j = 0; j < nc; ++j) {
            const int a = cs[j], mlen = cs[j + 1] - cs[j];
            for (int sgm = threadIdx.x; sgm < mlen; sgm += blockDim.x) {
                const uint32_t v0 = pv[a + sgm], v1 = pv[a + (sgm + 1 == mlen ? 0 : sgm + 1)];
                int x = (int)(v0 & 0xFFFFu), y = (int)(v0 >> 16);
                const int x1 = (int)(v1 & 0xFFFFu), y1 = (int)(v1 >> 16);
                const int sx = (x1 > x) - (x1 < x), sy = (y1 > y) - (y1 < y);
                const int steps = max(abs(x1 - x), abs(y1 - y));
                for (int t = 0; t <= steps; ++t) {
                    if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
                        uint8_t* px = image + ((size_t)y * W + x) * 3;
                        px[0] = col[0]; px[1] = col[1]; px[2] = col[2];
                    }
                    x += sx; y += sy;
                }
            }
        }
        __syncthreads();
    }
}
extern "C" int emia_overlay(uint8_t* image_bgr, int H, int W, const uint32_t* crops, const emia_inst_meta* meta,
                            const int64_t* crop_off, const int32_t* order, int64_t n, const int32_t* classes,
                            const uint8_t* colors_bgr, int32_t n_colors, const uint32_t* pts, const int64_t* pt_off,
                            const int32_t* cstart, int32_t cstart_stride, const int64_t* inst_cont_off, const int64_t* n_contours,
                            void* stream) {
    if (n < 0 || H <= 0 || W <= 0 || n_colors <= 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_overlay: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!image_bgr || !crops || !meta || !crop_off || !classes || !colors_bgr || !pts || !pt_off || !cstart || !n_contours ||
        (cstart_stride == 0 && !inst_cont_off))
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_overlay: %s", "null pointer");
    k_overlay<<<1, 1024, 0, (cudaStream_t)stream>>>(image_bgr, H, W, crops, meta, crop_off, order, (int)n, classes, colors_bgr, n_colors, pts,
                                                    pt_off, cstart, cstart_stride, inst_cont_off, n_contours);
    return emia_check_launch("emia_overlay launch: %s");
}

// ---- masked grey-level histogram (contrast d10 / d50 / d90, src/utils/measurements.py:195-215) ---------------------------
// gray = cv2.cvtColor(BGR2GRAY) for 8-bit images: (B * 3735 + G * 19235 + R * 9798 + 16384) >> 15 (OpenCV 4.13; verified
// exhaustively against cv2 on a 3-step colour grid); np.histogram(bins = 256,
// range = (0, 255)) puts the integer level v into bin v.  One CTA per instance, 256-bin histogram in shared memory.
__global__ void __launch_bounds__(128) k_gray_hist(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                                   const int64_t* __restrict__ crop_off, int64_t n, const uint8_t* __restrict__ image,
                                                   int H, int W, int channels, int32_t* __restrict__ hist) {
    __shared__ int s_h[256];
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        for (int k = threadIdx.x; k < 256; k += blockDim.x) s_h[k] = 0;
        __syncthreads();
        const emia_inst_meta m = meta[i];
        const uint32_t* crop = crops + crop_off[i];
        for (int k = threadIdx.x; k < m.ch * m.cw; k += blockDim.x) {
            uint32_t w = crop[k];
            const int r = k / m.cw, c = k - r * m.cw;
            const int y = m.ry0 + r;
            while (w) {
                const int b = __ffs((int)w) - 1;
                w &= w - 1;
                const int x = (m.wc0 + c) * 32 + b;
                if (x >= W || y >= H) continue;
                const uint8_t* px = image + ((size_t)y * W + x) * channels;
                const int g = (channels == 3) ? ((px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + 16384) >> 15) : px[0];
                atomicAdd(&s_h[g], 1);
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[i * 256 + k] = s_h[k];
        __syncthreads();
    }
}
extern "C" int emia_gray_hist(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n,
                              const uint8_t* image, int H, int W, int channels, int32_t* hist, void* stream) {
    if (n < 0 || H <= 0 || W <= 0 || (channels != 1 && channels != 3)) return emia_fail(EMIA_ERR_BAD_ARG, "emia_gray_hist: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !image || !hist) return emia_fail(EMIA_ERR_BAD_ARG, "emia_gray_hist: %s", "null pointer");
    k_gray_hist<<<(unsigned)(n < 65535 ? n : 65535), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, n, image, H, W, channels, hist);
    return emia_check_launch("emia_gray_hist launch: %s");
}

// ---- grey-level histogram of a whole image (calculate_image_quality_score, src/functions/inference.py:256-283: brightness =
// mean(gray) / 255, contrast = std(gray) / 128 — both follow exactly from the 256 integer counts) ------------------------------
__global__ void __launch_bounds__(256) k_image_gray_hist(const uint8_t* __restrict__ image, int64_t npix, int channels,
                                                         unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t* px = image + p * channels;
        const int g = (channels == 3) ? ((px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + 16384) >> 15) : px[0];
        atomicAdd(&s_h[g], 1u);
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)s_h[threadIdx.x]);
}
extern "C" int emia_image_gray_hist(const uint8_t* image, int H, int W, int channels, uint64_t* hist, void* stream) {
    if (H <= 0 || W <= 0 || (channels != 1 && channels != 3)) return emia_fail(EMIA_ERR_BAD_ARG, "emia_image_gray_hist: %s", "bad argument");
    if (!image || !hist) return emia_fail(EMIA_ERR_BAD_ARG, "emia_image_gray_hist: %s", "null pointer");
    cudaMemsetAsync(hist, 0, 256 * sizeof(uint64_t), (cudaStream_t)stream);
    const int64_t npix = (int64_t)H * W;
    int64_t blocks = (npix + 256 * 16 - 1) / (256 * 16);
    const int64_t cap = (int64_t)emia_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_image_gray_hist<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(image, npix, channels, (unsigned long long*)hist);
    return emia_check_launch("emia_image_gray_hist launch: %s");
}
