// emia_morpho_kernels.cuh — K5: external contours + morphometry (part of emia_kernels.cu).
// One thread per instance: border following is inherently sequential per component, so parallelism comes from
// the number of instances in flight (10^5..10^6 per launch); the crop (a few hundred bytes) stays in L1.
#pragma once

__device__ __forceinline__ EmiaBitView emia_make_view(const uint32_t* crops, const emia_inst_meta& m, int64_t off) {
    EmiaBitView v;
    v.bits = crops + off; v.pitch_words = m.cw; v.h = m.ch; v.wwords = m.cw;
    v.x_origin = m.wc0 * 32; v.y_origin = m.ry0;
    return v;
}

// ---- border following: one THREAD per instance ---------------------------------------------------------------------
#define EMIA_TRACE_THREADS 128

template <bool kStore>
__global__ void __launch_bounds__(EMIA_TRACE_THREADS) k_contour_trace(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off, int64_t n,
    uint32_t* __restrict__ marks, int64_t* __restrict__ n_contours, int64_t* __restrict__ n_points,
    int64_t* __restrict__ scratch_bytes, const int64_t* __restrict__ cont_off, const int64_t* __restrict__ pt_off,
    uint32_t* __restrict__ pts, int32_t* __restrict__ cstart) {
    const int64_t i = (int64_t)blockIdx.x * EMIA_TRACE_THREADS + threadIdx.x;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const int words = m.ch * m.cw;
    if (words <= 0) {
        if (kStore) cstart[cont_off[i] + i] = 0;
        else { n_contours[i] = 0; n_points[i] = 0; scratch_bytes[i] = 0; }
        return;
    }
    const int64_t off = crop_off[i];
    const EmiaBitView v = emia_make_view(crops, m, off);
    uint32_t* mk = marks + 2 * off;
    uint32_t* ng = mk + words;
    EmiaContourOut o;
    if (kStore) {
        o.pts = pts + pt_off[i]; o.cap_pts = (int)(pt_off[i + 1] - pt_off[i]);
        o.cstart = cstart + cont_off[i] + i; o.cap_contours = (int)(cont_off[i + 1] - cont_off[i]); o.store = 1;
    } else {
        o.pts = nullptr; o.cap_pts = 0; o.cstart = nullptr; o.cap_contours = 0; o.store = 0;
    }
    emia_find_external_contours(v, mk, ng, o);
    if (!kStore) {
        n_contours[i] = o.n_contours;
        n_points[i] = o.n_pts;
        scratch_bytes[i] = o.n_contours ? (int64_t)((emia_measure_scratch_bytes(o.max_len) + 15) & ~(size_t)15) : 0;
    }
}
#define EMIA_TRACE_SMEM_BYTES ((size_t)0)

// ---- hull pre-sort: one WARP per single-contour instance --------------------------------------------------------------
// The convex hull starts from the vertices sorted by (x, y, index).  A per-thread comparison sort in global memory is
// dominated by uncoalesced accesses (ncu/ablation: 83 % of the morphometry kernel), so the keys of the common case
// (one contour, <= EMIA_PRESORT_MAX vertices) are sorted beforehand by a bitonic network in shared memory, one warp per
// instance, and written where the hull expects them (the head of the instance's scratch block).
#define EMIA_PRESORT_MAX 256
#define EMIA_PRESORT_WARPS 8
__global__ void __launch_bounds__(EMIA_PRESORT_WARPS * 32) k_contour_presort(int64_t n, const int64_t* __restrict__ cont_off,
                                                                            const int64_t* __restrict__ pt_off,
                                                                            const int32_t* __restrict__ cstart, int cstart_stride,
                                                                            const int64_t* __restrict__ scratch_off,
                                                                            const uint32_t* __restrict__ pts, uint8_t* __restrict__ scratch) {
    __shared__ uint64_t s_keys[EMIA_PRESORT_WARPS][EMIA_PRESORT_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * EMIA_PRESORT_WARPS + warp;
    if (i >= n) return;
    const int64_t c0 = cont_off[i];
    if (cont_off[i + 1] - c0 != 1) return;
    const int32_t* cs = cstart_stride > 0 ? cstart + (size_t)i * cstart_stride : cstart + c0 + i;
    const int len = cs[1] - cs[0];
    if (len > EMIA_PRESORT_MAX || len < 2) {
        if (len == 1 && lane == 0) {
            const uint32_t p0 = pts[pt_off[i] + cs[0]];
            ((uint64_t*)(scratch + scratch_off[i]))[0] = EMIA_KEY(EMIA_PT_X(p0), EMIA_PT_Y(p0), 0);
        }
        return;
    }
    int N = 32;
    while (N < len) N <<= 1;
    uint64_t* k = s_keys[warp];
    const uint32_t* p = pts + pt_off[i] + cs[0];
    for (int t = lane; t < N; t += 32) {
        uint64_t key = ~0ull;
        if (t < len) { const uint32_t q = p[t]; key = EMIA_KEY(EMIA_PT_X(q), EMIA_PT_Y(q), t); }
        k[t] = key;
    }
    __syncwarp();
    for (int size = 2; size <= N; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (N >> 1); t += 32) {
                const int lo = ((t / stride) * (stride << 1)) + (t % stride);
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
            __syncwarp();
        }
    }
    uint64_t* dst = (uint64_t*)(scratch + scratch_off[i]);
    for (int t = lane; t < len; t += 32) dst[t] = k[t];
}

// ---- morphometry: one THREAD per instance over the stored vertex lists ---------------------------------------------
__global__ void __launch_bounds__(128) k_contour_measure(const emia_inst_meta* __restrict__ meta, int64_t n,
                                                         const int64_t* __restrict__ cont_off, const int64_t* __restrict__ pt_off,
                                                         const int64_t* __restrict__ scratch_off, double um_pix, double min_area,
                                                         const uint32_t* __restrict__ pts, const int32_t* __restrict__ cstart,
                                                         int cstart_stride, int presort_max, double* __restrict__ records,
                                                         int32_t* __restrict__ rec_inst, double* __restrict__ perim0,
                                                         uint8_t* __restrict__ scratch) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c0 = cont_off[i];
    const int nc = (int)(cont_off[i + 1] - c0);
    if (nc == 0) { perim0[i] = 0.0; return; }
    const int32_t* cs = cstart_stride > 0 ? cstart + (size_t)i * cstart_stride : cstart + c0 + i;
    const uint32_t* p = pts + pt_off[i];
    void* sc = scratch + scratch_off[i];
    for (int j = 0; j < nc; ++j) {
        const int k = nc - 1 - j;                     // OpenCV returns contours in reverse discovery order
        const uint32_t* cp = p + cs[k];
        const int len = cs[k + 1] - cs[k];
        double* rec = records + (size_t)(c0 + j) * EMIA_REC_FIELDS;
        emia_measure_contour(cp, len, um_pix, sc, rec, (nc == 1 && len <= presort_max) ? 1 : 0);
        rec[EMIA_REC_MEASURED] = (rec[EMIA_REC_AREA] >= min_area) ? 1.0 : 0.0;
        rec_inst[c0 + j] = (int32_t)i;
        if (j == 0) perim0[i] = rec[EMIA_REC_PERIMETER];
    }
}

extern "C" int emia_contour_count(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n,
                                  uint32_t* marks, int64_t* n_contours, int64_t* n_points, int64_t* scratch_bytes,
                                  void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_count: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !marks || !n_contours || !n_points || !scratch_bytes)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_count: %s", "null pointer");
    const unsigned grid = (unsigned)((n + EMIA_TRACE_THREADS - 1) / EMIA_TRACE_THREADS);
    k_contour_trace<false><<<grid, EMIA_TRACE_THREADS, EMIA_TRACE_SMEM_BYTES, (cudaStream_t)stream>>>(crops, meta, crop_off, n, marks, n_contours, n_points,
                                                                                  scratch_bytes, nullptr, nullptr, nullptr, nullptr);
    return emia_check_launch("emia_contour_count launch: %s");
}

extern "C" int emia_contour_measure(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n,
                                    uint32_t* marks, const int64_t* cont_off, const int64_t* pt_off,
                                    const int64_t* scratch_off, double um_pix, double min_area, uint32_t* pts,
                                    int32_t* cstart, double* records, int32_t* rec_inst, double* perim0,
                                    uint8_t* scratch, void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_measure: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !marks || !cont_off || !pt_off || !scratch_off || !pts || !cstart || !records ||
        !rec_inst || !perim0 || !scratch)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_measure: %s", "null pointer");
    const unsigned gridt = (unsigned)((n + EMIA_TRACE_THREADS - 1) / EMIA_TRACE_THREADS);
    k_contour_trace<true><<<gridt, EMIA_TRACE_THREADS, EMIA_TRACE_SMEM_BYTES, (cudaStream_t)stream>>>(crops, meta, crop_off, n, marks, nullptr, nullptr, nullptr,
                                                                                  cont_off, pt_off, pts, cstart);
    const unsigned grid = (unsigned)((n + 127) / 128);
    k_contour_presort<<<(unsigned)((n + EMIA_PRESORT_WARPS - 1) / EMIA_PRESORT_WARPS), EMIA_PRESORT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n, cont_off, pt_off, cstart, 0, scratch_off, pts, scratch);
    k_contour_measure<<<grid, 128, 0, (cudaStream_t)stream>>>(meta, n, cont_off, pt_off, scratch_off, um_pix, min_area, pts, cstart, 0,
                                                              EMIA_PRESORT_MAX, records, rec_inst, perim0, scratch);
    return emia_check_launch("emia_contour_measure launch: %s");
}


// ---- single-pass variant: vertices go to per-instance slabs of bounded capacity -------------------------------------
// cap_pts(i) = 4 * (ch + 32 * cw) + 32 (a blob's border is about 2 * (h + w) pixels); cstart slab of capc + 1 entries.
// Instances that exceed either capacity raise *overflow; the caller then re-runs the exact two-pass path.
__global__ void k_trace_caps(const emia_inst_meta* __restrict__ meta, int64_t n, int64_t* __restrict__ cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    cap[i] = (m.ch > 0 && m.cw > 0) ? (int64_t)(4 * (m.ch + 32 * m.cw) + 32) : 0;
}

__global__ void __launch_bounds__(EMIA_TRACE_THREADS) k_contour_trace_slab(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off, int64_t n,
    uint32_t* __restrict__ marks, const int64_t* __restrict__ pt_cap_off, int capc, uint32_t* __restrict__ pts,
    int32_t* __restrict__ cstart_slab, int64_t* __restrict__ n_contours, int64_t* __restrict__ scratch_bytes,
    int32_t* __restrict__ overflow) {
    const int64_t i = (int64_t)blockIdx.x * EMIA_TRACE_THREADS + threadIdx.x;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const int words = m.ch * m.cw;
    int32_t* cs = cstart_slab + (size_t)i * (capc + 1);
    if (words <= 0) { cs[0] = 0; n_contours[i] = 0; scratch_bytes[i] = 0; return; }
    const int64_t off = crop_off[i];
    const EmiaBitView v = emia_make_view(crops, m, off);
    uint32_t* mk = marks + 2 * off;
    uint32_t* ng = mk + words;
    EmiaContourOut o;
    o.pts = pts + pt_cap_off[i]; o.cap_pts = (int)(pt_cap_off[i + 1] - pt_cap_off[i]);
    o.cstart = cs; o.cap_contours = capc; o.store = 1;
    emia_find_external_contours(v, mk, ng, o);
    if (o.overflow) { atomicAdd(overflow, 1); n_contours[i] = 0; scratch_bytes[i] = 0; return; }
    n_contours[i] = o.n_contours;
    scratch_bytes[i] = o.n_contours ? (int64_t)((emia_measure_scratch_bytes(o.max_len) + 15) & ~(size_t)15) : 0;
}

extern "C" int emia_contour_trace_plan(const emia_inst_meta* meta, int64_t n, int64_t* pt_cap, void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_trace_plan: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!meta || !pt_cap) return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_trace_plan: %s", "null pointer");
    k_trace_caps<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(meta, n, pt_cap);
    return emia_check_launch("emia_contour_trace_plan launch: %s");
}

extern "C" int emia_contour_trace_slab(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n,
                                       uint32_t* marks, const int64_t* pt_cap_off, int32_t cap_contours, uint32_t* pts,
                                       int32_t* cstart_slab, int64_t* n_contours, int64_t* scratch_bytes, int32_t* overflow,
                                       void* stream) {
    if (n < 0 || cap_contours < 1) return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_trace_slab: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !marks || !pt_cap_off || !pts || !cstart_slab || !n_contours || !scratch_bytes || !overflow)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_trace_slab: %s", "null pointer");
    const unsigned grid = (unsigned)((n + EMIA_TRACE_THREADS - 1) / EMIA_TRACE_THREADS);
    k_contour_trace_slab<<<grid, EMIA_TRACE_THREADS, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, n, marks, pt_cap_off, cap_contours, pts,
                                                                             cstart_slab, n_contours, scratch_bytes, overflow);
    return emia_check_launch("emia_contour_trace_slab launch: %s");
}

extern "C" int emia_contour_measure_stored(const emia_inst_meta* meta, int64_t n, const int64_t* cont_off, const int64_t* pt_off,
                                           const int32_t* cstart, int32_t cstart_stride, const int64_t* scratch_off, double um_pix,
                                           double min_area, const uint32_t* pts, double* records, int32_t* rec_inst, double* perim0,
                                           uint8_t* scratch, void* stream) {
    if (n < 0 || cstart_stride < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_measure_stored: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!meta || !cont_off || !pt_off || !cstart || !scratch_off || !pts || !records || !rec_inst || !perim0 || !scratch)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_contour_measure_stored: %s", "null pointer");
    k_contour_presort<<<(unsigned)((n + EMIA_PRESORT_WARPS - 1) / EMIA_PRESORT_WARPS), EMIA_PRESORT_WARPS * 32, 0, (cudaStream_t)stream>>>(
        n, cont_off, pt_off, cstart, cstart_stride, scratch_off, pts, (uint8_t*)scratch);
    k_contour_measure<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(meta, n, cont_off, pt_off, scratch_off, um_pix, min_area, pts,
                                                                                  cstart, cstart_stride, EMIA_PRESORT_MAX, records, rec_inst, perim0, scratch);
    return emia_check_launch("emia_contour_measure_stored launch: %s");
}
