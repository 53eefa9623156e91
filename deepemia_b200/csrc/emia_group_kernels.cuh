// emia_group_kernels.cuh — K4: mask-IoU de-duplication and spatial constraints over G groups (part of emia_kernels.cu).
//
// Every operation has the same shape: (1) decide which list slots take part, (2) rank them into the order the
// reference visits them, (3) evaluate all candidate pairs in parallel into a per-group bit matrix of
// "a suppresses b" relations (in rank space), (4) replay the reference's sequential greedy loop with one warp per
// group on that bit matrix (32 ranks per word, one OR per word), (5) write the surviving list.
#pragma once

static size_t emia_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ int emia_find_group(const int32_t* __restrict__ cap_off, int G, int slot) {
    int lo = 0, hi = G;   // largest g with cap_off[g] <= slot
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cap_off[mid] <= slot) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ EmiaCropRef emia_crop_ref(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int inst) {
    const emia_inst_meta m = meta[inst];
    EmiaCropRef r;
    r.w = crops + crop_off[inst]; r.ry0 = m.ry0; r.wc0 = m.wc0; r.ch = m.ch; r.cw = m.cw;
    return r;
}

// ---- (1) participation -----------------------------------------------------------------------------------------
// mode 0: deduplicate_masks_smart pre-filter; mode 1: every live slot
__global__ void k_group_select(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                               const int32_t* __restrict__ in_idx, int L, int mode, const int32_t* __restrict__ bbox,
                               const int32_t* __restrict__ area, const double* __restrict__ perim0,
                               const int64_t* __restrict__ n_contours, double max_aspect_ratio, int32_t* __restrict__ ok) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    const int g = emia_find_group(cap_off, G, s);
    int v = 0;
    if (s - cap_off[g] < in_len[g]) {
        v = 1;
        if (mode == 0) {
            const int inst = in_idx[s];
            const int a = area[inst];
            if (a <= 0) v = 0;   // empty mask
            else {
                const int* b = bbox + 4 * inst;
                const int bw = b[3] - b[1] + 1, bh = b[2] - b[0] + 1;
                if (max_aspect_ratio > 0.0) {
                    const double asp = (double)max(bw, bh) / (double)min(bw, bh);
                    if (asp > max_aspect_ratio) v = 0;
                }
                if (v && n_contours[inst] > 0) {
                    const double per = perim0[inst];
                    if (per > 0) {
                        const double compactness = ((4 * M_PI) * (double)a) / (per * per);
                        if (compactness < 0.15) v = 0;
                    }
                }
            }
        }
    }
    ok[s] = v;
}

// ---- containment ---------------------------------------------------------------------------------------------------
// filter_by_containment_rules, one rule (child_class -> parent_class) at a time; rem[] persists across rules.
// child != parent: k_containment_fused (groups <= 1024 slots) or the sparse sweep of emia_group_sparse.cuh;
// child == parent is evaluated by a single thread per group (sequential semantics of the reference).
__global__ void k_containment_rule_selfclass(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                             const int64_t* __restrict__ crop_off, const int32_t* __restrict__ bbox,
                                             const int32_t* __restrict__ area, const int32_t* __restrict__ classes,
                                             const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                             const int32_t* __restrict__ in_idx, int cls, double thr, int32_t* __restrict__ rem) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int base = cap_off[g], len = in_len[g];
    // parents = instances of the class that were alive when the rule started and have a bbox (snapshot semantics
    // of the reference's `parent_centroids` list), skipped later if removed in the meantime
    for (int k = 0; k < len; ++k) {
        const int s = base + k;
        const int ic = in_idx[s];
        if (classes[ic] != cls || rem[s]) continue;
        const int ac = area[ic];
        const int* bc = bbox + 4 * ic;
        if (ac <= 0 || bc[0] < 0) { rem[s] = 1; continue; }
        const EmiaCropRef rc = emia_crop_ref(crops, meta, crop_off, ic);
        double best = 0.0;
        for (int j = 0; j < len; ++j) {
            const int t = base + j;
            const int ip = in_idx[t];
            if (classes[ip] != cls || rem[t]) continue;
            const int* bp = bbox + 4 * ip;
            if (!emia_bbox_overlap(bc, bp)) continue;
            const EmiaCropRef rp = emia_crop_ref(crops, meta, crop_off, ip);
            const double c = (double)emia_crop_inter(rc, rp) / (double)ac;
            if (c > best) best = c;
        }
        if (best < thr) rem[s] = 1;
    }
}

// survivors (rem == 0) in list order, one CTA per group (block-wide ballot scan: a single list of a whole micrograph has
// tens of thousands of slots)
__global__ void __launch_bounds__(1024) k_group_compact(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                                        const int32_t* __restrict__ in_idx, const int32_t* __restrict__ rem,
                                                        int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = cap_off[g], len = in_len[g];
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int k0 = 0; k0 < len; k0 += 1024) {
        const int k = k0 + threadIdx.x;
        const int keep = (k < len) && !rem[base + k];
        const int inst = (k < len) ? in_idx[base + k] : 0;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(b);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (keep) out_idx[base + before + __popc(b & ((1u << lane) - 1u))] = inst;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_w[w]; s_run += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_len[g] = s_run;
}

#include "emia_group_fused.cuh"
#include "emia_group_sparse.cuh"

extern "C" size_t emia_group_workspace_bytes(const int32_t* cap_off_host, int32_t G) {
    if (!cap_off_host || G < 0) return 0;
    return emia_sparse_ws_bytes((size_t)cap_off_host[G], (size_t)G);
}

// ---- host-side drivers ---------------------------------------------------------------------------------------------
static int emia_group_pipeline(int mode, const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                               const int32_t* bbox, const int32_t* area, const double* perim0, const int64_t* n_contours,
                               const float* scores, const int32_t* classes, const int32_t* cap_off, int32_t G,
                               const int32_t* in_len, const int32_t* in_idx, double thr, double max_aspect,
                               const int32_t* rule_active, const double* rule_max_iou, int num_classes, int32_t* out_len,
                               int32_t* out_idx, void* workspace, size_t workspace_bytes, cudaStream_t st, int L, int max_cap) {
    if (max_cap > 0 && max_cap <= EMIA_FUSED_MAX_CAP) {
        // every group fits one SM's shared memory: one CTA per group, no global workspace
        const size_t smem = emia_fused_smem_bytes(max_cap);
        emia_need_dyn_smem((const void*)k_group_fused, smem);
        k_group_fused<<<(unsigned)G, EMIA_FUSED_THREADS, smem, st>>>(crops, meta, crop_off, bbox, area, perim0, n_contours, scores, classes,
                                                                    cap_off, in_len, in_idx, mode, thr, max_aspect, rule_active, rule_max_iou,
                                                                    num_classes, max_cap, out_len, out_idx);
        return emia_check_launch("group op (fused) launch: %s");
    }
    // groups beyond one SM's shared memory (or max_cap unknown): the sparse path of emia_group_sparse.cuh
    EmiaSparseWs ws;
    if (emia_sparse_carve(workspace, workspace_bytes, (size_t)L, (size_t)G, &ws) != 0)
        return emia_fail(EMIA_ERR_WORKSPACE, "group op: %s", "workspace too small");
    cudaMemsetAsync(ws.nok, 0, (size_t)((unsigned char*)ws.edges - (unsigned char*)ws.nok), st);     // nok, nx, ecount, anyp
    const int T = 128;
    const unsigned gl = (unsigned)((L + T - 1) / T);
    const int rank_mode = (mode == 3) ? 0 : mode, pair_mode = (mode == 3) ? 1 : mode;
    k_group_select<<<gl, T, 0, st>>>(cap_off, G, in_len, in_idx, L, mode == 0 ? 0 : 1, bbox, area, perim0, n_contours, max_aspect, ws.ok);
    k_sp_fidx<<<(unsigned)G, 1024, 0, st>>>(cap_off, ws.ok, ws.fidx, ws.nok);
    k_sp_keys<<<gl, T, 0, st>>>(cap_off, G, in_idx, L, rank_mode, pair_mode == 2 ? 2 : 0, scores, classes, bbox, area, rule_active,
                                num_classes, ws.ok, ws.fidx, ws.k1, ws.k2, ws.kx, ws.nx);
    const unsigned gr = (unsigned)((L + EMIA_SP_THREADS - 1) / EMIA_SP_THREADS);
    const int big = max_cap > 0 ? max_cap : L;
    const unsigned gy = (unsigned)std::min(16, std::max(1, (big + EMIA_SP_TILE - 1) / EMIA_SP_TILE));
    cudaMemsetAsync(ws.racc, 0, (size_t)((unsigned char*)ws.k1 - (unsigned char*)ws.racc), st);        // racc, xacc
    if (rank_mode != 2 && big <= 65536)
        k_sp_rank<true><<<dim3(gr, gy), EMIA_SP_THREADS, 0, st>>>(cap_off, G, L, in_len, rank_mode != 1, ws.k1, ws.k2, ws.kx, ws.racc, ws.xacc);
    else
        k_sp_rank<false><<<dim3(gr, gy), EMIA_SP_THREADS, 0, st>>>(cap_off, G, L, in_len, rank_mode != 1, ws.k1, ws.k2, ws.kx, ws.racc, ws.xacc);
    k_sp_rank_finish<<<gl, T, 0, st>>>(cap_off, G, L, rank_mode != 1, ws.ok, ws.fidx, ws.kx, ws.racc, ws.xacc, ws.pos, ws.order, ws.xorder);
    const unsigned gp = (unsigned)(((size_t)L * 32 + EMIA_SP_THREADS - 1) / EMIA_SP_THREADS);
    k_sp_pairs<<<gp, EMIA_SP_THREADS, 0, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, G, in_idx, L, pair_mode, thr,
                                               rule_max_iou, nullptr, ws.xorder, ws.nx, ws.pos, ws.edges, ws.ecount, nullptr);
    k_sp_resolve<<<(unsigned)G, 1024, 0, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, in_idx, pair_mode, thr, rule_max_iou,
                                               rule_active, num_classes, ws.nok, ws.ok, ws.fidx, ws.pos, ws.order, ws.edges, ws.ecount,
                                               ws.status, ws.flagk, ws.flagu, mode == 0 ? 1 : 0, mode == 2 ? 0 : 1, out_len, out_idx);
    return emia_check_launch("group op launch: %s");
}

extern "C" int emia_dedup_smart(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                                const int32_t* bbox, const int32_t* area, const double* perim0, const int64_t* n_contours,
                                const float* scores, const int32_t* classes, const int32_t* cap_off, int32_t G, int32_t total_cap, int32_t max_cap,
                                const int32_t* in_len, const int32_t* in_idx, double iou_threshold, double max_aspect_ratio,
                                int32_t* out_len, int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (G < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_dedup_smart: %s", "bad G");
    if (G == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !area || !perim0 || !n_contours || !scores || !classes || !cap_off || !in_len ||
        !in_idx || !out_len || !out_idx || !workspace)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_dedup_smart: %s", "null pointer");
    const int L = total_cap;
    if (L == 0) { cudaMemsetAsync(out_len, 0, (size_t)G * 4, (cudaStream_t)stream); return EMIA_OK; }
    return emia_group_pipeline(0, crops, meta, crop_off, bbox, area, perim0, n_contours, scores, classes, cap_off, G, in_len, in_idx,
                               iou_threshold, max_aspect_ratio, nullptr, nullptr, 0, out_len, out_idx, workspace, workspace_bytes,
                               (cudaStream_t)stream, L, max_cap);
}

extern "C" int emia_dedup_inorder(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                                  const int32_t* bbox, const int32_t* area, const int32_t* cap_off, int32_t G, int32_t total_cap, int32_t max_cap,
                                  const int32_t* in_len, const int32_t* in_idx, double iou_threshold, int32_t* out_len,
                                  int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (G < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_dedup_inorder: %s", "bad G");
    if (G == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !area || !cap_off || !in_len || !in_idx || !out_len || !out_idx || !workspace)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_dedup_inorder: %s", "null pointer");
    const int L = total_cap;
    if (L == 0) { cudaMemsetAsync(out_len, 0, (size_t)G * 4, (cudaStream_t)stream); return EMIA_OK; }
    return emia_group_pipeline(1, crops, meta, crop_off, bbox, area, nullptr, nullptr, nullptr, nullptr, cap_off, G, in_len, in_idx,
                               iou_threshold, 0.0, nullptr, nullptr, 0, out_len, out_idx, workspace, workspace_bytes,
                               (cudaStream_t)stream, L, max_cap);
}

// score-sorted greedy de-dup with iou() (run_adaptive_multiscale_inference, src/functions/inference.py:1964-1978):
// visit in np.argsort(scores)[::-1] order, keep unless iou > thr with an already kept mask (any class); keep order out.
extern "C" int emia_dedup_sorted(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                                 const int32_t* bbox, const int32_t* area, const float* scores, const int32_t* cap_off, int32_t G,
                                 int32_t total_cap, int32_t max_cap, const int32_t* in_len, const int32_t* in_idx, double iou_threshold,
                                 int32_t* out_len, int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (G < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_dedup_sorted: %s", "bad G");
    if (G == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !area || !scores || !cap_off || !in_len || !in_idx || !out_len || !out_idx || !workspace)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_dedup_sorted: %s", "null pointer");
    const int L = total_cap;
    if (L == 0) { cudaMemsetAsync(out_len, 0, (size_t)G * 4, (cudaStream_t)stream); return EMIA_OK; }
    return emia_group_pipeline(3, crops, meta, crop_off, bbox, area, nullptr, nullptr, scores, nullptr, cap_off, G, in_len, in_idx,
                               iou_threshold, 0.0, nullptr, nullptr, 0, out_len, out_idx, workspace, workspace_bytes,
                               (cudaStream_t)stream, L, max_cap);
}

extern "C" int emia_overlap_rules(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                                  const int32_t* bbox, const int32_t* area, const float* scores, const int32_t* classes,
                                  const int32_t* cap_off, int32_t G, int32_t total_cap, int32_t max_cap, const int32_t* in_len, const int32_t* in_idx,
                                  const int32_t* rule_active, const double* rule_max_iou, int32_t num_classes,
                                  int32_t* out_len, int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (G < 0 || num_classes < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_overlap_rules: %s", "bad argument");
    if (G == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !area || !scores || !classes || !cap_off || !in_len || !in_idx || !rule_active ||
        !rule_max_iou || !out_len || !out_idx || !workspace)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_overlap_rules: %s", "null pointer");
    const int L = total_cap;
    if (L == 0) { cudaMemsetAsync(out_len, 0, (size_t)G * 4, (cudaStream_t)stream); return EMIA_OK; }
    return emia_group_pipeline(2, crops, meta, crop_off, bbox, area, nullptr, nullptr, scores, classes, cap_off, G, in_len, in_idx, 0.0,
                               0.0, rule_active, rule_max_iou, num_classes, out_len, out_idx, workspace, workspace_bytes,
                               (cudaStream_t)stream, L, max_cap);
}

extern "C" int emia_containment_rules(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                                      const int32_t* bbox, const int32_t* area, const int32_t* classes,
                                      const int32_t* cap_off, int32_t G, int32_t total_cap, int32_t max_cap, const int32_t* in_len, const int32_t* in_idx,
                                      const int32_t* child_class_host, const int32_t* parent_class_host, int32_t n_rules,
                                      double containment_threshold, int32_t* out_len, int32_t* out_idx, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    if (G < 0 || n_rules < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_containment_rules: %s", "bad argument");
    if (G == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !area || !classes || !cap_off || !in_len || !in_idx || !out_len || !out_idx ||
        !workspace || (n_rules > 0 && (!child_class_host || !parent_class_host)))
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_containment_rules: %s", "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = total_cap;
    if (L == 0) { cudaMemsetAsync(out_len, 0, (size_t)G * 4, st); return EMIA_OK; }
    EmiaSparseWs ws;
    if (emia_sparse_carve(workspace, workspace_bytes, (size_t)L, (size_t)G, &ws) != 0)
        return emia_fail(EMIA_ERR_WORKSPACE, "emia_containment_rules: %s", "workspace too small");
    const int T = 128;
    const unsigned gl = (unsigned)((L + T - 1) / T);
    int32_t* rem_a = ws.rem;       // state before the current rule
    int32_t* rem_b = ws.ok;        // state after (children of this rule added)
    cudaMemsetAsync(rem_a, 0, (size_t)L * 4, st);
    for (int r = 0; r < n_rules; ++r) {
        const int child = child_class_host[r], parent = parent_class_host[r];
        if (child == parent) {
            k_containment_rule_selfclass<<<(unsigned)((G + T - 1) / T), T, 0, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, G,
                                                                                   in_len, in_idx, child, containment_threshold, rem_a);
        } else {
            cudaMemcpyAsync(rem_b, rem_a, (size_t)L * 4, cudaMemcpyDeviceToDevice, st);
            if (max_cap > 0 && max_cap <= EMIA_FUSED_MAX_CAP) {
                const size_t smem = emia_fused_smem_bytes(max_cap);
                emia_need_dyn_smem((const void*)k_containment_fused, smem);
                k_containment_fused<<<(unsigned)G, EMIA_FUSED_THREADS, smem, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, in_len,
                                                                                  in_idx, child, parent, containment_threshold, max_cap, rem_a, rem_b);
            } else {
                // sparse sweep: roles + x_min keys -> x order -> (child, parent) pairs -> atomicMax(best[child]) -> decide
                int32_t* role = ws.fidx;
                cudaMemsetAsync(ws.nok, 0, (size_t)((unsigned char*)ws.edges - (unsigned char*)ws.nok), st);
                k_sp_contain_prepare<<<gl, T, 0, st>>>(cap_off, G, in_len, in_idx, L, classes, bbox, area, child, parent, rem_a, role, ws.best,
                                                       ws.kx, ws.nx, ws.anyp);
                const unsigned gr = (unsigned)((L + EMIA_SP_THREADS - 1) / EMIA_SP_THREADS);
                const int big = max_cap > 0 ? max_cap : L;
                const unsigned gy = (unsigned)std::min(16, std::max(1, (big + EMIA_SP_TILE - 1) / EMIA_SP_TILE));
                cudaMemsetAsync(ws.racc, 0, (size_t)((unsigned char*)ws.k1 - (unsigned char*)ws.racc), st);
                if (big <= 65536)
                    k_sp_rank<true><<<dim3(gr, gy), EMIA_SP_THREADS, 0, st>>>(cap_off, G, L, in_len, 0, ws.k1, ws.k2, ws.kx, ws.racc, ws.xacc);
                else
                    k_sp_rank<false><<<dim3(gr, gy), EMIA_SP_THREADS, 0, st>>>(cap_off, G, L, in_len, 0, ws.k1, ws.k2, ws.kx, ws.racc, ws.xacc);
                k_sp_rank_finish<<<gl, T, 0, st>>>(cap_off, G, L, 0, nullptr, nullptr, ws.kx, ws.racc, ws.xacc, nullptr, nullptr, ws.xorder);
                const unsigned gp = (unsigned)(((size_t)L * 32 + EMIA_SP_THREADS - 1) / EMIA_SP_THREADS);
                k_sp_pairs<<<gp, EMIA_SP_THREADS, 0, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, G, in_idx, L, 4, 0.0, nullptr,
                                                           role, ws.xorder, ws.nx, nullptr, nullptr, nullptr, ws.best);
                k_sp_contain_decide<<<gl, T, 0, st>>>(cap_off, G, in_idx, L, area, containment_threshold, role, ws.best, ws.anyp, rem_b);
            }
            int32_t* t = rem_a; rem_a = rem_b; rem_b = t;
        }
    }
    k_group_compact<<<(unsigned)G, 1024, 0, st>>>(cap_off, G, in_len, in_idx, rem_a, out_len, out_idx);
    return emia_check_launch("emia_containment_rules launch: %s");
}

__global__ void k_pair_counts(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                              const int64_t* __restrict__ crop_off, const int32_t* __restrict__ area,
                              const int32_t* __restrict__ pa, const int32_t* __restrict__ pb, int64_t n, int64_t* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int a = pa[k], b = pb[k];
    const EmiaCropRef ra = emia_crop_ref(crops, meta, crop_off, a);
    const EmiaCropRef rb = emia_crop_ref(crops, meta, crop_off, b);
    out[3 * k] = emia_crop_inter(ra, rb);
    out[3 * k + 1] = area[a];
    out[3 * k + 2] = area[b];
}

extern "C" int emia_pair_counts(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off,
                                const int32_t* area, const int32_t* pa, const int32_t* pb, int64_t n_pairs, int64_t* out,
                                void* stream) {
    if (n_pairs < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_pair_counts: %s", "bad n");
    if (n_pairs == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !area || !pa || !pb || !out) return emia_fail(EMIA_ERR_BAD_ARG, "emia_pair_counts: %s", "null pointer");
    k_pair_counts<<<(unsigned)((n_pairs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, area, pa, pb, n_pairs, out);
    return emia_check_launch("emia_pair_counts launch: %s");
}
