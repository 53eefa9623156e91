// emia_group_kernels.cuh — K4: mask-IoU de-duplication and spatial constraints over G groups (part of emia_kernels.cu).
//
// Every operation has the same shape: (1) decide which list slots take part, (2) rank them into the order the
// reference visits them, (3) evaluate all candidate pairs in parallel into a per-group bit matrix of
// "a suppresses b" relations (in rank space), (4) replay the reference's sequential greedy loop with one warp per
// group on that bit matrix (32 ranks per word, one OR per word), (5) write the surviving list.
#pragma once

struct EmiaGroupWs {
    int32_t* ok;       // L   slot takes part
    int32_t* fidx;     // L   index in the filtered list (position among ok slots of the group)
    int32_t* pos;      // L   rank (visit order) inside the group
    int32_t* order;    // L   order[cap_off[g] + rank] = slot
    int32_t* nok;      // G   number of ok slots per group
    int64_t* bm_off;   // G+1 word offset of each group's bit matrix (row stride = ceil(cap_g/32) words)
    int64_t* rm_off;   // G+1 word offset of each group's removed bitset
    uint32_t* bm;
    uint32_t* rm;
    int32_t* rem;      // L   per-slot removed flag (containment / compaction)
    size_t bm_words, rm_words;
};

static size_t emia_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t emia_group_workspace_bytes(const int32_t* cap_off_host, int32_t G) {
    if (!cap_off_host || G < 0) return 0;
    const size_t L = (size_t)cap_off_host[G];
    size_t bm = 0, rm = 0;
    for (int g = 0; g < G; ++g) {
        const size_t c = (size_t)(cap_off_host[g + 1] - cap_off_host[g]);
        const size_t w = (c + 31) / 32;
        bm += c * w;
        rm += w;
    }
    size_t b = 0;
    b += 5 * emia_align_up(L * 4 + 16, 256);                 // ok, fidx, pos, order, rem
    b += emia_align_up((size_t)G * 4 + 16, 256);             // nok
    b += 2 * emia_align_up(((size_t)G + 1) * 8, 256);        // bm_off, rm_off
    (void)rm;
    b += emia_align_up(bm * 4 + 16, 256) + emia_align_up((L / 32 + (size_t)G + 1) * 4 + 16, 256);
    return b + 512;
}

// carve the workspace; bm/rm word totals are recomputed on the device (k_group_offsets), the host only needs a
// safe upper bound for the memset, which is the remaining workspace.
static int emia_carve_ws(void* workspace, size_t bytes, size_t L, int G, EmiaGroupWs* ws, size_t* tail_bytes) {
    unsigned char* p = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    unsigned char* end = (unsigned char*)workspace + bytes;
    auto take = [&](size_t nbytes) -> void* { void* r = p; p += emia_align_up(nbytes, 256); return r; };
    ws->ok = (int32_t*)take(L * 4 + 16);
    ws->fidx = (int32_t*)take(L * 4 + 16);
    ws->pos = (int32_t*)take(L * 4 + 16);
    ws->order = (int32_t*)take(L * 4 + 16);
    ws->rem = (int32_t*)take(L * 4 + 16);
    ws->nok = (int32_t*)take((size_t)G * 4 + 16);
    ws->bm_off = (int64_t*)take(((size_t)G + 1) * 8);
    ws->rm_off = (int64_t*)take(((size_t)G + 1) * 8);
    if (p > end) return -1;
    ws->bm = (uint32_t*)p;       // bm then rm share the tail; split fixed by the device-computed offsets
    *tail_bytes = (size_t)(end - p);
    return 0;
}

__global__ void k_group_offsets(const int32_t* __restrict__ cap_off, int G, int64_t* __restrict__ bm_off, int64_t* __restrict__ rm_off) {
    if (blockIdx.x || threadIdx.x) return;
    int64_t b = 0, r = 0;
    for (int g = 0; g < G; ++g) {
        bm_off[g] = b; rm_off[g] = r;
        const int64_t c = cap_off[g + 1] - cap_off[g];
        const int64_t w = (c + 31) / 32;
        b += c * w; r += w;
    }
    bm_off[G] = b; rm_off[G] = r;
}

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

// one warp per group: fidx = exclusive count of ok slots in list order
__global__ void k_group_fidx(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ ok,
                             int32_t* __restrict__ fidx, int32_t* __restrict__ nok) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    int run = 0;
    for (int k0 = 0; k0 < cap; k0 += 32) {
        const int k = k0 + lane;
        const int v = (k < cap) ? ok[base + k] : 0;
        const unsigned b = __ballot_sync(0xffffffffu, v);
        if (k < cap) fidx[base + k] = run + __popc(b & ((1u << lane) - 1u));
        run += __popc(b);
    }
    if (lane == 0) nok[g] = run;
}

// ---- (2) rank --------------------------------------------------------------------------------------------------
// mode 0 (deduplicate_masks_smart): np.argsort(scores)[::-1] -> score descending, ties by filtered index descending
// mode 1 (in-order): list order
// mode 2 (overlap rules): stable sort by score descending inside each class; classes kept apart by class id
__global__ void k_group_rank(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_idx, int L, int mode,
                             const float* __restrict__ scores, const int32_t* __restrict__ classes,
                             const int32_t* __restrict__ ok, const int32_t* __restrict__ fidx, int32_t* __restrict__ pos,
                             int32_t* __restrict__ order) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    if (!ok[s]) { pos[s] = -1; return; }
    const int g = emia_find_group(cap_off, G, s);
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    int rank = 0;
    if (mode == 1) {
        rank = fidx[s];
    } else {
        const int inst = in_idx[s];
        const float sc = scores[inst];
        const int cl = (mode == 2) ? classes[inst] : 0;
        const int fi = fidx[s];
        for (int k = 0; k < cap; ++k) {
            const int t = base + k;
            if (!ok[t] || t == s) continue;
            const int it = in_idx[t];
            const float st = scores[it];
            bool before;
            if (mode == 0) before = (st > sc) || (st == sc && fidx[t] > fi);
            else {
                const int ct = classes[it];
                before = (ct < cl) || (ct == cl && ((st > sc) || (st == sc && fidx[t] < fi)));
            }
            rank += before;
        }
    }
    pos[s] = rank;
    order[base + rank] = s;
}

// ---- (3) pair relations -> bit matrix ----------------------------------------------------------------------------
// mode 0: smart de-dup  (same class, Q1 bbox test, IoU > thr)
// mode 1: in-order      (any class, IoU > thr)
// mode 2: overlap rules (same class with an active rule, IoU > rule_max_iou[class])
__global__ void k_group_pairs(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                              const int64_t* __restrict__ crop_off, const int32_t* __restrict__ bbox,
                              const int32_t* __restrict__ area, const int32_t* __restrict__ classes,
                              const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_idx, int L, int mode,
                              double thr, const int32_t* __restrict__ rule_active, const double* __restrict__ rule_max_iou,
                              int num_classes, const int32_t* __restrict__ ok, const int32_t* __restrict__ pos,
                              const int64_t* __restrict__ bm_off, uint32_t* __restrict__ bm) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L || !ok[s]) return;
    const int g = emia_find_group(cap_off, G, s);
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    const int stride = (cap + 31) >> 5;
    const int ia = in_idx[s];
    const int ca = (mode == 1) ? 0 : classes[ia];
    double th = thr;
    if (mode == 2) {
        if (ca < 0 || ca >= num_classes || !rule_active[ca]) return;
        th = rule_max_iou[ca];
    }
    const int aa = area[ia];
    if (aa <= 0) return;
    const int* ba = bbox + 4 * ia;
    const EmiaCropRef ra = emia_crop_ref(crops, meta, crop_off, ia);
    uint32_t* rows = bm + bm_off[g];
    const int pa = pos[s];
    for (int k = s - base + 1; k < cap; ++k) {
        const int t = base + k;
        if (!ok[t]) continue;
        const int ib = in_idx[t];
        if (mode != 1 && classes[ib] != ca) continue;
        const int* bb = bbox + 4 * ib;
        if (mode == 0) { if (!emia_bbox_overlap_q1(ba, bb)) continue; }
        if (!emia_bbox_overlap(ba, bb)) continue;       // no real overlap -> intersection 0 -> IoU 0
        const int ab = area[ib];
        if (ab <= 0) continue;
        const EmiaCropRef rb = emia_crop_ref(crops, meta, crop_off, ib);
        const int inter = emia_crop_inter(ra, rb);
        if (inter == 0) continue;
        const int uni = aa + ab - inter;
        if ((double)inter / (double)uni > th) {
            const int pb = pos[t];
            atomicOr(rows + (size_t)pa * stride + (pb >> 5), 1u << (pb & 31));
            atomicOr(rows + (size_t)pb * stride + (pa >> 5), 1u << (pa & 31));
        }
    }
}

// ---- (4) sequential greedy replay, one warp per group --------------------------------------------------------------
// q2 != 0: deduplicate_masks_smart's slice `sorted_indices[idx+1:]` (Q2) — the keeper at rank p with FILTERED INDEX idx
//          suppresses ranks q >= idx + 1;   q2 == 0: ranks q > p.
// ordered_out != 0: survivors are written in rank order (keep order); otherwise in list order.
__global__ void k_group_greedy(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_idx,
                               const int32_t* __restrict__ nok, const int32_t* __restrict__ order,
                               const int32_t* __restrict__ fidx, const int32_t* __restrict__ ok,
                               const int32_t* __restrict__ pos, const int64_t* __restrict__ bm_off,
                               const uint32_t* __restrict__ bm, const int64_t* __restrict__ rm_off, uint32_t* __restrict__ rm,
                               int q2, int ordered_out, int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    const int stride = (cap + 31) >> 5;
    const int n = nok[g];
    const uint32_t* rows = bm + bm_off[g];
    uint32_t* removed = rm + rm_off[g];
    const int words = (n + 31) >> 5;
    int nk = 0;
    for (int p = 0; p < n; ++p) {
        const uint32_t rw = removed[p >> 5];
        if ((rw >> (p & 31)) & 1u) continue;            // warp-uniform
        const int s = order[base + p];
        if (ordered_out && lane == 0) out_idx[base + nk] = in_idx[s];
        ++nk;
        const int first = q2 ? (fidx[s] + 1) : (p + 1); // suppress ranks >= first
        const uint32_t* row = rows + (size_t)p * stride;
        for (int w = lane; w < words; w += 32) {
            const int lo = w << 5;
            uint32_t mask;
            if (lo + 31 < first) mask = 0u;
            else if (lo >= first) mask = 0xffffffffu;
            else mask = 0xffffffffu << (first - lo);
            const uint32_t add = row[w] & mask;
            if (add) removed[w] |= add;
        }
        __syncwarp();
    }
    if (ordered_out) {
        if (lane == 0) out_len[g] = nk;
    } else {
        // survivors in list order
        int run = 0;
        for (int k0 = 0; k0 < cap; k0 += 32) {
            const int k = k0 + lane;
            int keep = 0;
            int inst = 0;
            if (k < cap && ok[base + k]) {
                const int p = pos[base + k];
                keep = !((removed[p >> 5] >> (p & 31)) & 1u);
                inst = in_idx[base + k];
            }
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (keep) out_idx[base + run + __popc(b & ((1u << lane) - 1u))] = inst;
            run += __popc(b);
        }
        if (lane == 0) out_len[g] = run;
    }
}

// ---- containment ---------------------------------------------------------------------------------------------------
// filter_by_containment_rules for ONE rule (child_class -> parent_class); rem[] persists across rules.
// child == parent is evaluated by a single thread per group (sequential semantics of the reference).
__global__ void k_containment_rule(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                   const int64_t* __restrict__ crop_off, const int32_t* __restrict__ bbox,
                                   const int32_t* __restrict__ area, const int32_t* __restrict__ classes,
                                   const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                   const int32_t* __restrict__ in_idx, int L, int child, int parent, double thr,
                                   const int32_t* __restrict__ rem_in, int32_t* __restrict__ rem_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    const int g = emia_find_group(cap_off, G, s);
    const int base = cap_off[g];
    const int len = in_len[g];
    if (s - base >= len) return;
    const int ic = in_idx[s];
    if (classes[ic] != child || rem_in[s]) return;
    // does the parent class occur at all in this group?
    bool any_parent = false;
    for (int k = 0; k < len && !any_parent; ++k) any_parent = (classes[in_idx[base + k]] == parent);
    if (!any_parent) { rem_out[s] = 1; return; }
    const int ac = area[ic];
    const int* bc = bbox + 4 * ic;
    if (ac <= 0 || bc[0] < 0) { rem_out[s] = 1; return; }
    const EmiaCropRef rc = emia_crop_ref(crops, meta, crop_off, ic);
    double best = 0.0;
    for (int k = 0; k < len; ++k) {
        const int t = base + k;
        const int ip = in_idx[t];
        if (classes[ip] != parent || rem_in[t]) continue;
        const int* bp = bbox + 4 * ip;
        if (!emia_bbox_overlap(bc, bp)) continue;
        const EmiaCropRef rp = emia_crop_ref(crops, meta, crop_off, ip);
        const int inter = emia_crop_inter(rc, rp);
        const double c = (double)inter / (double)ac;
        if (c > best) best = c;
    }
    if (best < thr) rem_out[s] = 1;
}

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

// survivors (rem == 0) in list order, one warp per group
__global__ void k_group_compact(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                const int32_t* __restrict__ in_idx, const int32_t* __restrict__ rem,
                                int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    const int base = cap_off[g], len = in_len[g];
    int run = 0;
    for (int k0 = 0; k0 < len; k0 += 32) {
        const int k = k0 + lane;
        const int keep = (k < len) && !rem[base + k];
        const int inst = (k < len) ? in_idx[base + k] : 0;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) out_idx[base + run + __popc(b & ((1u << lane) - 1u))] = inst;
        run += __popc(b);
    }
    if (lane == 0) out_len[g] = run;
}

#include "emia_group_fused.cuh"

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
        cudaFuncSetAttribute(k_group_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_group_fused<<<(unsigned)G, EMIA_FUSED_THREADS, smem, st>>>(crops, meta, crop_off, bbox, area, perim0, n_contours, scores, classes,
                                                                    cap_off, in_len, in_idx, mode, thr, max_aspect, rule_active, rule_max_iou,
                                                                    num_classes, max_cap, out_len, out_idx);
        return emia_check_launch("group op (fused) launch: %s");
    }
    EmiaGroupWs ws;
    size_t tail = 0;
    if (emia_carve_ws(workspace, workspace_bytes, (size_t)L, G, &ws, &tail) != 0)
        return emia_fail(EMIA_ERR_WORKSPACE, "group op: %s", "workspace too small");
    // bm and rm live in the tail: [bm | rm]; rm starts at the device-computed bm total, which the host bounds by
    // placing rm at the END of the tail (rm total words <= L/32 + G).
    const size_t rm_bytes = emia_align_up(((size_t)L / 32 + (size_t)G + 1) * 4 + 16, 256);
    if (tail < rm_bytes + 256) return emia_fail(EMIA_ERR_WORKSPACE, "group op: %s", "workspace too small");
    ws.rm = (uint32_t*)((unsigned char*)ws.bm + (tail - rm_bytes));
    cudaMemsetAsync(ws.bm, 0, tail, st);
    const int T = 128;
    const unsigned gl = (unsigned)((L + T - 1) / T);
    const unsigned gw = (unsigned)(((size_t)G * 32 + T - 1) / T);
    k_group_offsets<<<1, 1, 0, st>>>(cap_off, G, ws.bm_off, ws.rm_off);
    k_group_select<<<gl, T, 0, st>>>(cap_off, G, in_len, in_idx, L, mode == 0 ? 0 : 1, bbox, area, perim0, n_contours, max_aspect, ws.ok);
    k_group_fidx<<<gw, T, 0, st>>>(cap_off, G, ws.ok, ws.fidx, ws.nok);
    const int rank_mode = (mode == 3) ? 0 : mode, pair_mode = (mode == 3) ? 1 : mode;
    k_group_rank<<<gl, T, 0, st>>>(cap_off, G, in_idx, L, rank_mode, scores, classes, ws.ok, ws.fidx, ws.pos, ws.order);
    k_group_pairs<<<gl, T, 0, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, G, in_idx, L, pair_mode, thr, rule_active,
                                    rule_max_iou, num_classes, ws.ok, ws.pos, ws.bm_off, ws.bm);
    k_group_greedy<<<gw, T, 0, st>>>(cap_off, G, in_idx, ws.nok, ws.order, ws.fidx, ws.ok, ws.pos, ws.bm_off, ws.bm, ws.rm_off,
                                     ws.rm, mode == 0 ? 1 : 0, mode == 2 ? 0 : 1, out_len, out_idx);
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
    EmiaGroupWs ws;
    size_t tail = 0;
    if (emia_carve_ws(workspace, workspace_bytes, (size_t)L, G, &ws, &tail) != 0)
        return emia_fail(EMIA_ERR_WORKSPACE, "emia_containment_rules: %s", "workspace too small");
    const int T = 128;
    const unsigned gl = (unsigned)((L + T - 1) / T);
    const unsigned gw = (unsigned)(((size_t)G * 32 + T - 1) / T);
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
                cudaFuncSetAttribute(k_containment_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k_containment_fused<<<(unsigned)G, EMIA_FUSED_THREADS, smem, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, in_len,
                                                                                  in_idx, child, parent, containment_threshold, max_cap, rem_a, rem_b);
            } else
            k_containment_rule<<<gl, T, 0, st>>>(crops, meta, crop_off, bbox, area, classes, cap_off, G, in_len, in_idx, L, child, parent,
                                                 containment_threshold, rem_a, rem_b);
            int32_t* t = rem_a; rem_a = rem_b; rem_b = t;
        }
    }
    k_group_compact<<<gw, T, 0, st>>>(cap_off, G, in_len, in_idx, rem_a, out_len, out_idx);
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
