// emia_group_fused.cuh — K4 fast path: one CTA per group, the whole greedy operation in shared memory
// (part of emia_kernels.cu; semantics identical to the staged kernels of emia_group_kernels.cuh, which remain the path
// for groups larger than EMIA_FUSED_MAX_CAP, e.g. the global de-dup of a whole micrograph).
//
// A tile of ~500 instances is small enough that everything the reference's sequential loops touch fits in one SM's
// shared memory: per-member bbox / area / class / score (44 B each), the "a suppresses b" bit matrix in rank space
// (cap x cap/32 words, 46 KB for cap = 608) and a queue of bbox-overlapping candidate pairs.  Stages:
//   select -> filtered index (ballot scan) -> rank (counting, O(cap^2) shared-memory reads) ->
//   candidate pairs (bbox tests on shared memory, pushed to the queue) ->
//   mask IoU: ONE WARP PER CANDIDATE, lanes stride over the words of the crop intersection, popc + shuffle reduce
//   (coalesced reads of the bit-packed crops) -> bit matrix -> greedy replay by one warp with the removed set in registers.
#pragma once

#define EMIA_FUSED_MAX_CAP 1024
#define EMIA_FUSED_THREADS 512
#define EMIA_FUSED_QUEUE 4096

struct EmiaFusedSmem {
    int* inst; int4* bb; int4* geo; long long* coff; int* area; int* cls; float* score; int* fidx; int* pos; int* order;
    uint32_t* xs; uint32_t* bm; uint32_t* queue;
};
__host__ __device__ inline size_t emia_fused_smem_bytes(int cap) {
    const size_t c = (size_t)((cap + 31) & ~31);
    return c * (4 + 16 + 16 + 8 + 4 + 4 + 4 + 4 + 4 + 4) + c * (c / 32) * 4 + (size_t)EMIA_FUSED_QUEUE * 4 + 64 +
           (size_t)EMIA_FUSED_MAX_CAP * 4;
}
__device__ __forceinline__ EmiaFusedSmem emia_fused_carve(unsigned char* base, int cap) {
    const size_t c = (size_t)((cap + 31) & ~31);
    EmiaFusedSmem S;
    S.bb = (int4*)base; base += c * 16;
    S.geo = (int4*)base; base += c * 16;
    S.coff = (long long*)base; base += c * 8;
    S.inst = (int*)base; base += c * 4;
    S.area = (int*)base; base += c * 4;
    S.cls = (int*)base; base += c * 4;
    S.score = (float*)base; base += c * 4;
    S.fidx = (int*)base; base += c * 4;
    S.pos = (int*)base; base += c * 4;
    S.order = (int*)base; base += c * 4;
    S.queue = (uint32_t*)base; base += (size_t)EMIA_FUSED_QUEUE * 4;
    S.xs = (uint32_t*)base; base += (size_t)EMIA_FUSED_MAX_CAP * 4;
    S.bm = (uint32_t*)base;
    return S;
}

// popcount(a AND b) computed by a whole warp (all 32 lanes must call); result valid in every lane
__device__ __forceinline__ int emia_crop_inter_warp(const EmiaCropRef& a, const EmiaCropRef& b, int lane) {
    const int r0 = max(a.ry0, b.ry0), r1 = min(a.ry0 + a.ch, b.ry0 + b.ch);
    const int c0 = max(a.wc0, b.wc0), c1 = min(a.wc0 + a.cw, b.wc0 + b.cw);
    int cnt = 0;
    if (r0 < r1 && c0 < c1) {
        const int wc = c1 - c0, nw = (r1 - r0) * wc;
        for (int t = lane; t < nw; t += 32) {
            const int r = r0 + t / wc, c = c0 + t % wc;
            cnt += __popc(a.w[(size_t)(r - a.ry0) * a.cw + (c - a.wc0)] & b.w[(size_t)(r - b.ry0) * b.cw + (c - b.wc0)]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    return cnt;
}

// The same with EIGHT lanes per pair (four pairs per warp in flight: the loads of a pair are one dependent round trip, so
// more pairs in flight hide more latency).  geo = (ry0, wc0, ch, cw); all 32 lanes must call; `valid` per 8-lane group.
__device__ __forceinline__ int emia_crop_inter_oct(const uint32_t* __restrict__ crops, int4 ga, long long oa, int4 gb, long long ob,
                                                   bool valid, int sl) {
    int cnt = 0;
    if (valid) {
        const int r0 = max(ga.x, gb.x), r1 = min(ga.x + ga.z, gb.x + gb.z);
        const int c0 = max(ga.y, gb.y), c1 = min(ga.y + ga.w, gb.y + gb.w);
        if (r0 < r1 && c0 < c1) {
            const int wc = c1 - c0, nw = (r1 - r0) * wc;
            const uint32_t* pa = crops + oa;
            const uint32_t* pb = crops + ob;
            // four independent word pairs per trip: the loads of a trip are issued back to back (the crops sit in L2 / HBM and
            // a CTA has few warps, so memory-level parallelism inside the pair is what hides the latency)
            const int oa = -ga.x * ga.w - ga.y, ob = -gb.x * gb.w - gb.y;
            // tt / wc by a multiply-high: exact for tt < 2^16, 2 <= wc < 2^16 (two integer divisions per word were 15 % of the kernel's
            // instructions); wc == 1 has no 32-bit reciprocal
            const uint32_t inv_wc = wc > 1 ? (uint32_t)((0x100000000ull + (uint32_t)wc - 1u) / (uint32_t)wc) : 0u;
            for (int t = sl; t < nw; t += 32) {
                uint32_t wa[4], wb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tt = t + 8 * u;
                    const int rq = wc > 1 ? (int)__umulhi((uint32_t)tt, inv_wc) : tt;
                    const int r = r0 + rq, c = c0 + (tt - rq * wc);
                    const bool in = tt < nw;
                    wa[u] = in ? pa[r * ga.w + c + oa] : 0u;
                    wb[u] = in ? pb[r * gb.w + c + ob] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) cnt += __popc(wa[u] & wb[u]);
            }
        }
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    return cnt;
}

// mode 0: deduplicate_masks_smart, mode 1: in-order iou() de-dup, mode 2: overlap rules (see emia_group_kernels.cuh)
__global__ void __launch_bounds__(EMIA_FUSED_THREADS) k_group_fused(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off,
    const int32_t* __restrict__ bbox, const int32_t* __restrict__ area, const double* __restrict__ perim0,
    const int64_t* __restrict__ n_contours, const float* __restrict__ scores, const int32_t* __restrict__ classes,
    const int32_t* __restrict__ cap_off, const int32_t* __restrict__ in_len, const int32_t* __restrict__ in_idx, int mode,
    double thr, double max_aspect_ratio, const int32_t* __restrict__ rule_active, const double* __restrict__ rule_max_iou,
    int num_classes, int smem_cap, int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char fused_smem[];
    __shared__ int s_nok, s_qn;
    const EmiaFusedSmem S = emia_fused_carve(fused_smem, smem_cap);
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = EMIA_FUSED_THREADS / 32;
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    const int len = in_len[g];
    // mode 3 (score-sorted iou() de-dup): visit order of mode 0, pair relation of mode 1
    const int rank_mode = (mode == 3) ? 0 : mode;
    const int pair_mode = (mode == 3) ? 1 : mode;
    if (tid == 0) { s_qn = 0; s_nok = 0; }
    // ---- load + select
    for (int k = tid; k < cap; k += EMIA_FUSED_THREADS) {
        int ok = 0, inst = 0, a = 0, cl = 0;
        float sc = 0.f;
        int4 b = make_int4(-1, -1, -1, -1);
        if (k < len) {
            inst = in_idx[base + k];
            const emia_inst_meta m = meta[inst];
            S.geo[k] = make_int4(m.ry0, m.wc0, m.ch, m.cw);
            S.coff[k] = crop_off[inst];
            a = area[inst];
            b = ((const int4*)bbox)[inst];
            if (pair_mode != 1) cl = classes[inst];
            if (rank_mode != 1) sc = scores[inst];
            ok = 1;
            if (mode == 0) {
                if (a <= 0) ok = 0;
                else {
                    const int bw = b.w - b.y + 1, bh = b.z - b.x + 1;
                    if (max_aspect_ratio > 0.0) {
                        const double asp = (double)max(bw, bh) / (double)min(bw, bh);
                        if (asp > max_aspect_ratio) ok = 0;
                    }
                    if (ok && n_contours[inst] > 0) {
                        const double per = perim0[inst];
                        if (per > 0) {
                            const double compactness = ((4 * M_PI) * (double)a) / (per * per);
                            if (compactness < 0.15) ok = 0;
                        }
                    }
                }
            }
        }
        S.inst[k] = inst; S.area[k] = a; S.bb[k] = b; S.cls[k] = cl; S.score[k] = sc;
        S.fidx[k] = ok ? 0 : -1;
    }
    __syncthreads();
    // ---- filtered index: exclusive count of ok slots in list order (warp 0)
    if (warp == 0) {
        int run = 0;
        for (int k0 = 0; k0 < cap; k0 += 32) {
            const int k = k0 + lane;
            const int v = (k < cap) && S.fidx[k] == 0;
            const unsigned bmask = __ballot_sync(0xffffffffu, v);
            if (v) S.fidx[k] = run + __popc(bmask & ((1u << lane) - 1u));
            run += __popc(bmask);
        }
        if (lane == 0) s_nok = run;
    }
    __syncthreads();
    const int nok = s_nok;
    const int stride = (nok + 31) >> 5;
    // ---- rank: bitonic sort of (class, score descending, tie) keys in shared memory; the tie-break by filtered index is
    // a tie-break by slot (fidx is monotone in the slot).  Also clear the bit matrix and build the x-sorted sweep keys.
    for (int k = tid; k < nok * stride; k += EMIA_FUSED_THREADS) S.bm[k] = 0u;
    int P = 32;
    while (P < cap) P <<= 1;
    uint64_t* rk = (uint64_t*)S.queue;          // the queue is not in use yet (P * 8 <= 8 KB of its 16 KB)
    for (int k = tid; k < P; k += EMIA_FUSED_THREADS) {
        uint64_t key = ~0ull;
        uint32_t xk = 0xFFFFFFFFu;
        if (k < cap && S.fidx[k] >= 0) {
            if (rank_mode == 1) key = (uint64_t)k;
            else {
                const uint32_t bits = __float_as_uint(S.score[k]);
                const uint32_t u = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);      // order-preserving
                if (rank_mode == 0) key = ((uint64_t)(~u) << 16) | (uint64_t)(0xFFFFu - (uint32_t)k);
                else key = ((uint64_t)((uint32_t)(S.cls[k] + 0x8000) & 0xFFFFu) << 48) | ((uint64_t)(~u) << 16) | (uint64_t)k;
            }
            // takes part in the pair sweep?
            const int ca = S.cls[k];
            bool part = S.area[k] > 0 && S.bb[k].x >= 0;
            if (mode == 2 && (ca < 0 || ca >= num_classes || !rule_active[ca])) part = false;
            if (part) xk = ((uint32_t)S.bb[k].y << 16) | (uint32_t)k;                       // x_min | slot
        }
        rk[k] = key;
        S.xs[k] = xk;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int st = size >> 1; st > 0; st >>= 1) {
            for (int t = tid; t < (P >> 1); t += EMIA_FUSED_THREADS) {
                const int lo = ((t & ~(st - 1)) << 1) | (t & (st - 1));      // st is a power of two
                const int hi = lo + st;
                const bool up = ((lo & size) == 0);
                const uint64_t ka = rk[lo], kb = rk[hi];
                if ((ka > kb) == up) { rk[lo] = kb; rk[hi] = ka; }
                const uint32_t xa = S.xs[lo], xb = S.xs[hi];
                if ((xa > xb) == up) { S.xs[lo] = xb; S.xs[hi] = xa; }
            }
            __syncthreads();
        }
    }
    for (int k = tid; k < cap; k += EMIA_FUSED_THREADS) if (S.fidx[k] < 0) S.pos[k] = -1;
    for (int r = tid; r < nok; r += EMIA_FUSED_THREADS) {
        const uint32_t low = (uint32_t)(rk[r] & 0xFFFFu);
        const int k = (rank_mode == 0) ? (int)(0xFFFFu - low) : (int)low;
        S.pos[k] = r;
        S.order[r] = k;
    }
    __syncthreads();
    // ---- candidate pairs: sweep over the members sorted by x_min; member i only meets the members that start before its
    // x_max (the suppression relation is symmetric, so the enumeration order is free).  A pair that does not fit the queue
    // is evaluated by its thread.
    for (int i = tid; i < cap; i += EMIA_FUSED_THREADS) {
        const uint32_t xi = S.xs[i];
        if (xi == 0xFFFFFFFFu) continue;
        const int a = (int)(xi & 0xFFFFu);
        const int ca = S.cls[a];
        const double th = (mode == 2) ? rule_max_iou[ca] : thr;
        const int aa = S.area[a];
        const int4 ba4 = S.bb[a];
        const int ba[4] = {ba4.x, ba4.y, ba4.z, ba4.w};
        for (int j = i + 1; j < cap; ++j) {
            const uint32_t xj = S.xs[j];
            if (xj == 0xFFFFFFFFu || (int)(xj >> 16) > ba4.w) break;   // end of the list / starts right of a's x_max
            const int b = (int)(xj & 0xFFFFu);
            if (pair_mode != 1 && S.cls[b] != ca) continue;
            const int4 bb4 = S.bb[b];
            const int bb[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
            if (pair_mode == 0) { if (!emia_bbox_overlap_q1(ba, bb)) continue; }
            if (!emia_bbox_overlap(ba, bb)) continue;
            const int q = atomicAdd(&s_qn, 1);
            if (q < EMIA_FUSED_QUEUE) S.queue[q] = (uint32_t)a | ((uint32_t)b << 16);
            else {
                const EmiaCropRef ra = emia_crop_ref(crops, meta, crop_off, S.inst[a]);
                const EmiaCropRef rb = emia_crop_ref(crops, meta, crop_off, S.inst[b]);
                const int inter = emia_crop_inter(ra, rb);
                const int uni = aa + S.area[b] - inter;
                if (inter > 0 && (double)inter / (double)uni > th) {
                    const int pa = S.pos[a], pb = S.pos[b];
                    atomicOr(&S.bm[pa * stride + (pb >> 5)], 1u << (pb & 31));
                    atomicOr(&S.bm[pb * stride + (pa >> 5)], 1u << (pa & 31));
                }
            }
        }
    }
    __syncthreads();
    {
        const int nq = min(s_qn, EMIA_FUSED_QUEUE);
        const int sub = lane >> 3, sl = lane & 7;
        for (int q0 = warp * 4; q0 < nq; q0 += nwarps * 4) {
            const int q = q0 + sub;
            const bool valid = q < nq;
            const uint32_t e = valid ? S.queue[q] : 0u;
            const int a = (int)(e & 0xFFFFu), b = (int)(e >> 16);
            const int inter = emia_crop_inter_oct(crops, S.geo[a], S.coff[a], S.geo[b], S.coff[b], valid, sl);
            if (sl == 0 && inter > 0) {
                double th = thr;
                if (mode == 2) th = rule_max_iou[S.cls[a]];
                const int uni = S.area[a] + S.area[b] - inter;
                if ((double)inter / (double)uni > th) {
                    const int pa = S.pos[a], pb = S.pos[b];
                    atomicOr(&S.bm[pa * stride + (pb >> 5)], 1u << (pb & 31));
                    atomicOr(&S.bm[pb * stride + (pa >> 5)], 1u << (pa & 31));
                }
            }
        }
    }
    __syncthreads();
    // ---- ranks with any suppression edge.  The bit matrix is symmetric before the replay masks it, so a rank with an empty row
    // neither suppresses nor can be suppressed: it is kept unconditionally and the sequential replay only visits the others
    // (typically 10-20 % of a tile's members; the replay was 35 % of this kernel's stall samples).
    uint32_t* act = S.xs;                                      // the sweep keys are dead
    for (int p0 = warp * 32; p0 < nok; p0 += nwarps * 32) {
        const int p = p0 + lane;
        uint32_t any = 0u;
        if (p < nok) for (int w = 0; w < stride; ++w) any |= S.bm[p * stride + w];
        const unsigned b = __ballot_sync(0xffffffffu, any != 0u);
        if (lane == 0) act[p0 >> 5] = b;
    }
    __syncthreads();
    // ---- greedy replay (warp 0): lane w keeps word w of the removed set in a register (stride <= 32)
    if (warp == 0) {
        const int q2 = (mode == 0), ordered_out = (mode != 2);
        uint32_t removed = 0u, kept = 0u;                        // word `lane` of the removed set / of the visited-and-kept set
        const uint32_t my_act = (lane << 5) < nok ? act[lane] : 0u;
        for (int w = 0; (w << 5) < nok; ++w) {
            uint32_t aw = __shfl_sync(0xffffffffu, my_act, w);   // warp-uniform
            while (aw) {
                const int bit = __ffs((int)aw) - 1;
                aw &= aw - 1u;
                const int p = (w << 5) + bit;
                const uint32_t rw = __shfl_sync(0xffffffffu, removed, w);
                if ((rw >> bit) & 1u) continue;
                if (lane == w) kept |= 1u << bit;                    // decided now: a later Q2 "removal" of an earlier rank has no effect
                const int s = S.order[p];
                const int first = q2 ? (S.fidx[s] + 1) : (p + 1);    // suppress ranks >= first
                if (lane < stride) {
                    const int lo = lane << 5;
                    uint32_t mask;
                    if (lo + 31 < first) mask = 0u;
                    else if (lo >= first) mask = 0xffffffffu;
                    else mask = 0xffffffffu << (first - lo);
                    removed |= S.bm[p * stride + lane] & mask;
                }
            }
        }
        const uint32_t keepw = ~my_act | kept;                    // ranks without edges are kept unconditionally
        int nk = 0;
        if (ordered_out) {
            // survivors in rank order
            for (int p0 = 0; p0 < nok; p0 += 32) {
                const int p = p0 + lane;
                const uint32_t kw = __shfl_sync(0xffffffffu, keepw, p0 >> 5);
                const int keep = (p < nok) && ((kw >> lane) & 1u);
                const unsigned bmask = __ballot_sync(0xffffffffu, keep);
                if (keep) out_idx[base + nk + __popc(bmask & ((1u << lane) - 1u))] = S.inst[S.order[p]];
                nk += __popc(bmask);
            }
        }
        if (ordered_out) {
            if (lane == 0) out_len[g] = nk;
        } else {
            int run = 0;
            for (int k0 = 0; k0 < cap; k0 += 32) {
                const int k = k0 + lane;
                const int p = (k < cap) ? S.pos[k] : -1;
                const uint32_t kw = __shfl_sync(0xffffffffu, keepw, (p >= 0 ? p : 0) >> 5);
                const int keep = (p >= 0) && ((kw >> (p & 31)) & 1u);
                const unsigned bmask = __ballot_sync(0xffffffffu, keep);
                if (keep) out_idx[base + run + __popc(bmask & ((1u << lane) - 1u))] = S.inst[k];
                run += __popc(bmask);
            }
            if (lane == 0) out_len[g] = run;
        }
    }
}

// filter_by_containment_rules for ONE rule (child class != parent class), one CTA per group.
// rem_in / rem_out: per list slot, as in k_containment_rule (rem_out starts as a copy of rem_in).
__global__ void __launch_bounds__(EMIA_FUSED_THREADS) k_containment_fused(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off,
    const int32_t* __restrict__ bbox, const int32_t* __restrict__ area, const int32_t* __restrict__ classes,
    const int32_t* __restrict__ cap_off, const int32_t* __restrict__ in_len, const int32_t* __restrict__ in_idx, int child,
    int parent, double thr, int smem_cap, const int32_t* __restrict__ rem_in, int32_t* __restrict__ rem_out) {
    extern __shared__ __align__(16) unsigned char fused_smem[];
    __shared__ int s_any_parent, s_qn;
    const EmiaFusedSmem S = emia_fused_carve(fused_smem, smem_cap);
    int* best = S.fidx;      // per slot: largest intersection with a live parent
    int* role = S.pos;       // 1 = live child, 2 = live parent, 0 = neither
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = EMIA_FUSED_THREADS / 32;
    const int base = cap_off[g];
    const int len = in_len[g];
    if (tid == 0) { s_any_parent = 0; s_qn = 0; }
    __syncthreads();
    int any = 0;
    for (int k = tid; k < len; k += EMIA_FUSED_THREADS) {
        const int inst = in_idx[base + k];
        const int cl = classes[inst];
        const int rm = rem_in[base + k];
        S.inst[k] = inst; S.area[k] = area[inst]; S.bb[k] = ((const int4*)bbox)[inst];
        const emia_inst_meta m = meta[inst];
        S.geo[k] = make_int4(m.ry0, m.wc0, m.ch, m.cw);
        S.coff[k] = crop_off[inst];
        best[k] = 0;
        role[k] = rm ? 0 : (cl == child ? 1 : (cl == parent ? 2 : 0));
        any |= (cl == parent);          // presence of the parent class counts removed members too (reference: list scan)
    }
    if (__any_sync(0xffffffffu, any) && lane == 0) s_any_parent = 1;
    __syncthreads();
    const int any_parent = s_any_parent;
    // children that cannot be judged by intersection are decided here; the others and the live parents are sorted by x_min
    int P = 32;
    while (P < len) P <<= 1;
    for (int k = tid; k < P; k += EMIA_FUSED_THREADS) {
        uint32_t xk = 0xFFFFFFFFu;
        if (k < len) {
            if (role[k] == 1) {
                if (!any_parent || S.area[k] <= 0 || S.bb[k].x < 0) { rem_out[base + k] = 1; role[k] = 0; }
                else xk = ((uint32_t)S.bb[k].y << 16) | (uint32_t)k;
            } else if (role[k] == 2 && S.bb[k].x >= 0) xk = ((uint32_t)S.bb[k].y << 16) | (uint32_t)k;
        }
        S.xs[k] = xk;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int st = size >> 1; st > 0; st >>= 1) {
            for (int t = tid; t < (P >> 1); t += EMIA_FUSED_THREADS) {
                const int lo = ((t & ~(st - 1)) << 1) | (t & (st - 1));      // st is a power of two
                const int hi = lo + st;
                const bool up = ((lo & size) == 0);
                const uint32_t xa = S.xs[lo], xb = S.xs[hi];
                if ((xa > xb) == up) { S.xs[lo] = xb; S.xs[hi] = xa; }
            }
            __syncthreads();
        }
    }
    // sweep: every bbox-overlapping (child, parent) pair is met once, from the member with the smaller x_min
    for (int i = tid; i < len; i += EMIA_FUSED_THREADS) {
        const uint32_t xi = S.xs[i];
        if (xi == 0xFFFFFFFFu) continue;
        const int a = (int)(xi & 0xFFFFu);
        const int ra = role[a];
        const int4 ba4 = S.bb[a];
        const int ba[4] = {ba4.x, ba4.y, ba4.z, ba4.w};
        for (int j = i + 1; j < len; ++j) {
            const uint32_t xj = S.xs[j];
            if (xj == 0xFFFFFFFFu || (int)(xj >> 16) > ba4.w) break;
            const int b = (int)(xj & 0xFFFFu);
            if (role[b] == ra) continue;
            const int4 bb4 = S.bb[b];
            const int bb[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
            if (!emia_bbox_overlap(ba, bb)) continue;
            const int c = (ra == 1) ? a : b, p = (ra == 1) ? b : a;
            const int q = atomicAdd(&s_qn, 1);
            if (q < EMIA_FUSED_QUEUE) S.queue[q] = (uint32_t)c | ((uint32_t)p << 16);
            else {
                const EmiaCropRef rc = emia_crop_ref(crops, meta, crop_off, S.inst[c]);
                const EmiaCropRef rp = emia_crop_ref(crops, meta, crop_off, S.inst[p]);
                atomicMax(&best[c], emia_crop_inter(rc, rp));
            }
        }
    }
    __syncthreads();
    {
        const int nq = min(s_qn, EMIA_FUSED_QUEUE);
        const int sub = lane >> 3, sl = lane & 7;
        for (int q0 = warp * 4; q0 < nq; q0 += nwarps * 4) {
            const int q = q0 + sub;
            const bool valid = q < nq;
            const uint32_t e = valid ? S.queue[q] : 0u;
            const int c = (int)(e & 0xFFFFu), p = (int)(e >> 16);
            const int inter = emia_crop_inter_oct(crops, S.geo[c], S.coff[c], S.geo[p], S.coff[p], valid, sl);
            if (sl == 0 && inter > 0) atomicMax(&best[c], inter);
        }
    }
    __syncthreads();
    // max_p inter_p / ac == (max_p inter_p) / ac: the division by the child's own area is monotone
    for (int c = tid; c < len; c += EMIA_FUSED_THREADS) {
        if (role[c] != 1) continue;
        if ((double)best[c] / (double)S.area[c] < thr) rem_out[base + c] = 1;
    }
}
