// emia_group_sparse.cuh — K4 for groups that do not fit one SM's shared memory (the global de-dup of a whole micrograph:
// src/functions/inference.py:2472 sees every tile's instances of an 8192 x 8192 image in ONE list) — part of emia_kernels.cu.
//
// Same semantics as the fused kernels (emia_group_fused.cuh), different data structures: nothing here is O(cap^2) in memory.
//   select -> filtered index (block scan) -> sort keys ->
//   rank by counting (visit order AND x_min order in one pass over shared-memory key tiles) ->
//   candidate pairs by a sweep over the x_min-sorted members (one warp per member, the warp computes each mask
//   intersection cooperatively) -> SPARSE edge list (ranks a, b with IoU > thr) ->
//   resolve: the reference's sequential greedy loop is a DAG evaluation — rank q is kept iff no lower rank p with an edge to q
//   (and q >= first(p), the Q2 slice start) is kept — solved by fixed-point iteration over the edge list in one CTA per group
//   (the number of sweeps is the longest suppression chain, 2-4 for duplicate clusters), instead of a one-warp walk over all ranks.
// An edge list that overflows its capacity (EMIA_SP_EDGE_FACTOR edges per member: pathological inputs such as hundreds of
// identical masks) switches that group to a direct greedy loop that evaluates the pairs on the fly — slow, exact.
#pragma once

#define EMIA_SP_EDGE_FACTOR 16
#define EMIA_SP_TILE 1024
#define EMIA_SP_THREADS 256

struct EmiaSparseWs {
    int32_t *ok, *fidx, *pos, *order, *xorder, *rem, *best, *racc, *xacc;   // L each
    uint64_t *k1, *kx;                                         // L each
    uint32_t* k2;                                              // L
    int32_t *nok, *nx, *ecount, *anyp;                         // G each
    int2* edges;                                               // EMIA_SP_EDGE_FACTOR * L
    uint8_t *status, *flagk, *flagu;                           // L each
};

static size_t emia_sparse_ws_bytes(size_t L, size_t G) {
    size_t b = 0;
    b += 9 * emia_align_up(L * 4 + 16, 256);
    b += 2 * emia_align_up(L * 8 + 16, 256);
    b += emia_align_up(L * 4 + 16, 256);
    b += 4 * emia_align_up(G * 4 + 16, 256);
    b += emia_align_up((size_t)EMIA_SP_EDGE_FACTOR * L * 8 + 16, 256);
    b += 3 * emia_align_up(L + 16, 256);
    return b + 512;
}

static int emia_sparse_carve(void* workspace, size_t bytes, size_t L, size_t G, EmiaSparseWs* ws) {
    unsigned char* p = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    unsigned char* end = (unsigned char*)workspace + bytes;
    auto take = [&](size_t nbytes) -> void* { void* r = p; p += emia_align_up(nbytes, 256); return r; };
    ws->ok = (int32_t*)take(L * 4 + 16); ws->fidx = (int32_t*)take(L * 4 + 16); ws->pos = (int32_t*)take(L * 4 + 16);
    ws->order = (int32_t*)take(L * 4 + 16); ws->xorder = (int32_t*)take(L * 4 + 16); ws->rem = (int32_t*)take(L * 4 + 16);
    ws->best = (int32_t*)take(L * 4 + 16);
    // racc / xacc are contiguous (one memset clears both)
    ws->racc = (int32_t*)take(L * 4 + 16); ws->xacc = (int32_t*)take(L * 4 + 16);
    ws->k1 = (uint64_t*)take(L * 8 + 16); ws->kx = (uint64_t*)take(L * 8 + 16);
    ws->k2 = (uint32_t*)take(L * 4 + 16);
    // the four per-group counters are contiguous (one memset clears them)
    ws->nok = (int32_t*)take(G * 4 + 16); ws->nx = (int32_t*)take(G * 4 + 16); ws->ecount = (int32_t*)take(G * 4 + 16);
    ws->anyp = (int32_t*)take(G * 4 + 16);
    ws->edges = (int2*)take((size_t)EMIA_SP_EDGE_FACTOR * L * 8 + 16);
    ws->status = (uint8_t*)take(L + 16); ws->flagk = (uint8_t*)take(L + 16); ws->flagu = (uint8_t*)take(L + 16);
    return p > end ? -1 : 0;
}

// filtered index = exclusive count of ok slots in list order; one CTA per group
__global__ void __launch_bounds__(1024) k_sp_fidx(const int32_t* __restrict__ cap_off, const int32_t* __restrict__ ok,
                                                  int32_t* __restrict__ fidx, int32_t* __restrict__ nok) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int g = blockIdx.x;
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int k0 = 0; k0 < cap; k0 += 1024) {
        const int k = k0 + threadIdx.x;
        const int v = (k < cap) ? (ok[base + k] != 0) : 0;
        const unsigned b = __ballot_sync(0xffffffffu, v);
        if (lane == 0) s_w[warp] = __popc(b);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (k < cap) fidx[base + k] = before + __popc(b & ((1u << lane) - 1u));
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_w[w]; s_run += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) nok[g] = s_run;
}

__device__ __forceinline__ uint32_t emia_score_sortable(float f) {
    uint32_t bits = (f == 0.0f) ? 0u : __float_as_uint(f);          // -0.0 and +0.0 compare equal in the reference
    return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);       // larger float <=> larger unsigned
}

// sort keys.  Visit order: the slot with the LARGER (k1, k2) comes first.  x order: the slot with the SMALLER kx comes first.
//   rank_mode 0 (np.argsort(scores)[::-1]): score descending, ties by filtered index descending
//   rank_mode 1: list order (no keys)        rank_mode 2: class ascending, score descending, ties by filtered index ascending
// part_mode 0: participants of the pair sweep = ok slots with a non-empty mask; 2: additionally an active overlap rule.
__global__ void k_sp_keys(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_idx, int L, int rank_mode,
                          int part_mode, const float* __restrict__ scores, const int32_t* __restrict__ classes,
                          const int32_t* __restrict__ bbox, const int32_t* __restrict__ area, const int32_t* __restrict__ rule_active,
                          int num_classes, const int32_t* __restrict__ ok, const int32_t* __restrict__ fidx, uint64_t* __restrict__ k1,
                          uint32_t* __restrict__ k2, uint64_t* __restrict__ kx, int32_t* __restrict__ nx) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    uint64_t a = 0ull, x = ~0ull;
    uint32_t b = 0u;
    if (ok[s]) {
        const int g = emia_find_group(cap_off, G, s);
        const int inst = in_idx[s];
        const uint32_t fi = (uint32_t)fidx[s];
        if (rank_mode == 0) { a = (uint64_t)emia_score_sortable(scores[inst]) + 1ull; b = fi; }
        else if (rank_mode == 2) {
            const uint32_t cu = (uint32_t)classes[inst] + 0x80000000u;
            a = ((uint64_t)(~cu) << 32) | (uint64_t)emia_score_sortable(scores[inst]);
            a += 1ull;
            b = ~fi;
        } else { a = 1ull; }
        bool part = area[inst] > 0 && bbox[4 * inst] >= 0;
        if (part && part_mode == 2) {
            const int c = classes[inst];
            part = (c >= 0 && c < num_classes && rule_active[c]);
        }
        if (part) {
            x = ((uint64_t)(uint32_t)bbox[4 * inst + 1] << 32) | (uint64_t)(uint32_t)(s - cap_off[g]);
            atomicAdd(&nx[g], 1);
        }
    }
    k1[s] = a; k2[s] = b; kx[s] = x;
}

// rank by counting over shared-memory key tiles, split two ways for parallelism: blockIdx.x = 256 consecutive slots (which may
// belong to several groups: the tiles of every group the CTA touches are streamed through shared memory, a thread only counts
// in its own group's tiles), blockIdx.y = every gridDim.y-th key tile.  Partial counts are accumulated with atomicAdd into
// racc / xacc (cleared by the caller); k_sp_rank_finish turns them into pos / order / xorder.
// kNarrow: the keys are compared as 32-bit values in shared memory (score key; x_min << 16 | slot) — half the shared-memory
// traffic and compare instructions; valid when the visit order does not involve the class (rank modes 0 / 1) and every group has
// at most 65 536 slots.
template <bool kNarrow>
__global__ void __launch_bounds__(EMIA_SP_THREADS) k_sp_rank(const int32_t* __restrict__ cap_off, int G, int L,
                                                             const int32_t* __restrict__ in_len, int do_rank,
                                                             const uint64_t* __restrict__ k1, const uint32_t* __restrict__ k2,
                                                             const uint64_t* __restrict__ kx, int32_t* __restrict__ racc,
                                                             int32_t* __restrict__ xacc) {
    typedef typename std::conditional<kNarrow, uint32_t, uint64_t>::type key_t;
    __shared__ key_t s1[EMIA_SP_TILE];
    __shared__ key_t sx[EMIA_SP_TILE];
    __shared__ uint32_t s2[EMIA_SP_TILE];
    auto nar1 = [](uint64_t a) -> key_t { return kNarrow ? (key_t)(a > 0xFFFFFFFFull ? 0xFFFFFFFFull : a) : (key_t)a; };
    auto narx = [](uint64_t x) -> key_t {
        if (!kNarrow) return (key_t)x;
        return x == ~0ull ? (key_t)0xFFFFFFFFu : (key_t)((((uint32_t)(x >> 32)) << 16) | ((uint32_t)x & 0xFFFFu));
    };
    const int s_lo = blockIdx.x * EMIA_SP_THREADS;
    const int s_hi = min(L, s_lo + EMIA_SP_THREADS) - 1;
    const int s = s_lo + threadIdx.x;
    int my_g = -1;
    uint64_t r1 = 0ull, rx = ~0ull;
    uint32_t m2 = 0u;
    if (s < L) { my_g = emia_find_group(cap_off, G, s); r1 = do_rank ? k1[s] : 0ull; m2 = do_rank ? k2[s] : 0u; rx = kx[s]; }
    const bool mine_ok = do_rank && (r1 != 0ull), mine_part = (rx != ~0ull);
    if (!__syncthreads_or(mine_ok || mine_part)) return;            // a chunk of dead slots
    const key_t m1 = nar1(r1), mx = narx(rx);
    const int g_lo = emia_find_group(cap_off, G, s_lo), g_hi = emia_find_group(cap_off, G, s_hi);
    int rank = 0, xr = 0;
    for (int g = g_lo; g <= g_hi; ++g) {
        const int base = cap_off[g];
        const int len = in_len[g];                                   // slots beyond the live length carry null keys
        for (int t0 = blockIdx.y * EMIA_SP_TILE; t0 < len; t0 += gridDim.y * EMIA_SP_TILE) {
            const int cnt = min(EMIA_SP_TILE, len - t0);
            for (int j = threadIdx.x; j < cnt; j += EMIA_SP_THREADS) {
                if (do_rank) { s1[j] = nar1(k1[base + t0 + j]); s2[j] = k2[base + t0 + j]; }
                sx[j] = narx(kx[base + t0 + j]);
            }
            __syncthreads();
            if (my_g == g) {
                if (mine_ok) {
#pragma unroll 8
                    for (int j = 0; j < cnt; ++j) {
                        const key_t a = s1[j];
                        rank += (a > m1) || (a == m1 && s2[j] > m2);
                    }
                }
                if (mine_part) {
#pragma unroll 8
                    for (int j = 0; j < cnt; ++j) xr += (sx[j] < mx);
                }
            }
            __syncthreads();
        }
    }
    if (s >= L) return;
    if (rank) atomicAdd(&racc[s], rank);
    if (xr) atomicAdd(&xacc[s], xr);
}
__global__ void k_sp_rank_finish(const int32_t* __restrict__ cap_off, int G, int L, int do_rank, const int32_t* __restrict__ ok,
                                 const int32_t* __restrict__ fidx, const uint64_t* __restrict__ kx, const int32_t* __restrict__ racc,
                                 const int32_t* __restrict__ xacc, int32_t* __restrict__ pos, int32_t* __restrict__ order,
                                 int32_t* __restrict__ xorder) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    const bool part = (kx[s] != ~0ull);
    const bool live = pos ? (ok[s] != 0) : false;
    if (!part && !live) { if (pos) pos[s] = -1; return; }
    const int base = cap_off[emia_find_group(cap_off, G, s)];
    if (pos) {
        if (live) {
            const int r = do_rank ? racc[s] : fidx[s];
            pos[s] = r;
            order[base + r] = s;
        } else pos[s] = -1;
    }
    if (part) xorder[base + xacc[s]] = s;
}

// candidate pairs: one warp per member in x_min order; it meets the members that start before its x_max.
//   pair_mode 0: smart de-dup (same class, Q1 bbox test, IoU > thr)     1: iou() de-dup (any class, IoU > thr)
//   pair_mode 2: overlap rules (same class, IoU > rule_max_iou[class])   4: containment (child x parent -> atomicMax(best[child]))
__global__ void __launch_bounds__(EMIA_SP_THREADS) k_sp_pairs(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off,
    const int32_t* __restrict__ bbox, const int32_t* __restrict__ area, const int32_t* __restrict__ classes,
    const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_idx, int L, int pair_mode, double thr,
    const double* __restrict__ rule_max_iou, const int32_t* __restrict__ role, const int32_t* __restrict__ xorder,
    const int32_t* __restrict__ nx, const int32_t* __restrict__ pos, int2* __restrict__ edges, int32_t* __restrict__ ecount,
    int32_t* __restrict__ best) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= L) return;
    const int g = emia_find_group(cap_off, G, w);
    const int base = cap_off[g];
    const int i = w - base;
    const int n = nx[g];
    if (i >= n) return;
    const int sa = xorder[base + i];
    const int ia = in_idx[sa];
    const int4 ba4 = ((const int4*)bbox)[ia];
    const int ba[4] = {ba4.x, ba4.y, ba4.z, ba4.w};
    const int ca = (pair_mode == 0 || pair_mode == 2) ? classes[ia] : 0;
    const int ra_role = (pair_mode == 4) ? role[sa] : 0;
    const double th = (pair_mode == 2) ? rule_max_iou[ca] : thr;
    const int aa = area[ia];
    const EmiaCropRef ra = emia_crop_ref(crops, meta, crop_off, ia);
    const long long ecap = (long long)EMIA_SP_EDGE_FACTOR * (cap_off[g + 1] - base);
    int2* eg = edges + (long long)EMIA_SP_EDGE_FACTOR * base;
    for (int j0 = i + 1; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < n;
        int sb = 0, ib = 0;
        int4 bb4 = make_int4(-1, 0x7fffffff, -1, -1);
        if (valid) { sb = xorder[base + j]; ib = in_idx[sb]; bb4 = ((const int4*)bbox)[ib]; }
        const bool beyond = !valid || bb4.y > ba4.w;                 // starts right of a's x_max: so do all later members
        bool pass = !beyond;
        if (pass) {
            const int bb[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
            if (pair_mode == 0 || pair_mode == 2) pass = (classes[ib] == ca);
            if (pass && pair_mode == 4) pass = (role[sb] != ra_role);
            if (pass && pair_mode == 0) pass = emia_bbox_overlap_q1(ba, bb);
            if (pass) pass = emia_bbox_overlap(ba, bb);
        }
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int l = __ffs((int)m) - 1;
            m &= m - 1u;
            const int ib_l = __shfl_sync(0xffffffffu, ib, l);
            const int sb_l = __shfl_sync(0xffffffffu, sb, l);
            const EmiaCropRef rb = emia_crop_ref(crops, meta, crop_off, ib_l);
            const int inter = emia_crop_inter_warp(ra, rb, lane);
            if (lane == 0 && inter > 0) {
                if (pair_mode == 4) {
                    atomicMax(&best[(ra_role == 1) ? sa : sb_l], inter);
                } else {
                    const int uni = aa + area[ib_l] - inter;
                    if ((double)inter / (double)uni > th) {
                        const int e = atomicAdd(&ecount[g], 1);
                        if (e < ecap) eg[e] = make_int2(pos[sa], pos[sb_l]);
                    }
                }
            }
        }
        if (__any_sync(0xffffffffu, beyond)) break;
    }
}

// does `a` (slot sa) suppress slot sb under the pair relation?  (direct evaluation, used by the overflow fallback)
__device__ bool emia_sp_pair_over(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, const int32_t* bbox,
                                  const int32_t* area, const int32_t* classes, const int32_t* in_idx, int pair_mode, double thr,
                                  const double* rule_max_iou, const int32_t* rule_active, int num_classes, int sa, int sb) {
    const int ia = in_idx[sa], ib = in_idx[sb];
    const int* ba = bbox + 4 * ia;
    const int* bb = bbox + 4 * ib;
    double th = thr;
    if (pair_mode == 0 || pair_mode == 2) {
        const int ca = classes[ia];
        if (classes[ib] != ca) return false;
        if (pair_mode == 2) {
            if (ca < 0 || ca >= num_classes || !rule_active[ca]) return false;
            th = rule_max_iou[ca];
        }
    }
    if (pair_mode == 0 && !emia_bbox_overlap_q1(ba, bb)) return false;
    if (!emia_bbox_overlap(ba, bb)) return false;
    const int aa = area[ia], ab = area[ib];
    if (aa <= 0 || ab <= 0) return false;
    const int inter = emia_crop_inter(emia_crop_ref(crops, meta, crop_off, ia), emia_crop_ref(crops, meta, crop_off, ib));
    if (inter == 0) return false;
    return (double)inter / (double)(aa + ab - inter) > th;
}

// resolve: one CTA per group.  status[rank]: 0 undecided, 1 kept, 2 removed.
//   q2 != 0: a keeper with FILTERED INDEX idx suppresses ranks >= idx + 1 (deduplicate_masks_smart's slice, Q2); else ranks > its own.
//   ordered_out != 0: survivors in rank (keep) order, else in list order.
__global__ void __launch_bounds__(1024) k_sp_resolve(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off,
    const int32_t* __restrict__ bbox, const int32_t* __restrict__ area, const int32_t* __restrict__ classes,
    const int32_t* __restrict__ cap_off, const int32_t* __restrict__ in_idx, int pair_mode, double thr,
    const double* __restrict__ rule_max_iou, const int32_t* __restrict__ rule_active, int num_classes,
    const int32_t* __restrict__ nok, const int32_t* __restrict__ ok, const int32_t* __restrict__ fidx,
    const int32_t* __restrict__ pos, const int32_t* __restrict__ order, int2* __restrict__ edges, const int32_t* __restrict__ ecount,
    uint8_t* __restrict__ status_all, uint8_t* __restrict__ flagk_all, uint8_t* __restrict__ flagu_all, int q2, int ordered_out,
    int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    __shared__ int s_w[32];
    __shared__ int s_run;
    const int g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = cap_off[g], cap = cap_off[g + 1] - base;
    const int n = nok[g];
    uint8_t* status = status_all + base;
    uint8_t* flagk = flagk_all + base;
    uint8_t* flagu = flagu_all + base;
    const long long ecap = (long long)EMIA_SP_EDGE_FACTOR * cap;
    int2* eg = edges + (long long)EMIA_SP_EDGE_FACTOR * base;
    const int E_raw = ecount[g];
    for (int p = tid; p < n; p += 1024) status[p] = 0;
    __syncthreads();
    if ((long long)E_raw > ecap) {
        // the edge list overflowed: direct greedy, pairs evaluated on the fly (exact, slow; pathological inputs only)
        for (int p = 0; p < n; ++p) {
            if (status[p] == 2) { continue; }                          // uniform: written before the last barrier
            const int sa = order[base + p];
            const int first = q2 ? (fidx[sa] + 1) : (p + 1);
            for (int q = max(first, 0) + tid; q < n; q += 1024) {
                if (q == p || status[q] != 0 || q < p) continue;
                if (emia_sp_pair_over(crops, meta, crop_off, bbox, area, classes, in_idx, pair_mode, thr, rule_max_iou, rule_active,
                                      num_classes, sa, order[base + q]))
                    status[q] = 2;
            }
            __syncthreads();
            if (tid == 0) status[p] = 1;
            __syncthreads();
        }
    } else {
        const int E = E_raw;
        // orient the edges: lo -> hi in rank; an edge whose hi lies before the keeper's slice start never matters
        for (int e = tid; e < E; e += 1024) {
            const int2 ed = eg[e];
            const int lo = min(ed.x, ed.y), hi = max(ed.x, ed.y);
            const int first = q2 ? (fidx[order[base + lo]] + 1) : (lo + 1);
            eg[e] = (hi >= first && lo != hi) ? make_int2(lo, hi) : make_int2(-1, -1);
        }
        __syncthreads();
        for (;;) {
            for (int p = tid; p < n; p += 1024) { flagk[p] = 0; flagu[p] = 0; }
            __syncthreads();
            for (int e = tid; e < E; e += 1024) {
                const int2 ed = eg[e];
                if (ed.x < 0 || status[ed.y] != 0) continue;
                const uint8_t sl = status[ed.x];
                if (sl == 1) flagk[ed.y] = 1;
                else if (sl == 0) flagu[ed.y] = 1;
            }
            __syncthreads();
            int undecided = 0;
            for (int p = tid; p < n; p += 1024) {
                if (status[p] != 0) continue;
                if (flagk[p]) status[p] = 2;
                else if (!flagu[p]) status[p] = 1;
                else undecided = 1;
            }
            if (!__syncthreads_or(undecided)) break;
        }
    }
    // survivors
    if (tid == 0) s_run = 0;
    __syncthreads();
    const int total = ordered_out ? n : cap;
    for (int k0 = 0; k0 < total; k0 += 1024) {
        const int k = k0 + tid;
        int keep = 0, inst = 0;
        if (k < total) {
            if (ordered_out) { keep = (status[k] == 1); inst = in_idx[order[base + k]]; }
            else if (ok[base + k]) { keep = (status[pos[base + k]] == 1); inst = in_idx[base + k]; }
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(b);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (keep) out_idx[base + before + __popc(b & ((1u << lane) - 1u))] = inst;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_w[w]; s_run += t; }
        __syncthreads();
    }
    if (tid == 0) out_len[g] = s_run;
}

// ---- containment (child class != parent class) -----------------------------------------------------------------------
// role[s]: 1 = live child that can be judged by intersection, 2 = live parent with a bbox, 3 = live child that cannot, 0 = neither
__global__ void k_sp_contain_prepare(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                     const int32_t* __restrict__ in_idx, int L, const int32_t* __restrict__ classes,
                                     const int32_t* __restrict__ bbox, const int32_t* __restrict__ area, int child, int parent,
                                     const int32_t* __restrict__ rem_in, int32_t* __restrict__ role, int32_t* __restrict__ best,
                                     uint64_t* __restrict__ kx, int32_t* __restrict__ nx, int32_t* __restrict__ anyp) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    const int g = emia_find_group(cap_off, G, s);
    int r = 0;
    uint64_t x = ~0ull;
    if (s - cap_off[g] < in_len[g]) {
        const int inst = in_idx[s];
        const int cl = classes[inst];
        if (cl == parent) atomicOr(&anyp[g], 1);        // presence of the parent class counts removed members too (list scan)
        if (!rem_in[s]) {
            const bool has = area[inst] > 0 && bbox[4 * inst] >= 0;
            if (cl == child) r = has ? 1 : 3;
            else if (cl == parent && bbox[4 * inst] >= 0) r = 2;
            if (r == 1 || r == 2) {
                x = ((uint64_t)(uint32_t)bbox[4 * inst + 1] << 32) | (uint64_t)(uint32_t)(s - cap_off[g]);
                atomicAdd(&nx[g], 1);
            }
        }
    }
    role[s] = r; best[s] = 0; kx[s] = x;
}
__global__ void k_sp_contain_decide(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_idx, int L,
                                    const int32_t* __restrict__ area, double thr, const int32_t* __restrict__ role,
                                    const int32_t* __restrict__ best, const int32_t* __restrict__ anyp, int32_t* __restrict__ rem_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L) return;
    const int r = role[s];
    if (r != 1 && r != 3) return;
    const int g = emia_find_group(cap_off, G, s);
    if (r == 3 || !anyp[g]) { rem_out[s] = 1; return; }
    if ((double)best[s] / (double)area[in_idx[s]] < thr) rem_out[s] = 1;
}
