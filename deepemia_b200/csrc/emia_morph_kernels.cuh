// emia_morph_kernels.cuh — K2: bit-packed mask clean-up morphology (part of emia_kernels.cu).
//
// Replaces (reference): scipy.ndimage.binary_fill_holes + skimage erosion/dilation (3x3 cross; skimage 0.19.3 =
// scipy grey_erosion/grey_dilation, mode 'reflect' => out-of-frame neighbours are ignored) + skimage.measure.label as used by
//   postprocess_masks            src/utils/mask_utils.py:70-84     fill -> dilate -> erode -> first-come overlap removal -> >1 component => zero
//   process_masks_parallel       src/functions/inference.py:189-203 fill -> erode -> dilate
//   postprocess_masks_universal  src/functions/inference.py:1778-1806 fill -> erode [-> dilate], keep if sum >= min size
//
// One warp per instance.  The crop is copied into a padded plane (one extra row above/below, one extra WORD left/right)
// in SHARED memory (global workspace only for masks larger than ~1000 words); lanes own rows.  All operators are word-parallel shifts / AND / OR; fill-holes and the
// connected-component test are flood fills iterated to a fixed point (row-local run flooding by carry propagation, so the
// iteration count is the number of direction changes of the longest path, not its length).
#pragma once

struct EmiaPad {
    uint32_t* p;      // (ch + 2) x (cw + 2) words
    int rows, words;  // padded sizes
};
__device__ __forceinline__ uint32_t& emia_pad_at(const EmiaPad& P, int pr, int pc) { return P.p[pr * P.words + pc]; }

// bit mask of the pixels of padded word pc that lie inside the frame columns [0, W)
__device__ __forceinline__ uint32_t emia_valid_cols(int wc0, int pc, int W) {
    const int x0 = (wc0 + pc - 1) * 32;
    if (x0 + 32 <= 0 || x0 >= W) return 0u;
    uint32_t m = 0xffffffffu;
    if (x0 < 0) m = 0u;   // cannot happen partially: x0 is a multiple of 32
    if (x0 + 32 > W) m &= (W - x0 >= 32) ? 0xffffffffu : ((1u << (W - x0)) - 1u);
    return m;
}
// flood the seed bits s through the set bits of m inside one word, both directions (carry propagation)
__device__ __forceinline__ uint32_t emia_flood_word(uint32_t s, uint32_t m) {
    s &= m;
    // smear along runs of m: 5 doubling steps each way
    uint32_t r = s;
    uint32_t mm = m;
    r |= (r << 1) & mm; mm &= mm << 1;
    r |= (r << 2) & mm; mm &= mm << 2;
    r |= (r << 4) & mm; mm &= mm << 4;
    r |= (r << 8) & mm; mm &= mm << 8;
    r |= (r << 16) & mm;
    mm = m;
    r |= (r >> 1) & mm; mm &= mm >> 1;
    r |= (r >> 2) & mm; mm &= mm >> 2;
    r |= (r >> 4) & mm; mm &= mm >> 4;
    r |= (r >> 8) & mm; mm &= mm >> 8;
    r |= (r >> 16) & mm;
    return r;
}

// R <- flood of seeds (already in R) through ALLOWED, 4- or 8-connected, to a fixed point.
// Planes up to 32 words wide (every shared-memory plane): LANES OWN WORD COLUMNS and the rows are walked top-down, then
// bottom-up (Gauss-Seidel: a row sees the final value of the row before it, so one sweep carries a front across the whole
// plane; the number of sweeps is the number of direction changes of the longest path — 2-3 for particle-like shapes — instead
// of the number of rows).  Inside a row the front runs along the allowed runs by carry propagation in the word and by
// shuffles across word boundaries.  Wider planes (global workspace) keep the row-per-lane Jacobi sweep.
__device__ __forceinline__ uint32_t emia_flood_row(uint32_t in, uint32_t allow, int lane, int words) {
    uint32_t nv = emia_flood_word(in & allow, allow);
    for (;;) {
        uint32_t l = __shfl_up_sync(0xffffffffu, nv, 1) >> 31;
        uint32_t r = __shfl_down_sync(0xffffffffu, nv, 1) << 31;
        if (lane == 0) l = 0u;
        if (lane >= words - 1) r = 0u;
        const uint32_t add = (l | r) & allow & ~nv;
        if (!__any_sync(0xffffffffu, add != 0u)) break;
        nv = emia_flood_word(nv | add, allow);
    }
    return nv;
}
__device__ void emia_flood_cols(const EmiaPad& R, const EmiaPad& ALLOWED, int conn8, int lane) {
    const int words = R.words, rows = R.rows;
    const bool act = lane < words;
    for (;;) {
        bool changed = false;
        for (int dir = 0; dir < 2; ++dir) {
            uint32_t prev = 0u;                                   // final value of the row visited before this one
            for (int k = 0; k < rows; ++k) {
                const int pr = dir ? (rows - 1 - k) : k;
                const uint32_t allow = act ? emia_pad_at(ALLOWED, pr, lane) : 0u;
                const uint32_t cur = act ? emia_pad_at(R, pr, lane) : 0u;
                uint32_t from = prev;
                if (conn8) {
                    uint32_t l = prev << 1, r = prev >> 1;
                    const uint32_t pl = __shfl_up_sync(0xffffffffu, prev, 1), pn = __shfl_down_sync(0xffffffffu, prev, 1);
                    if (lane > 0) l |= pl >> 31;
                    if (lane < words - 1) r |= pn << 31;
                    from |= l | r;
                }
                const uint32_t nv = emia_flood_row(cur | from, allow, lane, words) | cur;
                if (act && nv != cur) { emia_pad_at(R, pr, lane) = nv; changed = true; }
                prev = act ? nv : 0u;
            }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, changed)) break;
    }
}
__device__ void emia_flood(const EmiaPad& R, const EmiaPad& ALLOWED, int conn8, int lane) {
    if (R.words <= 32) { emia_flood_cols(R, ALLOWED, conn8, lane); return; }
    for (;;) {
        bool changed = false;
        for (int pr = lane; pr < R.rows; pr += 32) {
            for (int pass = 0; pass < 2; ++pass) {
                for (int k = 0; k < R.words; ++k) {
                    const int pc = pass ? (R.words - 1 - k) : k;
                    const uint32_t allow = emia_pad_at(ALLOWED, pr, pc);
                    if (!allow) continue;
                    uint32_t cur = emia_pad_at(R, pr, pc);
                    uint32_t in = cur;
                    // from the rows above / below (previous sweep's values: Jacobi between rows, fine for a fixed point)
                    for (int d = -1; d <= 1; d += 2) {
                        const int qr = pr + d;
                        if (qr < 0 || qr >= R.rows) continue;
                        uint32_t v = emia_pad_at(R, qr, pc);
                        if (conn8) {
                            uint32_t l = v << 1, r = v >> 1;
                            if (pc > 0) l |= emia_pad_at(R, qr, pc - 1) >> 31;
                            if (pc + 1 < R.words) r |= emia_pad_at(R, qr, pc + 1) << 31;
                            v |= l | r;
                        }
                        in |= v;
                    }
                    // from the neighbouring words of this row
                    if (pc > 0) in |= emia_pad_at(R, pr, pc - 1) >> 31;
                    if (pc + 1 < R.words) in |= emia_pad_at(R, pr, pc + 1) << 31;
                    const uint32_t nv = emia_flood_word(in & allow, allow) | cur;
                    if (nv != cur) { emia_pad_at(R, pr, pc) = nv; changed = true; }
                }
            }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, changed)) break;
    }
}

// dst <- erosion / dilation of src with the 3x3 cross.  Out-of-frame neighbours are ignored; out-of-crop (in-frame) ones are 0.
__device__ void emia_cross_op(const EmiaPad& dst, const EmiaPad& src, int dilate, int ry0, int wc0, int H, int W, int lane) {
    for (int pr = lane; pr < src.rows; pr += 32) {
        const int y = ry0 + pr - 1;
        const bool row_ok = (y >= 0 && y < H);
        for (int pc = 0; pc < src.words; ++pc) {
            const uint32_t vc = emia_valid_cols(wc0, pc, W);
            uint32_t out = 0u;
            if (row_ok && vc) {
                const uint32_t c = emia_pad_at(src, pr, pc);
                uint32_t l = c << 1, r = c >> 1;
                if (pc > 0) l |= emia_pad_at(src, pr, pc - 1) >> 31;
                if (pc + 1 < src.words) r |= emia_pad_at(src, pr, pc + 1) << 31;
                uint32_t u = (pr > 0) ? emia_pad_at(src, pr - 1, pc) : 0u;
                uint32_t d = (pr + 1 < src.rows) ? emia_pad_at(src, pr + 1, pc) : 0u;
                // validity of the neighbours
                const uint32_t vl = (vc << 1) | ((pc > 0 ? emia_valid_cols(wc0, pc - 1, W) : 0u) >> 31);
                const uint32_t vr = (vc >> 1) | ((pc + 1 < src.words ? emia_valid_cols(wc0, pc + 1, W) : ((wc0 + pc) * 32 < W ? 0xffffffffu : 0u)) << 31);
                const bool up_ok = (y - 1 >= 0), dn_ok = (y + 1 < H);
                if (dilate) {
                    out = c | (l & vl) | (r & vr) | (up_ok ? u : 0u) | (dn_ok ? d : 0u);
                } else {
                    out = c & (l | ~vl) & (r | ~vr) & (up_ok ? u : 0xffffffffu) & (dn_ok ? d : 0xffffffffu);
                }
                out &= vc;
            }
            emia_pad_at(dst, pr, pc) = out;
        }
    }
    __syncwarp();
}

// true when some background pixel of M is enclosed by mask pixels both horizontally and vertically (a necessary condition for a
// hole).  Lanes own word columns (M.words <= 32); T is scratch (the running OR of the rows above).  ~25 instructions per row.
__device__ bool emia_may_have_holes(const EmiaPad& M, const EmiaPad& T, int lane) {
    const int words = M.words, rows = M.rows;
    const bool act = lane < words;
    uint32_t run = 0u;
    for (int r = 0; r < rows; ++r) {                       // T[r] = OR of rows 0 .. r-1
        if (act) emia_pad_at(T, r, lane) = run;
        run |= act ? emia_pad_at(M, r, lane) : 0u;
    }
    uint32_t below = 0u;
    bool found = false;
    for (int r = rows - 1; r >= 0; --r) {
        const uint32_t m = act ? emia_pad_at(M, r, lane) : 0u;
        const unsigned has = __ballot_sync(0xffffffffu, m != 0u);
        if (has) {
            const bool lower = (has & ((1u << lane) - 1u)) != 0u;            // a set pixel in a word column left of mine
            const bool higher = lane < 31 ? ((has >> (lane + 1)) != 0u) : false;
            uint32_t hull = 0xffffffffu;
            if (!lower) hull &= m ? (0xffffffffu << (__ffs((int)m) - 1)) : 0u;
            if (!higher) hull &= m ? (0xffffffffu >> __clz((int)m)) : 0u;
            const uint32_t above = act ? emia_pad_at(T, r, lane) : 0u;
            if (~m & hull & above & below) found = true;
        }
        below |= m;
    }
    __syncwarp();
    return __any_sync(0xffffffffu, found);
}

// true when every row of M holds at most one run of set pixels (then no background pixel has mask on both sides in its row,
// so the mask has no holes).  Lanes own rows; ~10 instructions per word.
__device__ bool emia_rows_single_run(const EmiaPad& M, int lane) {
    bool ok = true;
    for (int r = lane; r < M.rows; r += 32) {
        int f = 0x7fffffff, l = -1, c = 0;
        for (int w = 0; w < M.words; ++w) {
            const uint32_t v = emia_pad_at(M, r, w);
            if (v) {
                c += __popc(v);
                f = min(f, w * 32 + (__ffs((int)v) - 1));
                l = max(l, w * 32 + (31 - __clz((int)v)));
            }
        }
        if (c > 0 && c != l - f + 1) ok = false;
    }
    return !__any_sync(0xffffffffu, !ok);
}

// true when the mask in A is certainly ONE 8-connected component: every row is a single run of pixels, there is no empty row
// between non-empty rows, and the runs of consecutive rows touch (also diagonally).  Lanes own rows.  (Sufficient, not necessary.)
__device__ bool emia_single_blob(const EmiaPad& A, int lane) {
    const int words = A.words, rows = A.rows;
    bool ok = true;
    int prev_f = 0, prev_l = -1, state = 0;                // carried across 32-row chunks (lane 31 -> lane 0): last row's run, 0 = before, 1 = inside, 2 = after
    for (int r0 = 0; r0 < rows; r0 += 32) {
        const int r = r0 + lane;
        int f = 0x7fffffff, l = -1, c = 0;
        if (r < rows) {
            for (int w = 0; w < words; ++w) {
                const uint32_t v = emia_pad_at(A, r, w);
                if (v) {
                    c += __popc(v);
                    f = min(f, w * 32 + (__ffs((int)v) - 1));
                    l = max(l, w * 32 + (31 - __clz((int)v)));
                }
            }
        }
        const bool nonempty = c > 0;
        if (nonempty && c != l - f + 1) ok = false;          // more than one run in this row
        // the row above (previous lane, or the carry for lane 0)
        int pf = __shfl_up_sync(0xffffffffu, f, 1), pl = __shfl_up_sync(0xffffffffu, l, 1);
        const unsigned ne = __ballot_sync(0xffffffffu, nonempty);
        if (lane == 0) { pf = prev_f; pl = prev_l; }
        const bool prev_nonempty = lane == 0 ? (prev_l >= 0) : ((ne >> (lane - 1)) & 1u);
        if (nonempty && prev_nonempty && (f > pl + 1 || pf > l + 1)) ok = false;      // consecutive runs do not touch
        // empty rows between non-empty ones: the non-empty rows of the whole plane must be consecutive
        for (int k = 0; k < 32 && r0 + k < rows; ++k) {
            const bool n_k = (ne >> k) & 1u;
            if (state == 0 && n_k) state = 1;
            else if (state == 1 && !n_k) state = 2;
            else if (state == 2 && n_k) ok = false;
        }
        prev_f = __shfl_sync(0xffffffffu, f, 31); prev_l = __shfl_sync(0xffffffffu, l, 31);
    }
    return !__any_sync(0xffffffffu, !ok);
}

// Planes of one instance: shared memory when a padded plane fits EMIA_MORPH_SMEM_WORDS words (every particle-sized mask does:
// a 120 x 120-px mask is 122 x 6 = 732 words), else the caller's global workspace at pad_off[inst] (emia_morph_plan counts
// only those instances).  One warp per instance, EMIA_MORPH_WARPS warps per CTA, no CTA-wide barrier; k_morph is a persistent
// grid (a warp takes every gridDim.x * EMIA_MORPH_WARPS-th instance).
#define EMIA_MORPH_WARPS 6
#define EMIA_MORPH_SMEM_WORDS 1024
#define EMIA_MORPH_SMEM_BYTES (EMIA_MORPH_WARPS * 2 * EMIA_MORPH_SMEM_WORDS * 4)      // two planes per warp: 48 KB per CTA, 4 CTAs / SM
#define EMIA_MORPH_MAX_CTAS (148 * 4)                                               // persistent grid of k_morph
// a THIRD plane is only needed by the rare instances that go through the hole test / flood: it lives in a global scratch slot
// per resident warp (the first EMIA_MORPH_SCRATCH_WORDS words of `work`)
#define EMIA_MORPH_SCRATCH_WORDS (EMIA_MORPH_MAX_CTAS * EMIA_MORPH_WARPS * EMIA_MORPH_SMEM_WORDS)

// bbox / area of a crop-shaped output held in registers across the lanes (called by all 32 lanes)
struct EmiaStat { int a, ymin, xmin, ymax, xmax; };
__device__ __forceinline__ void emia_stat_init(EmiaStat& s) { s.a = 0; s.ymin = s.xmin = 0x7fffffff; s.ymax = s.xmax = -1; }
__device__ __forceinline__ void emia_stat_word(EmiaStat& s, uint32_t w, int y, int xword) {
    if (!w) return;
    s.a += __popc(w);
    s.ymin = min(s.ymin, y); s.ymax = max(s.ymax, y);
    s.xmin = min(s.xmin, xword * 32 + (__ffs((int)w) - 1));
    s.xmax = max(s.xmax, xword * 32 + (31 - __clz((int)w)));
}
__device__ __forceinline__ void emia_stat_store(EmiaStat s, int lane, int64_t inst, int32_t* bbox, int32_t* area) {
    for (int o = 16; o > 0; o >>= 1) {
        s.a += __shfl_xor_sync(0xffffffffu, s.a, o);
        s.ymin = min(s.ymin, __shfl_xor_sync(0xffffffffu, s.ymin, o)); s.xmin = min(s.xmin, __shfl_xor_sync(0xffffffffu, s.xmin, o));
        s.ymax = max(s.ymax, __shfl_xor_sync(0xffffffffu, s.ymax, o)); s.xmax = max(s.xmax, __shfl_xor_sync(0xffffffffu, s.xmax, o));
    }
    if (lane == 0) {
        if (area) area[inst] = s.a;
        if (bbox) ((int4*)bbox)[inst] = s.a > 0 ? make_int4(s.ymin, s.xmin, s.ymax, s.xmax) : make_int4(-1, -1, -1, -1);
    }
}

// ops: up to 4 operator codes per chain.  apply (optional) selects per instance: 0 = copied unchanged (the reference runs
// process_masks_parallel only on lists of more than two masks, src/functions/inference.py:1443), 1 = chain A (op0..op3),
// 2 = chain B (opb0..opb3: postprocess_masks_universal uses erosion only for small classes and an opening for the others,
// :1786-1796, so one launch serves both kinds of class).  bbox_out / area_out (optional): bbox / popcount of the result.
__global__ void __launch_bounds__(32 * EMIA_MORPH_WARPS) k_morph(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off, int64_t n,
    int H, int W, int op0, int op1, int op2, int op3, int opb0, int opb1, int opb2, int opb3, const int64_t* __restrict__ pad_off,
    uint32_t* __restrict__ work,
    const emia_inst_meta* __restrict__ meta_out, const int64_t* __restrict__ crop_off_out, uint32_t* __restrict__ crops_out,
    const int32_t* __restrict__ apply, int32_t* __restrict__ bbox_out, int32_t* __restrict__ area_out) {
    extern __shared__ uint32_t s_planes[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * EMIA_MORPH_WARPS + warp;
    for (int64_t inst = slot; inst < n; inst += (int64_t)gridDim.x * EMIA_MORPH_WARPS) {
        const emia_inst_meta m = meta[inst];
        const emia_inst_meta mo = meta_out[inst];
        EmiaStat st;
        emia_stat_init(st);
        if (m.ch <= 0 || m.cw <= 0) { emia_stat_store(st, lane, inst, bbox_out, area_out); continue; }
        const uint32_t* crop = crops + crop_off[inst];
        uint32_t* out = crops_out + crop_off_out[inst];
        const int sel = apply ? apply[inst] : 1;
        if (sel == 0) {
            // pass-through: the input bits in the output geometry (identical, or grown by one pixel), no planes involved
            for (int k = lane; k < mo.ch * mo.cw; k += 32) {
                const int r = k / mo.cw, c = k - r * mo.cw;
                const int ir = mo.ry0 + r - m.ry0, ic = mo.wc0 + c - m.wc0;
                const uint32_t v = ((unsigned)ir < (unsigned)m.ch && (unsigned)ic < (unsigned)m.cw) ? crop[(size_t)ir * m.cw + ic] : 0u;
                out[k] = v;
                emia_stat_word(st, v, mo.ry0 + r, mo.wc0 + c);
            }
            emia_stat_store(st, lane, inst, bbox_out, area_out);
            continue;
        }
        const int rows = m.ch + 2, words = m.cw + 2;
        const int plane = rows * words;
        const bool in_smem = plane <= EMIA_MORPH_SMEM_WORDS;
        uint32_t* base = in_smem ? (s_planes + warp * 2 * EMIA_MORPH_SMEM_WORDS) : (work + EMIA_MORPH_SCRATCH_WORDS + 3 * pad_off[inst]);
        EmiaPad A{base, rows, words}, B{base + plane, rows, words};
        EmiaPad C{in_smem ? (work + (size_t)slot * EMIA_MORPH_SMEM_WORDS) : (base + 2 * plane), rows, words};
        for (int k = lane; k < plane; k += 32) {
            const int pr = k / words, pc = k - pr * words;
            uint32_t v = 0u;
            if (pr >= 1 && pr <= m.ch && pc >= 1 && pc <= m.cw) v = crop[(size_t)(pr - 1) * m.cw + (pc - 1)];
            A.p[k] = v;
        }
        __syncwarp();
        const int ops[4] = {sel == 2 ? opb0 : op0, sel == 2 ? opb1 : op1, sel == 2 ? opb2 : op2, sel == 2 ? opb3 : op3};
        EmiaPad cur = A, other = B;
        for (int o = 0; o < 4; ++o) {
            const int op = ops[o];
            if (op == 0) break;
            if (op == EMIA_MORPH_FILL) {
                // Shortcut (exact): a hole pixel has mask pixels to its left AND right in its row and above AND below in its
                // column.  When no background pixel is enclosed that way — every blob-like particle — there is nothing to fill.
                if (emia_rows_single_run(cur, lane) || (words <= 32 && !emia_may_have_holes(cur, C, lane))) continue;
                // background reachable from the padded border (4-connected) ; holes = the rest of the background
                for (int k = lane; k < plane; k += 32) {
                    const int pr = k / words, pc = k - pr * words;
                    const bool ring = (pr == 0 || pr == rows - 1 || pc == 0 || pc == words - 1);
                    C.p[k] = ~cur.p[k];                       // allowed = background
                    other.p[k] = ring ? ~cur.p[k] : 0u;       // seeds
                }
                __syncwarp();
                emia_flood(other, C, 0, lane);
                for (int k = lane; k < plane; k += 32) {
                    const int pr = k / words, pc = k - pr * words;
                    const bool inner = (pr >= 1 && pr <= m.ch && pc >= 1 && pc <= m.cw);
                    other.p[k] = inner ? ~other.p[k] : 0u;    // mask | holes = everything the flood did not reach
                }
                __syncwarp();
            } else {
                emia_cross_op(other, cur, op == EMIA_MORPH_DILATE, m.ry0, m.wc0, H, W, lane);
            }
            EmiaPad t = cur; cur = other; other = t;
        }
        // output geometry: the input's, or (a chain that dilates first) the crop grown by one pixel towards every frame border —
        // out-of-frame neighbours are ignored by the erosion, so a closing can grow a mask that ends one pixel short of the border
        for (int k = lane; k < mo.ch * mo.cw; k += 32) {
            const int r = k / mo.cw, c = k - r * mo.cw;
            const uint32_t v = emia_pad_at(cur, mo.ry0 + r - (m.ry0 - 1), mo.wc0 + c - (m.wc0 - 1));
            out[k] = v;
            emia_stat_word(st, v, mo.ry0 + r, mo.wc0 + c);
        }
        emia_stat_store(st, lane, inst, bbox_out, area_out);
        __syncwarp();
    }
}

// first-come overlap removal + "more than one 8-connected component => zero" (postprocess_masks tail), one warp per list slot.
// The result (and its bbox / area when asked for) is written for list members only.
__global__ void __launch_bounds__(32 * EMIA_MORPH_WARPS) k_overlap_first_come(
    const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta, const int64_t* __restrict__ crop_off,
    const int32_t* __restrict__ bbox, const int32_t* __restrict__ cap_off, int G, int total_cap, const int32_t* __restrict__ in_len,
    const int32_t* __restrict__ in_idx, const int64_t* __restrict__ pad_off, uint32_t* __restrict__ work,
    uint32_t* __restrict__ crops_out, int32_t* __restrict__ bbox_out, int32_t* __restrict__ area_out) {
    extern __shared__ uint32_t s_planes[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * EMIA_MORPH_WARPS + warp;                       // list slot
    if (s >= total_cap) return;
    int lo = 0, hi = G;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cap_off[mid] <= s) lo = mid; else hi = mid; }
    const int g = lo, base = cap_off[g];
    if (s - base >= in_len[g]) return;
    const int inst = in_idx[s];
    const emia_inst_meta m = meta[inst];
    EmiaStat st;
    emia_stat_init(st);
    if (m.ch <= 0 || m.cw <= 0) { emia_stat_store(st, lane, inst, bbox_out, area_out); return; }
    const int rows = m.ch + 2, words = m.cw + 2, plane = rows * words;
    uint32_t* wb = (plane <= EMIA_MORPH_SMEM_WORDS) ? (s_planes + warp * 2 * EMIA_MORPH_SMEM_WORDS)
                                                      : (work + EMIA_MORPH_SCRATCH_WORDS + 3 * pad_off[inst]);
    EmiaPad A{wb, rows, words}, R{wb + plane, rows, words};
    const uint32_t* crop = crops + crop_off[inst];
    for (int k = lane; k < plane; k += 32) {
        const int pr = k / words, pc = k - pr * words;
        uint32_t v = 0u;
        if (pr >= 1 && pr <= m.ch && pc >= 1 && pc <= m.cw) v = crop[(size_t)(pr - 1) * m.cw + (pc - 1)];
        A.p[k] = v; R.p[k] = 0u;
    }
    __syncwarp();
    // remove what earlier list members cover: lanes test 32 earlier members' bboxes at once, the warp then ANDs out each hit
    const int4 bi = ((const int4*)bbox)[inst];
    const int bia[4] = {bi.x, bi.y, bi.z, bi.w};
    for (int k0 = 0; k0 < s - base; k0 += 32) {
        const int k = k0 + lane;
        int j = -1;
        if (k < s - base) {
            j = in_idx[base + k];
            const int4 bj = ((const int4*)bbox)[j];
            const int bja[4] = {bj.x, bj.y, bj.z, bj.w};
            if (!emia_bbox_overlap(bia, bja)) j = -1;
        }
        unsigned hits = __ballot_sync(0xffffffffu, j >= 0);
        while (hits) {
            const int l = __ffs((int)hits) - 1;
            hits &= hits - 1u;
            const int jj = __shfl_sync(0xffffffffu, j, l);
            const emia_inst_meta mj = meta[jj];
            const uint32_t* cj = crops + crop_off[jj];
            const int r0 = max(m.ry0, mj.ry0), r1 = min(m.ry0 + m.ch, mj.ry0 + mj.ch);
            const int c0 = max(m.wc0, mj.wc0), c1 = min(m.wc0 + m.cw, mj.wc0 + mj.cw);
            const int nw = (r1 - r0) * (c1 - c0);
            for (int t = lane; t < nw; t += 32) {
                const int r = r0 + t / (c1 - c0), c = c0 + t % (c1 - c0);
                emia_pad_at(A, r - m.ry0 + 1, c - m.wc0 + 1) &= ~cj[(size_t)(r - mj.ry0) * mj.cw + (c - mj.wc0)];
            }
            __syncwarp();
        }
    }
    // seed = first set pixel in raster order
    int first = 0x7fffffff;
    for (int k = lane; k < plane; k += 32) if (A.p[k]) { first = k; break; }
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    bool multi = false;
    if (first != 0x7fffffff && !emia_single_blob(A, lane)) {
        if (lane == 0) R.p[first] = A.p[first] & (0u - A.p[first]);   // lowest set bit
        __syncwarp();
        emia_flood(R, A, 1, lane);
        bool diff = false;
        for (int k = lane; k < plane; k += 32) diff |= (R.p[k] != A.p[k]);
        multi = __any_sync(0xffffffffu, diff);
    }
    uint32_t* out = crops_out + crop_off[inst];
    for (int k = lane; k < m.ch * m.cw; k += 32) {
        const int r = k / m.cw, c = k - r * m.cw;
        const uint32_t v = multi ? 0u : emia_pad_at(A, r + 1, c + 1);
        out[k] = v;
        emia_stat_word(st, v, m.ry0 + r, m.wc0 + c);
    }
    emia_stat_store(st, lane, inst, bbox_out, area_out);
}

// bbox + area of bit-packed crops (after morphology), one warp per instance
__global__ void __launch_bounds__(32) k_crop_stats(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                                   const int64_t* __restrict__ crop_off, int64_t n, int32_t* __restrict__ bbox,
                                                   int32_t* __restrict__ area) {
    const int lane = threadIdx.x;
    const int64_t inst = blockIdx.x;
    if (inst >= n) return;
    const emia_inst_meta m = meta[inst];
    const uint32_t* crop = crops + crop_off[inst];
    int a = 0, ymin = 0x7fffffff, xmin = 0x7fffffff, ymax = -1, xmax = -1;
    for (int k = lane; k < m.ch * m.cw; k += 32) {
        const uint32_t w = crop[k];
        if (!w) continue;
        const int r = k / m.cw, c = k - r * m.cw;
        a += __popc(w);
        ymin = min(ymin, m.ry0 + r); ymax = max(ymax, m.ry0 + r);
        xmin = min(xmin, (m.wc0 + c) * 32 + (__ffs((int)w) - 1));
        xmax = max(xmax, (m.wc0 + c) * 32 + (31 - __clz((int)w)));
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    }
    if (lane == 0) {
        area[inst] = a;
        ((int4*)bbox)[inst] = a > 0 ? make_int4(ymin, xmin, ymax, xmax) : make_int4(-1, -1, -1, -1);
    }
}

// padded plane sizes (words) per instance, for the caller's scan
__global__ void k_morph_plan(const emia_inst_meta* __restrict__ meta, int64_t n, int64_t* __restrict__ pad_words) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const emia_inst_meta m = meta[i];
    const int64_t plane = (m.ch > 0 && m.cw > 0) ? (int64_t)(m.ch + 2) * (m.cw + 2) : 0;
    pad_words[i] = plane > EMIA_MORPH_SMEM_WORDS ? plane : 0;     // smaller planes live in shared memory
}
// geometry of the result of a chain that may grow the mask by one pixel (see k_morph)
__global__ void k_morph_grow_plan(const emia_inst_meta* __restrict__ meta, int64_t n, int H, int W, emia_inst_meta* __restrict__ meta_out,
                                  int64_t* __restrict__ crop_words) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    emia_inst_meta m = meta[i];
    if (m.ch > 0 && m.cw > 0) {
        const int y0 = max(m.ry0 - 1, 0), y1 = min(m.ry0 + m.ch + 1, H);
        const int x0 = max(m.rx0 - 1, 0), x1 = min(m.rx1 + 1, W);
        m.ry0 = y0; m.ch = y1 - y0; m.rx0 = x0; m.rx1 = x1;
        m.wc0 = x0 >> 5; m.cw = ((x1 - 1) >> 5) - m.wc0 + 1;
    }
    meta_out[i] = m;
    crop_words[i] = (int64_t)m.ch * m.cw;
}
extern "C" int emia_morph_grow_plan(const emia_inst_meta* meta, int64_t n, int H, int W, emia_inst_meta* meta_out, int64_t* crop_words,
                                    void* stream) {
    if (n < 0 || H <= 0 || W <= 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph_grow_plan: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!meta || !meta_out || !crop_words) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph_grow_plan: %s", "null pointer");
    k_morph_grow_plan<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(meta, n, H, W, meta_out, crop_words);
    return emia_check_launch("emia_morph_grow_plan launch: %s");
}

extern "C" size_t emia_morph_scratch_words(void) { return (size_t)EMIA_MORPH_SCRATCH_WORDS; }
extern "C" int emia_morph_plan(const emia_inst_meta* meta, int64_t n, int64_t* pad_words, void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph_plan: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!meta || !pad_words) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph_plan: %s", "null pointer");
    k_morph_plan<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(meta, n, pad_words);
    return emia_check_launch("emia_morph_plan launch: %s");
}
static void emia_morph_smem_attr() {
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(k_morph, cudaFuncAttributeMaxDynamicSharedMemorySize, EMIA_MORPH_SMEM_BYTES);
        cudaFuncSetAttribute(k_overlap_first_come, cudaFuncAttributeMaxDynamicSharedMemorySize, EMIA_MORPH_SMEM_BYTES);
        done = true;
    }
}
extern "C" int emia_morph(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n, int H, int W,
                          const int32_t* ops_host, int32_t n_ops, const int64_t* pad_off, uint32_t* work,
                          const emia_inst_meta* meta_out, const int64_t* crop_off_out, uint32_t* crops_out,
                          const int32_t* apply_flag, int32_t* bbox_out, int32_t* area_out, void* stream) {
    // n_ops > 4: two chains — ops_host[0..3] (zero-padded) for apply_flag 1, ops_host[4..n_ops) for apply_flag 2
    int opsb[4] = {0, 0, 0, 0};
    if (n_ops > 4 && n_ops <= 8 && ops_host) {
        for (int i = 4; i < n_ops; ++i) {
            if (ops_host[i] < 0 || ops_host[i] > 3) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph: %s", "unknown operator");
            opsb[i - 4] = ops_host[i];
        }
        n_ops = 4;
        while (n_ops > 1 && ops_host[n_ops - 1] == 0) --n_ops;
    }
    if (n < 0 || n_ops < 1 || n_ops > 4 || !ops_host) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph: %s", "bad argument");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !pad_off || !work || !crops_out) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph: %s", "null pointer");
    if (!meta_out) { meta_out = meta; crop_off_out = crop_off; }
    if (!crop_off_out) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph: %s", "meta_out without crop_off_out");
    int ops[4] = {0, 0, 0, 0};
    for (int i = 0; i < n_ops; ++i) {
        if (ops_host[i] < 1 || ops_host[i] > 3) return emia_fail(EMIA_ERR_BAD_ARG, "emia_morph: %s", "unknown operator");
        ops[i] = ops_host[i];
    }
    emia_morph_smem_attr();
    const int64_t want = (n + EMIA_MORPH_WARPS - 1) / EMIA_MORPH_WARPS;
    k_morph<<<(unsigned)(want < EMIA_MORPH_MAX_CTAS ? want : EMIA_MORPH_MAX_CTAS), 32 * EMIA_MORPH_WARPS, EMIA_MORPH_SMEM_BYTES, (cudaStream_t)stream>>>(
        crops, meta, crop_off, n, H, W, ops[0], ops[1], ops[2], ops[3], opsb[0], opsb[1], opsb[2], opsb[3], pad_off, work, meta_out,
        crop_off_out, crops_out, apply_flag, bbox_out, area_out);
    return emia_check_launch("emia_morph launch: %s");
}
extern "C" int emia_overlap_first_come(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, const int32_t* bbox,
                                       const int32_t* cap_off, int32_t G, int32_t total_cap, const int32_t* in_len,
                                       const int32_t* in_idx, const int64_t* pad_off, uint32_t* work, uint32_t* crops_out,
                                       int32_t* bbox_out, int32_t* area_out, void* stream) {
    if (G < 0 || total_cap < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_overlap_first_come: %s", "bad argument");
    if (G == 0 || total_cap == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !cap_off || !in_len || !in_idx || !pad_off || !work || !crops_out)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_overlap_first_come: %s", "null pointer");
    emia_morph_smem_attr();
    k_overlap_first_come<<<(unsigned)((total_cap + EMIA_MORPH_WARPS - 1) / EMIA_MORPH_WARPS), 32 * EMIA_MORPH_WARPS, EMIA_MORPH_SMEM_BYTES,
                           (cudaStream_t)stream>>>(crops, meta, crop_off, bbox, cap_off, G, total_cap, in_len, in_idx, pad_off, work,
                                                   crops_out, bbox_out, area_out);
    return emia_check_launch("emia_overlap_first_come launch: %s");
}
extern "C" int emia_crop_stats(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, int64_t n, int32_t* bbox,
                               int32_t* area, void* stream) {
    if (n < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_crop_stats: %s", "bad n");
    if (n == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !bbox || !area) return emia_fail(EMIA_ERR_BAD_ARG, "emia_crop_stats: %s", "null pointer");
    k_crop_stats<<<(unsigned)n, 32, 0, (cudaStream_t)stream>>>(crops, meta, crop_off, n, bbox, area);
    return emia_check_launch("emia_crop_stats launch: %s");
}

// list members with area >= min_area, in list order (postprocess_masks_universal's `np.sum(final) >= min_crys_size`), one warp per group
__global__ void k_group_filter_area(const int32_t* __restrict__ cap_off, int G, const int32_t* __restrict__ in_len,
                                    const int32_t* __restrict__ in_idx, const int32_t* __restrict__ area, int min_area,
                                    int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= G) return;
    const int base = cap_off[g], len = in_len[g];
    int run = 0;
    for (int k0 = 0; k0 < len; k0 += 32) {
        const int k = k0 + lane;
        const int inst = (k < len) ? in_idx[base + k] : 0;
        const int keep = (k < len) && area[inst] >= min_area;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) out_idx[base + run + __popc(b & ((1u << lane) - 1u))] = inst;
        run += __popc(b);
    }
    if (lane == 0) out_len[g] = run;
}

// postprocess_masks' column gate (mask_utils.py:62-68): np.sum(masks, axis=(0,1)) is a vector over the W frame
// columns; K = number of columns whose total exceeds min_size; if K < len the list is truncated to its first K members
// (K == 0 => empty list).  One CTA per group, column totals in shared memory.
__global__ void __launch_bounds__(512) k_column_gate(const uint32_t* __restrict__ crops, const emia_inst_meta* __restrict__ meta,
                                                     const int64_t* __restrict__ crop_off, const int32_t* __restrict__ cap_off,
                                                     const int32_t* __restrict__ in_len, const int32_t* __restrict__ in_idx, int W,
                                                     int min_size, int32_t* __restrict__ out_len, int32_t* __restrict__ out_idx) {
    extern __shared__ int s_col[];
    __shared__ int s_cnt;
    const int g = blockIdx.x;
    const int base = cap_off[g], len = in_len[g];
    for (int x = threadIdx.x; x < W; x += blockDim.x) s_col[x] = 0;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int k = warp; k < len; k += nwarps) {
        const int inst = in_idx[base + k];
        const emia_inst_meta m = meta[inst];
        const uint32_t* crop = crops + crop_off[inst];
        // lane j counts bit j of one word column over the rows (every load is a warp-wide broadcast of one word)
        for (int c = 0; c < m.cw; ++c) {
            int cnt = 0;
#pragma unroll 4
            for (int r = 0; r < m.ch; ++r) cnt += (crop[(size_t)r * m.cw + c] >> lane) & 1u;
            const int x = (m.wc0 + c) * 32 + lane;
            if (cnt && x < W) atomicAdd(&s_col[x], cnt);
        }
    }
    __syncthreads();
    int mine = 0;
    for (int x = threadIdx.x; x < W; x += blockDim.x) mine += (s_col[x] > min_size);
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    const int K = s_cnt;
    const int nl = (K < len) ? K : len;
    for (int k = threadIdx.x; k < nl; k += blockDim.x) out_idx[base + k] = in_idx[base + k];
    if (threadIdx.x == 0) out_len[g] = nl;
}

extern "C" int emia_group_filter_area(const int32_t* cap_off, int32_t G, const int32_t* in_len, const int32_t* in_idx,
                                      const int32_t* area, int32_t min_area, int32_t* out_len, int32_t* out_idx, void* stream) {
    if (G < 0) return emia_fail(EMIA_ERR_BAD_ARG, "emia_group_filter_area: %s", "bad G");
    if (G == 0) return EMIA_OK;
    if (!cap_off || !in_len || !in_idx || !area || !out_len || !out_idx)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_group_filter_area: %s", "null pointer");
    k_group_filter_area<<<(unsigned)(((size_t)G * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(cap_off, G, in_len, in_idx, area,
                                                                                                   min_area, out_len, out_idx);
    return emia_check_launch("emia_group_filter_area launch: %s");
}
extern "C" int emia_column_gate(const uint32_t* crops, const emia_inst_meta* meta, const int64_t* crop_off, const int32_t* cap_off,
                                int32_t G, const int32_t* in_len, const int32_t* in_idx, int W, int32_t min_size,
                                int32_t* out_len, int32_t* out_idx, void* stream) {
    if (G < 0 || W <= 0 || W > 49152) return emia_fail(EMIA_ERR_BAD_ARG, "emia_column_gate: %s", "bad argument (W <= 49152)");
    if (G == 0) return EMIA_OK;
    if (!crops || !meta || !crop_off || !cap_off || !in_len || !in_idx || !out_len || !out_idx)
        return emia_fail(EMIA_ERR_BAD_ARG, "emia_column_gate: %s", "null pointer");
    const size_t smem = (size_t)W * 4;
    emia_need_dyn_smem((const void*)k_column_gate, smem);
    k_column_gate<<<(unsigned)G, 512, smem, (cudaStream_t)stream>>>(crops, meta, crop_off, cap_off, in_len, in_idx, W, min_size,
                                                                   out_len, out_idx);
    return emia_check_launch("emia_column_gate launch: %s");
}
