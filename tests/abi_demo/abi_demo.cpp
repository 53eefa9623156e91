// abi_demo.cpp — the fused hot path driven through the C ABI ALONE (include/emia.h + the CUDA runtime; no Python, no torch):
// what a C / C++ / cgo / JNI host would write.  Reads head outputs of T tiles from a binary file, runs
//   K1 paste -> K5a contours -> K4 deduplicate_masks_smart -> overlap rules -> containment rules -> K5b/c morphometry
// and writes the kept lists and the measurement records.  tests/test_gpu_abi_demo.py compares its output with the Python
// host layer (deepemia_b200/engine.py) on the same input: they must be identical.
//   usage: abi_demo <input.bin> <output.bin>
//   input : int32 T, H, W, n; int32 offs[T+1]; float probs[n*784]; float boxes[n*4]; float scores[n]; int32 classes[n]
//   output: int32 kept_len[T]; int32 kept_idx[n]; int64 n_rec; double records[n_rec*16]; int32 rec_inst[n_rec]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/emia.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
#define EK(x) do { int r_ = (x); if (r_ != EMIA_OK) { fprintf(stderr, "emia error %d at line %d: %s\n", r_, __LINE__, emia_last_error()); exit(3); } } while (0)

template <typename T> static T* dalloc(size_t n) { T* p = nullptr; CK(cudaMalloc(&p, (n ? n : 1) * sizeof(T))); return p; }
template <typename T> static T* upload(const std::vector<T>& v) { T* p = dalloc<T>(v.size()); if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return p; }
template <typename T> static T fetch(const T* d) { T v; CK(cudaMemcpy(&v, d, sizeof(T), cudaMemcpyDeviceToHost)); return v; }

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: abi_demo <input.bin> <output.bin>\n"); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("input"); return 1; }
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) return 1;
    const int T = hdr[0], H = hdr[1], W = hdr[2], n = hdr[3];
    std::vector<int32_t> offs(T + 1), classes(n);
    std::vector<float> probs((size_t)n * 784), boxes((size_t)n * 4), scores(n);
    if (fread(offs.data(), 4, T + 1, f) != (size_t)T + 1 || fread(probs.data(), 4, probs.size(), f) != probs.size() ||
        fread(boxes.data(), 4, boxes.size(), f) != boxes.size() || fread(scores.data(), 4, n, f) != (size_t)n ||
        fread(classes.data(), 4, n, f) != (size_t)n) { fprintf(stderr, "short input\n"); return 1; }
    fclose(f);
    cudaStream_t st; CK(cudaStreamCreate(&st));
    float *d_probs = upload(probs), *d_boxes = upload(boxes), *d_scores = upload(scores);
    int32_t *d_classes = upload(classes), *d_cap_off = upload(offs);
    void* scan_ws = dalloc<unsigned char>(emia_scan_workspace_bytes(n));
    const size_t scan_wb = emia_scan_workspace_bytes(n);

    // ---- K1
    emia_inst_meta* meta = dalloc<emia_inst_meta>(n);
    int64_t* crop_off = dalloc<int64_t>(n + 1);
    EK(emia_paste_plan(d_boxes, n, 1.f, 1.f, H, W, meta, crop_off, st));
    EK(emia_exclusive_scan_i64(crop_off, n, scan_ws, scan_wb, st));
    CK(cudaStreamSynchronize(st));
    const int64_t total_words = fetch(crop_off + n);
    uint32_t* crops = dalloc<uint32_t>(total_words);
    int32_t *bbox = dalloc<int32_t>((size_t)n * 4), *area = dalloc<int32_t>(n);
    EK(emia_paste_threshold_bitpack(d_probs, d_boxes, meta, crop_off, n, 1.f, 1.f, H, W, nullptr, 1, 8 * ((W + 255) / 256), crops, bbox, area, 2, nullptr, st));

    // ---- K5a: contours into per-instance slabs
    const int capc = 8;
    int64_t *pt_cap = dalloc<int64_t>(n + 1), *ncont = dalloc<int64_t>(n + 1), *scr_bytes = dalloc<int64_t>(n + 1);
    EK(emia_contour_trace_plan(meta, n, pt_cap, st));
    EK(emia_exclusive_scan_i64(pt_cap, n, scan_ws, scan_wb, st));
    CK(cudaStreamSynchronize(st));
    const int64_t cap_total = fetch(pt_cap + n);
    uint32_t *pts = dalloc<uint32_t>(cap_total), *marks = dalloc<uint32_t>(2 * total_words);
    int32_t *cstart = dalloc<int32_t>((size_t)n * (capc + 1) + 1), *overflow = dalloc<int32_t>(1);
    double* perim0 = dalloc<double>(n);
    CK(cudaMemsetAsync(overflow, 0, 4, st));
    EK(emia_contour_trace_slab(crops, meta, crop_off, n, marks, pt_cap, capc, pts, cstart, ncont, scr_bytes, overflow, perim0, nullptr, st));

    // ---- K4: de-dup 0.7, overlap rules (0: 0.30, 1: 0.50), containment 1 -> 0 at 0.95 (polyhipes_tommy)
    int max_cap = 0;
    for (int g = 0; g < T; ++g) max_cap = offs[g + 1] - offs[g] > max_cap ? offs[g + 1] - offs[g] : max_cap;
    std::vector<int32_t> len0(T), idx0(n);
    for (int g = 0; g < T; ++g) len0[g] = offs[g + 1] - offs[g];
    for (int i = 0; i < n; ++i) idx0[i] = i;
    int32_t *len_a = upload(len0), *idx_a = upload(idx0), *len_b = dalloc<int32_t>(T), *idx_b = dalloc<int32_t>(n);
    int32_t *len_c = dalloc<int32_t>(T), *idx_c = dalloc<int32_t>(n), *len_d = dalloc<int32_t>(T), *idx_d = dalloc<int32_t>(n);
    const size_t gwb = emia_group_workspace_bytes(offs.data(), T);
    void* gws = dalloc<unsigned char>(gwb);
    EK(emia_dedup_smart(crops, meta, crop_off, bbox, area, perim0, ncont, d_scores, d_classes, d_cap_off, T, n, max_cap, len_a, idx_a, 0.7, 0.0,
                        len_b, idx_b, gws, gwb, st));
    std::vector<int32_t> active = {1, 1};
    std::vector<double> max_iou = {0.30, 0.50};
    int32_t* d_active = upload(active);
    double* d_max_iou = upload(max_iou);
    EK(emia_overlap_rules(crops, meta, crop_off, bbox, area, d_scores, d_classes, d_cap_off, T, n, max_cap, len_b, idx_b, d_active, d_max_iou, 2,
                          len_c, idx_c, gws, gwb, st));
    const int32_t child[1] = {1}, parent[1] = {0};
    EK(emia_containment_rules(crops, meta, crop_off, bbox, area, d_classes, d_cap_off, T, n, max_cap, len_c, idx_c, child, parent, 1, 0.95,
                              len_d, idx_d, gws, gwb, st));
    CK(cudaStreamSynchronize(st));
    if (fetch(overflow) != 0) { fprintf(stderr, "slab overflow: this demo does not implement the exact two-pass fallback\n"); return 4; }

    // ---- K5b/c: morphometry of the survivors
    int32_t* item_inst = dalloc<int32_t>(n);
    int64_t *rec_off = dalloc<int64_t>(n + 1), *scr_off = dalloc<int64_t>(n + 1);
    CK(cudaMemsetAsync(rec_off, 0, (size_t)(n + 1) * 8, st));
    CK(cudaMemsetAsync(scr_off, 0, (size_t)(n + 1) * 8, st));
    EK(emia_list_measure_plan(d_cap_off, T, n, len_d, idx_d, ncont, scr_bytes, item_inst, rec_off, scr_off, st));
    EK(emia_exclusive_scan_i64(rec_off, n, scan_ws, scan_wb, st));
    EK(emia_exclusive_scan_i64(scr_off, n, scan_ws, scan_wb, st));
    CK(cudaStreamSynchronize(st));
    const int64_t n_rec = fetch(rec_off + n), n_scr = fetch(scr_off + n);
    double* records = dalloc<double>((size_t)n_rec * EMIA_REC_FIELDS);
    int32_t* rec_inst = dalloc<int32_t>(n_rec);
    uint8_t* scratch = dalloc<uint8_t>(n_scr + 16);
    const double min_area = 5.0 > H * W * 0.000005 * 0.05 ? 5.0 : H * W * 0.000005 * 0.05;
    EK(emia_contour_measure_list(n, item_inst, rec_off, scr_off, nullptr, pt_cap, cstart, capc + 1, 0.5, min_area, pts, records, rec_inst, scratch, nullptr, nullptr, st));
    CK(cudaStreamSynchronize(st));

    std::vector<int32_t> klen(T), kidx(n), rinst(n_rec);
    std::vector<double> rec((size_t)n_rec * EMIA_REC_FIELDS);
    CK(cudaMemcpy(klen.data(), len_d, (size_t)T * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(kidx.data(), idx_d, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (n_rec) {
        CK(cudaMemcpy(rec.data(), records, rec.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rinst.data(), rec_inst, (size_t)n_rec * 4, cudaMemcpyDeviceToHost));
    }
    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("output"); return 1; }
    fwrite(klen.data(), 4, T, o);
    fwrite(kidx.data(), 4, n, o);
    fwrite(&n_rec, 8, 1, o);
    fwrite(rec.data(), 8, rec.size(), o);
    fwrite(rinst.data(), 4, n_rec, o);
    fclose(o);
    printf("abi_demo: %d tiles, %d instances, %lld records, library version %d\n", T, n, (long long)n_rec, emia_version());
    return 0;
}
