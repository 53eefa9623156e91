// emia_common.cuh — shared host/device primitives for the deepEMIA B200 hot path.
//
// Every function in core/ is EMIA_HD so that the SAME source is (a) the body of the sm_100a kernels
// and (b) compilable by g++ into tests/hostsim (CPU-only CI harness that checks the device algorithms
// against OpenCV / torch before any GPU time is spent).  The product path never loads the host build.
//
// Floating-point discipline: the reference's float outputs go through discontinuous steps
// (>= 0.5 threshold, int truncation of box corners), so operation order and rounding matter.
//   * the CUDA library is compiled with -fmad=false and the host build with -ffp-contract=off:
//     a*b+c is NEVER contracted implicitly;
//   * where the oracle itself uses a fused multiply-add (torch CPU grid_sample) we call emia_fmaf().
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define EMIA_HD __host__ __device__ __forceinline__
#define EMIA_HD_NOINLINE __host__ __device__
#define EMIA_HD_COLD __host__ __device__ __noinline__      // rarely taken paths: keep them out of the hot path's register budget
#else
#define EMIA_HD inline
#define EMIA_HD_NOINLINE
#define EMIA_HD_COLD
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

EMIA_HD float emia_fmaf(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

EMIA_HD double emia_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

EMIA_HD int emia_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
EMIA_HD int emia_popcll(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
// index of lowest set bit (v != 0)
EMIA_HD int emia_ctz(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
// index of highest set bit (v != 0)
EMIA_HD int emia_msb(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
EMIA_HD int emia_min(int a, int b) { return a < b ? a : b; }
EMIA_HD int emia_max(int a, int b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------
// Bit-packed mask view.  Bit j (LSB first) of word w in a row is pixel x = x_origin + 32*w + j.
// A view addresses a rectangle of a frame:  rows [y0, y0+h), word columns starting at frame word wc0.
// ---------------------------------------------------------------------------------------------
struct EmiaBitView {
    const uint32_t* bits;  // first word of first row
    int pitch_words;       // words between consecutive rows
    int h;                 // rows
    int wwords;            // valid words per row
    int x_origin;          // frame x of bit 0 of word 0 (multiple of 32)
    int y_origin;          // frame y of row 0
};

EMIA_HD uint32_t emia_view_word(const EmiaBitView& v, int r, int c) {
    if ((unsigned)r >= (unsigned)v.h || (unsigned)c >= (unsigned)v.wwords) return 0u;
    return v.bits[(size_t)r * v.pitch_words + c];
}
// pixel test in LOCAL coordinates (lx in [0, 32*wwords), ly in [0,h)); outside -> 0
EMIA_HD int emia_view_px(const EmiaBitView& v, int lx, int ly) {
    if ((unsigned)ly >= (unsigned)v.h || (unsigned)lx >= (unsigned)(v.wwords * 32)) return 0;
    return (v.bits[(size_t)ly * v.pitch_words + (lx >> 5)] >> (lx & 31)) & 1u;
}
