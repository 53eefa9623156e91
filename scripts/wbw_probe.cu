// wbw_probe.cu — scratch micro-benchmark (not part of the product): how fast can B200 WRITE to HBM?
// Variants: cudaMemset, grid-stride 128-bit streaming stores, bulk shared->global copies (cp.async.bulk) of 16/32 KB pieces.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k_store(uint4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (MODE == 0) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(z.x), "r"(z.y), "r"(z.z), "r"(z.w) : "memory");
        else if (MODE == 1) p[i] = z;
        else asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(z.x), "r"(z.y), "r"(z.z), "r"(z.w) : "memory");
    }
}
// 256-bit stores (sm_100+: st.global.v4.b64), grid-stride over 32-byte units
__global__ void __launch_bounds__(256) k_store256(unsigned char* p, size_t n32) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const unsigned long long z = 0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride)
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p + i * 32), "l"(z), "l"(z), "l"(z), "l"(z) : "memory");
}
// frame-per-CTA with 256-bit stores
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_store_frames256(unsigned char* p, size_t nframes) {
    const unsigned long long z = 0ull;
    for (size_t f = blockIdx.x; f < nframes; f += gridDim.x) {
        unsigned char* fr = p + f * 131072;
#pragma unroll 4
        for (int q = threadIdx.x; q < 4096; q += THREADS)
            asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(fr + (size_t)q * 32), "l"(z), "l"(z), "l"(z), "l"(z) : "memory");
    }
}
// each CTA owns contiguous 128 KB "frames" (like the paste kernel): frame f -> CTA f % grid
template <int UNROLL>
__global__ void __launch_bounds__(256) k_store_frames(uint4* p, size_t nframes) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t f = blockIdx.x; f < nframes; f += gridDim.x) {
        uint4* fr = p + f * 8192;
#pragma unroll UNROLL
        for (int q = threadIdx.x; q < 8192; q += 256)
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(fr + q), "r"(z.x), "r"(z.y), "r"(z.z), "r"(z.w) : "memory");
    }
}
__global__ void __launch_bounds__(128) k_bulk(unsigned char* p, size_t nframes, int piece) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int k = threadIdx.x; k < piece / 4; k += blockDim.x) ((uint32_t*)sm)[k] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (size_t f = blockIdx.x; f < nframes; f += gridDim.x) {
            unsigned char* fr = p + f * 131072;
            for (int o = 0; o < 131072; o += piece)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(fr + o),
                             "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(piece) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    const size_t bytes = (size_t)24 << 30;
    unsigned char* d; CK(cudaMalloc(&d, bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t n4 = bytes / 16, nframes = bytes / 131072;
#define TIME(label, launch) do { for (int r = 0; r < 2; ++r) { launch; } CK(cudaDeviceSynchronize()); cudaEventRecord(a); for (int r = 0; r < 3; ++r) { launch; } cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b); printf("%-46s %8.1f GB/s\n", label, 3.0 * bytes / (ms * 1e-3) / 1e9); } while (0)
    TIME("cudaMemsetAsync", cudaMemsetAsync(d, 0, bytes));
    for (int per : {4, 8, 16}) {
        char l[96];
        snprintf(l, 96, "grid-stride st.cs.v4, %d CTAs/SM", per); TIME(l, (k_store<0><<<sms * per, 256>>>((uint4*)d, n4)));
        snprintf(l, 96, "grid-stride st.v4 (default), %d CTAs/SM", per); TIME(l, (k_store<1><<<sms * per, 256>>>((uint4*)d, n4)));
        snprintf(l, 96, "grid-stride st.wt.v4, %d CTAs/SM", per); TIME(l, (k_store<2><<<sms * per, 256>>>((uint4*)d, n4)));
        snprintf(l, 96, "frame-per-CTA st.cs.v4 unroll 8, %d CTAs/SM", per); TIME(l, (k_store_frames<8><<<sms * per, 256>>>((uint4*)d, nframes)));
        snprintf(l, 96, "frame-per-CTA st.cs.v4 unroll 32, %d CTAs/SM", per); TIME(l, (k_store_frames<32><<<sms * per, 256>>>((uint4*)d, nframes)));
    }
    for (int per : {2, 4, 8, 16, 32}) {
        char l[96];
        snprintf(l, 96, "grid-stride st.v4.b64 (256-bit), %d CTAs/SM", per); TIME(l, (k_store256<<<sms * per, 256>>>(d, bytes / 32)));
        snprintf(l, 96, "frame-per-CTA 256-bit, 256 thr, %d CTAs/SM", per); TIME(l, (k_store_frames256<256><<<sms * per, 256>>>(d, nframes)));
        snprintf(l, 96, "frame-per-CTA 256-bit, 128 thr, %d CTAs/SM", per); TIME(l, (k_store_frames256<128><<<sms * per, 128>>>(d, nframes)));
        snprintf(l, 96, "frame-per-CTA 256-bit, 1024 thr, %d CTAs/SM", per); TIME(l, (k_store_frames256<1024><<<sms * per, 1024>>>(d, nframes)));
    }
    for (int piece : {16384}) for (int per : {1, 2, 4}) {
        char l[96]; snprintf(l, 96, "bulk S2G piece %d B, %d CTAs/SM", piece, per);
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, piece);
        TIME(l, (k_bulk<<<sms * per, 128, piece>>>(d, nframes, piece)));
    }
    return 0;
}
