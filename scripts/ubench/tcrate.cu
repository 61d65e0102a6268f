// Micro-benchmark for the prefill GEMM's design question (Q8_0 scales change every 64 k-elements, so an int32 accumulator
// has to be promoted to fp32 every 2 K32 MMAs):
//   (a) how fast can accumulators LEAVE TMEM?  NW warps per SM loop tcgen05.ld 32x32b.x32 / .x64 / .x128 over the 512 columns
//       (no MMA, no arithmetic): bytes per clock per SM.  A 128 x 128 int32 group tile is 64 KB.
//   (b) what does the legacy register-accumulator path give on sm_100a?  NW warps per SM issue mma.sync.m16n8k32.s8 (and
//       m16n8k16.bf16 for scale) on register operands, 8 independent accumulators per warp: dense tera-ops/s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcrate tcrate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ int tmem_ld(uint32_t taddr);
template <>
__device__ __forceinline__ int tmem_ld<32>(uint32_t taddr) {
    int v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a ^= v[i];
    return a;
}
template <>
__device__ __forceinline__ int tmem_ld<8>(uint32_t taddr) {
    int v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return v[0] ^ v[1] ^ v[2] ^ v[3] ^ v[4] ^ v[5] ^ v[6] ^ v[7];
}
// two x32 loads in flight before one wait
template <>
__device__ __forceinline__ int tmem_ld<64>(uint32_t taddr) {
    int v[32], w[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
          "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
          "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
          "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
        : "r"(taddr + 32)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int a = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a ^= v[i] ^ w[i];
    return a;
}

// NW warps (multiple of 4) per CTA, one CTA per SM; warp w reads lanes 32 * (w % 4), columns cycling through 512
template <int X>
__global__ void __launch_bounds__(512, 1) k_tmem(int iters, long long* cyc, int* sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s + ((uint32_t) ((warp & 3) * 32) << 16);
    int acc = 0;
    const int nblk = 512 / X;
    int c = (warp >> 2) % nblk;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        acc ^= tmem_ld<X>(base + c * X);
        c = c + 1 == nblk ? 0 : c + 1;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678) *sink = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(512) : "memory");
}

// legacy tensor path: ACC independent m16n8k32 s8 accumulators per warp, register operands
template <int ACC>
__global__ void __launch_bounds__(1024, 1) k_imma(int iters, int* sink) {
    int c[ACC][4];
#pragma unroll
    for (int j = 0; j < ACC; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = threadIdx.x * 11, b1 = threadIdx.x * 13;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ACC; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int acc = 0;
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc ^= c[j][0] ^ c[j][1] ^ c[j][2] ^ c[j][3];
    if (acc == 0x12345678) *sink = acc;
}
template <int ACC>
__global__ void __launch_bounds__(1024, 1) k_hmma(int iters, int* sink) {
    float c[ACC][4];
#pragma unroll
    for (int j = 0; j < ACC; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
    uint32_t a0 = 0x3c003c00u, a1 = 0x3c003c00u, a2 = 0x3c003c00u, a3 = 0x3c003c00u, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ACC; ++j)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float acc = 0;
#pragma unroll
    for (int j = 0; j < ACC; ++j) acc += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    if (acc == 12345.678f) *sink = 1;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    long long* cyc; int* sink;
    cudaMalloc(&cyc, 8 * 256); cudaMalloc(&sink, 4);
    printf("SMs %d, clock %d MHz\n", sms, khz / 1000);
    printf("--- (a) tcgen05.ld rate, one CTA per SM, no MMA, no arithmetic\n");
    const int iters = 20000;
    for (int nw : {4, 8, 16}) {
        long long h[256];
        k_tmem<8><<<sms, nw * 32>>>(iters, cyc, sink); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
        printf("32x32b.x8      %2d warps: %7.1f B/clk/SM  (%lld cycles)\n", nw, (double) nw * iters * 8 * 128 / h[0], h[0]);
        k_tmem<32><<<sms, nw * 32>>>(iters, cyc, sink); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
        printf("32x32b.x32     %2d warps: %7.1f B/clk/SM  (%lld cycles)\n", nw, (double) nw * iters * 32 * 128 / h[0], h[0]);
        k_tmem<64><<<sms, nw * 32>>>(iters, cyc, sink); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8 * sms, cudaMemcpyDeviceToHost);
        printf("2 x 32x32b.x32 %2d warps: %7.1f B/clk/SM  (%lld cycles)\n", nw, (double) nw * iters * 64 * 128 / h[0], h[0]);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error after (a): %s\n", cudaGetErrorString(e));
    printf("--- (b) legacy mma.sync on register operands, one CTA per SM\n");
    const int it2 = 20000;
    for (int nw : {4, 8, 16, 32}) {
        float ms = time_ms([&] { k_imma<8><<<sms, nw * 32>>>(it2, sink); });
        printf("m16n8k32 s8   %2d warps x 8 acc: %8.1f dense TOPS (%.3f ms)\n", nw, 2.0 * 16 * 8 * 32 * 8 * it2 * nw * sms / (ms * 1e-3) / 1e12, ms);
        ms = time_ms([&] { k_imma<16><<<sms, nw * 32>>>(it2, sink); });
        printf("m16n8k32 s8   %2d warps x16 acc: %8.1f dense TOPS (%.3f ms)\n", nw, 2.0 * 16 * 8 * 32 * 16 * it2 * nw * sms / (ms * 1e-3) / 1e12, ms);
        ms = time_ms([&] { k_hmma<8><<<sms, nw * 32>>>(it2, sink); });
        printf("m16n8k16 bf16 %2d warps x 8 acc: %8.1f dense TFLOPS (%.3f ms)\n", nw, 2.0 * 16 * 8 * 16 * 8 * it2 * nw * sms / (ms * 1e-3) / 1e12, ms);
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 0;
}
