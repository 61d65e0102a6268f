// Micro-benchmark for the decode kernel's attention step (qwen3.c_b200/csrc/attn_core.cuh): how long does one block of
// 8 positions x 4 heads (attn_rows<4>) take per warp when W consumer warps of every SM run it back to back out of
// shared memory -- no ring, no hand-offs. The tiles phase of k_decode (8.1 us per layer at context 4096 on the 4B
// shape for 5.1 us of KV bytes) is instruction-issue bound; this isolates that cost so variants of attn_rows can be
// compared in seconds.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../qwen3.c_b200/csrc -o attn attn.cu
// run:   ./attn            (prints us per block for W = 1, 4, 8, 15 warps per SM)
#include <cstdio>
#include <cuda_runtime.h>

#include "attn_core.cuh"

void qw_set_error(const char*, ...) {}

constexpr int kPos = 112; // 4 tiles of 28 positions, as one group of the kernel
constexpr int kHW = 4;

template <int HW>
__global__ void __launch_bounds__(512, 1) k_attn(const float* __restrict__ kv, float* __restrict__ out, int warps, int reps, long long* cycles) {
    extern __shared__ __align__(16) float smem[]; // [kPos][128] K then [kPos][128] V
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2 * kPos * 128; i += blockDim.x) smem[i] = kv[i];
    __syncthreads();
    if (warp >= warps) return;
    float4 q[HW];
#pragma unroll
    for (int j = 0; j < HW; ++j) q[j] = make_float4(0.01f * (lane + j), -0.02f * j, 0.03f, 0.01f * lane);
    AttnState<HW> st;
    attn_state_reset(st);
    const int per = kPos / warps;                       // positions of this warp, as in consume_attn's even split
    const int p0 = per * warp, cnt = per < 8 ? per : 8;  // one block of <= 8 positions per repetition
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) attn_rows<HW>(smem + p0 * 128, cnt, smem + p0 * 128, cnt, kPos * 128, q, st, lane);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cycles[warp] = t1 - t0;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < HW; ++j) acc += st.acc[j].x + st.acc[j].w + st.l[j] + st.m[j];
    out[(blockIdx.x * 16 + warp) * 32 + lane] = acc;
}

int main() {
    const size_t kvn = 2 * kPos * 128;
    float* h = new float[kvn];
    for (size_t i = 0; i < kvn; ++i) h[i] = 0.001f * (float) ((i * 2654435761u) % 2001) - 1.0f;
    float *kv, *out;
    long long* cyc;
    cudaMalloc(&kv, kvn * 4);
    cudaMalloc(&out, 148 * 16 * 32 * 4);
    cudaMalloc(&cyc, 16 * 8);
    cudaMemcpy(kv, h, kvn * 4, cudaMemcpyHostToDevice);
    const int smem = (int) (kvn * 4);
    cudaFuncSetAttribute(k_attn<kHW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int reps = 200;
    for (int warps : {1, 4, 8, 15}) {
        k_attn<kHW><<<148, 512, smem>>>(kv, out, warps, reps, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[16] = {0};
        cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < warps; ++w) mx = c[w] > mx ? c[w] : mx;
        const int per = kPos / warps, cnt = per < 8 ? per : 8;
        printf("attn_rows<4>: %2d warps/SM, %d positions per block: %7.0f cycles = %.2f us per block (%s)\n", warps, cnt,
               (double) mx / reps, (double) mx / reps / (khz * 1e-3), e ? cudaGetErrorString(e) : "ok");
    }
    return 0;
}
