// Micro-benchmarks for the decode consumer: dp4a / I2F / LDS.128 issue rates and the GEMV inner
// loop in isolation (weights already in shared memory). One CTA per SM, 16 warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_dp4a(int* out, int iters, int chains) {
    int a = threadIdx.x, b = threadIdx.x * 3 + 1;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        c0 = __dp4a(a, b, c0);
        if (chains > 1) c1 = __dp4a(a, b, c1);
        if (chains > 2) { c2 = __dp4a(a, b, c2); c3 = __dp4a(a, b, c3); }
        a += c0 & 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (int) (t1 - t0);
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3;
}
__global__ void k_i2f(float* out, int iters) {
    int a = threadIdx.x;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        s0 += (float) (a + i); s1 += (float) (a + 2 * i); s2 += (float) (a - i); s3 += (float) (a ^ i);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) ((int*) out)[0] = (int) (t1 - t0);
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
// GEMV inner loop on a fake tile: rows x n, SG layout, x vector in smem. variant 0 = as in decode_mega.cu
template <int VAR>
__global__ void k_gemv(float* out, int n, int rows_per_warp, int reps) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int sgpr = (n + 255) / 256, groups = sgpr * 4;
    const int rb = sgpr * 272;
    uint8_t* xq = sm;
    uint8_t* tile = sm + ((rb + 127) & ~127);
    for (int i = threadIdx.x; i < rb * 17 / 4; i += blockDim.x) ((int*) sm)[i] = i * 2654435761u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rot = lane & 2;
    float total = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int rr = 0; rr < rows_per_warp; ++rr) {
            const uint8_t* row = tile + (size_t) ((warp + rr) & 15) * rb;
            if (VAR == 0) {
                float acc = 0;
                for (int G = lane; G < groups; G += 32) {
                    const int off = (G >> 2) * 272 + (G & 3) * 64;
                    int dot = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 wv = *(const int4*) (row + off + pc);
                        const int4 xv = *(const int4*) (xq + off + pc);
                        dot = __dp4a(wv.x, xv.x, dot); dot = __dp4a(wv.y, xv.y, dot);
                        dot = __dp4a(wv.z, xv.z, dot); dot = __dp4a(wv.w, xv.w, dot);
                    }
                    const int so = (G >> 2) * 272 + 256 + (G & 3) * 4;
                    acc += ((float) dot * *(const float*) (row + so)) * *(const float*) (xq + so);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                total += acc;
            } else if (VAR == 2) {
                const uint8_t* row2 = tile + (size_t) ((warp + rr + 1) & 15) * rb;
                float a0 = 0, a1 = 0;
                for (int G = lane; G < groups; G += 64) {
                    const int G2 = G + 32;
                    const bool has2 = G2 < groups;
                    const int off = (G >> 2) * 272 + (G & 3) * 64;
                    const int off2 = has2 ? (G2 >> 2) * 272 + (G2 & 3) * 64 : off;
                    int d00[4], d01[4], d10[4], d11[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 x0 = *(const int4*) (xq + off + pc);
                        const int4 x1 = *(const int4*) (xq + off2 + pc);
                        const int4 w00 = *(const int4*) (row + off + pc);
                        const int4 w01 = *(const int4*) (row + off2 + pc);
                        const int4 w10 = *(const int4*) (row2 + off + pc);
                        const int4 w11 = *(const int4*) (row2 + off2 + pc);
                        d00[i] = __dp4a(w00.x, x0.x, 0); d01[i] = __dp4a(w01.x, x1.x, 0); d10[i] = __dp4a(w10.x, x0.x, 0); d11[i] = __dp4a(w11.x, x1.x, 0);
                        d00[i] = __dp4a(w00.y, x0.y, d00[i]); d01[i] = __dp4a(w01.y, x1.y, d01[i]); d10[i] = __dp4a(w10.y, x0.y, d10[i]); d11[i] = __dp4a(w11.y, x1.y, d11[i]);
                        d00[i] = __dp4a(w00.z, x0.z, d00[i]); d01[i] = __dp4a(w01.z, x1.z, d01[i]); d10[i] = __dp4a(w10.z, x0.z, d10[i]); d11[i] = __dp4a(w11.z, x1.z, d11[i]);
                        d00[i] = __dp4a(w00.w, x0.w, d00[i]); d01[i] = __dp4a(w01.w, x1.w, d01[i]); d10[i] = __dp4a(w10.w, x0.w, d10[i]); d11[i] = __dp4a(w11.w, x1.w, d11[i]);
                    }
                    const int so = (G >> 2) * 272 + 256 + (G & 3) * 4;
                    const int so2 = (G2 >> 2) * 272 + 256 + (G2 & 3) * 4;
                    const float xs0 = *(const float*) (xq + so);
                    a0 += ((float) ((d00[0] + d00[1]) + (d00[2] + d00[3])) * *(const float*) (row + so)) * xs0;
                    a1 += ((float) ((d10[0] + d10[1]) + (d10[2] + d10[3])) * *(const float*) (row2 + so)) * xs0;
                    if (has2) {
                        const float xs1 = *(const float*) (xq + so2);
                        a0 += ((float) ((d01[0] + d01[1]) + (d01[2] + d01[3])) * *(const float*) (row + so2)) * xs1;
                        a1 += ((float) ((d11[0] + d11[1]) + (d11[2] + d11[3])) * *(const float*) (row2 + so2)) * xs1;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
                total += a0 + a1;
                ++rr;
            } else if (VAR == 3) {
                // four rows at once sharing x, two groups in flight (8 independent dp4a chains)
                const uint8_t* r1 = tile + (size_t) ((warp + rr + 1) & 15) * rb;
                const uint8_t* r2 = tile + (size_t) ((warp + rr + 2) & 15) * rb;
                const uint8_t* r3 = tile + (size_t) ((warp + rr + 3) & 15) * rb;
                float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                for (int G = lane; G < groups; G += 64) {
                    const int G2 = G + 32;
                    const bool has2 = G2 < groups;
                    const int off = (G >> 2) * 272 + (G & 3) * 64;
                    const int off2 = has2 ? (G2 >> 2) * 272 + (G2 & 3) * 64 : off;
                    int d[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 x0 = *(const int4*) (xq + off + pc);
                        const int4 x1 = *(const int4*) (xq + off2 + pc);
                        const uint8_t* rs[4] = {row, r1, r2, r3};
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int4 w0 = *(const int4*) (rs[r] + off + pc);
                            const int4 w1 = *(const int4*) (rs[r] + off2 + pc);
                            d[r][0] = __dp4a(w0.x, x0.x, d[r][0]); d[r][1] = __dp4a(w1.x, x1.x, d[r][1]);
                            d[r][0] = __dp4a(w0.y, x0.y, d[r][0]); d[r][1] = __dp4a(w1.y, x1.y, d[r][1]);
                            d[r][0] = __dp4a(w0.z, x0.z, d[r][0]); d[r][1] = __dp4a(w1.z, x1.z, d[r][1]);
                            d[r][0] = __dp4a(w0.w, x0.w, d[r][0]); d[r][1] = __dp4a(w1.w, x1.w, d[r][1]);
                        }
                    }
                    const int so = (G >> 2) * 272 + 256 + (G & 3) * 4;
                    const int so2 = (G2 >> 2) * 272 + 256 + (G2 & 3) * 4;
                    const float xs0 = *(const float*) (xq + so);
                    a0 += ((float) d[0][0] * *(const float*) (row + so)) * xs0;
                    a1 += ((float) d[1][0] * *(const float*) (r1 + so)) * xs0;
                    a2 += ((float) d[2][0] * *(const float*) (r2 + so)) * xs0;
                    a3 += ((float) d[3][0] * *(const float*) (r3 + so)) * xs0;
                    if (has2) {
                        const float xs1 = *(const float*) (xq + so2);
                        a0 += ((float) d[0][1] * *(const float*) (row + so2)) * xs1;
                        a1 += ((float) d[1][1] * *(const float*) (r1 + so2)) * xs1;
                        a2 += ((float) d[2][1] * *(const float*) (r2 + so2)) * xs1;
                        a3 += ((float) d[3][1] * *(const float*) (r3 + so2)) * xs1;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
                }
                total += a0 + a1 + a2 + a3;
                rr += 3;
            } else if (VAR == 4) {
                // two rows, groups = 32 * a + 8: full slots one group per lane, the last 8 groups with 4 lanes per group
                const uint8_t* row2 = tile + (size_t) ((warp + rr + 1) & 15) * rb;
                float a0 = 0, a1 = 0;
                const int full = (groups / 32) * 32;
                for (int base = 0; base < full; base += 32) {
                    const int G = base + lane;
                    const int off = (G >> 2) * 272 + (G & 3) * 64;
                    int d00 = 0, d01 = 0, d10 = 0, d11 = 0; // two half chains per row
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 x0 = *(const int4*) (xq + off + pc);
                        const int4 w00 = *(const int4*) (row + off + pc);
                        const int4 w10 = *(const int4*) (row2 + off + pc);
                        d00 = __dp4a(w00.x, x0.x, d00); d01 = __dp4a(w00.y, x0.y, d01); d10 = __dp4a(w10.x, x0.x, d10); d11 = __dp4a(w10.y, x0.y, d11);
                        d00 = __dp4a(w00.z, x0.z, d00); d01 = __dp4a(w00.w, x0.w, d01); d10 = __dp4a(w10.z, x0.z, d10); d11 = __dp4a(w10.w, x0.w, d11);
                    }
                    const int so = (G >> 2) * 272 + 256 + (G & 3) * 4;
                    const float xs0 = *(const float*) (xq + so);
                    a0 += ((float) (d00 + d01) * *(const float*) (row + so)) * xs0;
                    a1 += ((float) (d10 + d11) * *(const float*) (row2 + so)) * xs0;
                }
                if (groups - full == 8) {
                    const int G = full + (lane >> 2);
                    const int off = (G >> 2) * 272 + (G & 3) * 64 + (lane & 3) * 16;
                    const int4 x0 = *(const int4*) (xq + off);
                    const int4 w0 = *(const int4*) (row + off);
                    const int4 w1 = *(const int4*) (row2 + off);
                    int t0 = __dp4a(w0.x, x0.x, 0), t1 = __dp4a(w1.x, x0.x, 0);
                    int u0 = __dp4a(w0.y, x0.y, 0), u1 = __dp4a(w1.y, x0.y, 0);
                    t0 = __dp4a(w0.z, x0.z, t0); t1 = __dp4a(w1.z, x0.z, t1);
                    u0 = __dp4a(w0.w, x0.w, u0); u1 = __dp4a(w1.w, x0.w, u1);
                    t0 += u0; t1 += u1;
                    t0 += __shfl_xor_sync(0xffffffffu, t0, 1); t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
                    t0 += __shfl_xor_sync(0xffffffffu, t0, 2); t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
                    if ((lane & 3) == 0) {
                        const int so = (G >> 2) * 272 + 256 + (G & 3) * 4;
                        const float xs0 = *(const float*) (xq + so);
                        a0 += ((float) t0 * *(const float*) (row + so)) * xs0;
                        a1 += ((float) t1 * *(const float*) (row2 + so)) * xs0;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
                total += a0 + a1;
                ++rr;
            } else {
                // VAR 1: two rows at once sharing x, two groups in flight (4 independent dp4a chains)
                const uint8_t* row2 = tile + (size_t) ((warp + rr + 1) & 15) * rb;
                float a0 = 0, a1 = 0;
                for (int G = lane; G < groups; G += 64) {
                    const int G2 = G + 32;
                    const bool has2 = G2 < groups;
                    const int off = (G >> 2) * 272 + (G & 3) * 64;
                    const int off2 = has2 ? (G2 >> 2) * 272 + (G2 & 3) * 64 : off;
                    int d00 = 0, d01 = 0, d10 = 0, d11 = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 x0 = *(const int4*) (xq + off + pc);
                        const int4 x1 = *(const int4*) (xq + off2 + pc);
                        const int4 w00 = *(const int4*) (row + off + pc);
                        const int4 w01 = *(const int4*) (row + off2 + pc);
                        const int4 w10 = *(const int4*) (row2 + off + pc);
                        const int4 w11 = *(const int4*) (row2 + off2 + pc);
                        d00 = __dp4a(w00.x, x0.x, d00); d01 = __dp4a(w01.x, x1.x, d01); d10 = __dp4a(w10.x, x0.x, d10); d11 = __dp4a(w11.x, x1.x, d11);
                        d00 = __dp4a(w00.y, x0.y, d00); d01 = __dp4a(w01.y, x1.y, d01); d10 = __dp4a(w10.y, x0.y, d10); d11 = __dp4a(w11.y, x1.y, d11);
                        d00 = __dp4a(w00.z, x0.z, d00); d01 = __dp4a(w01.z, x1.z, d01); d10 = __dp4a(w10.z, x0.z, d10); d11 = __dp4a(w11.z, x1.z, d11);
                        d00 = __dp4a(w00.w, x0.w, d00); d01 = __dp4a(w01.w, x1.w, d01); d10 = __dp4a(w10.w, x0.w, d10); d11 = __dp4a(w11.w, x1.w, d11);
                    }
                    const int so = (G >> 2) * 272 + 256 + (G & 3) * 4;
                    const int so2 = (G2 >> 2) * 272 + 256 + (G2 & 3) * 4;
                    const float xs0 = *(const float*) (xq + so);
                    a0 += ((float) d00 * *(const float*) (row + so)) * xs0;
                    a1 += ((float) d10 * *(const float*) (row2 + so)) * xs0;
                    if (has2) {
                        const float xs1 = *(const float*) (xq + so2);
                        a0 += ((float) d01 * *(const float*) (row + so2)) * xs1;
                        a1 += ((float) d11 * *(const float*) (row2 + so2)) * xs1;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
                total += a0 + a1;
                ++rr;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) ((int*) out)[0] = (int) (t1 - t0);
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = total;
}

int main() {
    int* d; cudaMalloc(&d, 64 << 20);
    int h;
    for (int warps : {1, 4, 16}) for (int chains : {1, 2, 4}) {
        k_dp4a<<<148, warps * 32>>>(d, 4096, chains); cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
        printf("dp4a warps %2d chains %d: %.2f cycles per dp4a-warp-instr per SMSP-warp-set (%.2f cyc/iter)\n", warps, chains,
               (double) h / (4096.0 * (chains == 4 ? 4 : chains)) , (double) h / 4096.0);
    }
    for (int warps : {1, 16}) {
        k_i2f<<<148, warps * 32>>>((float*) d, 4096); cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
        printf("i2f+fadd warps %2d: %.2f cycles per iter (4 cvt + 4 fadd + int ops)\n", warps, (double) h / 4096.0);
    }
    cudaFuncSetAttribute(k_gemv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_gemv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_gemv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_gemv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_gemv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {2560, 4096, 9728}) for (int var : {1, 3, 4}) for (int warps : {15, 16}) {
        if (var == 4 && n != 2560) continue;
        const int sgpr = (n + 255) / 256, rb = sgpr * 272;
        const size_t smem = (size_t) rb * 17 + 256;
        const int rpw = 16, reps = 20;
        if (var == 0) k_gemv<0><<<148, warps * 32, smem>>>((float*) d, n, rpw, reps);
        else if (var == 1) k_gemv<1><<<148, warps * 32, smem>>>((float*) d, n, rpw, reps);
        else if (var == 3) k_gemv<3><<<148, warps * 32, smem>>>((float*) d, n, rpw, reps);
        else if (var == 4) k_gemv<4><<<148, warps * 32, smem>>>((float*) d, n, rpw, reps);
        else k_gemv<2><<<148, warps * 32, smem>>>((float*) d, n, rpw, reps);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
        const double bytes = (double) warps * rpw * reps * rb;
        printf("gemv var %d n %5d warps %2d: %8d cycles, %.1f cycles/row/warp, %.1f B/cycle/SM %s\n", var, n, warps, h,
               (double) h / (rpw * reps), bytes / h, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
