// Micro-benchmark: what does ISSUING a TMA bulk copy cost the issuing warp on B200? (k_decode_pw measured ~680 cycles per
// item on the warp's critical path.) NW warps per SM, lane 0 of each: expect_tx, cp.async.bulk of cb bytes, then either
// wait for it (serial) or keep `depth` in flight. Cycles (clock64) are accumulated around each instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmaissue tmaissue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512, 1) k(const uint8_t* big, size_t per, int cb, int depth, int iters, int lds_load, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[16 * 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 64; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* src = big + per * blockIdx.x + (size_t) warp * (per / nw / 4096 * 4096);
    uint8_t* region = smem + (size_t) warp * depth * cb;
    long long t_exp = 0, t_cp = 0, t_wait = 0, t_lds = 0;
    int acc = 0;
    const long long t_begin = clock64();
    for (int it = 0; it < iters + depth; ++it) {
        const int s = it % depth;
        const uint32_t bar = s_u32(&bars[warp * 4 + s]);
        if (it >= depth) { // wait for the copy issued `depth` iterations ago into this stage
            const long long t0 = clock64();
            uint32_t ok = 0;
            const uint32_t par = ((it - depth) / depth) & 1;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
            const long long t1 = clock64();
            t_wait += t1 - t0;
            if (lds_load) { // consume the stage like the GEMV does: LDS.128 by all lanes
                for (int o = lane * 16; o < cb; o += 512) { const int4 v = *(const int4*) (region + (size_t) s * cb + o); acc += v.x ^ v.w; }
                __syncwarp();
                t_lds += clock64() - t1;
            }
        }
        if (it < iters) {
            const long long t0 = clock64();
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cb) : "memory");
            const long long t1 = clock64();
            if (lane == 0) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(region + (size_t) s * cb)), "l"(src + (size_t) (it % 8) * cb), "r"(cb), "r"(bar) : "memory");
            const long long t2 = clock64();
            t_exp += t1 - t0;
            t_cp += t2 - t1;
        }
    }
    const long long t_all = clock64() - t_begin;
    if (lane == 0 && blockIdx.x == 0) {
        long long* o = out + warp * 8;
        o[0] = t_exp; o[1] = t_cp; o[2] = t_wait; o[3] = t_lds; o[4] = t_all; o[5] = acc;
    }
}
int main() {
    uint8_t* big; long long* out;
    const size_t per = 1 << 20;
    cudaMalloc(&big, per * 148); cudaMemset(big, 1, per * 148); cudaMalloc(&out, 16 * 8 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int nw, cb, depth, lds; };
    for (Cfg c : {Cfg{1, 5440, 1, 0}, Cfg{1, 5440, 2, 0}, Cfg{15, 5440, 1, 0}, Cfg{15, 5440, 2, 0}, Cfg{15, 5440, 2, 1}, Cfg{15, 2720, 4, 1}, Cfg{4, 28672, 1, 1}, Cfg{8, 10880, 2, 1}}) {
        const int iters = 400;
        long long h[16 * 8];
        for (int rep = 0; rep < 2; ++rep) {
            k<<<148, c.nw * 32, (size_t) c.nw * c.depth * c.cb>>>(big, per, c.cb, c.depth, iters, c.lds, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        double ex = 0, cp = 0, wt = 0, ld = 0, all = 0;
        for (int w = 0; w < c.nw; ++w) { ex += h[w * 8]; cp += h[w * 8 + 1]; wt += h[w * 8 + 2]; ld += h[w * 8 + 3]; all += h[w * 8 + 4]; }
        const double n = (double) iters * c.nw;
        printf("%2d warps x depth %d x %5d B%s: per copy  expect_tx %5.0f  cp.async.bulk %5.0f  wait %6.0f  lds %5.0f  | %6.0f cycles per copy and warp, %5.1f B/clk/SM\n", c.nw, c.depth, c.cb,
               c.lds ? " + LDS" : "      ", ex / n, cp / n, wt / n, ld / n, all / n, (double) c.cb * c.nw / (all / n));
    }
    return 0;
}
