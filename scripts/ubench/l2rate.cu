// Micro-benchmark: steady-state rate of TMA bulk copies into shared memory when the source is L2-resident vs HBM, for the
// per-warp pipelines of k_decode_pw (15 warps, each 2 copies of `cb` bytes in flight) and for one thread with 6 x 28 KB.
// Every CTA (one per SM) loops `reps` times over its own `per` bytes. per * 148 <= ~60 MB stays in L2 after the first pass;
// a large `per` (>= 2 MB per SM) streams from HBM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2rate l2rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
// NW warps, each with DEPTH stages of cb bytes; lane 0 of each warp issues and waits
__global__ void __launch_bounds__(512, 1) k(const uint8_t* big, size_t per, int cb, int depth, int reps, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[16 * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16 * 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (lane != 0) return;
    const uint8_t* src = big + per * blockIdx.x;
    const size_t ncopy = per / cb;                       // copies per pass over the region
    uint8_t* region = smem + (size_t) warp * depth * cb;
    const unsigned long long t0 = gtime();
    unsigned long long issued = 0, done = 0;
    const unsigned long long total = (unsigned long long) reps * ((ncopy - warp + nw - 1) / nw);
    // warp w takes copies w, w + nw, ... of every pass
    auto issue = [&](unsigned long long i) {
        const unsigned s = (unsigned) (i % depth);
        const size_t c = (warp + (i % ((ncopy - warp + nw - 1) / nw)) * nw) % ncopy;
        const uint32_t bar = s_u32(&bars[warp * 8 + s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(region + (size_t) s * cb)), "l"(src + c * cb), "r"(cb), "r"(bar) : "memory");
    };
    for (; issued < total && issued < (unsigned) depth; ++issued) issue(issued);
    for (; done < total; ++done) {
        wait(s_u32(&bars[warp * 8 + done % depth]), (unsigned) ((done / depth) & 1));
        if (issued < total) { issue(issued); ++issued; }
    }
    out[blockIdx.x * 16 + warp] = gtime() - t0;
}
int main() {
    uint8_t* big; unsigned long long* out;
    const size_t total = (size_t) 4 << 30;
    cudaMalloc(&big, total); cudaMemset(big, 1, total); cudaMalloc(&out, 148 * 16 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int nw, cb, depth; };
    for (Cfg c : {Cfg{15, 5440, 2}, Cfg{15, 2720, 4}, Cfg{15, 4096, 3}, Cfg{1, 28672, 6}, Cfg{4, 28672, 1}, Cfg{8, 10880, 2}}) {
        for (size_t per_kb : {256, 8192}) {            // 256 KB x 148 = 38 MB (L2 resident after pass 1), 8 MB x 148 = 1.2 GB (HBM)
            const size_t per = per_kb * 1024 / c.cb * c.cb;
            const int reps = per_kb == 256 ? 40 : 2;
            unsigned long long t[148 * 16], mx = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(out, 0, sizeof t);
                k<<<148, c.nw * 32, (size_t) c.nw * c.depth * c.cb>>>(big, per, c.cb, c.depth, reps, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(t, out, sizeof t, cudaMemcpyDeviceToHost);
            for (int i = 0; i < 148 * 16; ++i) mx = t[i] > mx ? t[i] : mx;
            const double bytes = (double) per * reps * 148;
            printf("%2d warps x %d x %5d B in flight (%3d KB/SM), region %4zu KB/SM (%s): %7.2f TB/s = %5.1f GB/s per SM\n", c.nw, c.depth, c.cb,
                   c.nw * c.depth * c.cb / 1024, per_kb, per_kb == 256 ? "L2" : "HBM", bytes / mx / 1e3, bytes / mx / 148);
        }
    }
    return 0;
}
