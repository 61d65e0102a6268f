// Micro-benchmark: does cp.async.bulk.prefetch.L2 make later bulk copies into shared memory faster on B200?
// Every CTA (one per SM) owns `per` bytes; optional prefetch in chunks of C bytes, a pause, then the timed
// copy of the region through a 6 x 28 KB shared-memory ring (one thread, like the decode producer).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(64, 1) k(const uint8_t* big, size_t per, int chunk, int pause_us, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) unsigned long long bars[6];
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint8_t* src = big + per * blockIdx.x;
    if (chunk > 0)
        for (size_t o = 0; o < per; o += chunk)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"((uint32_t) chunk) : "memory");
    const unsigned long long tw = gtime() + (unsigned long long) pause_us * 1000;
    while (gtime() < tw) {}
    const unsigned long long t0 = gtime();
    const unsigned n = (unsigned) (per / 28672);
    for (unsigned it = 0; it < n + 6; ++it) {
        const unsigned slot = it % 6, par = (it / 6) & 1;
        if (it >= 6) {
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s_u32(&bars[slot])), "r"(par ^ 1) : "memory");
        }
        if (it < n) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&bars[slot])), "r"(28672) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(ring + slot * 28672)), "l"(src + (size_t) it * 28672), "r"(28672), "r"(s_u32(&bars[slot])) : "memory");
        }
    }
    out[blockIdx.x] = gtime() - t0;
}
int main() {
    uint8_t* big; unsigned long long* out;
    const size_t total = (size_t) 2 << 30;
    cudaMalloc(&big, total); cudaMemset(big, 1, total); cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 28672);
    for (size_t per_kb : {112, 224, 336, 448, 672}) {
        const size_t per = per_kb * 1024 / 28672 * 28672;
        for (int chunk : {0, 4096, 16384, 28672}) {
            unsigned long long t[148], mx = 0;
            for (int rep = 0; rep < 3; ++rep) {
                // flush L2 between runs by touching another 512 MB
                cudaMemset(big + total / 2, rep, 512 << 20);
                k<<<148, 64, 6 * 28672>>>(big, per, chunk, 30, out);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(t, out, sizeof t, cudaMemcpyDeviceToHost);
            for (int i = 0; i < 148; ++i) mx = t[i] > mx ? t[i] : mx;
            printf("%4zu KB per SM (%5.1f MB total), prefetch chunk %5d: copy %7.2f us  -> %6.2f TB/s\n", per_kb, per * 148 / 1e6, chunk, mx / 1e3, per * 148.0 / mx / 1e3);
        }
    }
    return 0;
}
