// Micro-benchmark: how fast do TMA bulk copies into shared memory run on data that an L2 PREFETCH brought in (never read
// before), as opposed to data left in L2 by an earlier demand pass (l2rate.cu)? 148 CTAs, 15 warps x 2 x 5440 B in flight,
// ONE pass over `per` bytes per SM, every test on a fresh part of a 4 GB buffer (L2 flushed between tests).
//   cold      : no prefetch (HBM)
//   pf(T)     : one lane per CTA issues cp.async.bulk.prefetch.L2 for the CTA's whole region, all warps wait T us, then copy
//   pf-other  : the region is prefetched by ANOTHER kernel (CTA b + 74 asks for CTA b's bytes), 200 us earlier
//   demand    : the region was read once by the same copies before (what l2rate.cu measures)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pfrate pfrate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include <unistd.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
__device__ __forceinline__ void prefetch(const uint8_t* src, size_t bytes) {
    for (size_t o = 0; o < bytes; o += 32768) {
        const uint32_t n = (uint32_t) (bytes - o < 32768 ? bytes - o : 32768);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"(n) : "memory");
    }
}
__global__ void k_pf(const uint8_t* big, size_t per, int shift) {
    if (threadIdx.x == 0) prefetch(big + per * ((blockIdx.x + shift) % gridDim.x), per);
}
// pf_wait_ns < 0: no prefetch. how: 0 cp.async.bulk.prefetch.L2 (one lane) | 1 prefetch.global.L2 per 128-byte line (warp 15)
// | 2 ld.global.L2::128B one word per line | 3 cp.async 16 B per line with the L2::128B hint | 4 ld.global one word per 32-byte sector
// | 5 prefetch.global.L2::evict_last per line | 6 ld.global.L2::256B one word per 256 bytes
// | 7 TMA tensor load of a [64 rows][16 B] box out of a [N][256 B] view with L2 promotion 256B | 8 the same on a [N][128 B] view,
// promotion 128B | 9 cp.async.bulk.prefetch.tensor (full 256-byte rows) | 10 as 7 without promotion
struct Maps { CUtensorMap m256, m128, mfull, m256n; };
__global__ void __launch_bounds__(512, 1) k(const uint8_t* big, size_t per, int cb, long long pf_wait_ns, unsigned long long* out, int how, unsigned* sink,
                                             const __grid_constant__ Maps maps, const uint8_t* origin) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[16 * 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = 15, depth = 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* src = big + per * blockIdx.x;
    if (pf_wait_ns >= 0) {
        const unsigned long long t = gtime();
        __shared__ __align__(16) uint8_t dummy[32 * 16];
        unsigned acc = 0;
        if (warp == 15) {
            if (how == 0) {
                if (lane == 0) prefetch(src, per);
            } else if (how == 1) {
                for (size_t o = (size_t) lane * 128; o < per; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + o) : "memory");
            } else if (how == 5) {
                for (size_t o = (size_t) lane * 128; o < per; o += 32 * 128) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(src + o) : "memory");
            } else if (how == 2) {
                for (size_t o = (size_t) lane * 128; o < per; o += 32 * 128) {
                    unsigned v;
                    asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(src + o) : "memory");
                    acc ^= v;
                }
            } else if (how == 6) {
                for (size_t o = (size_t) lane * 256; o < per; o += 32 * 256) {
                    unsigned v;
                    asm volatile("ld.global.L2::256B.u32 %0, [%1];" : "=r"(v) : "l"(src + o) : "memory");
                    acc ^= v;
                }
            } else if (how == 3) {
                for (size_t o = (size_t) lane * 128; o < per; o += 32 * 128)
                    asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 16;" ::"r"(s_u32(dummy + lane * 16)), "l"(src + o) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else if (how >= 7 && how <= 10) {
                __shared__ __align__(128) uint8_t tdummy[64 * 16];
                __shared__ __align__(8) unsigned long long tbar;
                if (lane == 0) {
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&tbar)) : "memory");
                    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                    const size_t o0 = (size_t) (src - origin), o1 = o0 + per;
                    const int unit = how == 8 ? 128 : 256;
                    const CUtensorMap* mp = how == 7 ? &maps.m256 : how == 8 ? &maps.m128 : how == 9 ? &maps.mfull : &maps.m256n;
                    unsigned n = 0;
                    for (size_t r = o0 / unit; r * unit < o1; r += 64, ++n) {
                        if (how == 9) {
                            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(mp), "r"(0), "r"((int) r) : "memory");
                        } else {
                            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(s_u32(&tbar)), "r"(1024) : "memory");
                            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s_u32(tdummy)),
                                         "l"(mp), "r"(0), "r"((int) r), "r"(s_u32(&tbar)) : "memory");
                        }
                    }
                    if (how != 9) {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&tbar)) : "memory");
                        wait(s_u32(&tbar), 0);
                    }
                }
            } else if (how == 4) {
                for (size_t o = (size_t) lane * 32; o < per; o += 32 * 32) {
                    unsigned v;
                    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(src + o) : "memory");
                    acc ^= v;
                }
            }
        }
        while ((long long) (gtime() - t) < pf_wait_ns) {}
        if (how == 3) asm volatile("cp.async.wait_all;" ::: "memory");
        if (acc == 0x12345u) sink[0] = acc;
        __syncthreads();
    }
    if (lane != 0 || warp >= nw) return;
    const size_t ncopy = per / cb;
    uint8_t* region = smem + (size_t) warp * depth * cb;
    const unsigned long long t0 = gtime();
    unsigned long long issued = 0, done = 0;
    const unsigned long long total = (ncopy - warp + nw - 1) / nw;
    auto issue = [&](unsigned long long i) {
        const unsigned s = (unsigned) (i % depth);
        const size_t c = warp + i * nw;
        const uint32_t bar = s_u32(&bars[warp * 2 + s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(region + (size_t) s * cb)), "l"(src + c * cb), "r"(cb), "r"(bar) : "memory");
    };
    for (; issued < total && issued < (unsigned) depth; ++issued) issue(issued);
    for (; done < total; ++done) {
        wait(s_u32(&bars[warp * 2 + done % depth]), (unsigned) ((done / depth) & 1));
        if (issued < total) { issue(issued); ++issued; }
    }
    out[blockIdx.x * 16 + warp] = gtime() - t0;
}
int main() {
    uint8_t* big; unsigned long long* out; uint8_t* flush;
    const size_t total = (size_t) 4 << 30;
    cudaMalloc(&big, total); cudaMemset(big, 1, total); cudaMalloc(&out, 148 * 16 * 8); cudaMalloc(&flush, 512 << 20); unsigned* sink; cudaMalloc(&sink, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cb = 5440;
    Maps maps;
    {
        typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        Enc enc = (Enc) f;
        auto mk = [&](CUtensorMap* m, unsigned unit, unsigned box0, CUtensorMapL2promotion promo) {
            const cuuint64_t dims[2] = {unit, total / unit}; const cuuint64_t strides[1] = {unit};
            const cuuint32_t box[2] = {box0, 64}; const cuuint32_t estr[2] = {1, 1};
            CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, big, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc) printf("encode failed %d\n", (int) rc);
        };
        mk(&maps.m256, 256, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        mk(&maps.m128, 128, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
        mk(&maps.mfull, 256, 256, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        mk(&maps.m256n, 256, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    }
    size_t cursor = 0;
    auto fresh = [&](size_t bytes) { // a part of the buffer nobody has touched since the last flush
        if (cursor + bytes > total) cursor = 0;
        const uint8_t* p = big + cursor;
        cursor += (bytes + 4095) / 4096 * 4096;
        return p;
    };
    auto timed = [&](const char* name, const uint8_t* base, size_t per, long long pf_wait_ns, int how = 0) {
        unsigned long long t[148 * 16], mx = 0;
        cudaMemset(out, 0, sizeof t);
        k<<<148, 512, (size_t) 15 * 2 * cb>>>(base, per, cb, pf_wait_ns, out, how, sink, maps, big);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("error %s\n", cudaGetErrorString(e)); return; }
        cudaMemcpy(t, out, sizeof t, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 148 * 16; ++i) mx = t[i] > mx ? t[i] : mx;
        printf("%4zu KB/SM %-44s %6.2f us  %6.2f TB/s\n", per / 1024, name, mx / 1e3, (double) per * 148 / mx / 1e3);
    };
    for (size_t per_kb : {400}) {
        const size_t per = per_kb * 1024 / cb * cb;
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(flush, rep, 512 << 20); cudaDeviceSynchronize();
            timed("cold (HBM)", fresh(per * 148), per, -1);
            for (long long T : {0, 4000, 8000}) {
                char name[64]; snprintf(name, sizeof name, "prefetch, wait %lld us, copy", T / 1000);
                timed(name, fresh(per * 148), per, T);
            }
            const char* hows[] = {"bulk prefetch.L2", "prefetch.global.L2 per line", "ld.L2::128B word per line", "cp.async 16 B/line L2::128B", "ld word per sector", "prefetch.global.L2::evict_last", "ld.L2::256B word per 256 B",
                                  "TMA box 64 x 16 B of 256, promo 256B", "TMA box 64 x 16 B of 128, promo 128B", "TMA tensor prefetch 64 x 256 B", "TMA box 64 x 16 B of 256, no promo"};
            for (int how : {3, 7, 8, 9, 10})
                for (long long T : {8000, 16000}) {
                    char name[96]; snprintf(name, sizeof name, "%s, wait %lld us", hows[how], T / 1000);
                    timed(name, fresh(per * 148), per, T, how);
                }
            {
                const uint8_t* p = fresh(per * 148);
                k_pf<<<148, 32>>>(p, per, 74); cudaDeviceSynchronize(); usleep(200);
                timed("prefetched by another kernel", p, per, -1);
                timed("demand pass before (same data)", p, per, -1);
                timed("demand pass before, again", p, per, -1);
            }
        }
    }
    return 0;
}
