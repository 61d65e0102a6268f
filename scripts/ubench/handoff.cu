// Micro-benchmark: latency of handing a value / a vector from one CTA to others through L2 on B200,
// for the store and load flavours the persistent decode kernel could use.
//   pingpong : CTA a stores word i, CTA b polls it and answers; one-way latency = total / (2 * iters)
//   fanin    : all CTAs each store their ~17 floats of a 2560-float vector, then all CTAs poll the whole
//              vector (sentinel protocol), 64 hops chained; reports us per hop
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handoff handoff.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

constexpr uint32_t kSent = 0x7F808080u;

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <int ST>
__device__ __forceinline__ void store_u32(uint32_t* p, uint32_t v) {
    if (ST == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 1) asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 2) { *p = v; __threadfence(); }
    if (ST == 3) atomicExch(p, v);
    if (ST == 4) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 5) asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 6) asm volatile("st.global.wt.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (ST == 7) asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v - kSent) : "memory");
}
template <int LD>
__device__ __forceinline__ uint32_t load_u32(const uint32_t* p) {
    uint32_t v;
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (LD == 1) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (LD == 2) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (LD == 3) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 load_u4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <int ST, int LD>
__global__ void k_pingpong(uint32_t* ping, uint32_t* pong, int iters, int b, unsigned long long* out) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        const unsigned long long t0 = gtime();
        for (int i = 0; i < iters; ++i) {
            store_u32<ST>(ping + i * 32, 1u);
            while (load_u32<LD>(pong + i * 32) == kSent) {}
        }
        out[0] = gtime() - t0;
    } else if ((int) blockIdx.x == b) {
        for (int i = 0; i < iters; ++i) {
            while (load_u32<LD>(ping + i * 32) == kSent) {}
            store_u32<ST>(pong + i * 32, 1u);
        }
    }
}

// fan-in / fan-out of a D-float vector per hop; SLEEP: nanosleep between failed polls; FENCE: threadfence after the CTA's stores
template <int ST, int SLEEP, int FENCE>
__global__ void __launch_bounds__(512, 1) k_fanin(uint32_t* arena, int D, int hops, unsigned long long* out, float* sink) {
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = (int) ((long long) D * b / G), r1 = (int) ((long long) D * (b + 1) / G);
    float acc = 0.f;
    unsigned long long t0 = 0;
    if (tid == 0) t0 = gtime();
    if (warp >= 15) return;
    for (int h = 0; h < hops; ++h) {
        uint32_t* vec = arena + (size_t) h * D;
        // produce: rows r0.. in 2-row units by warp u % 15, lanes 0/1 (like the GEMV epilogue)
        {
            for (int u = warp; 2 * u < r1 - r0; u += 15) {
                const int row = r0 + 2 * u + lane;
                if (lane < 2 && row < r1) store_u32<ST>(vec + row, __float_as_uint(acc + (float) row));
            }
            if (FENCE == 1) __threadfence();
            if (FENCE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            if (FENCE == 3) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(arena + (size_t) hops * D + 32 * (b * 16 + warp)), "r"(0u) : "memory");
            if (FENCE == 4) asm volatile("membar.gl;" ::: "memory");
            if (FENCE == 5) asm volatile("fence.proxy.alias;" ::: "memory");
            // consume: poll the whole vector, lane owns 8 consecutive values of record warp, warp + 15
            for (int rec = warp; rec * 256 < D; rec += 15) {
                const uint32_t* q = vec + rec * 256 + lane * 8;
                uint4 a = load_u4(q), c = load_u4(q + 4);
                while (a.x == kSent || a.y == kSent || a.z == kSent || a.w == kSent) {
                    if (SLEEP) __nanosleep(SLEEP);
                    a = load_u4(q);
                }
                while (c.x == kSent || c.y == kSent || c.z == kSent || c.w == kSent) {
                    if (SLEEP) __nanosleep(SLEEP);
                    c = load_u4(q + 4);
                }
                acc += __uint_as_float(a.x) * 1e-9f + __uint_as_float(c.w) * 1e-9f;
            }
        }
        asm volatile("bar.sync 1, 480;" ::: "memory");
    }
    if (tid == 0) out[b] = gtime() - t0;
    sink[b * 512 + tid] = acc;
}

// the old way: stores, bar.sync, one release-add on a counter, poll the counter, then load the vector
__global__ void __launch_bounds__(512, 1) k_barrier(uint32_t* arena, unsigned long long* ctr, int D, int hops, unsigned long long* out, float* sink) {
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = (int) ((long long) D * b / G), r1 = (int) ((long long) D * (b + 1) / G);
    float acc = 0.f;
    unsigned long long t0 = 0;
    if (tid == 0) t0 = gtime();
    if (warp >= 15) return;
    for (int h = 0; h < hops; ++h) {
        uint32_t* vec = arena + (size_t) h * D;
        for (int u = warp; 2 * u < r1 - r0; u += 15) {
            const int row = r0 + 2 * u + lane;
            if (lane < 2 && row < r1) vec[row] = __float_as_uint(acc + (float) row);
        }
        asm volatile("bar.sync 1, 480;" ::: "memory");
        if (tid == 0) {
            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ull) : "memory");
            unsigned long long v;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
            } while (v < (unsigned long long) (h + 1) * G);
        }
        asm volatile("bar.sync 1, 480;" ::: "memory");
        for (int rec = warp; rec * 256 < D; rec += 15) {
            const uint32_t* q = vec + rec * 256 + lane * 8;
            const uint4 a = load_u4(q), c = load_u4(q + 4);
            acc += __uint_as_float(a.x) * 1e-9f + __uint_as_float(c.w) * 1e-9f;
        }
    }
    if (tid == 0) out[b] = gtime() - t0;
    sink[b * 512 + tid] = acc;
}

__global__ void k_fill(uint32_t* p, size_t n, uint32_t v) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

template <int ST, int LD>
void run_pp(const char* name, uint32_t* buf, unsigned long long* out, int b) {
    const int iters = 200;
    k_fill<<<64, 256>>>(buf, 2 * iters * 32, kSent);
    k_pingpong<ST, LD><<<148, 32>>>(buf, buf + iters * 32, iters, b, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long t = 0;
    cudaMemcpy(&t, out, 8, cudaMemcpyDeviceToHost);
    printf("pingpong %-34s cta 0 <-> %3d: %7.1f ns one way  %s\n", name, b, (double) t / (2.0 * iters), e ? cudaGetErrorString(e) : "");
}

template <int ST, int SLEEP, int FENCE>
void run_fan(const char* name, uint32_t* arena, unsigned long long* out, float* sink, int D) {
    const int hops = 64, G = 148;
    for (int rep = 0; rep < 2; ++rep) {
        k_fill<<<148, 256>>>(arena, (size_t) hops * D, kSent);
        void* args[] = {&arena, (void*) &D, (void*) &hops, &out, &sink};
        cudaLaunchCooperativeKernel((const void*) k_fanin<ST, SLEEP, FENCE>, dim3(G), dim3(512), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("fanin %s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    unsigned long long t[148];
    cudaMemcpy(t, out, sizeof t, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int i = 0; i < G; ++i) mx = t[i] > mx ? t[i] : mx;
    printf("fanin D=%5d %-40s %6.2f us per hop\n", D, name, (double) mx / hops / 1e3);
}


// ---------------------------------------------------------------- the same hand-off UNDER LOAD: warp 15 of every CTA streams
// 28 KB bulk copies from HBM into a 6-slot shared-memory ring at full rate while warps 0..14 hand vectors over.
// MODE 0: sentinel spin (+ fence)   1: counter barrier then load   2: counter as a hint (wait until all but SLACK
// warps arrived), then sentinel poll   3: sentinel with nanosleep(500) between polls
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
template <int MODE>
__global__ void __launch_bounds__(512, 1)
k_fanin_load(uint32_t* arena, unsigned* ctr, const uint8_t* big, size_t big_bytes, int D, int hops, int slack, unsigned long long* out, float* sink) {
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) unsigned long long bars[6];
    __shared__ volatile int done;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        done = 0;
        for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 15) {
        if (lane != 0) return;
        const size_t per = big_bytes / G / 28672 * 28672;
        const uint8_t* src = big + per * b;
        unsigned long long tiles = 0;
        const unsigned depth = ((slack >> 20) & 7) ? ((slack >> 20) & 7) : 6;
        const unsigned tb = ((slack >> 24) & 0xff) ? ((slack >> 24) & 0xff) * 1024u : 28672u; // bytes per copy
        for (unsigned it = 0; !done; ++it) {
            const unsigned slot = it % depth, par = (it / depth) & 1;
            if (!(slack & 0x20000) && it >= depth) { // wait for the previous copy into this slot
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s_u32(&bars[slot])), "r"(par ^ 1) : "memory");
            }
            if (!(slack & 0x20000)) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(&bars[slot])), "r"(tb) : "memory");
            if (slack & 0x20000) { // L2 prefetch only: nothing crosses the crossbar to this SM; paced by the clock
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t) (it % (per / 28672)) * 28672), "r"(tb) : "memory");
                const unsigned long long tw = gtime() + (unsigned long long) ((slack >> 20) & 7) * 100; // pace: depth field x 100 ns per copy
                while (gtime() < tw) {}
                tiles += tb;
                continue;
            }
            if (slack & 0x10000)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s_u32(ring + slot * 28672)),
                             "l"(src + (size_t) (it % (per / 28672)) * 28672), "r"(tb), "r"(s_u32(&bars[slot])), "l"(0x12F0000000000000ull) : "memory");
            else
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(ring + slot * 28672)),
                             "l"(src + (size_t) (it % (per / 28672)) * 28672), "r"(tb), "r"(s_u32(&bars[slot])) : "memory");
            tiles += tb;
        }
        out[G + b] = tiles;
        return;
    }
    const int r0 = (int) ((long long) D * b / G), r1 = (int) ((long long) D * (b + 1) / G);
    float acc = 0.f;
    const unsigned long long t0 = gtime();
    for (int h = 0; h < hops; ++h) {
        uint32_t* vec = arena + (size_t) h * D;
        for (int u = warp; 2 * u < r1 - r0; u += 15) {
            const int row = r0 + 2 * u + lane;
            if (lane < 2 && row < r1) store_u32<0>(vec + row, __float_as_uint(acc + (float) row));
        }
        __threadfence();
        if (MODE == 1 || MODE == 2) {
            if (lane == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ctr + h * 32), "r"(1u) : "memory");
            const unsigned target = (unsigned) G * 15u - (MODE == 2 ? (unsigned) (slack & 0xffff) : 0u);
            if (lane == 0) while (load_u32<0>(ctr + h * 32) < target) {}
            __syncwarp();
        }
        for (int rec = warp; rec * 256 < D; rec += 15) {
            const uint32_t* q = vec + rec * 256 + lane * 8;
            uint4 a = load_u4(q), c = load_u4(q + 4);
            while (a.x == kSent || a.y == kSent || a.z == kSent || a.w == kSent) {
                if (MODE == 3) __nanosleep(500);
                a = load_u4(q);
            }
            while (c.x == kSent || c.y == kSent || c.z == kSent || c.w == kSent) {
                if (MODE == 3) __nanosleep(500);
                c = load_u4(q + 4);
            }
            acc += __uint_as_float(a.x) * 1e-9f + __uint_as_float(c.w) * 1e-9f;
        }
        asm volatile("bar.sync 1, 480;" ::: "memory");
    }
    if (tid == 0) {
        out[b] = gtime() - t0;
        done = 1;
    }
    sink[b * 512 + tid] = acc;
}

template <int MODE>
void run_load(const char* name, uint32_t* arena, unsigned* ctr, const uint8_t* big, size_t big_bytes, unsigned long long* out, float* sink, int D, int slack) {
    const int hops = 64, G = 148;
    cudaFuncSetAttribute(k_fanin_load<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 28672);
    unsigned long long t[296];
    for (int rep = 0; rep < 2; ++rep) {
        k_fill<<<148, 256>>>(arena, (size_t) hops * D, kSent);
        cudaMemset(ctr, 0, hops * 32 * 4);
        void* args[] = {&arena, &ctr, &big, &big_bytes, (void*) &D, (void*) &hops, (void*) &slack, &out, &sink};
        cudaLaunchCooperativeKernel((const void*) k_fanin_load<MODE>, dim3(G), dim3(512), args, 6 * 28672, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("load %s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(t, out, sizeof t, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0, tiles = 0;
    for (int i = 0; i < G; ++i) { mx = t[i] > mx ? t[i] : mx; tiles += t[G + i]; }
    printf("under load D=%5d %-44s %6.2f us per hop   (background stream %.2f TB/s)\n", D, name, (double) mx / hops / 1e3,
           (double) tiles / (double) mx / 1e3);
}

int main() {
    uint32_t* buf;
    unsigned long long *out, *ctr;
    float* sink;
    cudaMalloc(&buf, 64 << 20);
    cudaMalloc(&out, 148 * 8);
    cudaMalloc(&ctr, 8);
    cudaMalloc(&sink, 148 * 512 * 4);
    for (int b : {1, 74, 147}) {
        run_pp<0, 0>("st.relaxed.gpu / ld.relaxed.gpu", buf, out, b);
        run_pp<1, 1>("st.volatile / ld.volatile", buf, out, b);
        run_pp<2, 0>("st + threadfence / ld.relaxed", buf, out, b);
        run_pp<3, 0>("atomicExch / ld.relaxed", buf, out, b);
        run_pp<4, 2>("st.release / ld.acquire", buf, out, b);
        run_pp<5, 3>("st.cg / ld.cg", buf, out, b);
        run_pp<6, 0>("st.wt / ld.relaxed", buf, out, b);
        run_pp<7, 0>("red.add / ld.relaxed", buf, out, b);
    }
    for (int D : {2560, 9728}) {
        run_fan<0, 0, 0>("st.relaxed.gpu, spin", buf, out, sink, D);
        run_fan<0, 0, 1>("st.relaxed.gpu + __threadfence", buf, out, sink, D);
        run_fan<0, 0, 2>("st.relaxed.gpu + fence.acq_rel.gpu", buf, out, sink, D);
        run_fan<0, 0, 3>("st.relaxed.gpu + red.release.gpu dummy", buf, out, sink, D);
        run_fan<0, 0, 4>("st.relaxed.gpu + membar.gl", buf, out, sink, D);
        run_fan<4, 0, 0>("st.release.gpu each store", buf, out, sink, D);
    }
    {
        uint8_t* big;
        const size_t big_bytes = (size_t) 4 << 30;
        unsigned* ctr;
        cudaMalloc(&big, big_bytes);
        cudaMalloc(&ctr, 64 * 32 * 4);
        cudaMemset(big, 1, big_bytes);
        unsigned long long* out2;
        cudaMalloc(&out2, 296 * 8);
        for (int D : {2560}) {
            char nm[96];
            for (int pace : {7, 6, 5, 4, 3}) {
                snprintf(nm, sizeof nm, "sentinel spin, L2 prefetch 28 KB / %d00 ns", pace);
                run_load<0>(nm, buf, ctr, big, big_bytes, out2, sink, D, 0x20000 | (pace << 20) | (28 << 24));
            }
            run_load<0>("sentinel spin, stream depth 2 x 28 KB", buf, ctr, big, big_bytes, out2, sink, D, (2 << 20) | (28 << 24));
        }
    }
    return 0;
}
