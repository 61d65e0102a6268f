// prefill_gemm.cu -- Q8_0 group-scaled GEMM for prompt prefill on the 5th-gen tensor cores.
//
//   out[t][i] = sum over groups g of ((float) dot_g[t][i] * ws[i][g]) * xs[t][g]
//   dot_g[t][i] = exact int32 dot of the 64 int8 codes of group g          (reference forward.c:79-101,
//                                                                            applied to T tokens at once)
//
// The reference has no prefill: it runs forward() once per prompt token (completion.c:57-66). This
// kernel computes the same matmul for T tokens in one pass, with the reference's arithmetic: the
// int32 group dot is exact (tcgen05.mma kind::i8, s8 x s8 -> s32), each group's term is formed as
// ((float) dot * ws) * xs and the terms are folded into an fp32 accumulator left to right in group
// order -- so every output is BIT-IDENTICAL to the reference matmul of that token.
//
// Two kernels live here. k_prefill_gemm_p (further down) is the one forward_prefill uses: persistent, tokens on the UMMA M
// dimension, tile shapes that fill whole rounds of SMs. k_prefill_gemm is the round-1 kernel, kept selectable (QWEN_GEMM_V=1)
// and as the reference point of profiles/r2_prefill_summary.md. Both produce the same bits.
//
// Structure of k_prefill_gemm (one CTA per 128 weight rows x 128 tokens tile, warp specialised):
//   warp 0  : TMA producer. Per group: a 128-row x 64-byte box of W straight out of the SG layout
//             (cp.async.bulk.tensor.2d, SWIZZLE_64B), the matching 128-token x 64-byte box of the
//             int8 activations, and the 128 activation scales of the group (bulk copy).
//   warp 1  : MMA issuer (one elected lane): 2 x tcgen05.mma.cta_group::1.kind::i8 (M=128, N=128,
//             K=32) per group into one of four TMEM accumulator buffers, tcgen05.commit to mbarriers.
//   warps 2-17: epilogue. tcgen05.ld the group's 128x128 int32 tile (warp w reads TMEM lanes
//             32*(w%4).., 32 columns), scale-promote to fp32 in registers, release the TMEM
//             buffer; after the last group store out[t][i].
// Because the scales change every 64 k-elements on both operands the accumulator must leave TMEM
// every 2 MMAs: the CUDA-core promotion (3 instructions per output per group), not the tensor pipe, is the
// ceiling. Block-scaled MMA kinds do not apply (UE8M0 / E4M3 scale formats only).
// Measured on B200 (4B w1/w3, T = 512): 153 us with 8 epilogue warps, 132 us with 16; of those ~45 us are the
// promotion arithmetic (skipping it: 96 us) and the MMAs themselves are free (not issuing them: -7 %). Deeper
// TMEM / shared-memory pipelines (2 -> 4 accumulators, 6 -> 12 stages), 128-byte SWIZZLE_128B operand rows
// (two groups per stage) and prefetching the weight scales changed nothing or lost a few per cent.
#include <algorithm>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace {

constexpr int kTileM = 128;   // weight rows per tile  (UMMA M, TMEM lanes)
constexpr int kTileN = 128;   // tokens per tile       (UMMA N, TMEM columns); 64 for grids that would leave SMs idle
constexpr int kStages = 12;
constexpr int kAccBufs = 4;   // TMEM accumulator buffers: 4 x 128 columns = all of TMEM
constexpr int kABytes = kTileM * 64;
constexpr int kStageBytes = kABytes + kTileN * 64 + 1024; // A, B (sized for the wide tile), scales; stages stay 1024-aligned
constexpr int kMaxEpiWarps = 16; // wide tile: 4 per scheduler, one warp's TMEM load / barrier wait hides behind the others' arithmetic
constexpr int kThreadsG = 32 * (2 + kMaxEpiWarps);
constexpr unsigned long long kTimeoutNsG = 2000000000ull;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long g_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mb_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a wrong descriptor must not hang the GPU
__device__ __forceinline__ bool mb_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    if (mb_try(bar, parity)) return true;
    unsigned long long t0 = 0;
    for (unsigned spin = 1;; ++spin) {
        if (mb_try(bar, parity)) return true;
        if ((spin & 255u) == 0) {
            if (*(volatile int*) err) return false;
            const unsigned long long now = g_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTimeoutNsG) {
                atomicExch(err, code);
                return false;
            }
        }
    }
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// K-major, SWIZZLE_64B shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 512 B between
// 8-row groups | version 1 [46,48) | layout type 4 = SWIZZLE_64B [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    return (uint64_t) ((saddr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): D = s32 (2 @ bit 4), A,B = signed 8 bit
// (1 @ bits 7 and 10), both K-major, N>>3 @ bit 17, M>>4 @ bit 24
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
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
}

// Packed fp32x2 multiplies for the promotion epilogue (sm_100 FMUL2: two IEEE fp32 products per instruction).
// CAREFUL: ptxas (CUDA 12.9) contracts mul.rn.f32x2 + add.rn.f32x2 -- and even fma(x, y, -0) + fma(t, 1, acc) --
// into ONE FFMA2, which rounds once where the reference rounds twice (forward.c:94-96), -fmad=false or not.
// So only the two multiplies are packed and the accumulation stays two scalar FADDs: the SASS must show
// 2 FMUL2 + 2 FADD per pair of tokens (and no FFMA2 / FFMA);
// tests/test_gpu_parity.py::test_prefill_matmul_batch_bit_identical_to_reference_matmul is the guard.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// (float) v for |v| < 2^22 without I2F: the integer is added to the bit pattern of 1.5 * 2^23 on the integer pipe, where one
// ulp is 1, and the offset is taken off again by an exact fp32 subtraction. Two at a time. MEASURED SLOWER than I2FP in
// k_prefill_gemm_p (a 16-output chunk 400-450 cycles against 300-350: VIADD + FADD2 cost more than the half-rate I2FP they
// replace; the pipes do not overlap, a chunk costs the SUM of its instructions' cycles). Kept for the record, unused.
__device__ __forceinline__ unsigned long long cvt2_exact(int a, int b) {
    return add2(pack2(__int_as_float(a + 0x4B400000), __int_as_float(b + 0x4B400000)), pack2(-12582912.0f, -12582912.0f));
}

struct GemmParams {
    const uint8_t* w;      // SG layout, rows x n
    const float* xsT;      // [groups][Tpad] activation scales, transposed
    float* out;            // [T][d]
    int32_t* dots;         // optional [T][d][groups] exact int32 group dots (test hook)
    int d, n, T, Tpad;
    int* err;
};

// TN = tokens per tile (128, or 64 when 128-token tiles would not fill the SMs); 4 * TN / 32 epilogue warps
template <int TN>
__global__ void __launch_bounds__(kThreadsG, 1)
k_prefill_gemm(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_64B operands want their 512-byte atoms aligned; do not rely on where static shared ends
    uint8_t* smem = smem_raw + ((1024u - (s_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[2 * kStages + 2 * kAccBufs];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kEpiWarps = 4 * TN / 32, kBBytes = TN * 64, kSBytes = TN * 4;
    const int row0 = blockIdx.x * kTileM, t0 = blockIdx.y * TN;
    const int groups = p.n / 64;
    const uint32_t full0 = s_u32(&bars[0]), empty0 = s_u32(&bars[kStages]);
    const uint32_t tfull0 = s_u32(&bars[2 * kStages]), tempty0 = s_u32(&bars[2 * kStages + kAccBufs]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mb_init(full0 + 8 * s, 1);
            mb_init(empty0 + 8 * s, 1 + kEpiWarps); // MMA retire + every epilogue warp done with the stage's scales
        }
        for (int b = 0; b < kAccBufs; ++b) {
            mb_init(tfull0 + 8 * b, 1);
            mb_init(tempty0 + 8 * b, kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) { // TMEM: kAccBufs accumulator buffers x 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_s)), "n"(kAccBufs * TN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int g = 0; g < groups; ++g) {
                const int s = g % kStages;
                const uint32_t par = (g / kStages) & 1;
                if (!mb_wait(empty0 + 8 * s, par ^ 1, p.err, 1)) break;
                const uint32_t base = s_u32(smem + (size_t) s * kStageBytes);
                mb_expect(full0 + 8 * s, kABytes + kBBytes + kSBytes);
                tma_2d(base, &map_w, (g >> 2) * QW_SG_BYTES + (g & 3) * 64, row0, full0 + 8 * s);
                tma_2d(base + kABytes, &map_x, g * 64, t0, full0 + 8 * s);
                bulk_1d(base + kABytes + kTileN * 64, p.xsT + (size_t) g * p.Tpad + t0, kSBytes, full0 + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_i8(kTileM, TN);
            for (int g = 0; g < groups; ++g) {
                const int s = g % kStages, b = g % kAccBufs;
                if (!mb_wait(tempty0 + 8 * b, ((g / kAccBufs) & 1) ^ 1, p.err, 2)) break; // epilogue drained this buffer
                if (!mb_wait(full0 + 8 * s, (g / kStages) & 1, p.err, 3)) break;    // operands landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = s_u32(smem + (size_t) s * kStageBytes);
                const uint32_t d = tmem_base + b * TN;
#pragma unroll
                for (int k = 0; k < 2; ++k) // 64 codes = 2 x K32; the k-th 32-byte slice of every swizzled row
                    umma_i8(d, umma_desc_sw64(base + 32 * k), umma_desc_sw64(base + kABytes + 32 * k), idesc, k);
                umma_commit(empty0 + 8 * s);  // smem stage reusable once these MMAs retire
                umma_commit(tfull0 + 8 * b);  // accumulator ready for the epilogue
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: scale-promote per group
        const int ew = warp - 2;
        if (ew < kEpiWarps) {
        // a warp may only touch TMEM lanes 32*(warp_id % 4) .. +31 (hardware rule, CTA warp id): warp%4 picks the
        // lane quarter (32 weight rows), (warp - 2) / 4 the 32-token column block
        const int lane_blk = warp & 3, col_blk = ew >> 2;
        const int i = row0 + lane_blk * 32 + lane;          // weight row of this thread
        const bool row_ok = i < p.d;
        const uint8_t* wrow = p.w + (size_t) (row_ok ? i : 0) * qw_row_bytes(p.n);
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
        bool ok = true;
        for (int g = 0; g < groups && ok; ++g) {
            const int s = g % kStages, b = g % kAccBufs;
            const float wsc = row_ok ? __ldg(reinterpret_cast<const float*>(wrow + (g >> 2) * QW_SG_BYTES + 256 + (g & 3) * 4)) : 0.0f;
            ok = mb_wait(tfull0 + 8 * b, (g / kAccBufs) & 1, p.err, 4);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const float* xs = reinterpret_cast<const float*>(smem + (size_t) s * kStageBytes + kABytes + kTileN * 64) + col_blk * 32;
            const uint32_t taddr = tmem_base + ((uint32_t) (lane_blk * 32) << 16) + b * TN + col_blk * 32;
            int v0[32];
            tmem_ld32(taddr, v0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mb_arrive(tempty0 + 8 * b); // accumulator buffer may be overwritten
            if (p.dots && row_ok) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int t = t0 + col_blk * 32 + c;
                    if (t < p.T) p.dots[((size_t) t * p.d + i) * groups + g] = v0[c];
                }
            }
            // acc += ((float) dot * ws) * xs, the reference's two roundings and one add (forward.c:94-96)
            const unsigned long long ws2 = pack2(wsc, wsc);
            const unsigned long long* xs2 = reinterpret_cast<const unsigned long long*>(xs);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float lo, hi; // the two multiplies packed (FMUL2), the add scalar: ptxas cannot contract across the two forms
                unpack2(mul2(mul2(pack2((float) v0[2 * c], (float) v0[2 * c + 1]), ws2), xs2[c]), lo, hi);
                acc[2 * c] = __fadd_rn(acc[2 * c], lo);
                acc[2 * c + 1] = __fadd_rn(acc[2 * c + 1], hi);
            }
            __syncwarp();
            if (lane == 0) mb_arrive(empty0 + 8 * s); // done with this stage's activation scales
        }
        if (ok && row_ok) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int t = t0 + col_blk * 32 + c;
                if (t < p.T) p.out[(size_t) t * p.d + i] = acc[c];
            }
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccBufs * TN) : "memory");
    }
}


// =====================================================================================================================
// k_prefill_gemm_p -- the persistent form (round 2). What the measurements said about the kernel above (QWEN_GEMM_DBG,
// profiles/r2_prefill_gemm_components.log; scripts/ubench/tcrate.cu): with the TMA copies, the MMAs, the TMEM loads, the
// scale loads AND the arithmetic all switched off, the 4B w1/w3 problem still takes 79 of its 143 us -- 608 one-tile CTAs in
// 4.1 "waves" (the fifth holds 16 CTAs), each paying TMEM allocation, barrier set-up and an un-overlapped 64 KB output
// store, and 34 mbarrier arrivals per group. TMEM itself is not the limit (tcgen05.ld: 386 B/clk/SM with 16 warps = 170
// cycles for a 128 x 128 int32 tile), and the legacy register-accumulator path (mma.sync m16n8k32.s8) peaks at 1140 TOPS
// before any epilogue. So the structure changed, not the instruction mix:
//   * ONE CTA per SM walks a static list of tiles; TMEM and the barriers are set up once; the producer and the MMA issuer
//     run ahead into the next tile while the epilogue warps store the finished one.
//   * TOKENS are the UMMA M dimension (128 TMEM lanes = 128 tokens, one token per epilogue thread) and WEIGHT ROWS the N
//     dimension, because N may be any multiple of 16 up to 256: the host picks N so that the tile count fills whole rounds
//     of 148 SMs (4B, T = 512: w1/w3 N = 176 -> 444 tiles = 3 x 148; qkv N = 176 -> 140 tiles; wo / w2 N = 80 -> 128).
//   * scales: the weight scales of the tile's rows arrive by TMA as a [N rows][4 groups] box once per 4 groups; a SCALE WARP
//     (warp 18) transposes the box into a [4 groups][N] table and puts the 128 tokens' activation scales of the same four
//     groups next to it (two table slots) -- the epilogue warps pay one barrier wait per four groups, one LDS per group for
//     the token's scale and one broadcast LDS.128 per 4 outputs for the weight scales. (First version: every epilogue warp
//     transposed its own columns and prefetched its own activation scales: ~760 cycles per four groups on all 16 warps.)
//   * a shared-memory stage is released by the MMA's commit alone and an accumulator buffer by one arrival per warp:
//     1 + 16 barrier arrivals per group instead of 34.
// Arithmetic per output and group as above: ((float) dot * ws) * xs, added left to right in group order (bit-identical
// to the reference matmul of every token).
constexpr int kPM = 128;                       // tokens per tile
constexpr int kPMaxN = 192;                    // weight rows per tile, at most (6 units of 8 columns per epilogue warp: 48 accumulators)
constexpr int kPStages = 8;
constexpr int kPABytes = kPM * 64;             // activation codes of one group: 128 tokens x 64 B
constexpr int kPStageBytes = kPABytes + kPMaxN * 64; // 20 KB, keeps stages 1024-aligned
constexpr int kPScSlots = 3;
constexpr int kPScBytes = kPMaxN * 16;         // [N rows][4 scales]
constexpr int kPEpiWarps = 16;
constexpr int kPThreads = 32 * (3 + kPEpiWarps); // producer, MMA issuer, 16 epilogue warps, scale warp
constexpr int kPTblSlots = 2;
constexpr int kPTblBytes = 4 * kPMaxN * 4 + 4 * kPM * 4; // [4 groups][N weight scales] + [4 groups][128 activation scales]
constexpr int kPBarOff = kPStages * kPStageBytes + kPScSlots * kPScBytes + kPTblSlots * kPTblBytes; // mbarriers + the TMEM base word
constexpr int kPBars = 2 * kPStages + 2 * kPScSlots + 2 * 4 + 2 * kPTblSlots;
constexpr int kPSmem = kPBarOff + 8 * kPBars + 16 + 1024;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

struct GemmPParams {
    const float* xsT;      // [groups][Tpad]
    float* out;            // [T][d]
    int32_t* dots;         // optional [T][d][groups]
    int d, n, T, Tpad;
    int tok_tiles, tiles;  // tiles = row tiles x token tiles, token tile fastest (CTAs that run together share weight rows)
    int* err;
    long long* prof;       // QWEN_GEMM_PROF=1: clock64 stamps of CTA 0's first 64 groups, [3 roles][64][8] (normally NULL)
};

template <int U>
__device__ __forceinline__ void tmem_ld_unit(uint32_t taddr, int (&v)[U]);
template <>
__device__ __forceinline__ void tmem_ld_unit<4>(uint32_t taddr, int (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_unit<8>(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// U = columns per tcgen05.ld (4 for N <= 128, 8 above), UC = loads per epilogue warp and group: N = 4 * UC * U weight rows per
// tile, everything about the tile shape a compile-time constant (a first version with run-time unit counts spent more
// issue slots on predicates and index arithmetic than on the promotion: 347 instructions per 48 outputs).
template <int U, int UC, bool DOTS, bool EXACT, bool PROF>
__global__ void __launch_bounds__(kPThreads, 1)
k_prefill_gemm_p(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_s, const GemmPParams p) {
    constexpr int N = 4 * UC * U;                 // weight rows per tile
    constexpr int CW = UC * U;                    // columns per epilogue warp
    constexpr uint32_t NBUF = 512 / N < 4 ? 512 / N : 4; // TMEM accumulator buffers
    static_assert(N % 16 == 0 && N <= kPMaxN && CW <= 64, "tile shape");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (s_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sc_ring = smem + (size_t) kPStages * kPStageBytes;
    uint8_t* tbl_ring = sc_ring + kPScSlots * kPScBytes;
    // the barriers live in the dynamic shared memory too: their addresses are the (register-resident) base plus constants. As a
    // static __shared__ array ptxas rebuilt their window address from SR_CgaCtaId (an S2R, tens of cycles) in front of every wait.
    volatile uint32_t& tmem_base_s = *reinterpret_cast<volatile uint32_t*>(smem + kPBarOff + 8 * kPBars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = p.n / 64;
    const uint32_t full0 = s_u32(smem) + kPBarOff, empty0 = full0 + 8 * kPStages;
    const uint32_t scfull0 = empty0 + 8 * kPStages, scempty0 = scfull0 + 8 * kPScSlots;
    const uint32_t tfull0 = scempty0 + 8 * kPScSlots, tempty0 = tfull0 + 8 * 4;
    const uint32_t wtfull0 = tempty0 + 8 * 4, wtempty0 = wtfull0 + 8 * kPTblSlots;
    if (PROF && threadIdx.x == 0) p.prof[3 * 64 * 8 + 4 * blockIdx.x] = (long long) g_ns();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kPStages; ++s) {
            mb_init(full0 + 8 * s, 1);
            mb_init(empty0 + 8 * s, 1); // the MMAs that read the stage have retired
        }
        for (int s = 0; s < kPScSlots; ++s) {
            mb_init(scfull0 + 8 * s, 1);
            mb_init(scempty0 + 8 * s, 1); // the scale warp has transposed the box
        }
        for (int s = 0; s < kPTblSlots; ++s) {
            mb_init(wtfull0 + 8 * s, 1);
            mb_init(wtempty0 + 8 * s, kPEpiWarps);
        }
        for (int b = 0; b < 4; ++b) {
            mb_init(tfull0 + 8 * b, 1);
            mb_init(tempty0 + 8 * b, kPEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(full0 + 8 * kPBars), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (PROF && threadIdx.x == 0) p.prof[3 * 64 * 8 + 4 * blockIdx.x + 1] = (long long) g_ns();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0, sc_it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < p.tiles && ok; tile += gridDim.x) {
                const int row0 = (tile / p.tok_tiles) * N, t0 = (tile % p.tok_tiles) * kPM;
                for (int g = 0; g < groups; ++g, ++it) {
                    if ((g & 3) == 0) { // the weight scales of groups g .. g + 3: 16 bytes of every row of the tile
                        const uint32_t sl = sc_it % kPScSlots;
                        if (!(ok = mb_wait(scempty0 + 8 * sl, ((sc_it / kPScSlots) & 1) ^ 1, p.err, 5))) break;
                        mb_expect(scfull0 + 8 * sl, (uint32_t) N * 16u);
                        tma_2d(s_u32(sc_ring + sl * kPScBytes), &map_s, (g >> 2) * QW_SG_BYTES + 256, row0, scfull0 + 8 * sl);
                        ++sc_it;
                    }
                    const uint32_t s = it % kPStages;
                    long long* pr = (PROF && blockIdx.x == 0 && it < 64) ? p.prof + (64 + it) * 8 : nullptr;
                    if (pr) pr[0] = clock64();
                    if (!(ok = mb_wait(empty0 + 8 * s, ((it / kPStages) & 1) ^ 1, p.err, 1))) break;
                    if (pr) pr[1] = clock64();
                    const uint32_t base = s_u32(smem + (size_t) s * kPStageBytes);
                    mb_expect(full0 + 8 * s, (uint32_t) (kPABytes + N * 64));
                    tma_2d(base, &map_x, g * 64, t0, full0 + 8 * s);
                    tma_2d(base + kPABytes, &map_w, (g >> 2) * QW_SG_BYTES + (g & 3) * 64, row0, full0 + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer: the whole warp walks the loop (no divergent
        // region around the uniform-datapath instructions), one elected lane issues; the shared-memory descriptors of a stage
        // are the descriptors of stage 0 plus a constant
        const uint32_t idesc = umma_idesc_i8(kPM, N);
        const uint64_t da0 = umma_desc_sw64(s_u32(smem)), db0 = umma_desc_sw64(s_u32(smem) + kPABytes);
        uint32_t it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.tiles && ok; tile += gridDim.x) {
            for (int g = 0; g < groups; ++g, ++it) {
                const uint32_t s = it % kPStages, b = it % NBUF;
                long long* pr = (PROF && blockIdx.x == 0 && it < 64 && lane == 0) ? p.prof + it * 8 : nullptr;
                if (pr) pr[0] = clock64();
                // every lane waits on the same barriers, so `ok` is warp-uniform except after a time-out (then the launch is lost anyway)
                if (!(ok = mb_wait(tempty0 + 8 * b, ((it / NBUF) & 1) ^ 1, p.err, 2))) break; // the epilogue has read this buffer
                if (pr) pr[1] = clock64();
                if (!(ok = mb_wait(full0 + 8 * s, (it / kPStages) & 1, p.err, 3))) break;     // operands landed
                if (pr) pr[2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t so = (uint64_t) ((s * (uint32_t) kPStageBytes) >> 4);
                    const uint32_t dcol = tmem_base + b * (uint32_t) N;
                    // 64 codes = 2 x K32: D[token][row] (+)= X[token][k] * W[row][k]; the second K32 slice sits 32 bytes further
                    umma_i8(dcol, da0 + so, db0 + so, idesc, 0);
                    umma_i8(dcol, da0 + so + 2, db0 + so + 2, idesc, 1);
                    umma_commit(empty0 + 8 * s);
                    umma_commit(tfull0 + 8 * b);
                }
                __syncwarp();
                if (pr) pr[3] = clock64();
            }
        }
    } else if (warp == 2 + kPEpiWarps) {
        // ------------------------------------------------------------ scale warp: one table per four groups
        uint32_t k = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.tiles && ok; tile += gridDim.x) {
            const float* xcol = p.xsT + (tile % p.tok_tiles) * kPM + 4 * lane; // this lane's 4 tokens (xsT rows are padded to Tpad)
            for (int g = 0; g < groups; g += 4, ++k) {
                float4 x4[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    x4[kk] = g + kk < groups ? __ldg(reinterpret_cast<const float4*>(xcol + (size_t) (g + kk) * p.Tpad)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t sl = k % kPScSlots, ts = k % kPTblSlots;
                if (!(ok = mb_wait(scfull0 + 8 * sl, (k / kPScSlots) & 1, p.err, 6))) break;           // the box has landed
                if (!(ok = mb_wait(wtempty0 + 8 * ts, ((k / kPTblSlots) & 1) ^ 1, p.err, 7))) break;   // the table slot has been read
                const float4* box = reinterpret_cast<const float4*>(sc_ring + sl * kPScBytes);
                float* tw = reinterpret_cast<float*>(tbl_ring + ts * kPTblBytes);
                float* tx = tw + 4 * kPMaxN;
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 32) {
                    const int c = c0 + lane;
                    if (N % 32 == 0 || c < N) {
                        const float4 s4 = box[c];
                        tw[c] = s4.x; tw[kPMaxN + c] = s4.y; tw[2 * kPMaxN + c] = s4.z; tw[3 * kPMaxN + c] = s4.w;
                    }
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) *reinterpret_cast<float4*>(tx + kk * kPM + 4 * lane) = x4[kk];
                __syncwarp();
                if (lane == 0) {
                    mb_arrive(wtfull0 + 8 * ts);
                    mb_arrive(scempty0 + 8 * sl);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: one token per thread, CW columns per warp
        const int ew = warp - 2;
        const int q = warp & 3, j = ew >> 2;             // TMEM lane quarter (hardware rule: CTA warp id % 4), column block
        const int cbeg = j * CW;                         // this warp's columns of the tile
        uint32_t it = 0, sc_it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.tiles && ok; tile += gridDim.x) {
            const int row0 = (tile / p.tok_tiles) * N, t0 = (tile % p.tok_tiles) * kPM;
            const int t = t0 + q * 32 + lane;
            const bool t_ok = t < p.T;
            float acc[CW];
#pragma unroll
            for (int c = 0; c < CW; ++c) acc[c] = 0.0f;
            for (int g = 0; g < groups; ++g, ++it) {
                const int gk = g & 3;
                long long* pr = (PROF && blockIdx.x == 0 && it < 64 && warp == 2 && lane == 0) ? p.prof + (128 + it) * 8 : nullptr;
                if (pr) pr[0] = clock64();
                const uint32_t ts = sc_it % kPTblSlots;
                if (gk == 0 && !(ok = mb_wait(wtfull0 + 8 * ts, (sc_it / kPTblSlots) & 1, p.err, 8))) break; // the scale table of groups g .. g + 3
                const float* tw = reinterpret_cast<const float*>(tbl_ring + ts * kPTblBytes);
                const float xsc = tw[4 * kPMaxN + gk * kPM + q * 32 + lane]; // this token's activation scale
                const float* wg = tw + gk * kPMaxN + cbeg;                   // the weight scales of this warp's columns
                const uint32_t b = it % NBUF;
                if (pr) pr[1] = clock64();
                if (!(ok = mb_wait(tfull0 + 8 * b, (it / NBUF) & 1, p.err, 4))) break;
                if (pr) pr[2] = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + b * (uint32_t) N + cbeg;
                const unsigned long long xs2 = pack2(xsc, xsc);
                // TMEM -> registers 16 columns (CH loads) per round trip; the compiler keeps one chunk's registers, so the next
                // chunk's loads are issued behind this chunk's arithmetic; the buffer goes back to the MMA issuer as soon as the
                // last load has landed
                constexpr int CH = 16 / U, NCH = (UC + CH - 1) / CH;
                int v[2][CH][U];
#pragma unroll
                for (int h = 0; h < CH; ++h)
                    if (h < UC) tmem_ld_unit<U>(taddr + h * U, v[0][h]);
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (pr && k < 2) pr[3 + k] = clock64();
                    if (k + 1 < NCH) {
#pragma unroll
                        for (int h = 0; h < CH; ++h)
                            if ((k + 1) * CH + h < UC) tmem_ld_unit<U>(taddr + ((k + 1) * CH + h) * U, v[(k + 1) & 1][h]);
                    } else {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mb_arrive(tempty0 + 8 * b);
                    }
#pragma unroll
                    for (int h = 0; h < CH; ++h) {
                        if (k * CH + h < UC) {
                            const int (&vv)[U] = v[k & 1][h];
                            if (DOTS && t_ok) { // test hook: a separate instantiation, the product kernel carries no such code
#pragma unroll
                                for (int e = 0; e < U; ++e) {
                                    const int row = row0 + cbeg + (k * CH + h) * U + e;
                                    if (row < p.d) p.dots[((size_t) t * p.d + row) * groups + g] = vv[e];
                                }
                            }
#pragma unroll
                            for (int e = 0; e < U; e += 4) {
                                const int c = (k * CH + h) * U + e;
                                const float4 w4 = *reinterpret_cast<const float4*>(wg + c);
                                float lo, hi;
                                if (EXACT) { // two packed multiplies, scalar adds (see the note at mul2): the reference's three roundings
                                    unpack2(mul2(mul2(pack2((float) vv[e], (float) vv[e + 1]), pack2(w4.x, w4.y)), xs2), lo, hi);
                                    acc[c] = __fadd_rn(acc[c], lo);
                                    acc[c + 1] = __fadd_rn(acc[c + 1], hi);
                                    unpack2(mul2(mul2(pack2((float) vv[e + 2], (float) vv[e + 3]), pack2(w4.z, w4.w)), xs2), lo, hi);
                                    acc[c + 2] = __fadd_rn(acc[c + 2], lo);
                                    acc[c + 3] = __fadd_rn(acc[c + 3], hi);
                                } else { // acc = fma((float) dot * ws, xs, acc): one rounding fewer per term, one FMA-pipe pass fewer per output
                                    unpack2(fma2(mul2(pack2((float) vv[e], (float) vv[e + 1]), pack2(w4.x, w4.y)), xs2, pack2(acc[c], acc[c + 1])), lo, hi);
                                    acc[c] = lo;
                                    acc[c + 1] = hi;
                                    unpack2(fma2(mul2(pack2((float) vv[e + 2], (float) vv[e + 3]), pack2(w4.z, w4.w)), xs2, pack2(acc[c + 2], acc[c + 3])), lo, hi);
                                    acc[c + 2] = lo;
                                    acc[c + 3] = hi;
                                }
                            }
                        }
                    }
                }
                if (gk == 3 || g == groups - 1) { // done with this table
                    __syncwarp();
                    if (lane == 0) mb_arrive(wtempty0 + 8 * ts);
                    ++sc_it;
                }
                if (pr) pr[7] = clock64();
            }
            if (PROF && warp == 2 && lane == 0 && tile == (int) blockIdx.x) p.prof[3 * 64 * 8 + 4 * blockIdx.x + 2] = (long long) g_ns();
            if (ok && t_ok) { // the tile is done: the stores overlap the next tile's main loop
                float* orow = p.out + (size_t) t * p.d + row0 + cbeg;
                if (row0 + cbeg + CW <= p.d && (p.d & 3) == 0) {
#pragma unroll
                    for (int c = 0; c < CW; c += 4) *reinterpret_cast<float4*>(orow + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
                } else {
#pragma unroll
                    for (int c = 0; c < CW; ++c)
                        if (row0 + cbeg + c < p.d) orow[c] = acc[c];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PROF && threadIdx.x == 0) p.prof[3 * 64 * 8 + 4 * blockIdx.x + 3] = (long long) g_ns();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

typedef void (*GemmPKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmPParams);
// weight rows per tile -> kernel: N = 16 .. 192 in steps of 16 (U = 4 columns per tcgen05.ld, or 8 where N / 4 is a multiple of 8)
template <bool DOTS, bool EXACT, bool PROF = false>
GemmPKernel gemm_p_kernel(int N) {
    switch (N) {
        case 16: return k_prefill_gemm_p<4, 1, DOTS, EXACT, PROF>;
        case 32: return k_prefill_gemm_p<4, 2, DOTS, EXACT, PROF>;
        case 48: return k_prefill_gemm_p<4, 3, DOTS, EXACT, PROF>;
        case 64: return k_prefill_gemm_p<4, 4, DOTS, EXACT, PROF>;
        case 80: return k_prefill_gemm_p<4, 5, DOTS, EXACT, PROF>;
        case 96: return k_prefill_gemm_p<4, 6, DOTS, EXACT, PROF>;
        case 112: return k_prefill_gemm_p<4, 7, DOTS, EXACT, PROF>;
        case 128: return k_prefill_gemm_p<4, 8, DOTS, EXACT, PROF>;
        case 144: return k_prefill_gemm_p<4, 9, DOTS, EXACT, PROF>;
        case 160: return k_prefill_gemm_p<8, 5, DOTS, EXACT, PROF>;
        case 176: return k_prefill_gemm_p<4, 11, DOTS, EXACT, PROF>;
        case 192: return k_prefill_gemm_p<8, 6, DOTS, EXACT, PROF>;
        default: return nullptr;
    }
}
constexpr int kPShapes[] = {16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 176, 192};
// weight rows per tile for a d-row matrix and tok_tiles token tiles on `sms` SMs: minimise rounds x (N + a per-group fixed
// cost worth ~24 rows), rounds = ceil(tiles / SMs); ties go to the smaller N (more SMs busy)
int gemm_tile_rows(int d, int tok_tiles, int sms) {
    int bestN = 16;
    long best = -1;
    for (int N : kPShapes) {
        const long tiles = (long) ((d + N - 1) / N) * tok_tiles, rounds = (tiles + sms - 1) / sms;
        const long cost = rounds * (N + 24);
        if (best < 0 || cost < best) {
            best = cost;
            bestN = N;
        }
    }
    return bestN;
}

// Measured int8 tensor peak (MEASURED_PEAKS.json has no int8 figure, "the builder must measure it on the
// box"): every CTA issues `iters` back-to-back tcgen05.mma.cta_group::1.kind::i8 M128 N256 K32 on resident shared-memory
// operands into one TMEM accumulator -- no loads, no promotion, no epilogue. The operand bytes are whatever shared memory
// holds; integer MMAs have no data-dependent timing.
__global__ void __launch_bounds__(128, 1) k_int8_peak(int iters, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (s_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 64 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (uint32_t) (i & 3);
    if (threadIdx.x == 0) {
        mb_init(s_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_s)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = umma_idesc_i8(128, 256);
        const uint32_t a = s_u32(smem), b = s_u32(smem) + 128 * 64;
        for (int i = 0; i < iters; ++i) umma_i8(tmem_base, umma_desc_sw64(a + 32 * (i & 1)), umma_desc_sw64(b + 32 * (i & 1)), idesc, 1);
        umma_commit(s_u32(&bar));
        mb_wait(s_u32(&bar), 0, err, 9);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_fn() {
    static EncodeTiled fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiled) f;
    }
    return fn;
}

int make_map(CUtensorMap* m, const void* base, uint64_t row_bytes, uint64_t rows, uint32_t box_rows, uint32_t box_bytes = 64,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_64B) {
    EncodeTiled enc = encode_fn();
    if (!enc) {
        qw_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return -1;
    }
    const cuuint64_t dims[2] = {row_bytes, rows};
    const cuuint64_t strides[1] = {row_bytes};
    const cuuint32_t box[2] = {box_bytes, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        qw_set_error("cuTensorMapEncodeTiled failed (%d) for %llu x %llu", (int) rc, (unsigned long long) row_bytes,
                     (unsigned long long) rows);
        return -1;
    }
    return 0;
}

} // namespace

// Device-side entry: W in SG layout (d rows, n columns), xq [T][n] int8 (n % 64 == 0, row pitch n, 16-byte
// aligned), xsT [n/64][Tpad] fp32, out [T][d]. dots may be NULL. Asynchronous on `st`; *err_dev must be 0.
int qw_prefill_gemm(const uint8_t* w, const int8_t* xq, const float* xsT, float* out, int32_t* dots, int d, int n, int T,
                    int Tpad, int* err_dev, cudaStream_t st, float* ms_out) {
    if (n % 64 || n % 16 || T <= 0 || d <= 0) {
        qw_set_error("prefill gemm: bad shape d=%d n=%d T=%d", d, n, T);
        return -2;
    }
    CUtensorMap mw, mx, ms;
    static int sms = 0, variant = -1;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (variant < 0) {
        const char* e = getenv("QWEN_GEMM_V"); // 1: the one-tile-per-CTA kernel of round 1; default: the persistent kernel
        variant = e ? atoi(e) : 2;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_out) {
        QW_CUDA(cudaEventCreate(&e0));
        QW_CUDA(cudaEventCreate(&e1));
    }
    if (variant != 1) {
        // Persistent kernel: weight rows per tile chosen so that the tiles fill whole rounds of SMs (gemm_tile_rows)
        const int tok_tiles = (T + kPM - 1) / kPM;
        const int bestN = gemm_tile_rows(d, tok_tiles, sms);
        static int forceN = -1;
        if (forceN < 0) {
            const char* e = getenv("QWEN_GEMM_N");
            forceN = e ? atoi(e) : 0;
        }
        const int N = forceN > 0 ? forceN : bestN;
        static long long* prof = nullptr;
        static int want_prof = -1;
        static int prof_skip = 0; // QWEN_GEMM_PROF=k: dump the stamps of the k-th launch of the process
        if (want_prof < 0) {
            want_prof = getenv("QWEN_GEMM_PROF") ? 1 : 0;
            prof_skip = want_prof ? std::max(0, atoi(getenv("QWEN_GEMM_PROF")) - 1) : 0;
        }
        const bool prof_now = want_prof == 1 && prof_skip-- <= 0;
        if (prof_now && !prof) {
            cudaMalloc((void**) &prof, (3 * 64 * 8 + 4 * 160) * 8);
            cudaMemset(prof, 0, (3 * 64 * 8 + 4 * 160) * 8);
        }
        static int exact = -1;
        if (exact < 0) {
            const char* e = getenv("QWEN_GEMM_EXACT"); // 0: fma fold (one rounding fewer per term)
            exact = e ? atoi(e) : 1;
        }
        GemmPKernel kern = dots ? gemm_p_kernel<true, true>(N) : exact ? gemm_p_kernel<false, true>(N) : gemm_p_kernel<false, false>(N);
        if (prof_now && !dots) kern = gemm_p_kernel<false, true, true>(N); // the stamped instantiation (exact fold)
        if (!kern) {
            qw_set_error("prefill gemm: no kernel for %d weight rows per tile", N);
            return -2;
        }
        if (make_map(&mx, xq, (uint64_t) n, (uint64_t) T, kPM) || make_map(&mw, w, qw_row_bytes(n), (uint64_t) d, (uint32_t) N)
            || make_map(&ms, w, qw_row_bytes(n), (uint64_t) d, (uint32_t) N, 16, CU_TENSOR_MAP_SWIZZLE_NONE))
            return -1;
        static bool attr_p = false;
        if (!attr_p) {
            for (int n_ : kPShapes) {
                QW_CUDA(cudaFuncSetAttribute((const void*) gemm_p_kernel<false, true>(n_), cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
                QW_CUDA(cudaFuncSetAttribute((const void*) gemm_p_kernel<false, false>(n_), cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
                QW_CUDA(cudaFuncSetAttribute((const void*) gemm_p_kernel<true, true>(n_), cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
                QW_CUDA(cudaFuncSetAttribute((const void*) gemm_p_kernel<false, true, true>(n_), cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
            }
            attr_p = true;
        }
        GemmPParams pp{xsT, out, dots, d, n, T, Tpad, tok_tiles, ((d + N - 1) / N) * tok_tiles, err_dev, prof_now ? prof : nullptr};
        const int grid = std::min(sms, pp.tiles);
        if (ms_out) QW_CUDA(cudaEventRecord(e0, st));
        kern<<<grid, kPThreads, kPSmem, st>>>(mx, mw, ms, pp);
        QW_CUDA(cudaGetLastError());
        if (prof_now) { // one dump per process
            want_prof = 2;
            cudaStreamSynchronize(st);
            static long long h[3 * 64 * 8 + 4 * 160];
            cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
            const long long z = h[0];
            {
                const long long* c = h + 3 * 64 * 8;
                long long t0 = c[0], t1 = 0, setup = 0, first = 0, last_end_min = 1ll << 62;
                for (int b = 0; b < grid; ++b) {
                    t0 = std::min(t0, c[4 * b]);
                    t1 = std::max(t1, c[4 * b + 3]);
                    last_end_min = std::min(last_end_min, c[4 * b + 3]);
                    setup = std::max(setup, c[4 * b + 1] - c[4 * b]);
                    first = std::max(first, c[4 * b + 2] - c[4 * b]);
                }
                {
                    std::vector<long long> sp;
                    for (int b = 0; b < grid; ++b) sp.push_back(c[4 * b + 3] - c[4 * b]);
                    std::sort(sp.begin(), sp.end());
                    fprintf(stderr, "[gemm prof] per-CTA lifetime ns: min %lld p25 %lld median %lld p75 %lld max %lld\n", sp[0], sp[grid / 4], sp[grid / 2], sp[3 * grid / 4], sp[grid - 1]);
                }
                fprintf(stderr, "[gemm prof] CTA spans (globaltimer ns): first start -> last end %lld, earliest end %lld, start skew %lld, setup <= %lld, first tile done <= %lld\n",
                        t1 - t0, last_end_min - t0, [&] { long long m = 0; for (int b = 0; b < grid; ++b) m = std::max(m, c[4 * b] - t0); return m; }(), setup, first);
            }
            fprintf(stderr, "[gemm prof] d=%d n=%d T=%d N=%d tiles=%d (cycles since the MMA issuer's first stamp)\n", d, n, T, N, pp.tiles);
            for (int i = 0; i < 20; ++i) {
                const long long *m = h + i * 8, *pd = h + (64 + i) * 8, *ep = h + (128 + i) * 8;
                fprintf(stderr, "[gemm prof] g%02d mma top %6lld tempty %6lld full %6lld commit %6lld | prod top %6lld empty %6lld | epi top %6lld table %6lld tfull %6lld ld0 %6lld ld1 %6lld | - %6lld - %6lld | done %6lld\n",
                        i, m[0] - z, m[1] - z, m[2] - z, m[3] - z, pd[0] - z, pd[1] - z, ep[0] - z, ep[1] - z, ep[2] - z, ep[3] - z,
                        ep[4] ? ep[4] - z : 0, ep[5] ? ep[5] - z : 0, ep[6] ? ep[6] - z : 0, ep[7] - z);
            }
        }
    } else {
    // 128-token tiles unless 64-token tiles finish sooner: time ~ waves x cost of one tile, and a 64-token tile costs
    // ~0.7 of a 128-token one (measured). E.g. 1.7B wo / w2 at T = 512: 64 wide tiles on 148 SMs vs 128 narrow ones
    // in one wave -> narrow; 4B w2: 80 wide tiles in one wave vs 160 narrow ones in two -> wide.
    const int rt = (d + kTileM - 1) / kTileM;
    const int waves_w = (rt * ((T + kTileN - 1) / kTileN) + sms - 1) / sms, waves_n = (rt * ((T + 63) / 64) + sms - 1) / sms;
    const bool wide = waves_w * 100 <= waves_n * 72;
    const int TN = wide ? kTileN : 64;
    if (make_map(&mw, w, qw_row_bytes(n), (uint64_t) d, kTileM) || make_map(&mx, xq, (uint64_t) n, (uint64_t) T, TN)) return -1;
    static bool attr = false;
    const size_t smem = (size_t) kStages * kStageBytes + 1024;
    if (!attr) {
        QW_CUDA(cudaFuncSetAttribute(k_prefill_gemm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        QW_CUDA(cudaFuncSetAttribute(k_prefill_gemm<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr = true;
    }
    GemmParams p{w, xsT, out, dots, d, n, T, Tpad, err_dev};
    const dim3 grid((d + kTileM - 1) / kTileM, (T + TN - 1) / TN);
    if (ms_out) QW_CUDA(cudaEventRecord(e0, st));
    if (wide) k_prefill_gemm<128><<<grid, kThreadsG, smem, st>>>(mw, mx, p);
    else k_prefill_gemm<64><<<grid, kThreadsG, smem, st>>>(mw, mx, p);
    QW_CUDA(cudaGetLastError());
    }
    if (ms_out) {
        QW_CUDA(cudaEventRecord(e1, st));
        QW_CUDA(cudaEventSynchronize(e1));
        QW_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return 0;
}

// Host-in / host-out wrapper: the batched form of matmul() (reference forward.c:79-101 for T tokens).
// xq [T][n], xs [T][n/64], wq [d][n], ws [d][n/64] in checkpoint layout; out [T][d]; dots optional
// [T][d][n/64]. reps > 1 re-runs the kernel for timing; *ms (optional) = best device time of one run.
extern "C" int qwen_cuda_matmul_batch(float* out, int32_t* dots, const int8_t* xq, const float* xs, const int8_t* wq,
                                      const float* ws, int n, int d, int T, int reps, float* ms) {
    if (qwen_cuda_device_count() <= 0) {
        qw_set_error("no CUDA device: this library has no CPU path");
        return -1;
    }
    if (n <= 0 || n % 64 || d <= 0 || T <= 0) {
        qw_set_error("matmul_batch: n must be a positive multiple of 64");
        return -2;
    }
    const int groups = n / 64, Tpad = (T + kTileN - 1) / kTileN * kTileN;
    int8_t *dwq = nullptr, *dxq = nullptr;
    float *dws = nullptr, *dxsT = nullptr, *dout = nullptr;
    uint8_t* dw = nullptr;
    int32_t* ddots = nullptr;
    int* derr = nullptr;
    int rc = -1;
    float* xsT = (float*) calloc((size_t) groups * Tpad, sizeof(float));
    for (int t = 0; t < T && xsT; ++t)
        for (int g = 0; g < groups; ++g) xsT[(size_t) g * Tpad + t] = xs[(size_t) t * groups + g];
    do {
        if (!xsT) break;
        if (cudaMalloc(&dwq, (size_t) d * n) || cudaMalloc(&dws, (size_t) d * groups * 4) || cudaMalloc(&dw, qw_row_bytes(n) * d)
            || cudaMalloc(&dxq, (size_t) Tpad * n) || cudaMalloc(&dxsT, (size_t) groups * Tpad * 4)
            || cudaMalloc(&dout, (size_t) T * d * 4) || cudaMalloc(&derr, 4)
            || (dots && cudaMalloc(&ddots, (size_t) T * d * groups * 4))) {
            qw_set_error("matmul_batch: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaMemset(derr, 0, 4);
        cudaMemset(dxq, 0, (size_t) Tpad * n);
        cudaMemcpy(dwq, wq, (size_t) d * n, cudaMemcpyHostToDevice);
        cudaMemcpy(dws, ws, (size_t) d * groups * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dxq, xq, (size_t) T * n, cudaMemcpyHostToDevice);
        cudaMemcpy(dxsT, xsT, (size_t) groups * Tpad * 4, cudaMemcpyHostToDevice);
        launch_repack(dwq, dws, n, 0, n, d, dw, 0, 1, 0);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            qw_set_error("matmul_batch: repack failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        float best = 1e30f;
        bool bad = false;
        for (int r = 0; r < (reps > 0 ? reps : 1); ++r) {
            float t = 0;
            if (qw_prefill_gemm(dw, dxq, dxsT, dout, r == 0 ? ddots : nullptr, d, n, T, Tpad, derr, 0, &t)) {
                bad = true;
                break;
            }
            if (t < best) best = t;
        }
        if (bad) break;
        if (cudaDeviceSynchronize() != cudaSuccess) {
            qw_set_error("matmul_batch: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        int herr = 0;
        cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
        if (herr) {
            qw_set_error("matmul_batch: pipeline wait %d timed out (descriptor / barrier bug)", herr);
            break;
        }
        cudaMemcpy(out, dout, (size_t) T * d * 4, cudaMemcpyDeviceToHost);
        if (dots) cudaMemcpy(dots, ddots, (size_t) T * d * groups * 4, cudaMemcpyDeviceToHost);
        if (ms) *ms = best;
        rc = 0;
    } while (0);
    free(xsT);
    cudaFree(dwq); cudaFree(dws); cudaFree(dw); cudaFree(dxq); cudaFree(dxsT); cudaFree(dout); cudaFree(derr); cudaFree(ddots);
    return rc;
}

// Measured dense int8 tensor throughput of this device (tera-ops/s, 2 ops per MAC): see k_int8_peak. Best of `reps` runs.
extern "C" int qwen_cuda_int8_peak(int iters, int reps, float* tops) {
    if (qwen_cuda_device_count() <= 0) {
        qw_set_error("no CUDA device: this library has no CPU path");
        return -1;
    }
    if (iters < 1 || reps < 1 || !tops) return -2;
    int dev = 0, sms = 0;
    QW_CUDA(cudaGetDevice(&dev));
    QW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int* derr = nullptr;
    QW_CUDA(cudaMalloc(&derr, 4));
    QW_CUDA(cudaMemset(derr, 0, 4));
    const size_t smem = (128 + 256) * 64 + 1024;
    cudaEvent_t e0, e1;
    QW_CUDA(cudaEventCreate(&e0));
    QW_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r <= reps; ++r) { // run 0 warms up
        QW_CUDA(cudaEventRecord(e0, 0));
        k_int8_peak<<<sms, 128, smem, 0>>>(iters, derr);
        QW_CUDA(cudaEventRecord(e1, 0));
        QW_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        QW_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    int herr = 0;
    QW_CUDA(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    cudaFree(derr);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (herr) {
        qw_set_error("int8 peak: the MMA chain did not retire (wait %d timed out)", herr);
        return -1;
    }
    *tops = (float) (2.0 * 128 * 256 * 32 * (double) iters * sms / (best * 1e-3) / 1e12);
    return 0;
}

// Host-only test hook (no GPU needed): the tile shape qw_prefill_gemm picks for a d-row matrix and T tokens on `sms` SMs.
// Returns the weight rows per tile N; *tiles = row tiles x token tiles.
extern "C" int qwen_cuda_debug_gemm_plan(int d, int T, int sms, int* tiles) {
    if (d <= 0 || T <= 0 || sms <= 0) return -2;
    const int tok_tiles = (T + kPM - 1) / kPM;
    const int N = gemm_tile_rows(d, tok_tiles, sms);
    if (tiles) *tiles = ((d + N - 1) / N) * tok_tiles;
    return N;
}
