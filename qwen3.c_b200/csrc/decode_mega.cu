// decode_mega.cu -- the whole decode step (reference src/forward.c:225-350) as ONE
// persistent cooperative sm_100a kernel.
//
// Why one kernel: at batch 1 every op is a matrix-vector product whose only cost is
// streaming its weights from HBM once. A layer is ~107 MB (Qwen3-4B) = ~16 us of HBM
// time split over 7 dependent GEMVs; a launch-per-op design spends about as long in
// launch gaps and pipeline fill/drain as it does streaming. Here one CTA per SM stays
// resident for the whole token:
//
//   * warp 16 (one elected lane) is the PRODUCER. It walks the step's fixed schedule of
//     weight tiles and KV-cache chunks and copies each into a shared-memory ring with
//     cp.async.bulk (TMA bulk copy, completion on an mbarrier). It never waits for
//     activations, so it runs up to one ring (6-7 x 28 KB per SM, ~29 MB chip-wide)
//     AHEAD of the math, straight through phase boundaries and grid barriers. HBM
//     stays busy while the consumers synchronise.
//   * warps 0..15 are CONSUMERS. Per phase they build the quantised activation vector
//     in shared memory (RMSNorm + Q8_0 quantise, fused), then eat tiles from the ring:
//     a half-warp takes one 272-byte super-group record, 4 x dp4a per lane, int32 group
//     dots by shuffle, fp32 scaling exactly as the reference does, and one thread per
//     row folds the group terms left to right -- the reference's own order, so a GEMV
//     is bit-identical to the reference for identical inputs.
//   * phases are separated by a grid-wide barrier (one atomic + spin per CTA).
//
// Work split: every matrix is split by contiguous row ranges over the CTAs (no split-K,
// no atomics -> deterministic). Attention is split-KV over (kv head, 28-position chunk)
// units with an online-softmax merge in a small combine phase.
//
// Every wait in this file has a wall-clock timeout that raises a sticky error flag
// instead of hanging the GPU.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kSlotBytes = 28672;           // one ring slot: 105 SG records or 2 x 28 KV rows
constexpr int kChunk = 28;                  // KV positions per attention unit (2*28*512 B = one slot)
constexpr int kMaxSlots = 8;
constexpr int kPartFloats = (kSlotBytes / QW_SG_BYTES) * 4; // group terms of one tile
constexpr unsigned long long kTimeoutNs = 4000000000ull;
constexpr int kProfSlots = 16; // per layer: stamps after each phase step (CTA-local, thread 0)

struct MegaParams {
    int D, Hdl, L, Hl, KVHl, Pl, Kl, Vl, S, kv_mul;
    int pos, token, layers_run;
    const int* token_dev;
    const uint8_t *w_qkv, *w_o, *w_13, *w_2, *w_cls, *w_emb;
    size_t s_qkv, s_o, s_13, s_2;
    const float *att_norm, *ffn_norm, *out_norm, *q_norm, *k_norm, *rope_cos, *rope_sin;
    float *k_cache, *v_cache;
    float *x, *qkv, *att, *h, *logits;
    int8_t* att_q;
    float* att_s;
    float *part_m, *part_l, *part_acc;
    unsigned long long* bar;
    unsigned long long bar_base;
    int* err;
    unsigned long long* prof; // optional [CTA][kProfSlots] globaltimer stamps (debug)
    int nslot, off_xq, off_xs, off_scr, off_part, off_misc, off_bar;
};

struct MegaState {
    int8_t* att_q = nullptr;
    float* att_s = nullptr;
    float *part_m = nullptr, *part_l = nullptr, *part_acc = nullptr;
    int grid = 0, nslot = 0;
    unsigned long long* prof = nullptr;
    size_t smem = 0;
    int off_xq, off_xs, off_scr, off_part, off_misc, off_bar;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
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
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct Shared {
    uint8_t* ring;
    int8_t* xq;
    float* xs;
    float* scr;
    float* part;
    float* misc;
    uint32_t full, empty; // shared-space addresses of the barrier arrays
    volatile int* abort_flag;
};

__device__ __forceinline__ void mbar_wait(const Shared& sh, const MegaParams& p, uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = gtime_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (*sh.abort_flag) return;
        if (gtime_ns() - t0 > kTimeoutNs) {
            *sh.abort_flag = code;
            *p.err = code;
            return;
        }
    }
}

// grid-wide barrier among the consumer threads of every CTA
__device__ __forceinline__ void grid_barrier(const Shared& sh, const MegaParams& p, int& nbar) {
    __threadfence();
    bar_consumers();
    if (threadIdx.x == 0) {
        const unsigned long long target = p.bar_base + (unsigned long long) (nbar + 1) * gridDim.x;
        __threadfence();
        atomicAdd(p.bar, 1ull);
        if (ld_acquire_u64(p.bar) < target) {
            const unsigned long long t0 = gtime_ns();
            while (ld_acquire_u64(p.bar) < target) {
                if (*sh.abort_flag) break;
                if (gtime_ns() - t0 > kTimeoutNs) {
                    *sh.abort_flag = 100 + nbar;
                    *p.err = 100 + nbar;
                    break;
                }
            }
        }
        __threadfence();
    }
    bar_consumers();
    ++nbar;
}

// ---------------------------------------------------------------- schedule (shared by producer and consumers)
struct MatPhase {
    const uint8_t* base;
    int rows, n, gran;
};
__device__ __forceinline__ void cta_rows(const MatPhase& m, int& r0, int& r1) {
    const long long units = m.rows / m.gran;
    r0 = (int) (units * blockIdx.x / gridDim.x) * m.gran;
    r1 = (int) (units * (blockIdx.x + 1) / gridDim.x) * m.gran;
}
__device__ __forceinline__ int rows_per_tile(const MatPhase& m) {
    int rt = kSlotBytes / (int) qw_row_bytes(m.n);
    rt -= rt % m.gran;
    return rt < m.gran ? m.gran : rt;
}
__device__ __forceinline__ MatPhase ph_qkv(const MegaParams& p, int l) { return {p.w_qkv + l * p.s_qkv, p.Pl + 2 * p.Kl, p.D, 1}; }
__device__ __forceinline__ MatPhase ph_o(const MegaParams& p, int l) { return {p.w_o + l * p.s_o, p.D, p.Pl, 1}; }
__device__ __forceinline__ MatPhase ph_13(const MegaParams& p, int l) { return {p.w_13 + l * p.s_13, 2 * p.Hdl, p.D, 2}; }
__device__ __forceinline__ MatPhase ph_2(const MegaParams& p, int l) { return {p.w_2 + l * p.s_2, p.D, p.Hdl, 1}; }
__device__ __forceinline__ MatPhase ph_cls(const MegaParams& p) { return {p.w_cls, p.Vl, p.D, 1}; }

__device__ __forceinline__ void attn_units(const MegaParams& p, int& nc, int& u0, int& u1) {
    nc = p.pos / kChunk + 1;
    const long long U = (long long) p.KVHl * nc;
    u0 = (int) (U * blockIdx.x / gridDim.x);
    u1 = (int) (U * (blockIdx.x + 1) / gridDim.x);
}

// ---------------------------------------------------------------- producer
__device__ void produce_mat(const Shared& sh, const MegaParams& p, const MatPhase& m, unsigned& it) {
    int r0, r1;
    cta_rows(m, r0, r1);
    const int rt = rows_per_tile(m);
    const size_t rb = qw_row_bytes(m.n);
    for (int r = r0; r < r1; r += rt, ++it) {
        const int nr = min(rt, r1 - r);
        const unsigned slot = it % p.nslot, par = (it / p.nslot) & 1;
        mbar_wait(sh, p, sh.empty + slot * 8, par ^ 1, 1);
        const uint32_t bytes = (uint32_t) (nr * rb);
        mbar_expect_tx(sh.full + slot * 8, bytes);
        bulk_g2s(smem_u32(sh.ring + (size_t) slot * kSlotBytes), m.base + (size_t) r * rb, bytes, sh.full + slot * 8);
    }
}

__device__ void produce_attn(const Shared& sh, const MegaParams& p, int l, unsigned& it) {
    int nc, u0, u1;
    attn_units(p, nc, u0, u1);
    for (int u = u0; u < u1; ++u, ++it) {
        const int kvh = u / nc, c = u % nc;
        const int p0 = c * kChunk;
        const int cnt = min(p.pos, p0 + kChunk) - p0; // slot `pos` itself is produced by this step
        const unsigned slot = it % p.nslot, par = (it / p.nslot) & 1;
        mbar_wait(sh, p, sh.empty + slot * 8, par ^ 1, 2);
        if (cnt > 0) {
            const size_t off = (((size_t) l * p.KVHl + kvh) * p.S + p0) * 128;
            const uint32_t bytes = (uint32_t) cnt * 512u;
            const uint32_t dst = smem_u32(sh.ring + (size_t) slot * kSlotBytes);
            mbar_expect_tx(sh.full + slot * 8, 2 * bytes);
            bulk_g2s(dst, p.k_cache + off, bytes, sh.full + slot * 8);
            bulk_g2s(dst + kChunk * 512, p.v_cache + off, bytes, sh.full + slot * 8);
        } else {
            mbar_arrive(sh.full + slot * 8);
        }
    }
}

__device__ void producer(const Shared& sh, const MegaParams& p) {
    unsigned it = 0;
    for (int l = 0; l < p.layers_run; ++l) {
        produce_mat(sh, p, ph_qkv(p, l), it);
        produce_attn(sh, p, l, it);
        produce_mat(sh, p, ph_o(p, l), it);
        produce_mat(sh, p, ph_13(p, l), it);
        produce_mat(sh, p, ph_2(p, l), it);
    }
    produce_mat(sh, p, ph_cls(p), it);
}

// ---------------------------------------------------------------- consumer: GEMV over ring tiles
// Epi is called by exactly one thread per unit of `gran` rows with the folded values.
template <int GRAN, class Epi>
__device__ void consume_mat(const Shared& sh, const MegaParams& p, const MatPhase& m, unsigned& it, Epi epi) {
    int r0, r1;
    cta_rows(m, r0, r1);
    const int rt = rows_per_tile(m);
    const int sgpr = qw_sg_per_row(m.n);
    const int groups = m.n / 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l16 = lane & 15, half = lane >> 4, grp = l16 >> 2;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int hw = warp * 2 + half;
    for (int r = r0; r < r1; r += rt, ++it) {
        const int nr = min(rt, r1 - r);
        const unsigned slot = it % p.nslot, par = (it / p.nslot) & 1;
        mbar_wait(sh, p, sh.full + slot * 8, par, 3);
        const uint8_t* tile = sh.ring + (size_t) slot * kSlotBytes;
        float* part = sh.part + (it & 1) * kPartFloats;
        const int nsg = nr * sgpr;
        for (int s = hw; s < nsg; s += 2 * kConsumerWarps) {
            const int c = s % sgpr;
            const uint8_t* rec = tile + (size_t) s * QW_SG_BYTES;
            const int4 wv = *reinterpret_cast<const int4*>(rec + 16 * l16);
            const int4 xv = *reinterpret_cast<const int4*>(sh.xq + c * 256 + 16 * l16);
            int dot = dot16(wv, xv);
            dot += __shfl_xor_sync(hmask, dot, 1);
            dot += __shfl_xor_sync(hmask, dot, 2);
            if ((l16 & 3) == 0)
                part[s * 4 + grp] = q8_term(dot, *reinterpret_cast<const float*>(rec + 256 + 4 * grp), sh.xs[c * 4 + grp]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sh.empty + slot * 8); // slot may be refilled
        bar_consumers();
        // fold: one thread per unit, group terms left to right from 0.0f (reference forward.c:86-99)
        for (int rr = 0; rr < nr; rr += GRAN) {
            const int g = (r - r0 + rr) / GRAN;
            if ((g & (kConsumerWarps - 1)) == warp && ((g >> 4) & 31) == lane) {
                float v[GRAN];
#pragma unroll
                for (int k = 0; k < GRAN; ++k) {
                    const float* t = part + (size_t) (rr + k) * sgpr * 4;
                    float acc = 0.0f;
                    for (int j = 0; j < groups; ++j) acc = __fadd_rn(acc, t[j]);
                    v[k] = acc;
                }
                epi(r + rr, v);
            }
        }
    }
}

// ---------------------------------------------------------------- consumer: prologues
// x (fp32, D) -> RMSNorm with weights w -> Q8_0 codes + scales in shared memory.
// layer0: the residual stream starts as the dequantised embedding row (forward.c:237).
__device__ void prologue_norm_quant(const Shared& sh, const MegaParams& p, const float* __restrict__ w, bool from_embedding) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    float* sx = sh.scr;
    if (from_embedding) {
        const int tok = p.token_dev ? *p.token_dev : p.token;
        const uint8_t* row = p.w_emb + (size_t) tok * qw_row_bytes(D);
        for (int i = tid; i < D; i += kConsumerThreads) {
            const uint8_t* rec = row + (size_t) (i >> 8) * QW_SG_BYTES;
            const int within = i & 255;
            const float v = __fmul_rn((float) reinterpret_cast<const int8_t*>(rec)[within],
                                      reinterpret_cast<const float*>(rec + 256)[within >> 6]);
            sx[i] = v;
            if ((i / kConsumerThreads) % gridDim.x == blockIdx.x) p.x[i] = v; // each element stored once chip-wide
        }
    } else {
        for (int i = tid; i < D; i += kConsumerThreads) sx[i] = __ldcg(p.x + i);
    }
    bar_consumers();
    float ss = 0.0f;
    for (int i = tid; i < D; i += kConsumerThreads) ss = __fadd_rn(ss, __fmul_rn(sx[i], sx[i]));
    ss = warp_sum(ss);
    if (lane == 0) sh.misc[warp] = ss;
    bar_consumers();
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) tot = __fadd_rn(tot, sh.misc[i]);
    const float r = rms_rscale(tot, D);
    const int groups = D / 64, pad_groups = qw_pad_cols(D) / 64;
    for (int g = warp; g < pad_groups; g += kConsumerWarps) {
        if (g < groups) {
            const int i0 = g * 64 + lane, i1 = i0 + 32;
            const float a = __fmul_rn(__ldg(w + i0), __fmul_rn(r, sx[i0]));
            const float b = __fmul_rn(__ldg(w + i1), __fmul_rn(r, sx[i1]));
            const float scale = q8_scale(warp_max(fmaxf(fabsf(a), fabsf(b))));
            sh.xq[i0] = (int8_t) q8_code(a, scale);
            sh.xq[i1] = (int8_t) q8_code(b, scale);
            if (lane == 0) sh.xs[g] = scale;
        } else {
            sh.xq[g * 64 + lane] = 0;
            sh.xq[g * 64 + 32 + lane] = 0;
            if (lane == 0) sh.xs[g] = 0.0f;
        }
    }
    bar_consumers();
}

// fp32 vector in global memory (written by other CTAs) -> Q8_0 codes + scales in shared memory
__device__ void prologue_quant_global(const Shared& sh, const float* src, int n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = n / 64, pad_groups = qw_pad_cols(n) / 64;
    for (int g = warp; g < pad_groups; g += kConsumerWarps) {
        if (g < groups) {
            const float a = __ldcg(src + g * 64 + lane), b = __ldcg(src + g * 64 + 32 + lane);
            const float scale = q8_scale(warp_max(fmaxf(fabsf(a), fabsf(b))));
            sh.xq[g * 64 + lane] = (int8_t) q8_code(a, scale);
            sh.xq[g * 64 + 32 + lane] = (int8_t) q8_code(b, scale);
            if (lane == 0) sh.xs[g] = scale;
        } else {
            sh.xq[g * 64 + lane] = 0;
            sh.xq[g * 64 + 32 + lane] = 0;
            if (lane == 0) sh.xs[g] = 0.0f;
        }
    }
    bar_consumers();
}

// already-quantised vector (attention output) from global -> shared
__device__ void prologue_load_codes(const Shared& sh, const int8_t* q, const float* s, int n) {
    const int tid = threadIdx.x;
    const int pad = qw_pad_cols(n);
    for (int i = tid; i < pad / 16; i += kConsumerThreads) {
        int4 v = make_int4(0, 0, 0, 0);
        if (i * 16 < n) v = __ldcg(reinterpret_cast<const int4*>(q) + i);
        reinterpret_cast<int4*>(sh.xq)[i] = v;
    }
    for (int i = tid; i < pad / 64; i += kConsumerThreads) sh.xs[i] = i < n / 64 ? __ldcg(s + i) : 0.0f;
    bar_consumers();
}

// ---------------------------------------------------------------- consumer: attention
// scratch layout inside sh.scr (floats): raw[1024] | sq[1024] | ssc[8*28] | so[128]
__device__ void head_norm_rope(float* dst, const float* raw, const float* g, const MegaParams& p, float r, int i) {
    // called by 128 threads of one head after raw[] and r are ready; writes dst[i]
    const int j = i & 63;
    const float c = __ldg(p.rope_cos + (size_t) p.pos * 64 + j), s = __ldg(p.rope_sin + (size_t) p.pos * 64 + j);
    const float a = __fmul_rn(__ldg(g + j), __fmul_rn(r, raw[j]));
    const float b = __fmul_rn(__ldg(g + j + 64), __fmul_rn(r, raw[j + 64]));
    dst[i] = (i < 64) ? __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, s)) : __fadd_rn(__fmul_rn(a, s), __fmul_rn(b, c));
}
// sequential sum of squares over 128 values: the reference's order (forward.c:16-19), so the
// per-head q/k norms are bit-identical to the reference for identical inputs
__device__ float head_rscale(const float* raw) {
    float ss = 0.0f;
    for (int i = 0; i < 128; ++i) ss = __fadd_rn(ss, __fmul_rn(raw[i], raw[i]));
    return rms_rscale(ss, 128);
}

__device__ void consume_attn(const Shared& sh, const MegaParams& p, int l, unsigned& it) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kv_mul = p.kv_mul;
    float* raw = sh.scr;
    float* sq = sh.scr + 1024;
    float* ssc = sh.scr + 2048;
    float* s_m = sh.misc + 32;
    float* s_l = sh.misc + 40;
    float* s_scale = sh.misc + 48;
    float* s_r = sh.misc + 56;
    int nc, u0, u1;
    attn_units(p, nc, u0, u1);
    const float* gq = p.q_norm + (size_t) l * 128;
    const float* gk = p.k_norm + (size_t) l * 128;
    const float inv = sqrtf(128.0f);
    float acc[2] = {0.0f, 0.0f};
    int cur = -1;
    auto flush = [&](int kvh) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int idx = tid + k * kConsumerThreads;
            if (idx < kv_mul * 128) {
                const int j = idx >> 7, d = idx & 127, h = kvh * kv_mul + j;
                const size_t slot = (size_t) h * gridDim.x + blockIdx.x;
                p.part_acc[slot * 128 + d] = acc[k];
                if (d == 0) {
                    p.part_m[slot] = s_m[j];
                    p.part_l[slot] = s_l[j];
                }
            }
        }
    };
    for (int u = u0; u < u1; ++u, ++it) {
        const int kvh = u / nc, c = u % nc;
        const int p0 = c * kChunk;
        const bool last = (c == nc - 1);
        const int cnt = min(p.pos + 1, p0 + kChunk) - p0;
        if (kvh != cur) {
            if (cur >= 0) flush(cur);
            bar_consumers();
            for (int i = tid; i < kv_mul * 128; i += kConsumerThreads) raw[i] = __ldcg(p.qkv + (size_t) kvh * kv_mul * 128 + i);
            bar_consumers();
            if (tid < kv_mul) s_r[tid] = head_rscale(raw + tid * 128);
            bar_consumers();
            for (int i = tid; i < kv_mul * 128; i += kConsumerThreads)
                head_norm_rope(sq + (i & ~127), raw + (i & ~127), gq, p, s_r[i >> 7], i & 127);
            if (tid < kv_mul) {
                s_m[tid] = -INFINITY;
                s_l[tid] = 0.0f;
            }
            acc[0] = acc[1] = 0.0f;
            cur = kvh;
        }
        const unsigned slot = it % p.nslot, par = (it / p.nslot) & 1;
        mbar_wait(sh, p, sh.full + slot * 8, par, 4);
        float* Kt = reinterpret_cast<float*>(sh.ring + (size_t) slot * kSlotBytes);
        float* Vt = Kt + kChunk * 128;
        bar_consumers(); // sq ready; previous unit's PV reads of ssc are done
        if (last) {
            // this step's own K/V row: norm + rope K, raw V; into the tile and into the cache
            const int il = p.pos - p0;
            const size_t coff = (((size_t) l * p.KVHl + kvh) * p.S + p.pos) * 128;
            if (tid < 128) raw[tid] = __ldcg(p.qkv + p.Pl + (size_t) kvh * 128 + tid);
            else if (tid < 256) {
                const float v = __ldcg(p.qkv + p.Pl + p.Kl + (size_t) kvh * 128 + (tid - 128));
                Vt[il * 128 + (tid - 128)] = v;
                p.v_cache[coff + (tid - 128)] = v;
            }
            bar_consumers();
            if (tid == 0) s_r[8] = head_rscale(raw);
            bar_consumers();
            if (tid < 128) {
                head_norm_rope(Kt + il * 128, raw, gk, p, s_r[8], tid);
                p.k_cache[coff + tid] = Kt[il * 128 + tid];
            }
            bar_consumers();
        }
        // scores (forward.c:156-165)
        for (int i = warp; i < cnt; i += kConsumerWarps) {
            const float4 kv = *reinterpret_cast<const float4*>(Kt + i * 128 + lane * 4);
            for (int j = 0; j < kv_mul; ++j) {
                const float4 qv = *reinterpret_cast<const float4*>(sq + j * 128 + lane * 4);
                float d = __fmul_rn(qv.x, kv.x);
                d = __fmaf_rn(qv.y, kv.y, d);
                d = __fmaf_rn(qv.z, kv.z, d);
                d = __fmaf_rn(qv.w, kv.w, d);
                d = warp_sum(d);
                if (lane == 0) ssc[j * kChunk + i] = __fdiv_rn(d, inv);
            }
        }
        bar_consumers();
        // online softmax update, one warp per query head
        if (warp < kv_mul) {
            const int j = warp;
            const float s = lane < cnt ? ssc[j * kChunk + lane] : -INFINITY;
            const float m_old = s_m[j];
            const float m_new = fmaxf(m_old, warp_max(s));
            const float e = lane < cnt ? expf(__fsub_rn(s, m_new)) : 0.0f;
            if (lane < cnt) ssc[j * kChunk + lane] = e;
            const float lsum = warp_sum(e);
            if (lane == 0) {
                const float sc = expf(__fsub_rn(m_old, m_new));
                s_scale[j] = sc;
                s_l[j] = __fmaf_rn(s_l[j], sc, lsum);
                s_m[j] = m_new;
            }
        }
        bar_consumers();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int idx = tid + k * kConsumerThreads;
            if (idx < kv_mul * 128) {
                const int j = idx >> 7, d = idx & 127;
                float a = __fmul_rn(acc[k], s_scale[j]);
                for (int i = 0; i < cnt; ++i) a = __fmaf_rn(ssc[j * kChunk + i], Vt[i * 128 + d], a);
                acc[k] = a;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sh.empty + slot * 8);
    }
    if (cur >= 0) flush(cur);
}

// merge the split-KV partials of each head, write fp32 att (debug) and its Q8_0 codes
__device__ void combine_attn(const Shared& sh, const MegaParams& p) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* so = sh.scr + 2048 + 8 * kChunk;
    const int nc = p.pos / kChunk + 1;
    const long long U = (long long) p.KVHl * nc;
    const int G = gridDim.x;
    for (int h = blockIdx.x; h < p.Hl; h += G) {
        const int kvh = h / p.kv_mul;
        const long long ulo = (long long) kvh * nc, uhi = ulo + nc;
        int blo = (int) (ulo * G / U) - 1, bhi = (int) (uhi * G / U) + 1;
        blo = max(blo, 0);
        bhi = min(bhi, G - 1);
        if (tid < 128) {
            float M = -INFINITY;
            for (int b = blo; b <= bhi; ++b) {
                const long long s0 = U * b / G, s1 = U * (b + 1) / G;
                if (max(s0, ulo) < min(s1, uhi)) M = fmaxf(M, __ldcg(p.part_m + (size_t) h * G + b));
            }
            float Lsum = 0.0f, A = 0.0f;
            for (int b = blo; b <= bhi; ++b) {
                const long long s0 = U * b / G, s1 = U * (b + 1) / G;
                if (max(s0, ulo) < min(s1, uhi)) {
                    const size_t slot = (size_t) h * G + b;
                    const float w = expf(__fsub_rn(__ldcg(p.part_m + slot), M));
                    Lsum = __fmaf_rn(__ldcg(p.part_l + slot), w, Lsum);
                    A = __fmaf_rn(__ldcg(p.part_acc + slot * 128 + tid), w, A);
                }
            }
            const float o = __fdiv_rn(A, Lsum);
            so[tid] = o;
            p.att[(size_t) h * 128 + tid] = o;
        }
        bar_consumers();
        if (warp < 2) {
            const float a = so[warp * 64 + lane], b = so[warp * 64 + 32 + lane];
            const float scale = q8_scale(warp_max(fmaxf(fabsf(a), fabsf(b))));
            p.att_q[(size_t) h * 128 + warp * 64 + lane] = (int8_t) q8_code(a, scale);
            p.att_q[(size_t) h * 128 + warp * 64 + 32 + lane] = (int8_t) q8_code(b, scale);
            if (lane == 0) p.att_s[h * 2 + warp] = scale;
        }
        bar_consumers();
    }
}

// ---------------------------------------------------------------- consumer main
__device__ __forceinline__ void stamp(const MegaParams& p, int l, int k) {
    if (p.prof && threadIdx.x == 0)
        p.prof[((size_t) blockIdx.x * (p.L + 1) + l) * kProfSlots + k] = gtime_ns();
}

__device__ void consumer(const Shared& sh, const MegaParams& p) {
    unsigned it = 0;
    int nbar = 0;
    for (int l = 0; l < p.layers_run; ++l) {
        stamp(p, l, 0);
        // --- attention block (forward.c:254-298)
        prologue_norm_quant(sh, p, p.att_norm + (size_t) l * p.D, l == 0);
        stamp(p, l, 1);
        consume_mat<1>(sh, p, ph_qkv(p, l), it, [&](int row, const float* v) { p.qkv[row] = v[0]; });
        stamp(p, l, 2);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 3);
        consume_attn(sh, p, l, it);
        stamp(p, l, 4);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 5);
        combine_attn(sh, p);
        stamp(p, l, 6);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 7);
        prologue_load_codes(sh, p.att_q, p.att_s, p.Pl);
        consume_mat<1>(sh, p, ph_o(p, l), it,
                       [&](int row, const float* v) { p.x[row] = __fadd_rn(__ldcg(p.x + row), v[0]); });
        stamp(p, l, 8);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 9);
        // --- feed-forward block (forward.c:303-338)
        prologue_norm_quant(sh, p, p.ffn_norm + (size_t) l * p.D, false);
        stamp(p, l, 10);
        consume_mat<2>(sh, p, ph_13(p, l), it,
                       [&](int row, const float* v) { p.h[row >> 1] = __fmul_rn(silu_ref(v[0]), v[1]); });
        stamp(p, l, 11);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 12);
        prologue_quant_global(sh, p.h, p.Hdl);
        stamp(p, l, 13);
        consume_mat<1>(sh, p, ph_2(p, l), it,
                       [&](int row, const float* v) { p.x[row] = __fadd_rn(__ldcg(p.x + row), v[0]); });
        stamp(p, l, 14);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 15);
    }
    stamp(p, p.L, 0);
    // --- final norm + classifier (forward.c:344-348)
    prologue_norm_quant(sh, p, p.out_norm, p.layers_run == 0);
    stamp(p, p.L, 1);
    consume_mat<1>(sh, p, ph_cls(p), it, [&](int row, const float* v) { p.logits[row] = v[0]; });
    stamp(p, p.L, 2);
}

__global__ void __launch_bounds__(kThreads, 1) k_decode(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int abort_flag;
    Shared sh;
    sh.ring = smem;
    sh.xq = reinterpret_cast<int8_t*>(smem + p.off_xq);
    sh.xs = reinterpret_cast<float*>(smem + p.off_xs);
    sh.scr = reinterpret_cast<float*>(smem + p.off_scr);
    sh.part = reinterpret_cast<float*>(smem + p.off_part);
    sh.misc = reinterpret_cast<float*>(smem + p.off_misc);
    sh.full = smem_u32(smem + p.off_bar);
    sh.empty = sh.full + kMaxSlots * 8;
    sh.abort_flag = &abort_flag;
    if (threadIdx.x == 0) {
        abort_flag = 0;
        for (int s = 0; s < p.nslot; ++s) {
            mbar_init(sh.full + s * 8, 1);
            mbar_init(sh.empty + s * 8, kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= kConsumerThreads) {
        if (threadIdx.x == kConsumerThreads) producer(sh, p);
        return;
    }
    consumer(sh, p);
}

} // namespace

// ---------------------------------------------------------------- host side
static MegaState* state_of(QwenCudaCtx* c) { return reinterpret_cast<MegaState*>(c->mega); }

int qw_mega_init(QwenCudaCtx* c) {
    MegaState* st = new MegaState();
    c->mega = st;
    if (c->tp_size != 1) { // tensor-parallel contexts use the per-op path for now
        c->path = 1;
        return 0;
    }
    const int amax = qw_pad_cols(std::max(c->D, std::max(c->Pl, c->Hdl)));
    if (qw_row_bytes(amax) > (size_t) kSlotBytes || 2 * qw_row_bytes(c->D) > (size_t) kSlotBytes) {
        qw_set_error("persistent decode kernel: a weight row (%d columns) does not fit one %d-byte ring slot", amax, kSlotBytes);
        return -1;
    }
    int dev_smem = 0, coop = 0;
    QW_CUDA(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    QW_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device));
    if (!coop) {
        qw_set_error("device does not support cooperative launch");
        return -1;
    }
    int off = 0;
    auto take = [&](int bytes) {
        const int o = off;
        off += (bytes + 127) & ~127;
        return o;
    };
    // everything except the ring first, then give the ring all remaining slots
    const int xq_b = amax, xs_b = amax / 64 * 4;
    const int scr_b = std::max(c->D * 4, (2048 + 8 * kChunk + 128) * 4);
    const int part_b = 2 * kPartFloats * 4, misc_b = 512, bar_b = 2 * kMaxSlots * 8;
    const int fixed = ((xq_b + 127) & ~127) + ((xs_b + 127) & ~127) + ((scr_b + 127) & ~127) + ((part_b + 127) & ~127)
                      + ((misc_b + 127) & ~127) + ((bar_b + 127) & ~127);
    const int avail = dev_smem - fixed - 1024; // 1 KB left for static shared + driver reserve
    st->nslot = std::min(kMaxSlots, avail / kSlotBytes);
    if (st->nslot < 2) {
        qw_set_error("persistent decode kernel: not enough shared memory for a 2-slot ring (%d bytes free)", avail);
        return -1;
    }
    take(st->nslot * kSlotBytes);
    st->off_xq = take(xq_b);
    st->off_xs = take(xs_b);
    st->off_scr = take(scr_b);
    st->off_part = take(part_b);
    st->off_misc = take(misc_b);
    st->off_bar = take(bar_b);
    st->smem = off;
    QW_CUDA(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) st->smem));
    int per_sm = 0;
    QW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_decode, kThreads, st->smem));
    if (per_sm < 1) {
        qw_set_error("persistent decode kernel does not fit on an SM (smem %zu)", st->smem);
        return -1;
    }
    st->grid = c->num_sms;
    QW_CUDA(cudaMalloc((void**) &st->att_q, qw_pad_cols(c->Pl)));
    QW_CUDA(cudaMalloc((void**) &st->att_s, (size_t) qw_pad_cols(c->Pl) / 64 * 4));
    QW_CUDA(cudaMalloc((void**) &st->part_m, (size_t) c->Hl * st->grid * 4));
    QW_CUDA(cudaMalloc((void**) &st->part_l, (size_t) c->Hl * st->grid * 4));
    QW_CUDA(cudaMalloc((void**) &st->part_acc, (size_t) c->Hl * st->grid * 128 * 4));
    QW_CUDA(cudaMemset(c->bar_counter, 0, 8));
    c->bar_epoch = 0;
    c->path = 0;
    return 0;
}

void qw_mega_free(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st) return;
    void* bufs[] = {st->att_q, st->att_s, st->part_m, st->part_l, st->part_acc, st->prof};
    for (void* b : bufs)
        if (b) cudaFree(b);
    delete st;
    c->mega = nullptr;
}

int qw_decode_mega(QwenCudaCtx* c, int token, const int* token_dev, int pos) {
    MegaState* st = state_of(c);
    if (!st || !st->grid) {
        qw_set_error("persistent decode kernel is not initialised for this context");
        return -4;
    }
    MegaParams p;
    p.D = c->D; p.Hdl = c->Hdl; p.L = c->L; p.Hl = c->Hl; p.KVHl = c->KVHl; p.Pl = c->Pl; p.Kl = c->Kl; p.Vl = c->Vl;
    p.S = c->S; p.kv_mul = c->Hl / c->KVHl;
    p.pos = pos; p.token = token; p.token_dev = token_dev;
    p.layers_run = (c->layers_run >= 0 && c->layers_run <= c->L) ? c->layers_run : c->L;
    p.w_qkv = c->w_qkv; p.w_o = c->w_o; p.w_13 = c->w_13; p.w_2 = c->w_2; p.w_cls = c->w_cls; p.w_emb = c->w_emb;
    p.s_qkv = c->w_qkv_stride; p.s_o = c->w_o_stride; p.s_13 = c->w_13_stride; p.s_2 = c->w_2_stride;
    p.att_norm = c->att_norm; p.ffn_norm = c->ffn_norm; p.out_norm = c->out_norm; p.q_norm = c->q_norm; p.k_norm = c->k_norm;
    p.rope_cos = c->rope_cos; p.rope_sin = c->rope_sin;
    p.k_cache = c->k_cache; p.v_cache = c->v_cache;
    p.x = c->x; p.qkv = c->qkv; p.att = c->att; p.h = c->h; p.logits = c->logits;
    p.att_q = st->att_q; p.att_s = st->att_s;
    p.part_m = st->part_m; p.part_l = st->part_l; p.part_acc = st->part_acc;
    p.bar = c->bar_counter; p.bar_base = c->bar_epoch;
    p.err = c->err_flag;
    p.prof = st->prof;
    p.nslot = st->nslot; p.off_xq = st->off_xq; p.off_xs = st->off_xs; p.off_scr = st->off_scr; p.off_part = st->off_part;
    p.off_misc = st->off_misc; p.off_bar = st->off_bar;
    const int nbar = 6 * p.layers_run;
    c->bar_epoch += (unsigned long long) nbar * st->grid;
    void* args[] = {&p};
    QW_CUDA(cudaLaunchCooperativeKernel((const void*) k_decode, dim3(st->grid), dim3(kThreads), args, st->smem, c->stream));
    return 0;
}

int qw_decode_mega_launches(const QwenCudaCtx*) { return 1; }

// debug: per-CTA phase timestamps of the NEXT steps; read back with qw_mega_profile_read
int qw_mega_profile_enable(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st || !st->grid) return -1;
    const size_t n = (size_t) st->grid * (c->L + 1) * kProfSlots;
    if (!st->prof) QW_CUDA(cudaMalloc((void**) &st->prof, n * 8));
    QW_CUDA(cudaMemset(st->prof, 0, n * 8));
    QW_CUDA(cudaDeviceSynchronize());
    return (int) n;
}
int qw_mega_profile_read(QwenCudaCtx* c, unsigned long long* host, size_t max_elems) {
    MegaState* st = state_of(c);
    if (!st || !st->prof) return -1;
    size_t n = (size_t) st->grid * (c->L + 1) * kProfSlots;
    if (n > max_elems) n = max_elems;
    QW_CUDA(cudaStreamSynchronize(c->stream));
    QW_CUDA(cudaMemcpy(host, st->prof, n * 8, cudaMemcpyDeviceToHost));
    return st->grid;
}
