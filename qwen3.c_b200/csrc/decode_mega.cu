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

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

#ifndef QW_GEMV_PRED
#define QW_GEMV_PRED 0
#endif

namespace {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kSlotBytes = 28672;           // one ring slot: 105 SG records or 2 x 28 KV rows
constexpr int kChunk = 28;                  // KV positions per attention unit (2*28*512 B = one slot)
constexpr int kMaxSlots = 8;
constexpr int kMaxGrid = 256;
constexpr unsigned long long kTimeoutNs = 4000000000ull;
constexpr int kTileLog = 8192;
constexpr int kProfSlots = 16; // per layer: stamps after each phase step (CTA-local, thread 0)

struct MegaParams {
    int D, Hdl, L, Hl, KVHl, Pl, Kl, Vl, S, kv_mul;
    int pos, token, layers_run;
    int perm; // CTA -> row-block permutation multiplier (coprime to the grid)
    int copy_split; // debug: issue each tile as this many bulk copies
    int dbg_mode; // 0 normal; 1 consumers skip the GEMV math (ring throughput test); 2 skip attention math too
    const int* token_dev;
    const uint8_t *w_qkv, *w_o, *w_13, *w_2, *w_cls, *w_emb;
    size_t s_qkv, s_o, s_13, s_2;
    const float *att_norm, *ffn_norm, *out_norm, *q_norm, *k_norm, *rope_cos, *rope_sin;
    float *k_cache, *v_cache;
    float *x, *qkv, *att, *h, *logits;
    uint8_t* att_q; // attention output, quantised, SG layout (row of Pl columns)
    float *part_m, *part_l, *part_acc;
    unsigned long long* bar;
    unsigned long long bar_base;
    int* err;
    unsigned long long* tlog; // optional [4][kTileLog] per-tile stamps of CTA 0 (debug)
    int tlog_warp;
    unsigned long long* prof; // optional [CTA][kProfSlots] globaltimer stamps (debug)
    int nslot, off_xq, off_scr, off_misc, off_bar;
};

struct MegaState {
    uint8_t* att_q = nullptr;
    float *part_m = nullptr, *part_l = nullptr, *part_acc = nullptr;
    int grid = 0, nslot = 0, dbg_mode = 0, copy_split = 1, perm = 1;
    unsigned long long* prof = nullptr;
    unsigned long long* tlog = nullptr;
    int tlog_warp = 0;
    size_t smem = 0;
    int off_xq, off_scr, off_misc, off_bar;
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
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct Shared {
    uint8_t* ring;
    uint8_t* xq;   // activation vector in SG layout: [256 codes][4 scales] records
    float* scr;
    float* misc;
    uint32_t full, empty; // shared-space addresses of the barrier arrays
    volatile int* abort_flag;
};

__device__ __forceinline__ void mbar_wait(const Shared& sh, const MegaParams& p, uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (unsigned spin = 1;; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spin & 255u) == 0) { // look at the clock only now and then: the common case is a short wait
            if (*sh.abort_flag) return;
            const unsigned long long now = gtime_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTimeoutNs) {
                *sh.abort_flag = code;
                *p.err = code;
                return;
            }
        }
    }
}

// Grid-wide barrier among the consumer threads of every CTA: bar.sync orders the CTA's writes
// before thread 0's release-add on one global counter; thread 0 polls it with ld.acquire and a
// second bar.sync hands the other CTAs' writes to every thread here (readers use ld.cg /
// ld.acquire afterwards, never L1). ~1.5 us on 148 CTAs, all of it L2 round trips. A variant
// with one flag word per CTA and 148 polling threads per CTA was measured and is 3x SLOWER
// (148 x 5 polled lines hot-spot the L2 slices), so the single counter stays.
__device__ __forceinline__ void grid_barrier(const Shared& sh, const MegaParams& p, int& nbar) {
    bar_consumers();
    if (threadIdx.x == 0) {
        const unsigned long long target = p.bar_base + (unsigned long long) (nbar + 1) * gridDim.x;
        red_release_add_u64(p.bar, 1ull);
        unsigned long long t0 = 0;
        for (unsigned spin = 1; ld_acquire_u64(p.bar) < target; ++spin) {
            if ((spin & 1023u) == 0) {
                const unsigned long long now = gtime_ns();
                if (t0 == 0) t0 = now;
                if (*sh.abort_flag || now - t0 > kTimeoutNs) {
                    *sh.abort_flag = 100 + nbar;
                    *p.err = 100 + nbar;
                    break;
                }
            }
        }
    }
    bar_consumers();
    ++nbar;
}

// ---------------------------------------------------------------- schedule (shared by producer and consumers)
struct MatPhase {
    const uint8_t* base;
    int rows, n, gran;
    int perm = 1;
};
__device__ __forceinline__ void cta_rows(const MatPhase& m, int& r0, int& r1) {
    const long long units = m.rows / m.gran;
    // CTA -> row-block map: a multiplicative permutation (perm coprime to the grid). With the identity
    // map the GEMV phases ran 8-10 % slower and a fixed third of the CTAs arrived 2-4 us late at every
    // barrier (measured, profiles/r1_k_decode_summary.md); any scattering permutation removes that.
    const int b = (int) ((blockIdx.x * (unsigned) m.perm) % gridDim.x);
    r0 = (int) (units * b / gridDim.x) * m.gran;
    r1 = (int) (units * (b + 1) / gridDim.x) * m.gran;
}
__device__ __forceinline__ int rows_per_tile(const MatPhase& m) {
    int rt = kSlotBytes / (int) qw_row_bytes(m.n);
    if (rt >= 2) rt &= ~1; // whole 2-row units per tile
    return rt < m.gran ? m.gran : rt;
}
__device__ __forceinline__ MatPhase ph_qkv(const MegaParams& p, int l) { return {p.w_qkv + l * p.s_qkv, p.Pl + 2 * p.Kl, p.D, 1, p.perm}; }
__device__ __forceinline__ MatPhase ph_o(const MegaParams& p, int l) { return {p.w_o + l * p.s_o, p.D, p.Pl, 1, p.perm}; }
__device__ __forceinline__ MatPhase ph_13(const MegaParams& p, int l) { return {p.w_13 + l * p.s_13, 2 * p.Hdl, p.D, 2, p.perm}; }
__device__ __forceinline__ MatPhase ph_2(const MegaParams& p, int l) { return {p.w_2 + l * p.s_2, p.D, p.Hdl, 1, p.perm}; }
__device__ __forceinline__ MatPhase ph_cls(const MegaParams& p) { return {p.w_cls, p.Vl, p.D, 1, p.perm}; }

// index of the unit block this CTA attends over (also the slot of its partial results)
__device__ __forceinline__ int attn_block(const MegaParams& p) {
    return (int) ((blockIdx.x * (unsigned) p.perm) % gridDim.x);
}
__device__ __forceinline__ void attn_units(const MegaParams& p, int& nc, int& u0, int& u1) {
    nc = p.pos / kChunk + 1;
    const long long U = (long long) p.KVHl * nc;
    const int b = attn_block(p);
    u0 = (int) (U * b / gridDim.x);
    u1 = (int) (U * (b + 1) / gridDim.x);
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
        if (p.tlog && blockIdx.x == 0 && it < kTileLog) p.tlog[it] = gtime_ns();
        mbar_expect_tx(sh.full + slot * 8, bytes);
        const uint32_t dst = smem_u32(sh.ring + (size_t) slot * kSlotBytes);
        const uint8_t* src = m.base + (size_t) r * rb;
        const uint32_t piece = ((bytes / p.copy_split) + 15u) & ~15u;
        for (uint32_t o = 0; o < bytes; o += piece) bulk_g2s(dst + o, src + o, min(piece, bytes - o), sh.full + slot * 8);
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
        if (p.tlog && blockIdx.x == 0 && it < kTileLog) p.tlog[it] = gtime_ns();
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
// Work unit = 2 consecutive rows, owned by ONE warp (unit u of the CTA's range -> warp u % 16).
// Lane L owns Q8_0 groups L, L+32, ... of both rows: a group is 64 codes = 4 x LDS.128 of W per
// row and 4 x LDS.128 of x, shared by the two rows (x sits in shared memory in the same record
// layout, so one offset serves both). dp4a has ~24 cycles of dependent latency on this part
// (measured, scripts/ubench), so each lane keeps FOUR independent dp4a chains in flight: two
// rows x two groups. The 16-byte pieces are read in a lane-rotated order so every quarter-warp
// hits 8 distinct 16-byte bank groups despite the 64-byte lane stride (0 excess wavefronts in
// ncu). Per group the exact int32 dot is scaled as ((float) dot * ws) * xs (forward.c:94-96) and
// added in fp32; lanes are combined by a shuffle tree.
// No CTA-wide barrier per tile: warps meet only at the ring's mbarriers, so with one unit per
// tile (n = 9728) different warps work on different ring slots at the same time.
// KIND 0: out[row] = v      KIND 1: x[row] += v (residual)      KIND 2: h[row/2] = silu(v0) * v1
template <int KIND>
__device__ void consume_mat(const Shared& sh, const MegaParams& p, const MatPhase& m, unsigned& it, float* out) {
    int r0, r1;
    cta_rows(m, r0, r1);
    const int rt = rows_per_tile(m);
    const int sgpr = qw_sg_per_row(m.n);
    const size_t rb = (size_t) sgpr * QW_SG_BYTES;
    const int groups = sgpr * 4; // padded groups carry zero codes and zero scales: they add +0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rot = (lane & 2);
    const int nrows = r1 - r0;
    const int upt = (rt + 1) / 2;             // units per tile (rt is 1 only when a row fills the slot)
    const int rows_pu = rt >= 2 ? 2 : 1;      // rows per unit
    const int total = (nrows + rows_pu - 1) / rows_pu;
    for (int t0 = 0; t0 < total; t0 += upt, ++it) {
        const unsigned slot = it % p.nslot, par = (it / p.nslot) & 1;
#ifdef QW_TLOG
        const bool logme = p.tlog && blockIdx.x == 0 && warp == p.tlog_warp && lane == 0 && it < kTileLog;
#else
        constexpr bool logme = false;
#endif
        if (logme) p.tlog[kTileLog + it] = gtime_ns();
        mbar_wait(sh, p, sh.full + slot * 8, par, 3);
        if (logme) p.tlog[2 * kTileLog + it] = gtime_ns();
        const uint8_t* tile = sh.ring + (size_t) slot * kSlotBytes;
        const int t1 = p.dbg_mode >= 1 ? t0 : min(t0 + upt, total);
        int ufirst = (warp - t0) % kConsumerWarps; // first unit >= t0 owned by this warp (u % 15 == warp)
        if (ufirst < 0) ufirst += kConsumerWarps;
        for (int u = t0 + ufirst; u < t1; u += kConsumerWarps) {
            const int lr = (u - t0) * rows_pu;          // first row of the unit inside the tile
            const int grow = r0 + u * rows_pu;          // its global row
            const bool two = rows_pu == 2 && grow + 1 < r1;
            const uint8_t* rowa = tile + (size_t) lr * rb;
            const uint8_t* rowb = two ? rowa + rb : rowa;
            float xres = 0.0f;
            if (KIND == 1 && lane < 2 && (lane == 0 || two)) xres = __ldcg(out + grow + lane); // hide the L2 round trip
            float acca = 0.0f, accb = 0.0f;
            for (int G = lane; G < groups; G += 64) {
                const int G2 = G + 32;
                const bool has2 = G2 < groups;
                const int off = (G >> 2) * QW_SG_BYTES + (G & 3) * 64;
#if QW_GEMV_PRED
                const int off2 = (G2 >> 2) * QW_SG_BYTES + (G2 & 3) * 64;
                int da0 = 0, da1 = 0, db0 = 0, db1 = 0;
                if (has2) {
#pragma unroll 2
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 x0 = *reinterpret_cast<const int4*>(sh.xq + off + pc);
                        const int4 x1 = *reinterpret_cast<const int4*>(sh.xq + off2 + pc);
                        const int4 a0 = *reinterpret_cast<const int4*>(rowa + off + pc);
                        const int4 a1 = *reinterpret_cast<const int4*>(rowa + off2 + pc);
                        const int4 b0 = *reinterpret_cast<const int4*>(rowb + off + pc);
                        const int4 b1 = *reinterpret_cast<const int4*>(rowb + off2 + pc);
                        da0 = __dp4a(a0.x, x0.x, da0); da1 = __dp4a(a1.x, x1.x, da1); db0 = __dp4a(b0.x, x0.x, db0); db1 = __dp4a(b1.x, x1.x, db1);
                        da0 = __dp4a(a0.y, x0.y, da0); da1 = __dp4a(a1.y, x1.y, da1); db0 = __dp4a(b0.y, x0.y, db0); db1 = __dp4a(b1.y, x1.y, db1);
                        da0 = __dp4a(a0.z, x0.z, da0); da1 = __dp4a(a1.z, x1.z, da1); db0 = __dp4a(b0.z, x0.z, db0); db1 = __dp4a(b1.z, x1.z, db1);
                        da0 = __dp4a(a0.w, x0.w, da0); da1 = __dp4a(a1.w, x1.w, da1); db0 = __dp4a(b0.w, x0.w, db0); db1 = __dp4a(b1.w, x1.w, db1);
                    }
                } else { // lanes without a second group issue no loads for it (a partial warp LDS costs fewer wavefronts)
                    int dc0 = 0, dc1 = 0; // split each row's chain in two to keep four chains in flight
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = ((i + rot) & 3) * 16;
                        const int4 x0 = *reinterpret_cast<const int4*>(sh.xq + off + pc);
                        const int4 a0 = *reinterpret_cast<const int4*>(rowa + off + pc);
                        const int4 b0 = *reinterpret_cast<const int4*>(rowb + off + pc);
                        da0 = __dp4a(a0.x, x0.x, da0); dc0 = __dp4a(a0.y, x0.y, dc0); db0 = __dp4a(b0.x, x0.x, db0); dc1 = __dp4a(b0.y, x0.y, dc1);
                        da0 = __dp4a(a0.z, x0.z, da0); dc0 = __dp4a(a0.w, x0.w, dc0); db0 = __dp4a(b0.z, x0.z, db0); dc1 = __dp4a(b0.w, x0.w, dc1);
                    }
                    da0 += dc0;
                    db0 += dc1;
                }
#else
                const int off2 = has2 ? (G2 >> 2) * QW_SG_BYTES + (G2 & 3) * 64 : off;
                int da0 = 0, da1 = 0, db0 = 0, db1 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int pc = ((i + rot) & 3) * 16;
                    const int4 x0 = *reinterpret_cast<const int4*>(sh.xq + off + pc);
                    const int4 x1 = *reinterpret_cast<const int4*>(sh.xq + off2 + pc);
                    const int4 a0 = *reinterpret_cast<const int4*>(rowa + off + pc);
                    const int4 a1 = *reinterpret_cast<const int4*>(rowa + off2 + pc);
                    const int4 b0 = *reinterpret_cast<const int4*>(rowb + off + pc);
                    const int4 b1 = *reinterpret_cast<const int4*>(rowb + off2 + pc);
                    da0 = __dp4a(a0.x, x0.x, da0); da1 = __dp4a(a1.x, x1.x, da1); db0 = __dp4a(b0.x, x0.x, db0); db1 = __dp4a(b1.x, x1.x, db1);
                    da0 = __dp4a(a0.y, x0.y, da0); da1 = __dp4a(a1.y, x1.y, da1); db0 = __dp4a(b0.y, x0.y, db0); db1 = __dp4a(b1.y, x1.y, db1);
                    da0 = __dp4a(a0.z, x0.z, da0); da1 = __dp4a(a1.z, x1.z, da1); db0 = __dp4a(b0.z, x0.z, db0); db1 = __dp4a(b1.z, x1.z, db1);
                    da0 = __dp4a(a0.w, x0.w, da0); da1 = __dp4a(a1.w, x1.w, da1); db0 = __dp4a(b0.w, x0.w, db0); db1 = __dp4a(b1.w, x1.w, db1);
                }
#endif
                const int so = (G >> 2) * QW_SG_BYTES + 256 + (G & 3) * 4;
                const float xs0 = *reinterpret_cast<const float*>(sh.xq + so);
                acca = __fadd_rn(acca, q8_term(da0, *reinterpret_cast<const float*>(rowa + so), xs0));
                accb = __fadd_rn(accb, q8_term(db0, *reinterpret_cast<const float*>(rowb + so), xs0));
                if (has2) {
                    const int so2 = (G2 >> 2) * QW_SG_BYTES + 256 + (G2 & 3) * 4;
                    const float xs1 = *reinterpret_cast<const float*>(sh.xq + so2);
                    acca = __fadd_rn(acca, q8_term(da1, *reinterpret_cast<const float*>(rowa + so2), xs1));
                    accb = __fadd_rn(accb, q8_term(db1, *reinterpret_cast<const float*>(rowb + so2), xs1));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acca = __fadd_rn(acca, __shfl_xor_sync(0xffffffffu, acca, o));
                accb = __fadd_rn(accb, __shfl_xor_sync(0xffffffffu, accb, o));
            }
            if (KIND == 2) {
                if (lane == 0) out[grow >> 1] = __fmul_rn(silu_ref(acca), accb);
            } else if (lane < 2 && (lane == 0 || two)) {
                const float v = lane == 0 ? acca : accb;
                out[grow + lane] = KIND == 1 ? __fadd_rn(xres, v) : v;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sh.empty + slot * 8); // this warp is done with the slot
        if (logme) p.tlog[3 * kTileLog + it] = gtime_ns();
    }
}

// ---------------------------------------------------------------- consumer: prologues
// store one quantised group (64 codes + scale) into the shared activation vector (SG layout)
__device__ __forceinline__ void put_group(uint8_t* xq, int g, int lane, float a, float b) {
    const float scale = q8_scale(warp_max(fmaxf(fabsf(a), fabsf(b))));
    int8_t* codes = reinterpret_cast<int8_t*>(xq + (g >> 2) * QW_SG_BYTES + (g & 3) * 64);
    codes[lane] = (int8_t) q8_code(a, scale);
    codes[lane + 32] = (int8_t) q8_code(b, scale);
    if (lane == 0) *reinterpret_cast<float*>(xq + (g >> 2) * QW_SG_BYTES + 256 + (g & 3) * 4) = scale;
}
__device__ __forceinline__ void zero_group(uint8_t* xq, int g, int lane) {
    int8_t* codes = reinterpret_cast<int8_t*>(xq + (g >> 2) * QW_SG_BYTES + (g & 3) * 64);
    codes[lane] = 0;
    codes[lane + 32] = 0;
    if (lane == 0) *reinterpret_cast<float*>(xq + (g >> 2) * QW_SG_BYTES + 256 + (g & 3) * 4) = 0.0f;
}

constexpr int kMaxGroupsPerWarpNorm = 5; // D <= 5120 (Qwen3-32B); checked at init

// x (fp32, D) -> RMSNorm with weights w -> Q8_0 codes + scales in shared memory (forward.c:254-259).
// Warp w owns groups w, w+16, ...; all its loads are issued up front (one L2 round trip).
// from_embedding: the residual stream starts as the dequantised embedding row (forward.c:237).
__device__ void prologue_norm_quant(const Shared& sh, const MegaParams& p, const float* __restrict__ w, bool from_embedding) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = p.D;
    const int groups = D / 64, pad_groups = qw_pad_cols(D) / 64;
    float va[kMaxGroupsPerWarpNorm], vb[kMaxGroupsPerWarpNorm];
    if (from_embedding) {
        const int tok = p.token_dev ? *p.token_dev : p.token;
        const uint8_t* row = p.w_emb + (size_t) tok * qw_row_bytes(D);
#pragma unroll
        for (int k = 0; k < kMaxGroupsPerWarpNorm; ++k) {
            const int g = warp + k * kConsumerWarps;
            va[k] = vb[k] = 0.0f;
            if (g < groups) {
                const uint8_t* rec = row + (size_t) (g >> 2) * QW_SG_BYTES;
                const float sc = *reinterpret_cast<const float*>(rec + 256 + (g & 3) * 4);
                const int8_t* codes = reinterpret_cast<const int8_t*>(rec + (g & 3) * 64);
                va[k] = __fmul_rn((float) codes[lane], sc);
                vb[k] = __fmul_rn((float) codes[lane + 32], sc);
                if (blockIdx.x == 0) { // the residual stream lives in global memory; CTA 0 seeds it
                    p.x[g * 64 + lane] = va[k];
                    p.x[g * 64 + 32 + lane] = vb[k];
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < kMaxGroupsPerWarpNorm; ++k) {
            const int g = warp + k * kConsumerWarps;
            va[k] = vb[k] = 0.0f;
            if (g < groups) {
                va[k] = __ldcg(p.x + g * 64 + lane);
                vb[k] = __ldcg(p.x + g * 64 + 32 + lane);
            }
        }
    }
    // norm weights: issued now so their latency overlaps the reduction (no L1 in this kernel: 1-2 us under load)
    float wa[kMaxGroupsPerWarpNorm], wb[kMaxGroupsPerWarpNorm];
#pragma unroll
    for (int k = 0; k < kMaxGroupsPerWarpNorm; ++k) {
        const int g = warp + k * kConsumerWarps;
        wa[k] = wb[k] = 0.0f;
        if (g < groups) {
            wa[k] = __ldg(w + g * 64 + lane);
            wb[k] = __ldg(w + g * 64 + 32 + lane);
        }
    }
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxGroupsPerWarpNorm; ++k) ss = __fadd_rn(ss, __fadd_rn(__fmul_rn(va[k], va[k]), __fmul_rn(vb[k], vb[k])));
    ss = warp_sum(ss);
    if (lane == 0) sh.misc[warp] = ss;
    bar_consumers();
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; ++i) tot = __fadd_rn(tot, sh.misc[i]);
    const float r = rms_rscale(tot, D);
#pragma unroll
    for (int k = 0; k < kMaxGroupsPerWarpNorm; ++k) {
        const int g = warp + k * kConsumerWarps;
        if (g < groups) {
            const float a = __fmul_rn(wa[k], __fmul_rn(r, va[k]));
            const float b = __fmul_rn(wb[k], __fmul_rn(r, vb[k]));
            put_group(sh.xq, g, lane, a, b);
        } else if (g < pad_groups) {
            zero_group(sh.xq, g, lane);
        }
    }
    bar_consumers();
}

// fp32 vector in global memory (written by other CTAs) -> Q8_0 codes + scales in shared memory
__device__ void prologue_quant_global(const Shared& sh, const float* src, int n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = n / 64, pad_groups = qw_pad_cols(n) / 64;
    constexpr int kBatch = 10; // Hd 9728 -> 152 groups -> 9.5 per warp: one round trip
    for (int g0 = warp; g0 < pad_groups; g0 += kBatch * kConsumerWarps) {
        float va[kBatch], vb[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int g = g0 + k * kConsumerWarps;
            va[k] = vb[k] = 0.0f;
            if (g < groups) {
                va[k] = __ldcg(src + g * 64 + lane);
                vb[k] = __ldcg(src + g * 64 + 32 + lane);
            }
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int g = g0 + k * kConsumerWarps;
            if (g < groups) put_group(sh.xq, g, lane, va[k], vb[k]);
            else if (g < pad_groups) zero_group(sh.xq, g, lane);
        }
    }
    bar_consumers();
}

// already-quantised vector (attention output, SG layout in global) -> shared
__device__ void prologue_load_codes(const Shared& sh, const uint8_t* q, int n) {
    const int bytes = (int) qw_row_bytes(n);
    for (int i = threadIdx.x; i < bytes / 16; i += kConsumerThreads)
        reinterpret_cast<int4*>(sh.xq)[i] = __ldcg(reinterpret_cast<const int4*>(q) + i);
    bar_consumers();
}

// ---------------------------------------------------------------- consumer: attention
// scratch (floats, inside sh.scr): raw[1280] | sq[1280] | scores[4 groups][8 heads][32]
constexpr int kScrRaw = 0, kScrQ = 1280, kScrS = 2560;
constexpr int kScrFloats = kScrS + 4 * 8 * 32;

// RMSNorm weight + RoPE for element i of a 128-wide head (forward.c:267-280, 104-118);
// cos/sin come from the host-computed table so the angles are the reference's bit for bit.
struct RopeCoef { // this thread's constants for element i of any head: loaded once per attention phase
    float c, s, glo, ghi;
};
__device__ __forceinline__ RopeCoef rope_coef(const float* g, const MegaParams& p, int i) {
    const int j = i & 63;
    return {__ldg(p.rope_cos + (size_t) p.pos * 64 + j), __ldg(p.rope_sin + (size_t) p.pos * 64 + j), __ldg(g + j), __ldg(g + j + 64)};
}
__device__ __forceinline__ float head_norm_rope(const float* raw, const RopeCoef& k, float r, int i) {
    const int j = i & 63;
    const float a = __fmul_rn(k.glo, __fmul_rn(r, raw[j]));
    const float b = __fmul_rn(k.ghi, __fmul_rn(r, raw[j + 64]));
    return (i < 64) ? __fsub_rn(__fmul_rn(a, k.c), __fmul_rn(b, k.s)) : __fadd_rn(__fmul_rn(a, k.s), __fmul_rn(b, k.c));
}
// sum of squares over 128 values by one warp (4 per lane, then the shuffle tree)
__device__ __forceinline__ float head_rscale_warp(const float* raw, int lane) {
    const float4 v = *reinterpret_cast<const float4*>(raw + lane * 4);
    float ss = __fmul_rn(v.x, v.x);
    ss = __fmaf_rn(v.y, v.y, ss);
    ss = __fmaf_rn(v.z, v.z, ss);
    ss = __fmaf_rn(v.w, v.w, ss);
    return rms_rscale(warp_sum(ss), 128);
}

// Split-KV attention for the units of this CTA. The 16 consumer warps form 4 groups of 4 warps;
// each group takes whole 28-position tiles (unit i of a segment -> group i % 4) and runs the three
// stages of a tile -- scores, online-softmax update, P.V -- synchronised by its OWN named barrier
// (128 threads), so four tiles are in flight per SM and no CTA-wide barrier sits in the tile loop.
//   scores : 8 lanes per position (lane covers float4 columns sub, sub+8, sub+16, sub+24 -> every
//            quarter-warp reads 128 contiguous bytes of K), all KV_MUL heads, 3-step shuffle tree;
//   softmax: warp w of the group owns heads w, w+4; lane = position;
//   P.V    : thread = (head, 4 output dims), float4 accumulator in registers.
// A segment = the CTA's units of one KV head; its query heads (and, if the CTA owns the last
// chunk, this step's own K/V row: RMSNorm + RoPE, written to the cache for later steps) are
// prepared once per segment with CTA-wide barriers, and the 4 group states are merged at its end.
constexpr int kAttnGroups = 4, kAttnGT = 128;
__device__ __forceinline__ void bar_group(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kAttnGT) : "memory");
}

template <int KV_MUL>
__device__ void consume_attn(const Shared& sh, const MegaParams& p, int l, unsigned& it) {
    constexpr int NACC = (KV_MUL * 32 + kAttnGT - 1) / kAttnGT; // float4 accumulators per thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 2, gw = warp & 3, gt = tid & (kAttnGT - 1);
    float* raw = sh.scr + kScrRaw;   // [KV_MUL*128 q | 128 k | 128 v] raw projections; reused as flush scratch
    float* sq = sh.scr + kScrQ;      // [KV_MUL*128] normalised + rotated q, then [128] k, [128] v of this step
    float* sc = sh.scr + kScrS + grp * (8 * 32); // this group's scores / probabilities [head][32]
    float* s_r = sh.misc + 16;
    float* g_m = sh.misc + 64 + grp * 32;  // running max per head
    float* g_l = g_m + 8;                  // running sum per head
    float* g_scale = g_m + 16;             // rescale factor of the last update
    int nc, u0, u1;
    attn_units(p, nc, u0, u1);
    const float* gq = p.q_norm + (size_t) l * 128;
    const float* gk = p.k_norm + (size_t) l * 128;
    const float inv = sqrtf(128.0f);
    const unsigned it_base = it;
    float4 acc[NACC];
    // element e = tid & 127 is the one this thread normalises/rotates for every head (the strides are
    // multiples of 128): fetch its cos/sin and norm weights now, off the critical path (no L1 here)
    const RopeCoef kq = rope_coef(gq, p, tid & 127), kk = rope_coef(gk, p, tid & 127);

    // one tile: cnt positions, K rows at Kt, V rows at Vt (shared memory)
    auto tile = [&](const float* Kt, const float* Vt, int cnt) {
        // ---- scores (forward.c:156-165)
        const int sub = gt & 7;
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
            const int pos = r * 16 + (gt >> 3);
            const int pc = min(pos, cnt - 1); // idle lanes recompute a valid row: uniform control flow for the shuffles
            float d[KV_MUL];
#pragma unroll
            for (int j = 0; j < KV_MUL; ++j) d[j] = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 kf = *reinterpret_cast<const float4*>(Kt + pc * 128 + (sub + 8 * i) * 4);
#pragma unroll
                for (int j = 0; j < KV_MUL; ++j) {
                    const float4 qf = *reinterpret_cast<const float4*>(sq + j * 128 + (sub + 8 * i) * 4);
                    d[j] = __fmaf_rn(qf.x, kf.x, d[j]);
                    d[j] = __fmaf_rn(qf.y, kf.y, d[j]);
                    d[j] = __fmaf_rn(qf.z, kf.z, d[j]);
                    d[j] = __fmaf_rn(qf.w, kf.w, d[j]);
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < KV_MUL; ++j) d[j] = __fadd_rn(d[j], __shfl_xor_sync(0xffffffffu, d[j], o));
            }
            if (sub == 0 && pos < cnt) {
#pragma unroll
                for (int j = 0; j < KV_MUL; ++j) sc[j * 32 + pos] = __fdiv_rn(d[j], inv); // score / sqrtf(head_dim)
            }
        }
        bar_group(grp);
        // ---- online softmax update: warp gw owns heads gw, gw+4; lane = position
        for (int j = gw; j < KV_MUL; j += 4) {
            const float s = lane < cnt ? sc[j * 32 + lane] : -INFINITY;
            const float m_old = g_m[j];
            const float m_new = fmaxf(m_old, warp_max(s));
            const float e = lane < cnt ? expf(__fsub_rn(s, m_new)) : 0.0f;
            if (lane < cnt) sc[j * 32 + lane] = e;
            const float lsum = warp_sum(e);
            if (lane == 0) {
                const float scl = (m_old == -INFINITY) ? 0.0f : expf(__fsub_rn(m_old, m_new));
                g_scale[j] = scl;
                g_l[j] = __fmaf_rn(g_l[j], scl, lsum);
                g_m[j] = m_new;
            }
        }
        bar_group(grp);
        // ---- P.V: thread = (head j, dims 4*d4 .. 4*d4+3)
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            const int idx = gt + k * kAttnGT;
            if (idx < KV_MUL * 32) {
                const int j = idx >> 5, d4 = idx & 31;
                const float scl = g_scale[j];
                float4 a = acc[k];
                a.x = __fmul_rn(a.x, scl); a.y = __fmul_rn(a.y, scl); a.z = __fmul_rn(a.z, scl); a.w = __fmul_rn(a.w, scl);
                for (int i = 0; i < cnt; ++i) {
                    const float pw = sc[j * 32 + i];
                    const float4 vv = *reinterpret_cast<const float4*>(Vt + i * 128 + d4 * 4);
                    a.x = __fmaf_rn(pw, vv.x, a.x); a.y = __fmaf_rn(pw, vv.y, a.y);
                    a.z = __fmaf_rn(pw, vv.z, a.z); a.w = __fmaf_rn(pw, vv.w, a.w);
                }
                acc[k] = a;
            }
        }
        bar_group(grp); // sc / g_scale are rewritten by this group's next tile
    };

    int u = u0;
    while (u < u1) {
        const int kvh = u / nc;
        const int seg_end = min(u1, (kvh + 1) * nc);
        const bool own_last = seg_end == (kvh + 1) * nc; // the segment contains the KV head's last chunk
        // ---- segment prologue (CTA-wide): q heads, and this step's K/V row if we own the last chunk
        const int nraw = KV_MUL * 128 + (own_last ? 256 : 0);
        bar_consumers();
        for (int i = tid; i < nraw; i += kConsumerThreads) {
            const float* src = i < KV_MUL * 128 ? p.qkv + (size_t) kvh * KV_MUL * 128 + i
                               : i < KV_MUL * 128 + 128 ? p.qkv + p.Pl + (size_t) kvh * 128 + (i - KV_MUL * 128)
                                                        : p.qkv + p.Pl + p.Kl + (size_t) kvh * 128 + (i - KV_MUL * 128 - 128);
            raw[i] = __ldcg(src);
        }
        bar_consumers();
        if (warp < KV_MUL + (own_last ? 1 : 0)) {
            const float r = head_rscale_warp(raw + warp * 128, lane);
            if (lane == 0) s_r[warp] = r;
        }
        bar_consumers();
        for (int i = tid; i < nraw; i += kConsumerThreads) {
            const int hd = i >> 7, e = i & 127;
            if (hd < KV_MUL) {
                sq[i] = head_norm_rope(raw + hd * 128, kq, s_r[hd], e);
            } else {
                const size_t coff = (((size_t) l * p.KVHl + kvh) * p.S + p.pos) * 128;
                if (hd == KV_MUL) {
                    const float kx = head_norm_rope(raw + hd * 128, kk, s_r[hd], e);
                    sq[i] = kx;
                    p.k_cache[coff + e] = kx;
                } else {
                    sq[i] = raw[i];
                    p.v_cache[coff + e] = raw[i];
                }
            }
        }
        if (gt < 8) {
            g_m[gt] = -INFINITY;
            g_l[gt] = 0.0f;
            g_scale[gt] = 0.0f;
        }
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        bar_consumers();
        // ---- tiles of the segment: every warp walks every tile (ring protocol), its group works on every 4th
        for (int uu = u; uu < seg_end; ++uu) {
            const unsigned itx = it_base + (unsigned) (uu - u0);
            const unsigned slot = itx % p.nslot, par = (itx / p.nslot) & 1;
            mbar_wait(sh, p, sh.full + slot * 8, par, 4);
            if (((uu - u) & (kAttnGroups - 1)) == grp) {
                const int p0 = (uu - kvh * nc) * kChunk;
                const int cnt = min(p.pos, p0 + kChunk) - p0; // cached positions in the tile
                const float* Kt = reinterpret_cast<const float*>(sh.ring + (size_t) slot * kSlotBytes);
                if (cnt > 0) tile(Kt, Kt + kChunk * 128, cnt);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sh.empty + slot * 8);
        }
        if (own_last && grp == 0) tile(sq + KV_MUL * 128, sq + KV_MUL * 128 + 128, 1); // this step's own position
        // ---- merge the 4 group states (online-softmax merge) and publish (m, l, acc) for the combine phase
        bar_consumers();
        float* red = raw; // [grp][head][132]; raw/sq are dead now
        constexpr int HP = KV_MUL < 4 ? KV_MUL : 4; // heads per pass (scratch is 2560 floats)
#pragma unroll
        for (int h0 = 0; h0 < KV_MUL; h0 += HP) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) {
                const int idx = gt + k * kAttnGT;
                const int j = idx >> 5, d4 = idx & 31;
                if (idx < KV_MUL * 32 && j >= h0 && j < h0 + HP)
                    *reinterpret_cast<float4*>(red + (grp * HP + (j - h0)) * 132 + d4 * 4) = acc[k];
            }
            if (gt < HP) {
                red[(grp * HP + gt) * 132 + 128] = g_m[h0 + gt];
                red[(grp * HP + gt) * 132 + 129] = g_l[h0 + gt];
            }
            bar_consumers();
            if (tid < HP * 128) {
                const int j = tid >> 7, d = tid & 127;
                float M = -INFINITY;
#pragma unroll
                for (int g = 0; g < kAttnGroups; ++g) M = fmaxf(M, red[(g * HP + j) * 132 + 128]);
                float L = 0.0f, A = 0.0f;
#pragma unroll
                for (int g = 0; g < kAttnGroups; ++g) {
                    const float mg = red[(g * HP + j) * 132 + 128];
                    const float e = (mg == -INFINITY) ? 0.0f : expf(__fsub_rn(mg, M));
                    L = __fmaf_rn(red[(g * HP + j) * 132 + 129], e, L);
                    A = __fmaf_rn(red[(g * HP + j) * 132 + d], e, A);
                }
                const int h = kvh * KV_MUL + h0 + j;
                const size_t slot = (size_t) h * gridDim.x + attn_block(p);
                p.part_acc[slot * 128 + d] = A;
                if (d == 0) {
                    p.part_m[slot] = M;
                    p.part_l[slot] = L;
                }
            }
            bar_consumers();
        }
        u = seg_end;
    }
    it = it_base + (unsigned) (u1 - u0);
}

__device__ void consume_attn_dispatch(const Shared& sh, const MegaParams& p, int l, unsigned& it) {
    switch (p.kv_mul) {
        case 1: consume_attn<1>(sh, p, l, it); break;
        case 2: consume_attn<2>(sh, p, l, it); break;
        case 4: consume_attn<4>(sh, p, l, it); break;
        default: consume_attn<8>(sh, p, l, it); break;
    }
}

// merge the split-KV partials of each head (online-softmax merge), write fp32 att (debug
// read-back) and its Q8_0 codes straight into the SG-layout vector the wo GEMV loads.
__device__ void combine_attn(const Shared& sh, const MegaParams& p) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* mm = sh.scr;          // [kMaxGrid]
    float* ww = sh.scr + 256;    // [kMaxGrid]
    float* ll = sh.scr + 512;    // [kMaxGrid]
    float* red = sh.scr + 768;   // [4][128]
    float* so = sh.scr + 1280;   // [128]
    const int nc = p.pos / kChunk + 1;
    const long long U = (long long) p.KVHl * nc;
    const int G = gridDim.x;
    for (int h = blockIdx.x; h < p.Hl; h += G) {
        const int kvh = h / p.kv_mul;
        const long long ulo = (long long) kvh * nc, uhi = ulo + nc;
        const int blo = max((int) (ulo * G / U) - 1, 0), bhi = min((int) (uhi * G / U) + 1, G - 1);
        const int nb = bhi - blo + 1;
        // all global loads of this head are issued up front (one L2 round trip): m, l by the first nb
        // threads, the nb partial accumulators by (dim, part) threads, 4 parts x up to 8 each
        const int d = tid & 127, part = tid >> 7;
        constexpr int kMaxPer = 8;
        float av[kMaxPer];
#pragma unroll
        for (int k = 0; k < kMaxPer; ++k) {
            const int i = part + 4 * k;
            av[k] = i < nb ? __ldcg(p.part_acc + ((size_t) h * G + blo + i) * 128 + d) : 0.0f;
        }
        float my_m = -INFINITY, my_l = 0.0f;
        if (tid < nb) {
            const int b = blo + tid;
            const long long s0 = U * b / G, s1 = U * (b + 1) / G;
            if (max(s0, ulo) < min(s1, uhi)) {
                my_m = __ldcg(p.part_m + (size_t) h * G + b);
                my_l = __ldcg(p.part_l + (size_t) h * G + b);
            }
        }
        bar_consumers(); // scratch free
        if (tid < nb) {
            mm[tid] = my_m;
            ll[tid] = my_l;
        }
        bar_consumers();
        float M = -INFINITY;
        for (int i = 0; i < nb; ++i) M = fmaxf(M, mm[i]);
        if (tid < nb) ww[tid] = (mm[tid] == -INFINITY) ? 0.0f : expf(__fsub_rn(mm[tid], M));
        bar_consumers();
        {
            float A = 0.0f;
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) {
                const int i = part + 4 * k;
                if (i < nb && ww[i] != 0.0f) A = __fmaf_rn(av[k], ww[i], A); // stale slots have weight 0 and are skipped
            }
            for (int i = part + 4 * kMaxPer; i < nb; i += 4) { // very long merges (few KV heads per GPU)
                const float w = ww[i];
                if (w != 0.0f) A = __fmaf_rn(__ldcg(p.part_acc + ((size_t) h * G + blo + i) * 128 + d), w, A);
            }
            red[part * 128 + d] = A;
        }
        bar_consumers();
        if (tid < 128) {
            float Lsum = 0.0f;
            for (int i = 0; i < nb; ++i) Lsum = __fmaf_rn(ll[i], ww[i], Lsum);
            const float A = __fadd_rn(__fadd_rn(red[tid], red[128 + tid]), __fadd_rn(red[256 + tid], red[384 + tid]));
            const float o = __fdiv_rn(A, Lsum);
            so[tid] = o;
            p.att[(size_t) h * 128 + tid] = o;
        }
        bar_consumers();
        if (warp < 2) put_group(p.att_q, h * 2 + warp, lane, so[warp * 64 + lane], so[warp * 64 + 32 + lane]);
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
        consume_mat<0>(sh, p, ph_qkv(p, l), it, p.qkv);
        stamp(p, l, 2);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 3);
        consume_attn_dispatch(sh, p, l, it);
        stamp(p, l, 4);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 5);
        combine_attn(sh, p);
        stamp(p, l, 6);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 7);
        prologue_load_codes(sh, p.att_q, p.Pl);
        consume_mat<1>(sh, p, ph_o(p, l), it, p.x);
        stamp(p, l, 8);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 9);
        // --- feed-forward block (forward.c:303-338)
        prologue_norm_quant(sh, p, p.ffn_norm + (size_t) l * p.D, false);
        stamp(p, l, 10);
        consume_mat<2>(sh, p, ph_13(p, l), it, p.h);
        stamp(p, l, 11);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 12);
        prologue_quant_global(sh, p.h, p.Hdl);
        stamp(p, l, 13);
        consume_mat<1>(sh, p, ph_2(p, l), it, p.x);
        stamp(p, l, 14);
        grid_barrier(sh, p, nbar);
        stamp(p, l, 15);
    }
    stamp(p, p.L, 0);
    // --- final norm + classifier (forward.c:344-348)
    prologue_norm_quant(sh, p, p.out_norm, p.layers_run == 0);
    stamp(p, p.L, 1);
    consume_mat<0>(sh, p, ph_cls(p), it, p.logits);
    stamp(p, p.L, 2);
}

__global__ void __launch_bounds__(kThreads, 1) k_decode(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int abort_flag;
    Shared sh;
    sh.ring = smem;
    sh.xq = smem + p.off_xq;
    sh.scr = reinterpret_cast<float*>(smem + p.off_scr);
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
    if (qw_pad_cols(c->D) / 64 > kMaxGroupsPerWarpNorm * kConsumerWarps) {
        qw_set_error("persistent decode kernel: dim %d exceeds the fused RMSNorm prologue's %d columns", c->D,
                     kMaxGroupsPerWarpNorm * kConsumerWarps * 64);
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
    const int xq_b = (int) qw_row_bytes(amax);
    const int scr_b = kScrFloats * 4;
    const int misc_b = 1024, bar_b = 2 * kMaxSlots * 8;
    const int fixed = ((xq_b + 127) & ~127) + ((scr_b + 127) & ~127) + ((misc_b + 127) & ~127) + ((bar_b + 127) & ~127);
    const int avail = dev_smem - fixed - 1024; // 1 KB left for static shared + driver reserve
    st->nslot = std::min(kMaxSlots, avail / kSlotBytes);
    if (const char* e = getenv("QWEN_MEGA_NSLOT")) st->nslot = std::max(2, std::min(st->nslot, atoi(e)));
    if (const char* e = getenv("QWEN_MEGA_MODE")) st->dbg_mode = atoi(e);
    if (const char* e = getenv("QWEN_MEGA_SPLIT")) st->copy_split = std::max(1, atoi(e));
    if (getenv("QWEN_MEGA_VERBOSE")) fprintf(stderr, "[mega] nslot %d mode %d split %d smem %d\n", st->nslot, st->dbg_mode, st->copy_split, fixed);
    if (st->nslot < 2) {
        qw_set_error("persistent decode kernel: not enough shared memory for a 2-slot ring (%d bytes free)", avail);
        return -1;
    }
    take(st->nslot * kSlotBytes);
    st->off_xq = take(xq_b);
    st->off_scr = take(scr_b);
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
    {
        auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
        int k = std::max(1, st->grid / 3);
        while (gcd(k, st->grid) != 1) ++k;
        st->perm = k;
        if (const char* e = getenv("QWEN_MEGA_PERM")) { const int v = atoi(e); if (v > 0 && gcd(v, st->grid) == 1) st->perm = v; }
    }
    QW_CUDA(cudaMalloc((void**) &st->att_q, qw_row_bytes(c->Pl)));
    QW_CUDA(cudaMemset(st->att_q, 0, qw_row_bytes(c->Pl))); // pad groups stay zero
    QW_CUDA(cudaMalloc((void**) &st->part_m, (size_t) c->Hl * st->grid * 4));
    QW_CUDA(cudaMalloc((void**) &st->part_l, (size_t) c->Hl * st->grid * 4));
    QW_CUDA(cudaMalloc((void**) &st->part_acc, (size_t) c->Hl * st->grid * 128 * 4));
    QW_CUDA(cudaMemset(c->bar_counter, 0, 512 * 8)); // one flag word per CTA
    c->bar_epoch = 0;
    c->path = 0;
    return 0;
}

void qw_mega_free(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st) return;
    void* bufs[] = {st->tlog, st->att_q, st->part_m, st->part_l, st->part_acc, st->prof};
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
    p.att_q = st->att_q;
    p.part_m = st->part_m; p.part_l = st->part_l; p.part_acc = st->part_acc;
    p.bar = c->bar_counter; p.bar_base = c->bar_epoch;
    p.err = c->err_flag;
    p.dbg_mode = st->dbg_mode;
    p.perm = st->perm;
    p.copy_split = st->copy_split;
    p.prof = st->prof;
    p.tlog = st->tlog;
    p.tlog_warp = st->tlog_warp;
    p.nslot = st->nslot; p.off_xq = st->off_xq; p.off_scr = st->off_scr;
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
int qw_mega_tlog(QwenCudaCtx* c, int warp, unsigned long long* host) {
    MegaState* st = state_of(c);
    if (!st || !st->grid) return -1;
    if (!host) { // enable
        if (!st->tlog) QW_CUDA(cudaMalloc((void**) &st->tlog, 4 * kTileLog * 8));
        QW_CUDA(cudaMemset(st->tlog, 0, 4 * kTileLog * 8));
        QW_CUDA(cudaDeviceSynchronize());
        st->tlog_warp = warp;
        return kTileLog;
    }
    QW_CUDA(cudaStreamSynchronize(c->stream));
    QW_CUDA(cudaMemcpy(host, st->tlog, 4 * kTileLog * 8, cudaMemcpyDeviceToHost));
    return kTileLog;
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
