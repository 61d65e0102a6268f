// decode_mega.cu -- the whole decode step (reference src/forward.c:225-350) as ONE
// persistent sm_100a kernel, one CTA per SM, organised as a DATAFLOW: there is no grid barrier.
//
// Why one kernel: at batch 1 every op is a matrix-vector product whose only cost is
// streaming its weights from HBM once. A layer is ~107 MB (Qwen3-4B) = ~16 us of HBM
// time split over 7 dependent GEMVs; a launch-per-op design spends about as long in
// launch gaps and pipeline fill/drain as it does streaming.
//
//   * warp 15 (one elected lane) is the PRODUCER. It walks the step's fixed schedule of
//     weight tiles and KV-cache chunks and copies each into a shared-memory ring with
//     cp.async.bulk (TMA bulk copy, completion on an mbarrier). It never waits for
//     activations, so it runs up to one ring (7 x 28 KB per SM, ~29 MB chip-wide) AHEAD
//     of the math, straight through phase boundaries. HBM stays busy while the consumers
//     wait for each other's results.
//   * warps 0..14 are CONSUMERS. Per phase they build the quantised activation vector in
//     shared memory (RMSNorm + Q8_0 quantise, fused), then eat tiles from the ring.
//   * Phases hand vectors to each other through a FLOW ARENA in global memory (L2): every
//     word of the arena starts as a sentinel bit pattern that no computation can produce
//     (a signalling NaN whose low bytes are the int8 code -128, which the quantiser never
//     emits). A producer simply stores its result; a consumer polls the words it needs with
//     ld.relaxed.gpu until none is the sentinel. No flags, no counters: one store
//     latency plus one L2 round trip per hand-off, and a CTA starts a phase the moment ITS
//     inputs exist. Results leave an SM as ONE TMA bulk store per CTA (cp.async.bulk from a
//     shared-memory staging area: GEMV rows and attention partials are contiguous in the arena),
//     so no warp waits in a fence and the stores do not queue behind the polling loads. Every (layer, vector) has its own slot in the arena, so each word is
//     written exactly once per launch and there is no write-after-read hazard inside a launch;
//     two arenas alternate between launches and each launch refills the other one with the
//     sentinel for the next launch (stream order makes that refill complete before it is used).
//
// Work split: every matrix is split by contiguous row ranges over the CTAs (no split-K,
// no atomics -> deterministic, bit-identical run to run). Attention is split-KV: the
// (kv head, 28-position chunk) tiles are spread evenly over all CTAs, each warp runs a whole
// online-softmax over its share of the positions with no cross-warp synchronisation, the CTA merges its
// 15 warp states in shared memory and publishes one partial per head; (head, half) combine tasks
// merge the partials of a head and write the attention output already quantised for wo.
//
// Tensor parallelism (k_decode<KV_MUL, true>, one process per GPU): the all-reduce after wo / w2 is
// part of the dataflow. Every rank's arenas are mapped into every process (cudaIpc, NVLink peer
// access); the wo / w2 epilogue bulk-stores the CTA's partial sums into slot tp_rank of EVERY rank's
// arena, and the next prologue adds the tp slots in rank order to the residual stream that each CTA
// keeps in shared memory -- bit-identical x on all ranks, no collective call, no extra launch.
//
// Every wait in this file has a wall-clock timeout that raises a sticky error flag
// instead of hanging the GPU.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "attn_core.cuh"

namespace {

constexpr int kConsumerWarps = 15;         // + the producer warp = 16 warps = 512 threads: 128 registers per thread
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kSlotBytes = 28672;           // one ring slot: 105 SG records or 2 x 28 KV rows
constexpr int kChunk = 28;                  // KV positions per attention tile (2*28*512 B = one slot)
constexpr int kMaxSlots = 8;
constexpr int kMaxGrid = 256;
constexpr unsigned long long kTimeoutNs = 4000000000ull; // also covers the launch skew between tensor-parallel ranks
constexpr int kMaxTp = 8;
constexpr int kProfSlots = 16;              // per layer: stamps after each phase step (CTA-local, thread 0)
constexpr uint32_t kSent = 0x7F808080u;     // "not written yet": sNaN as fp32, contains code -128 as int8x4
constexpr int kGroupTiles = 4;              // attention tiles the 15 warps work on together
constexpr int kPartStride = 132;            // floats per published attention partial: acc[128], m, l, pad

// one weight matrix of the step's schedule, filled on the host (no divisions left for the kernel)
struct MatDesc {
    const uint8_t* base; // layer 0
    size_t stride;       // bytes between layers
    int rows, n, gran;   // rows, columns, granularity of the per-CTA row split
    int rt;              // rows per ring tile
    int stage;           // 1: results are collected in shared memory and leave as ONE bulk store per CTA (see consume_mat)
    int C, cr;           // per-warp streaming variant: chunks per row and records per chunk (decode_pw.cuh)
    int kind;            // epilogue: 0 out[row] = v, 1 out[row] = resid[row] + v, 2 out[row/2] = silu(v0) * v1,
                         // 3 tensor parallel: this rank's partial v goes to slot tp_rank of EVERY rank's arena (peer stores)
};

struct MegaParams {
    int D, Hdl, L, Hl, KVHl, Pl, Kl, Vl, S, kv_mul;
    int pos, token, layers_run;
    int l_begin;             // debug: first layer to run (0 normally); the residual stream then starts from x_inject
    const float* x_inject;   // debug: replaces the embedding row as x entering layer l_begin (nullptr normally)
    uint8_t* dbg_codes;      // debug: CTA 0 copies the Q8_0 activation vector of every GEMV here, [4 * L + 1][dbg_stride]
    int dbg_stride;
    int perm;       // CTA -> row-block permutation multiplier (coprime to the grid)
    int dbg_mode;   // 0 normal; 1 consumers skip the GEMV math (ring throughput test)
    int pw_late;    // per-warp variant: 1 = a phase's first weight items are issued after its hand-off polls, not before
    int pw_pub;     // per-warp variant: 0 = results leave as one TMA bulk store per CTA, 1 = plain stores by warp 0 + fence
    int l2_cls;     // bytes of this CTA's classifier rows to prefetch into L2 ahead of the classifier phase
    int inflight;   // producer: at most this many bulk-copy tiles in flight per SM (0 = the ring depth)
    int l2_ahead;   // 0: no L2 prefetch; else prefetch the next sub-phase (and this many 32 KB pieces of the classifier)
    const int* token_dev;
    MatDesc mat[5]; // 0 wq|wk|wv, 1 wo, 2 w1/w3 interleaved, 3 w2, 4 classifier
    const uint8_t* w_emb;
    const float *att_norm, *ffn_norm, *out_norm, *q_norm, *k_norm, *rope_cos, *rope_sin;
    float *k_cache, *v_cache;
    float* flow;        // this launch's arena: every word is kSent until its producer stores it
    float* flow_other;  // the other arena: refilled with kSent by this launch, used by the next
    float* flow_x0;     // [D] inside `flow`: the dequantised embedding row (residual stream entering layer 0)
    size_t flow_layer_words, flow_words;
    int o_xa, o_xb, o_qkv, o_h, o_attq, o_part, part_slots; // word offsets inside a layer's block
    float *att, *logits;
    int* err;
    unsigned long long* prof; // optional [CTA][L+1][kProfSlots] globaltimer stamps (debug)
    int nslot, off_xq, off_scr, off_misc, off_bar;
    // tensor parallelism (one process per GPU): the all-reduce after wo / w2 is fused into the GEMV epilogue as
    // peer stores over NVLink into every rank's flow arena; readers add the tp partials in rank order
    int tp, tp_rank;
    int attn_ga;        // blocks that take part in attention (<= grid): min(grid, cap * KVHl)
    int part_stride;    // words between the per-rank partial vectors of xa / xb
    int off_xres;       // shared memory: the fp32 residual stream, kept by every CTA (tp > 1 only)
    float* peer_flow[kMaxTp]; // this launch's arena on every rank (own entry = flow)
};

struct MegaState {
    float* arena[2] = {nullptr, nullptr};
    size_t layer_words = 0, words = 0, x0_off = 0;
    int o_xa = 0, o_xb = 0, o_qkv = 0, o_h = 0, o_attq = 0, o_part = 0, part_slots = 0;
    unsigned launches = 0;
    int last_layers = 0;
    int grid = 0, nslot = 0, dbg_mode = 0, perm = 1, l2_ahead = 0, inflight = 0;
    unsigned long long* prof = nullptr;
    size_t smem = 0;
    int off_xq, off_scr, off_misc, off_bar, off_xres = 0;
    int attn_ga = 0, part_stride = 0;
    float* peer[2][kMaxTp] = {};  // mapped arenas of every rank (cudaIpc), [parity][rank]; own entries = arena[parity]
    bool peers_open = false;
    bool pw = false;              // per-warp streaming variant (decode_pw.cuh): single-GPU contexts
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
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(0x12F0000000000000ull) // evict-first: a tile is read once
                 : "memory");
}
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

// flow arena accesses: always at L2 (gpu scope), never cached in L1
__device__ __forceinline__ uint32_t ldf_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldf_u4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stf_f32(float* p, float v) {
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stf_u32(void* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stf_f4(float* p, const float4& v) {
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// tensor-parallel partials: written by peer GPUs over NVLink, so system scope on both sides
__device__ __forceinline__ void stf_sys_f32(float* p, float v) {
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 ldf_sys_u4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// "push my stores out now": see consume_mat. QW_FENCE=1 tries the lighter acq_rel fence instead of __threadfence()'s sc fence.
__device__ __forceinline__ void flush_stores() {
#if defined(QW_FENCE) && QW_FENCE == 1
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
    __threadfence();
#endif
}
__device__ __forceinline__ bool unset4(const uint4& v) { return v.x == kSent || v.y == kSent || v.z == kSent || v.w == kSent; }

struct Shared {
    uint8_t* ring;
    uint8_t* xq;   // activation vector in SG layout: [256 codes][4 scales] records
    float* xres;   // tp > 1: the residual stream x[D], element i owned by the thread that quantises it
    float* scr;
    float* misc;
    uint32_t full, empty; // shared-space addresses of the barrier arrays
    volatile int* abort_flag;
};

// Slow paths are deliberately NOT inlined: the kernel's code must stay small enough for the SM's
// instruction cache (an earlier build inlined the timeout logic at ~40 sites: 577 KB of SASS, every
// phase change ran cold code fetched through a saturated L2).
__device__ __noinline__ void mbar_wait_slow(volatile int* abort_flag, int* err, uint32_t bar, uint32_t parity, int code) {
    unsigned long long t0 = 0;
    for (unsigned n = 1; !mbar_try_wait(bar, parity); ++n) {
        if ((n & 255u) == 0) {
            if (*abort_flag) return;
            const unsigned long long now = gtime_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTimeoutNs) {
                *abort_flag = code;
                *err = code;
                return;
            }
        }
    }
}
__device__ __forceinline__ void mbar_wait(const Shared& sh, const MegaParams& p, uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(sh.abort_flag, p.err, bar, parity, code);
}
// shared tail of the inline polling loops: called every 256 rounds; true = give up (abort raised here or elsewhere)
__device__ __noinline__ bool poll_timed_out(volatile int* abort_flag, int* err, unsigned long long& t0, int code) {
    if (*abort_flag) return true;
    const unsigned long long now = gtime_ns();
    if (t0 == 0) t0 = now;
    if (now - t0 > kTimeoutNs) {
        *abort_flag = code;
        *err = code;
        return true;
    }
    return false;
}
// poll until the 4 (or 1) words at q are all written
__device__ __noinline__ uint4 poll4_slow(volatile int* abort_flag, int* err, const void* q, int code) {
    unsigned long long t0 = 0;
    uint4 v = ldf_u4(q);
    for (unsigned n = 1; unset4(v); ++n) {
        if ((n & 255u) == 0) {
            if (*abort_flag) break;
            const unsigned long long now = gtime_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTimeoutNs) {
                *abort_flag = code;
                *err = code;
                break;
            }
        }
        v = ldf_u4(q);
    }
    return v;
}
__device__ __noinline__ uint32_t poll1_slow(volatile int* abort_flag, int* err, const void* q, int code) {
    unsigned long long t0 = 0;
    uint32_t v = ldf_u32(q);
    for (unsigned n = 1; v == kSent; ++n) {
        if ((n & 255u) == 0) {
            if (*abort_flag) break;
            const unsigned long long now = gtime_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTimeoutNs) {
                *abort_flag = code;
                *err = code;
                break;
            }
        }
        v = ldf_u32(q);
    }
    return v;
}
__device__ __forceinline__ uint4 poll4(const Shared& sh, const MegaParams& p, const void* q, int code) {
    return poll4_slow(sh.abort_flag, p.err, q, code);
}
__device__ __forceinline__ float poll1(const Shared& sh, const MegaParams& p, const float* q, int code) {
    const uint32_t v = ldf_u32(q);
    return __uint_as_float(v != kSent ? v : poll1_slow(sh.abort_flag, p.err, q, code));
}
__device__ __forceinline__ float4 as_f4(const uint4& v) {
    return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ float* flow_layer(const MegaParams& p, int l) { return p.flow + (size_t) l * p.flow_layer_words; }

// Position in the shared-memory ring: slot index and the mbarrier phase parity of this pass over the ring. Advanced by
// one per tile; computing it as it % nslot, (it / nslot) & 1 put a runtime integer division (I2F + MUFU.RCP + fix-up,
// ~40 dependent instructions) into every tile iteration of every warp: 5 % of all stall samples in ncu.
struct RingPos {
    unsigned slot, par;
    __device__ __forceinline__ void next(int nslot) {
        if (++slot == (unsigned) nslot) {
            slot = 0;
            par ^= 1u;
        }
    }
};

// ---------------------------------------------------------------- schedule (shared by producer and consumers)
// The step is a fixed list of phases: 4 per layer (QKV, WO, W1/W3, W2) and the classifier. Producer and
// consumers walk it with ONE copy of each routine (runtime descriptors, no per-phase template
// instances): the whole kernel has to fit the SM's 32 KB instruction cache or every phase change
// fetches cold code through an L2 that is busy streaming weights.
__device__ __forceinline__ int my_block(int perm) { return (int) ((blockIdx.x * (unsigned) perm) % gridDim.x); }
__device__ __forceinline__ void cta_rows(const MatDesc& m, int perm, int& r0, int& r1) {
    const unsigned units = (unsigned) m.rows / (unsigned) m.gran; // units * grid < 2^31 for every supported shape
    // CTA -> row-block map: a multiplicative permutation (perm coprime to the grid). With the identity
    // map the GEMV phases ran 8-10 % slower and a fixed third of the CTAs finished 2-4 us late in every
    // phase (measured, profiles/r1_k_decode_summary.md); any scattering permutation removes that.
    const unsigned b = (unsigned) my_block(perm);
    r0 = (int) (units * b / gridDim.x) * m.gran;
    r1 = (int) (units * (b + 1) / gridDim.x) * m.gran;
}

// Attention work split. Every block serves ONE kv head: kv head h gets the blocks [G*h/KVH, G*(h+1)/KVH) (18 or
// 19 of 148 for 8 kv heads) and block j of its n attends over the cached positions [pos*j/n, pos*(j+1)/n) -- an
// even split to the position, streamed as tiles of up to 28 rows that start at the block's own first position.
// The last block of a kv head also takes this step's own position. (Splitting whole 28-position tiles evenly
// over all blocks left one block per kv head with two segments -- two query preparations, two merges, two
// publications -- and 8 vs 9 tiles elsewhere; every combine task waited for the slowest.)
struct AttnSplit {
    int kvh, j, n;   // kv head, index of this block among the kv head's n blocks
    int p_lo, p_hi;  // cached positions [p_lo, p_hi) of this block
    bool active;     // false: this block takes no part in attention (tensor-parallel ranks with few kv heads
                     // use at most p.attn_ga blocks, so a combine task never merges more than ~42 partials)
};
__device__ __forceinline__ AttnSplit attn_split(const MegaParams& p) {
    const int G = p.attn_ga, KVH = p.KVHl, b = my_block(p.perm);
    if (b >= G) return AttnSplit{0, 0, 0, 0, 0, false};
    int h = b * KVH / G;
    while ((h + 1) * G / KVH <= b) ++h;
    const int f0 = h * G / KVH;
    AttnSplit a;
    a.kvh = h;
    a.j = b - f0;
    a.n = (h + 1) * G / KVH - f0;
    a.p_lo = (int) ((long long) p.pos * a.j / a.n);
    a.p_hi = (int) ((long long) p.pos * (a.j + 1) / a.n);
    a.active = true;
    return a;
}

__device__ __forceinline__ void stamp(const MegaParams& p, int l, int k) {
    if (p.prof && threadIdx.x == 0)
        p.prof[((size_t) blockIdx.x * (p.L + 1) + l) * kProfSlots + k] = gtime_ns();
}

// ---------------------------------------------------------------- producer
// L2 prefetch of a contiguous byte range (the hardware takes it in pieces of <= 32 KB here)
__device__ __forceinline__ void prefetch_l2(const uint8_t* src, size_t bytes) {
#pragma unroll 1
    for (size_t o = 0; o < bytes; o += 32768) {
        const uint32_t n = (uint32_t) min((size_t) 32768, bytes - o);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + o), "r"(n) : "memory");
    }
}
// prefetch everything this CTA will read in sub-phase sp (5 * layer + {0 QKV, 1 attention, 2 WO, 3 W13, 4 W2};
// 5 * layers_run = classifier): one contiguous range per matrix, one K and one V range per KV-head segment
__device__ __noinline__ void prefetch_subphase(const MegaParams& p, int sp) {
    const int nsp = 5 * p.layers_run;
    if (sp > nsp) return;
    const int l = sp / 5, k = sp == nsp ? 5 : sp % 5;
    if (k == 1) {
        const AttnSplit a = attn_split(p);
        if (a.active && a.p_hi > a.p_lo) {
            const size_t off = (((size_t) l * p.KVHl + a.kvh) * p.S + a.p_lo) * 128;
            prefetch_l2(reinterpret_cast<const uint8_t*>(p.k_cache + off), (size_t) (a.p_hi - a.p_lo) * 512);
            prefetch_l2(reinterpret_cast<const uint8_t*>(p.v_cache + off), (size_t) (a.p_hi - a.p_lo) * 512);
        }
    } else {
        const MatDesc& m = p.mat[k == 0 ? 0 : k == 5 ? 4 : k - 1];
        int r0, r1;
        cta_rows(m, p.perm, r0, r1);
        const size_t rb = qw_row_bytes(m.n);
        size_t bytes = (size_t) (r1 - r0) * rb;
        if (k == 5) bytes = min(bytes, (size_t) p.l2_cls); // the classifier is far larger than L2's share
        prefetch_l2(m.base + (k == 5 ? 0 : (size_t) l * m.stride) + (size_t) r0 * rb, bytes);
    }
}

// The producer copies tile `it` of the schedule into shared-memory slot it % nslot as soon as the consumers
// have released the slot (evict-first L2 hint: once in shared memory a tile is dead in L2).
// Optional (QWEN_MEGA_L2AHEAD=1, default off): on ENTERING a sub-phase also prefetch the whole next sub-phase
// of this CTA into L2 with cp.async.bulk.prefetch.L2. Measured on B200 (scripts/ubench/l2pf.cu,
// handoff.cu): HBM->L2 prefetch traffic does not slow the hand-offs the way streaming into shared memory
// does, but a bulk copy that hits L2 is no faster than one that streams from HBM -- 148 SMs x 6 copies of
// 28 KB in flight top out at 7.5 TB/s either way -- so the ring refills no sooner and the step time does
// not improve (2.07 ms vs 2.14 ms with the prefetch on).
// Optional cap on the bulk copies in flight (QWEN_MEGA_INFLIGHT): before issuing tile i the producer waits until tile
// i - cap has landed. A full ring's worth of copies issued at once at every phase boundary (5 x 28 KB x 148 SMs = 20 MB)
// keeps the L2 miss queues full for ~3 us, and the hand-off polls of that moment queue behind them.
__device__ __forceinline__ void producer_throttle(const Shared& sh, const MegaParams& p, const RingPos& rp, unsigned issued) {
    if (p.inflight <= 0 || issued < (unsigned) p.inflight) return;
    const unsigned back = (unsigned) p.inflight;
    const unsigned slot = rp.slot >= back ? rp.slot - back : rp.slot + (unsigned) p.nslot - back;
    const unsigned par = rp.slot >= back ? rp.par : rp.par ^ 1u;
    mbar_wait(sh, p, sh.full + slot * 8, par, 5);
}
__device__ void producer(const Shared& sh, const MegaParams& p) {
    RingPos rp{0u, 0u};
    const int nph = 4 * p.layers_run;
    int sp = 0;
    unsigned issued = 0;
#pragma unroll 1
    for (int ph = 4 * p.l_begin; ph <= nph; ++ph) {
        const int l = ph >> 2, k = ph == nph ? 4 : (ph & 3);
        if (k == 1) { // the layer's KV tiles come between QKV and WO
            if (p.l2_ahead) prefetch_subphase(p, sp + 1);
            ++sp;
            const AttnSplit a = attn_split(p);
#pragma unroll 1
            for (int p0 = a.p_lo; p0 < a.p_hi; p0 += kChunk, rp.next(p.nslot)) {
                const int cnt = min(kChunk, a.p_hi - p0);
                const unsigned slot = rp.slot, par = rp.par;
                mbar_wait(sh, p, sh.empty + slot * 8, par ^ 1, 2);
                producer_throttle(sh, p, rp, issued++);
                const size_t off = (((size_t) l * p.KVHl + a.kvh) * p.S + p0) * 128;
                const uint32_t bytes = (uint32_t) cnt * 512u;
                const uint32_t dst = smem_u32(sh.ring + (size_t) slot * kSlotBytes);
                mbar_expect_tx(sh.full + slot * 8, 2 * bytes);
                bulk_g2s(dst, p.k_cache + off, bytes, sh.full + slot * 8);
                bulk_g2s(dst + kChunk * 512, p.v_cache + off, bytes, sh.full + slot * 8);
            }
        }
        if (p.l2_ahead) prefetch_subphase(p, sp + 1);
        ++sp;
        const MatDesc& m = p.mat[k];
        int r0, r1;
        cta_rows(m, p.perm, r0, r1);
        const size_t rb = qw_row_bytes(m.n);
        const uint8_t* base = m.base + (size_t) l * m.stride;
#pragma unroll 1
        for (int r = r0; r < r1; r += m.rt, rp.next(p.nslot)) {
            const int nr = min(m.rt, r1 - r);
            const unsigned slot = rp.slot, par = rp.par;
            mbar_wait(sh, p, sh.empty + slot * 8, par ^ 1, 1);
            producer_throttle(sh, p, rp, issued++);
            const uint32_t bytes = (uint32_t) (nr * rb);
            mbar_expect_tx(sh.full + slot * 8, bytes);
            bulk_g2s(smem_u32(sh.ring + (size_t) slot * kSlotBytes), base + (size_t) r * rb, bytes, sh.full + slot * 8);
        }
    }
}

// ---------------------------------------------------------------- consumer: GEMV over ring tiles
// Work unit = 2 consecutive rows, owned by ONE warp (unit u of the CTA's range -> warp u % 15).
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
// kind 0: out[row] = v      kind 1: out[row] = resid[row] + v      kind 2: out[row/2] = silu(v0) * v1
// kind 3 (tensor parallel, wo / w2): v is this rank's partial sum over its column window; lane 2q + j stores row j
// of the unit into slot tp_rank of rank q's arena at BYTE offset `out` (st.relaxed.sys over NVLink; the
// own rank is one of the q). That IS the all-reduce: the readers add the tp slots in rank order (prologue_quant).
// `out` is a flow-arena vector (or the logits): each element is stored exactly once.
// one Q8_0 group (64 codes + scale at group index G of the record layout) of TWO weight rows against the activation
// vector: two dp4a chains of 8 per row (dp4a: ~24 cycles dependent latency), then ((float) dot * ws) * xs (forward.c:94-96)
__device__ __forceinline__ void gemv_step2(const uint8_t* xq, const uint8_t* rowa, const uint8_t* rowb, int G, int rot, float& acca, float& accb) {
    const int off = (G >> 2) * QW_SG_BYTES + (G & 3) * 64;
    int da0 = 0, da1 = 0, db0 = 0, db1 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pc = ((i + rot) & 3) * 16;
        const int4 x = *reinterpret_cast<const int4*>(xq + off + pc);
        const int4 a = *reinterpret_cast<const int4*>(rowa + off + pc);
        const int4 b = *reinterpret_cast<const int4*>(rowb + off + pc);
        if (i < 2) {
            da0 = __dp4a(a.x, x.x, da0); db0 = __dp4a(b.x, x.x, db0); da0 = __dp4a(a.y, x.y, da0); db0 = __dp4a(b.y, x.y, db0);
            da0 = __dp4a(a.z, x.z, da0); db0 = __dp4a(b.z, x.z, db0); da0 = __dp4a(a.w, x.w, da0); db0 = __dp4a(b.w, x.w, db0);
        } else {
            da1 = __dp4a(a.x, x.x, da1); db1 = __dp4a(b.x, x.x, db1); da1 = __dp4a(a.y, x.y, da1); db1 = __dp4a(b.y, x.y, db1);
            da1 = __dp4a(a.z, x.z, da1); db1 = __dp4a(b.z, x.z, db1); da1 = __dp4a(a.w, x.w, da1); db1 = __dp4a(b.w, x.w, db1);
        }
    }
    const int so = (G >> 2) * QW_SG_BYTES + 256 + (G & 3) * 4;
    const float xs = *reinterpret_cast<const float*>(xq + so);
    acca = __fadd_rn(acca, q8_term(da0 + da1, *reinterpret_cast<const float*>(rowa + so), xs));
    accb = __fadd_rn(accb, q8_term(db0 + db1, *reinterpret_cast<const float*>(rowb + so), xs));
}
// the same for ONE row: four chains of 4
__device__ __forceinline__ float gemv_step1(const uint8_t* xq, const uint8_t* row, int G, int rot) {
    const int off = (G >> 2) * QW_SG_BYTES + (G & 3) * 64;
    int d[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pc = ((i + rot) & 3) * 16;
        const int4 x = *reinterpret_cast<const int4*>(xq + off + pc);
        const int4 a = *reinterpret_cast<const int4*>(row + off + pc);
        d[i] = __dp4a(a.x, x.x, d[i]); d[i] = __dp4a(a.y, x.y, d[i]); d[i] = __dp4a(a.z, x.z, d[i]); d[i] = __dp4a(a.w, x.w, d[i]);
    }
    const int so = (G >> 2) * QW_SG_BYTES + 256 + (G & 3) * 4;
    return q8_term((d[0] + d[1]) + (d[2] + d[3]), *reinterpret_cast<const float*>(row + so), *reinterpret_cast<const float*>(xq + so));
}

template <bool TP>
__device__ __forceinline__ void consume_mat(const Shared& sh, const MegaParams& p, const MatDesc& m, int /*layer*/, RingPos& rp, float* out, const float* resid) {
    int r0, r1;
    cta_rows(m, p.perm, r0, r1);
    const int rt = m.rt;
    const int KIND = m.kind;
    const int sgpr = qw_sg_per_row(m.n);
    const size_t rb = (size_t) sgpr * QW_SG_BYTES;
    const int groups = sgpr * 4; // padded groups carry zero codes and zero scales: they add +0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rot = (lane & 2);
    // column split of a row over the lanes: nfull steps of 32 groups (lane = group), then a tail of rem groups
    const int nfull = groups >> 5, rem = groups & 31;
    const bool tail_split = rem > 0 && rem <= 16;
    const bool t_rowb = tail_split && lane >= rem;
    const int tG = rem == 0 ? -1 : tail_split ? (lane < 2 * rem ? nfull * 32 + (t_rowb ? lane - rem : lane) : -1) : (lane < rem ? nfull * 32 + lane : -1);
    const int nrows = r1 - r0;
    const int upt = (rt + 1) / 2;             // units per tile (rt is 1 only when a row fills the slot)
    const int rows_pu = rt >= 2 ? 2 : 1;      // rows per unit
    const int total = (nrows + rows_pu - 1) / rows_pu;
    const bool staged = m.stage != 0;
    float* stage = sh.scr;                    // free during the GEMV phases (attention / combine scratch)
    int u = warp; // this warp's next unit
    float pa0 = 0.0f, pb0 = 0.0f, xres = 0.0f; // lane partials of a unit whose reduction waits for the next unit
    int grow0 = 0;
    bool two0 = false, pend = false;
#ifdef QW_UNITPROF
    long long up_wait = 0, up_math = 0, up_epi = 0, up_units = 0, up_t, up_bar = 0, up_store = 0;
    const long long up_entry = clock64();
#define UP_BEGIN() up_t = clock64()
#define UP_END(acc) do { const long long up_n = clock64(); acc += up_n - up_t; up_t = up_n; } while (0)
#else
#define UP_BEGIN()
#define UP_END(acc)
#endif
#pragma unroll 1
    for (int t0 = 0; t0 < total; t0 += upt, rp.next(p.nslot)) {
        const int t1 = min(t0 + upt, total);
        // Every warp visits every tile, owner of a unit or not: a warp that skipped tiles could no longer tell the phases of
        // a slot's mbarrier apart (parity waits alias once a waiter is more than one phase away) -- tried, it reads tiles
        // that have not landed. The visit (wait, __syncwarp, arrive) costs ~400 cycles: 28 % of the w1/w3 phase.
        const unsigned slot = rp.slot, par = rp.par;
        UP_BEGIN();
        mbar_wait(sh, p, sh.full + slot * 8, par, 3);
        UP_END(up_wait);
        const uint8_t* tile = sh.ring + (size_t) slot * kSlotBytes;
        bool released = false;
        if (p.dbg_mode >= 1) // ring throughput test: skip the math
            while (u < t1) u += kConsumerWarps;
#pragma unroll 1
        for (; u < t1; u += kConsumerWarps) { // this warp's units u = warp, warp + 15, ... fall into the tiles in order
            const int lr = (u - t0) * rows_pu;          // first row of the unit inside the tile
            const int grow = r0 + u * rows_pu;          // its global row
            const bool two = rows_pu == 2 && grow + 1 < r1;
            const uint8_t* rowa = tile + (size_t) lr * rb;
            const uint8_t* rowb = two ? rowa + rb : rowa;
            // residual element of the row this lane will store (see the epilogue): lane 16 * (unit of the pair) + 8 * (row
            // of the unit); issued early, it hides the L2 round trip
            if (KIND == 1 && (lane & 7) == 0 && ((lane & 16) != 0) == pend && ((lane & 8) == 0 || two))
                xres = __uint_as_float(ldf_u32(resid + grow + ((lane >> 3) & 1)));
            float acca = 0.0f, accb = 0.0f;
            int s = 0;
#pragma unroll 1
            for (; s + 2 <= nfull; s += 2) { // two full 32-group steps at a time: four dp4a chains of 16
                const int G = s * 32 + lane, G2 = G + 32;
                const int off = (G >> 2) * QW_SG_BYTES + (G & 3) * 64;
                const int off2 = (G2 >> 2) * QW_SG_BYTES + (G2 & 3) * 64;
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
                const int so = (G >> 2) * QW_SG_BYTES + 256 + (G & 3) * 4;
                const int so2 = (G2 >> 2) * QW_SG_BYTES + 256 + (G2 & 3) * 4;
                const float xs0 = *reinterpret_cast<const float*>(sh.xq + so);
                const float xs1 = *reinterpret_cast<const float*>(sh.xq + so2);
                acca = __fadd_rn(acca, q8_term(da0, *reinterpret_cast<const float*>(rowa + so), xs0));
                accb = __fadd_rn(accb, q8_term(db0, *reinterpret_cast<const float*>(rowb + so), xs0));
                acca = __fadd_rn(acca, q8_term(da1, *reinterpret_cast<const float*>(rowa + so2), xs1));
                accb = __fadd_rn(accb, q8_term(db1, *reinterpret_cast<const float*>(rowb + so2), xs1));
            }
            // an odd full step, then the tail of rem < 32 groups when more than 16 lanes have one: one group per lane for
            // both rows (real branches: lanes without a group issue no shared-memory wavefronts; one copy of the code)
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const int G = k == 0 ? (s < nfull ? s * 32 + lane : -1) : (tail_split ? -1 : tG);
                if (G >= 0) gemv_step2(sh.xq, rowa, rowb, G, rot, acca, accb);
            }
            // a tail of <= 16 groups: lane l < rem takes the group for row a, lane rem + l for row b (n = 2560: 8 + 8 lanes
            // instead of a second full-width step of which 24 lanes only reloaded their first group: 37 % fewer wavefronts)
            if (tail_split && tG >= 0) {
                const float t = gemv_step1(sh.xq, t_rowb ? rowb : rowa, tG, rot);
                if (t_rowb) accb = __fadd_rn(accb, t); else acca = __fadd_rn(acca, t);
            }
            UP_END(up_math);
            if (u + kConsumerWarps >= t1) {
                // this warp's last unit in the tile: every byte it needs from the ring slot has been consumed by a
                // dp4a, so the slot goes back to the producer BEFORE the cross-lane reduction and the epilogue
                __syncwarp();
                if (lane == 0) mbar_arrive(sh.empty + slot * 8);
                released = true;
            }
            // Cross-lane reduction and epilogue, for TWO units (4 rows) at a time: the first unit's lane partials wait in
            // registers (measured: the 10-shuffle tree + epilogue per unit cost 580-710 cycles against 1200 for the dot
            // products). A transposing butterfly folds 4 values per lane in 6 shuffles: after xor 16 / xor 8 a lane keeps
            // row 2 * (lane >> 4) + ((lane >> 3) & 1), after xor 4, 2, 1 all 8 lanes of that row's group hold its sum.
            const bool last_unit = u + kConsumerWarps >= total;
            if (!pend && !last_unit) {
                pa0 = acca; pb0 = accb; grow0 = grow; two0 = two;
                pend = true;
            } else {
                // rows: 0, 1 = the pending unit (or this one when nothing is pending), 2, 3 = this unit after a pending one
                const float v0 = pend ? pa0 : acca, v1 = pend ? pb0 : accb, v2 = pend ? acca : 0.0f, v3 = pend ? accb : 0.0f;
                const bool hi = (lane & 16) != 0, mid = (lane & 8) != 0;
                float t0 = __fadd_rn(hi ? v2 : v0, __shfl_xor_sync(0xffffffffu, hi ? v0 : v2, 16));
                float t1 = __fadd_rn(hi ? v3 : v1, __shfl_xor_sync(0xffffffffu, hi ? v1 : v3, 16));
                float r = __fadd_rn(mid ? t1 : t0, __shfl_xor_sync(0xffffffffu, mid ? t0 : t1, 8));
                r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
                r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
                r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
                // this lane's row: unit `hi` of the pair (0 = the older one), row `mid` of the unit
                const int ug = (pend && !hi) ? grow0 : grow;              // first row of the lane's unit
                const bool utwo = (pend && !hi) ? two0 : two;
                const bool live = (!hi || pend) && (!mid || utwo);        // the lane's row exists
                if (KIND == 2) {
                    const float gate = __shfl_xor_sync(0xffffffffu, r, 8); // the w3 row of the pair sits 8 lanes up
                    if ((lane & 15) == 0 && live) {
                        const float hv = __fmul_rn(silu_ref(r), gate);
                        if (staged) stage[(ug - r0) >> 1] = hv; else stf_f32(out + (ug >> 1), hv);
                    }
                } else if (TP && KIND == 3 && staged) {
                    if ((lane & 7) == 0 && live) stage[ug + (mid ? 1 : 0) - r0] = r;
                } else if (TP && KIND == 3) {
                    if ((lane & 7) < p.tp && live) { // lane q of the row's group of 8 stores to rank q
                        char* base = reinterpret_cast<char*>(p.peer_flow[0]);
#pragma unroll
                        for (int q = 1; q < kMaxTp; ++q) // static indices: a dynamic one makes a local copy of the parameter array
                            if ((lane & 7) == q) base = reinterpret_cast<char*>(p.peer_flow[q]);
                        stf_sys_f32(reinterpret_cast<float*>(base + reinterpret_cast<size_t>(out)) + ug + (mid ? 1 : 0), r);
                    }
                } else if ((lane & 7) == 0 && live) {
                    const float o = KIND == 1 ? __fadd_rn(xres, r) : r;
                    if (staged) stage[ug + (mid ? 1 : 0) - r0] = o; else stf_f32(out + ug + (mid ? 1 : 0), o);
                }
                pend = false;
            }
            UP_END(up_epi);
#ifdef QW_UNITPROF
            ++up_units;
#endif
        }
        if (!released) { // no unit of this warp in the tile
            __syncwarp();
            if (lane == 0) mbar_arrive(sh.empty + slot * 8);
        }
    }
#ifdef QW_UNITPROF
    const long long up_loop_end = clock64();
#endif
    // push this warp's results out NOW: without a fence the stores sit in the SM's write path for
    // microseconds (measured, scripts/ubench/handoff.cu: 2.7 us per hand-off without, 1.05 us with)
    if (staged) {
        // The CTA's rows are contiguous in the output vector: once every warp has put its results into shared memory, one
        // thread sends them with a single TMA bulk store. The async proxy writes straight to L2 -- no SM store queue shared
        // with the polling loads, and no warp waits in a fence (13 % of all stall samples in ncu with per-warp stores + fence).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        UP_BEGIN();
        bar_consumers();
        UP_END(up_bar);
        if (TP && KIND == 3) {
            // tensor parallel: one bulk store per rank, issued by threads 0 .. tp-1, straight into slot tp_rank of that
            // rank's arena over NVLink (`out` is the byte offset of the slot) -- the sending half of the fused all-reduce
            if ((int) threadIdx.x < p.tp && nrows > 0) {
                char* base = reinterpret_cast<char*>(p.peer_flow[0]);
#pragma unroll
                for (int q = 1; q < kMaxTp; ++q) // static indices: a dynamic one makes a local copy of the parameter array
                    if ((int) threadIdx.x == q) base = reinterpret_cast<char*>(p.peer_flow[q]);
                float* dst = reinterpret_cast<float*>(base + reinterpret_cast<size_t>(out)) + r0;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stage)), "r"(nrows * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else if (threadIdx.x == 0 && nrows > 0) {
            const int nout = KIND == 2 ? nrows >> 1 : nrows;
            float* dst = out + (KIND == 2 ? r0 >> 1 : r0);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stage)), "r"(nout * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the staging area may be rewritten after the next CTA barrier
        }
    } else if (TP && KIND == 3) {
        __threadfence_system();
        bar_consumers(); // the next prologue rewrites sh.xq: no warp may still be reading it (the staged branch has its barrier)
    } else {
        flush_stores();
        bar_consumers();
    }
#ifdef QW_UNITPROF
    UP_END(up_store);
    if (p.prof && lane == 0) {
        const int mi = (int) (&m - p.mat);
        unsigned long long* up = p.prof + (size_t) gridDim.x * (p.L + 1) * kProfSlots + ((size_t) (blockIdx.x * 16 + warp) * 5 + mi) * 8;
        up[0] += up_wait; up[1] += up_math; up[2] += up_epi; up[3] += up_units; up[4] += up_bar; up[5] += up_store;
        up[6] += up_loop_end - up_entry - up_wait - up_math - up_epi;
    }
#endif
}

// ---------------------------------------------------------------- consumer: prologues
// exact Q8_0 codes of 8 values (reference q8.c:27-28), packed; out of line: it runs for ~1 value in 500
__device__ __noinline__ uint2 q8_pack8_exact(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, float scale) {
    const uint32_t w0 = (uint32_t) (q8_code(v0, scale) & 0xff) | (uint32_t) (q8_code(v1, scale) & 0xff) << 8
                        | (uint32_t) (q8_code(v2, scale) & 0xff) << 16 | (uint32_t) (q8_code(v3, scale) & 0xff) << 24;
    const uint32_t w1 = (uint32_t) (q8_code(v4, scale) & 0xff) | (uint32_t) (q8_code(v5, scale) & 0xff) << 8
                        | (uint32_t) (q8_code(v6, scale) & 0xff) << 16 | (uint32_t) (q8_code(v7, scale) & 0xff) << 24;
    return make_uint2(w0, w1);
}

// One warp quantises one 256-column record: lane owns 8 consecutive values v[0..7] (group = lane / 8
// of the record). Group absmax by 3 shuffles; scale = absmax / 127 with the reference's IEEE division.
// Codes: the reference computes roundf(x / scale) (IEEE division, half away from zero, clamp). Here
// t = x * rcp(scale) is within 2^-21 relative of the true quotient (|t| <= 127.01), so rintf(t) is the
// reference's code unless t lies within 1e-3 of a rounding boundary k + 0.5 -- then (and for NaN or a
// scale too small to invert) all 8 values take the exact path. Packed codes go out with one 8-byte STS
// per lane. live = the lane's group exists (pad groups get zero codes, zero scale).
__device__ __forceinline__ void quant_record(uint8_t* xq, int rec, int lane, const float (&v)[8], bool live) {
    float amax = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) amax = fmaxf(amax, fabsf(v[i]));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
    const float scale = q8_scale(amax);
    const float rinv = __frcp_rn(scale);
    bool bad = !(rinv <= 3.0e38f);
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float t = __fmul_rn(v[i], rinv);
        const float r = rintf(t);
        bad |= !(fabsf(__fsub_rn(t, r)) <= 0.499f);
        const uint32_t c = (uint32_t) ((int) r & 0xff) << (8 * (i & 3));
        if (i < 4) w0 |= c; else w1 |= c;
    }
    uint2 w = make_uint2(w0, w1);
    if (bad) w = q8_pack8_exact(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], scale);
    uint8_t* r = xq + (size_t) rec * QW_SG_BYTES;
    *reinterpret_cast<uint2*>(r + lane * 8) = live ? w : make_uint2(0u, 0u);
    if ((lane & 7) == 0) *reinterpret_cast<float*>(r + 256 + (lane >> 3) * 4) = live ? scale : 0.0f;
}

constexpr int kRecBatch = 3; // records per warp and pass: 45 records = 11520 columns per pass

// poll until the 4 words at q (written by a peer GPU) are all there
__device__ __noinline__ uint4 poll4_sys_slow(volatile int* abort_flag, int* err, const void* q, int code) {
    unsigned long long t0 = 0;
    uint4 v = ldf_sys_u4(q);
    for (unsigned n = 1; unset4(v); ++n) {
        if ((n & 255u) == 0) {
            if (*abort_flag) break;
            const unsigned long long now = gtime_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTimeoutNs) {
                *abort_flag = code;
                *err = code;
                break;
            }
        }
        v = ldf_sys_u4(q);
    }
    return v;
}

// Receiving half of the fused all-reduce (tensor parallel): x += part[0] + ... + part[nparts-1] for the elements of
// the residual stream this thread owns (records warp, warp + 15, ...; 8 columns per lane), in rank order so that all
// ranks hold bit-identical x. The partials were pushed into this rank's arena by every rank's wo / w2 epilogue.
// Out of line: its registers must not add to the pressure of the kernel's main body.
__device__ __noinline__ void tp_gather_x(float* xres, const float* src, int nparts, int part_stride, int recs, int groups,
                                         volatile int* abort_flag, int* err) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int rec = warp; rec < recs; rec += kConsumerWarps) {
        if (rec * 4 + (lane >> 3) >= groups) continue;
        const int col = rec * 256 + lane * 8;
        float4* xr = reinterpret_cast<float4*>(xres + col);
        float4 x0 = xr[0], x1 = xr[1];
#pragma unroll 1
        for (int q0 = 0; q0 < nparts; q0 += 4) { // four ranks' partials in flight at once
            uint4 a[4], b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (q0 + j < nparts) {
                    const float* q = src + (size_t) (q0 + j) * part_stride + col;
                    a[j] = ldf_sys_u4(q);
                    b[j] = ldf_sys_u4(q + 4);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (q0 + j < nparts) {
                    const float* q = src + (size_t) (q0 + j) * part_stride + col;
                    if (unset4(a[j])) a[j] = poll4_sys_slow(abort_flag, err, q, 15);
                    if (unset4(b[j])) b[j] = poll4_sys_slow(abort_flag, err, q + 4, 15);
                    x0.x = __fadd_rn(x0.x, __uint_as_float(a[j].x)); x0.y = __fadd_rn(x0.y, __uint_as_float(a[j].y));
                    x0.z = __fadd_rn(x0.z, __uint_as_float(a[j].z)); x0.w = __fadd_rn(x0.w, __uint_as_float(a[j].w));
                    x1.x = __fadd_rn(x1.x, __uint_as_float(b[j].x)); x1.y = __fadd_rn(x1.y, __uint_as_float(b[j].y));
                    x1.z = __fadd_rn(x1.z, __uint_as_float(b[j].z)); x1.w = __fadd_rn(x1.w, __uint_as_float(b[j].w));
                }
        }
        xr[0] = x0;
        xr[1] = x1;
    }
}

// fp32 flow vector src[n] (written by other CTAs, polled) -> optional RMSNorm with weights nw
// (forward.c:254-259; nw == nullptr: none) -> Q8_0 codes + scales in shared memory (q8.c:5-30).
// Warp w owns records w, w + 15, w + 30 of a pass; all loads of a pass are issued up front (one L2
// round trip). With RMSNorm the vector must fit one pass (checked at init).
// nparts >= 0 (tensor parallel, n == D): the vector is the residual stream, which every CTA keeps in shared
// memory (sh.xres, each element owned by the thread that quantises it): x += part[0] + ... + part[nparts-1],
// the partial sums every rank pushed into this rank's arena (src + q * part_stride), added in rank order so
// that all ranks hold bit-identical x. This is the receiving half of the fused all-reduce (forward.c:295-298,
// 335-338 are the residual adds it replaces).
template <bool TP>
__device__ __forceinline__ void prologue_quant(const Shared& sh, const MegaParams& p, const float* src, int n, const float* __restrict__ nw, int nparts) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = n / 64, recs = qw_sg_per_row(n);
#pragma unroll 1
    for (int pass0 = 0; pass0 < recs; pass0 += kRecBatch * kConsumerWarps) { // uniform over the CTA: it contains a barrier
        const int r0 = pass0 + warp;
        float v[kRecBatch][8];
        bool live[kRecBatch];
#pragma unroll
        for (int k = 0; k < kRecBatch; ++k) {
            const int rec = r0 + k * kConsumerWarps;
            live[k] = rec < recs && rec * 4 + (lane >> 3) < groups;
        }
        if (!TP || nparts < 0) {
            uint4 a[kRecBatch], b[kRecBatch];
#pragma unroll
            for (int k = 0; k < kRecBatch; ++k) {
                const int rec = r0 + k * kConsumerWarps;
                if (live[k]) {
                    a[k] = ldf_u4(src + rec * 256 + lane * 8);
                    b[k] = ldf_u4(src + rec * 256 + lane * 8 + 4);
                }
            }
            // poll: re-issue ALL still-unset pieces together each round (one L2 round trip per round). Polling them one
            // after the other cost up to six serial round trips after the data had landed, because every piece's first
            // load was issued before its producer had stored.
            {
                unsigned long long t_start = 0;
#pragma unroll 1
                for (unsigned n = 1;; ++n) {
                    bool busy = false;
#pragma unroll
                    for (int k = 0; k < kRecBatch; ++k)
                        if (live[k]) busy |= unset4(a[k]) | unset4(b[k]);
                    if (!busy) break;
                    if ((n & 255u) == 0 && poll_timed_out(sh.abort_flag, p.err, t_start, 10)) break;
#pragma unroll
                    for (int k = 0; k < kRecBatch; ++k) {
                        const float* q = src + (r0 + k * kConsumerWarps) * 256 + lane * 8;
                        if (live[k] && unset4(a[k])) a[k] = ldf_u4(q);
                        if (live[k] && unset4(b[k])) b[k] = ldf_u4(q + 4);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kRecBatch; ++k) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k][i] = 0.0f;
                if (live[k]) {
                    v[k][0] = __uint_as_float(a[k].x); v[k][1] = __uint_as_float(a[k].y);
                    v[k][2] = __uint_as_float(a[k].z); v[k][3] = __uint_as_float(a[k].w);
                    v[k][4] = __uint_as_float(b[k].x); v[k][5] = __uint_as_float(b[k].y);
                    v[k][6] = __uint_as_float(b[k].z); v[k][7] = __uint_as_float(b[k].w);
                }
            }
        } else {
            if (nparts > 0) tp_gather_x(sh.xres, src, nparts, p.part_stride, recs, groups, sh.abort_flag, p.err);
#pragma unroll
            for (int k = 0; k < kRecBatch; ++k) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k][i] = 0.0f;
                if (!live[k]) continue;
                const int col = (r0 + k * kConsumerWarps) * 256 + lane * 8;
                const float4* xr = reinterpret_cast<const float4*>(sh.xres + col);
                const float4 x0 = xr[0], x1 = xr[1];
                v[k][0] = x0.x; v[k][1] = x0.y; v[k][2] = x0.z; v[k][3] = x0.w;
                v[k][4] = x1.x; v[k][5] = x1.y; v[k][6] = x1.z; v[k][7] = x1.w;
            }
        }
        if (nw) {
            float ss = 0.0f;
#pragma unroll
            for (int k = 0; k < kRecBatch; ++k)
#pragma unroll
                for (int i = 0; i < 8; ++i) ss = __fmaf_rn(v[k][i], v[k][i], ss);
            ss = warp_sum(ss);
            if (lane == 0) sh.misc[warp] = ss;
            bar_consumers();
            float tot = 0.0f;
#pragma unroll
            for (int i = 0; i < kConsumerWarps; ++i) tot = __fadd_rn(tot, sh.misc[i]);
            const float r = rms_rscale(tot, n);
#pragma unroll
            for (int k = 0; k < kRecBatch; ++k) {
                if (!live[k]) continue;
                const float* g = nw + (r0 + k * kConsumerWarps) * 256 + lane * 8;
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(g)), g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
                v[k][0] = __fmul_rn(g0.x, __fmul_rn(r, v[k][0])); v[k][1] = __fmul_rn(g0.y, __fmul_rn(r, v[k][1]));
                v[k][2] = __fmul_rn(g0.z, __fmul_rn(r, v[k][2])); v[k][3] = __fmul_rn(g0.w, __fmul_rn(r, v[k][3]));
                v[k][4] = __fmul_rn(g1.x, __fmul_rn(r, v[k][4])); v[k][5] = __fmul_rn(g1.y, __fmul_rn(r, v[k][5]));
                v[k][6] = __fmul_rn(g1.z, __fmul_rn(r, v[k][6])); v[k][7] = __fmul_rn(g1.w, __fmul_rn(r, v[k][7]));
            }
        }
#pragma unroll
        for (int k = 0; k < kRecBatch; ++k) {
            const int rec = r0 + k * kConsumerWarps;
            if (rec < recs) quant_record(sh.xq, rec, lane, v[k], live[k]);
        }
    }
    bar_consumers(); // xq complete; also protects sh.misc against the next prologue
}

// already-quantised flow vector (attention output, SG layout: codes and scales, polled) -> shared.
// Words of pad groups (columns >= n) are never written by anyone: they are zero here.
__device__ void prologue_load_codes(const Shared& sh, const MegaParams& p, const uint8_t* q, int n) {
    const int pieces = (int) (qw_row_bytes(n) / 16), groups = n / 64;
    for (int i = threadIdx.x; i < pieces; i += kConsumerThreads) {
        const int rec = i / 17, k = i % 17; // 17 x 16 B per record: 16 code pieces, then the 4 scales
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (k < 16) {
            if (rec * 4 + (k >> 2) < groups) v = poll4(sh, p, q + (size_t) i * 16, 12);
        } else if (rec * 4 + 3 < groups) {
            v = poll4(sh, p, q + (size_t) i * 16, 12);
        } else { // a record with pad groups: poll only the live scales
            uint32_t s[4] = {0u, 0u, 0u, 0u};
            for (int g = 0; g < 4; ++g)
                if (rec * 4 + g < groups) s[g] = __float_as_uint(poll1(sh, p, reinterpret_cast<const float*>(q + (size_t) i * 16) + g, 12));
            v = make_uint4(s[0], s[1], s[2], s[3]);
        }
        reinterpret_cast<uint4*>(sh.xq)[i] = v;
    }
    bar_consumers();
}

// ---------------------------------------------------------------- consumer: attention
// RMSNorm weight + RoPE for a 128-wide head (forward.c:267-280, 104-118); cos/sin come from the
// host-computed table so the angles are the reference's bit for bit. One warp per head: lane owns
// the pairs (lane, lane + 64) and (lane + 32, lane + 96).
__device__ __forceinline__ void head_norm_rope_warp(float* dst, const float x[4], const float* g, const MegaParams& p, int lane) {
    // x[0..3] = raw[lane], raw[lane + 32], raw[lane + 64], raw[lane + 96]
    float ss = __fmul_rn(x[0], x[0]);
    ss = __fmaf_rn(x[1], x[1], ss);
    ss = __fmaf_rn(x[2], x[2], ss);
    ss = __fmaf_rn(x[3], x[3], ss);
    const float r = rms_rscale(warp_sum(ss), 128);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = lane + 32 * k; // 0..63
        const float c = __ldg(p.rope_cos + (size_t) p.pos * 64 + j), s = __ldg(p.rope_sin + (size_t) p.pos * 64 + j);
        const float a = __fmul_rn(__ldg(g + j), __fmul_rn(r, x[k]));
        const float b = __fmul_rn(__ldg(g + j + 64), __fmul_rn(r, x[2 + k]));
        dst[j] = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, s));
        dst[j + 64] = __fadd_rn(__fmul_rn(a, s), __fmul_rn(b, c));
    }
}

// merge another warp's dumped state (shared memory, [HW][kPartStride]) into this warp's
template <int HW>
__device__ __forceinline__ void attn_merge_in(AttnState<HW>& st, const float* slot, int lane) {
#pragma unroll
    for (int j = 0; j < HW; ++j) {
        const float m2 = slot[j * kPartStride + 128], l2 = slot[j * kPartStride + 129];
        const float4 a2 = *reinterpret_cast<const float4*>(slot + j * kPartStride + lane * 4);
        const float M = fmaxf(st.m[j], m2);
        const float wa = (st.m[j] == -INFINITY) ? 0.0f : expf(__fsub_rn(st.m[j], M));
        const float wb = (m2 == -INFINITY) ? 0.0f : expf(__fsub_rn(m2, M));
        st.acc[j].x = __fmaf_rn(a2.x, wb, __fmul_rn(st.acc[j].x, wa));
        st.acc[j].y = __fmaf_rn(a2.y, wb, __fmul_rn(st.acc[j].y, wa));
        st.acc[j].z = __fmaf_rn(a2.z, wb, __fmul_rn(st.acc[j].z, wa));
        st.acc[j].w = __fmaf_rn(a2.w, wb, __fmul_rn(st.acc[j].w, wa));
        st.l[j] = __fmaf_rn(l2, wb, __fmul_rn(st.l[j], wa));
        st.m[j] = M;
    }
}
template <int HW>
__device__ __forceinline__ void attn_dump(const AttnState<HW>& st, float* slot, int lane) {
#pragma unroll
    for (int j = 0; j < HW; ++j) {
        *reinterpret_cast<float4*>(slot + j * kPartStride + lane * 4) = st.acc[j];
        if (lane == 0) {
            slot[j * kPartStride + 128] = st.m[j];
            slot[j * kPartStride + 129] = st.l[j];
        }
    }
}

// shared scratch (floats, inside sh.scr): the segment's heads sq[KV_MUL*128 q | 128 k | 128 v]; after the
// tiles the same memory holds up to 8 dumped warp states for the merge tree.
constexpr int kScrFloats = 8 * 4 * kPartStride;

// Split-KV attention over this CTA's positions of its kv head (attn_split). The query heads (and, on the kv
// head's last block, this step's own K/V row: RMSNorm + RoPE, written to the cache for later steps) are
// prepared first, one warp per head. Warp w
// serves head group w % NHG; the warps of a head group split the segment's cached positions into
// equal contiguous ranges (a range may straddle two tiles), so the load is balanced at any context
// length. Every warp walks every tile for the ring protocol.
template <int KV_MUL>
__device__ void consume_attn(const Shared& sh, const MegaParams& p, int l, RingPos& rp) {
    constexpr int HW = KV_MUL < 4 ? KV_MUL : 4; // heads per warp
    constexpr int NHG = KV_MUL / HW;            // head groups per KV head
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sq = sh.scr;
    float* fl = flow_layer(p, l);
    const float* qkv = fl + p.o_qkv;
    const float* gq = p.q_norm + (size_t) l * 128;
    const float* gk = p.k_norm + (size_t) l * 128;
    const AttnSplit as = attn_split(p);
    if (!as.active) return; // this block has no share of the attention (no tiles, no partial expected from it)
    const int kvh = as.kvh, my_slot = as.j;
    const int npos = as.p_hi - as.p_lo;                          // cached positions of this block
    const int ntile = (npos + kChunk - 1) / kChunk;
    const bool own_last = as.j == as.n - 1;                      // this block also takes the step's own position
    const int hg = warp % NHG;                                   // this warp's head group
    const int wi = warp / NHG;                                   // its index among the warps of the group
    const int wn = (kConsumerWarps - hg + NHG - 1) / NHG;        // warps in the group

    if (npos == 0 && !own_last) {
        // nothing to attend over here (short context): the kv head's combine tasks still expect this block's (m, l)
        if (warp == 0 && lane < KV_MUL) {
            float* dst = fl + p.o_part + ((size_t) (kvh * p.part_slots + my_slot) * KV_MUL + lane) * kPartStride;
            stf_f32(dst + 128, -INFINITY);
            stf_f32(dst + 129, 0.0f);
            flush_stores();
        }
        return;
    }
    {
        // ---- segment prologue: warp j < KV_MUL prepares query head j; warps KV_MUL, KV_MUL+1 this step's K, V row
        bar_consumers(); // previous users of the scratch are done
        if (warp < KV_MUL + (own_last ? 2 : 0)) {
            const float* src = warp < KV_MUL ? qkv + (size_t) (kvh * KV_MUL + warp) * 128
                               : warp == KV_MUL ? qkv + p.Pl + (size_t) kvh * 128
                                                : qkv + p.Pl + p.Kl + (size_t) kvh * 128;
            float x[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = __uint_as_float(ldf_u32(src + lane + 32 * k));
            {
                unsigned long long t_start = 0;
#pragma unroll 1
                for (unsigned n = 1;; ++n) { // all four words re-polled together (see prologue_quant)
                    bool busy = false;
#pragma unroll
                    for (int k = 0; k < 4; ++k) busy |= __float_as_uint(x[k]) == kSent;
                    if (!busy) break;
                    if ((n & 255u) == 0 && poll_timed_out(sh.abort_flag, p.err, t_start, 13)) break;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (__float_as_uint(x[k]) == kSent) x[k] = __uint_as_float(ldf_u32(src + lane + 32 * k));
                }
            }
            float* dst = sq + warp * 128;
            if (warp <= KV_MUL) {
                head_norm_rope_warp(dst, x, warp < KV_MUL ? gq : gk, p, lane);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = x[k];
            }
            if (warp >= KV_MUL) { // this step's K / V row goes into the cache for later steps
                __syncwarp();
                float* cache = (warp == KV_MUL ? p.k_cache : p.v_cache) + (((size_t) l * p.KVHl + kvh) * p.S + p.pos) * 128;
                *reinterpret_cast<float4*>(cache + lane * 4) = *reinterpret_cast<const float4*>(dst + lane * 4);
            }
        }
        bar_consumers();
        stamp(p, l, 1);
        // ---- tiles, in groups of kGroupTiles: the warps of a head group split the cached positions of a
        // group into equal contiguous ranges (<= 2 tiles each), so all warps work on every group at once and
        // the ring keeps flowing (splitting the whole segment contiguously serialised the warps behind it)
        float4 q[HW];
#pragma unroll
        for (int j = 0; j < HW; ++j) q[j] = *reinterpret_cast<const float4*>(sq + (hg * HW + j) * 128 + lane * 4);
        AttnState<HW> st;
        attn_state_reset(st);
#pragma unroll 1
        for (int g0 = 0; g0 < ntile; g0 += kGroupTiles) {
            const int gt = min(kGroupTiles, ntile - g0);
            const int gn = max(0, min(npos - g0 * kChunk, gt * kChunk));      // cached positions in the group
            const int my0 = gn * wi / wn, my1 = gn * (wi + 1) / wn;           // this warp's, relative to the group
            const float* Ka = nullptr;
            int cnta = 0;
            unsigned prev_slot = 0;
#pragma unroll 1
            for (int t = 0; t < gt; ++t, rp.next(p.nslot)) { // every warp walks every tile, in order
                const unsigned slot = rp.slot;
                mbar_wait(sh, p, sh.full + slot * 8, rp.par, 4);
                const int a = max(my0, t * kChunk) - t * kChunk, b = min(my1, (t + 1) * kChunk) - t * kChunk;
                const float* Kt = reinterpret_cast<const float*>(sh.ring + (size_t) slot * kSlotBytes) + a * 128;
                const bool more = my1 > (t + 1) * kChunk && t + 1 < gt; // my range continues in the next tile
                if (b > a && more) { // first of two tiles: remember it, its slot is released below
                    Ka = Kt;
                    cnta = b - a;
                    prev_slot = slot;
                    continue;
                }
                if (b > a) attn_rows<HW>(Ka ? Ka : Kt, Ka ? cnta : b - a, Kt, (Ka ? cnta : 0) + b - a, kChunk * 128, q, st, lane);
                __syncwarp();
                if (lane == 0) {
                    if (Ka) mbar_arrive(sh.empty + prev_slot * 8);
                    mbar_arrive(sh.empty + slot * 8);
                }
                Ka = nullptr;
            }
        }
        // this step's own position, from shared memory: the last warp of each head group takes it
        if (own_last && wi == wn - 1)
            attn_one_row<HW>(*reinterpret_cast<const float4*>(sq + KV_MUL * 128 + lane * 4),
                             *reinterpret_cast<const float4*>(sq + (KV_MUL + 1) * 128 + lane * 4), q, st);
        // ---- merge the warp states pairwise through shared memory (w <- w + half keeps w % NHG)
        bar_consumers(); // sq is dead from here
        stamp(p, l, 9);
#pragma unroll 1
        for (int half = 8; half >= NHG; half >>= 1) {
            if (warp >= half && warp < 2 * half) attn_dump<HW>(st, sh.scr + (warp - half) * (HW * kPartStride), lane);
            bar_consumers();
            if (warp < half && warp + half < kConsumerWarps) attn_merge_in<HW>(st, sh.scr + warp * (HW * kPartStride), lane);
            bar_consumers();
        }
        // ---- publish (m, l, acc) of this CTA for the segment's heads: the KV_MUL x 132 floats are contiguous in the arena,
        // so they are assembled in shared memory and leave as one TMA bulk store (no per-warp fence; see consume_mat)
        if (warp < NHG) {
            float* dst = sh.scr + (warp * HW) * kPartStride;
#pragma unroll
            for (int j = 0; j < HW; ++j) {
                *reinterpret_cast<float4*>(dst + j * kPartStride + lane * 4) = st.acc[j];
                if (lane == 0) {
                    dst[j * kPartStride + 128] = st.m[j];
                    dst[j * kPartStride + 129] = st.l[j];
                    dst[j * kPartStride + 130] = 0.0f; // pad words: never read
                    dst[j * kPartStride + 131] = 0.0f;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        bar_consumers();
        if (threadIdx.x == 0) {
            float* gdst = fl + p.o_part + ((size_t) (kvh * p.part_slots + my_slot) * KV_MUL) * kPartStride;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(sh.scr)), "r"(KV_MUL * kPartStride * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // combine_attn reuses the scratch after its barrier
        }
    }
}

// Combine (CTA-wide): task t = (head h, half hf) merges the published partials of the head (online-
// softmax merge over the CTAs that attended its KV head), divides by the sum (forward.c:60-75), writes
// the fp32 result (debug read-back) and the Q8_0 group -- 64 codes + scale -- straight into the flow
// vector the wo GEMV loads. Task t runs on CTA t % grid with ALL its warps: thread (dim = tid & 63,
// slot group = tid >> 6) takes slots sg, sg + 7, ...; every load of the task is in flight at once, so
// the task costs one L2 round trip plus a shared-memory reduction. Every block of a KV head publishes
// (m, l) -- blocks without tiles publish (-inf, 0) -- so no slot of the range stays unwritten.
__device__ __forceinline__ void combine_attn(const Shared& sh, const MegaParams& p, int l) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntask = 2 * p.Hl;
    const unsigned G = gridDim.x;
    float* fl = flow_layer(p, l);
    float* mm = sh.scr;        // [<= kMaxGrid] slot maxima, then slot weights
    float* ll = sh.scr + 256;  // [<= kMaxGrid] slot sums
    float* red = sh.scr + 512; // [7][64] partial outputs per slot group
#pragma unroll 1
    for (int t = blockIdx.x; t < ntask; t += G) {
        const int h = t >> 1, hf = t & 1, kvh = h / p.kv_mul;
        const int nb = (kvh + 1) * p.attn_ga / p.KVHl - kvh * p.attn_ga / p.KVHl; // blocks of the kv head: all of them publish
        const float* part = fl + p.o_part + ((size_t) kvh * p.part_slots * p.kv_mul + h % p.kv_mul) * kPartStride;
        const size_t ss = (size_t) p.kv_mul * kPartStride; // between slots
        const int d = hf * 64 + (tid & 63), sg = tid >> 6;
        bar_consumers(); // scratch free (attention merge / previous task)
        // Thread (dim, slot group) needs the accumulator word of slots sg, sg + 7, sg + 14 -- but a slot whose maximum is
        // -inf never gets an accumulator, so the thread polls (m, acc) of each of its slots TOGETHER and is done with a
        // slot once m is there and either m == -inf or acc is there. Everything the task needs is in flight at once and is
        // noticed one round trip after it lands (polling acc only after m had landed cost up to 3 serial round trips).
        float a[3], ms[3];
        bool pend[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pend[k] = sg < 7 && sg + 7 * k < nb;
            a[k] = 0.0f;
            ms[k] = -INFINITY;
        }
        bool pend_ml = tid < nb;
        float my_m = -INFINITY, my_l = 0.0f;
        {
            unsigned long long t_start = 0;
#pragma unroll 1
            for (unsigned n = 1;; ++n) {
                uint32_t wm[3], wa[3], w_m = 0, w_l = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (pend[k]) {
                        wm[k] = ldf_u32(part + (sg + 7 * k) * ss + 128);
                        wa[k] = ldf_u32(part + (sg + 7 * k) * ss + d);
                    }
                if (pend_ml) {
                    w_m = ldf_u32(part + tid * ss + 128);
                    w_l = ldf_u32(part + tid * ss + 129);
                }
                bool busy = false;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (pend[k]) {
                        const float mk = __uint_as_float(wm[k]);
                        if (wm[k] != kSent && (mk == -INFINITY || wa[k] != kSent)) {
                            ms[k] = mk;
                            a[k] = mk == -INFINITY ? 0.0f : __uint_as_float(wa[k]);
                            pend[k] = false;
                        } else {
                            busy = true;
                        }
                    }
                if (pend_ml) {
                    if (w_m != kSent && w_l != kSent) {
                        my_m = __uint_as_float(w_m);
                        my_l = __uint_as_float(w_l);
                        pend_ml = false;
                    } else {
                        busy = true;
                    }
                }
                if (!busy) break;
                if ((n & 255u) == 0) {
                    if (*sh.abort_flag) break;
                    const unsigned long long now = gtime_ns();
                    if (t_start == 0) t_start = now;
                    if (now - t_start > kTimeoutNs) {
                        *sh.abort_flag = 14;
                        *p.err = 14;
                        break;
                    }
                }
            }
        }
        if (tid < nb) {
            mm[tid] = my_m;
            ll[tid] = my_l;
        }
        bar_consumers();
        float M = -INFINITY;
#pragma unroll 1
        for (int i = 0; i < nb; ++i) M = fmaxf(M, mm[i]);
        if (sg < 7) {
            float A = 0.0f;
#pragma unroll 1
            for (int i0 = sg; i0 < nb; i0 += 21) {
                if (i0 != sg) { // slots beyond the first 21 of the kv head (tensor-parallel ranks with 1-2 kv heads): three loads in flight
                    uint32_t w[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int i = i0 + 7 * k;
                        w[k] = (i < nb && mm[i] != -INFINITY) ? ldf_u32(part + i * ss + d) : 0u;
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int i = i0 + 7 * k;
                        if (w[k] == kSent) w[k] = poll1_slow(sh.abort_flag, p.err, part + i * ss + d, 14);
                        a[k] = __uint_as_float(w[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = i0 + 7 * k;
                    if (i >= nb) continue;
                    const float m = mm[i];
                    if (m == -INFINITY) continue; // nothing attended there: its accumulator is not even written
                    A = __fmaf_rn(a[k], expf(__fsub_rn(m, M)), A);
                }
            }
            red[sg * 64 + (tid & 63)] = A;
        }
        bar_consumers();
        if (warp == 0) {
            float A0 = 0.0f, A1 = 0.0f, Ls = 0.0f;
#pragma unroll
            for (int s7 = 0; s7 < 7; ++s7) {
                A0 = __fadd_rn(A0, red[s7 * 64 + lane]);
                A1 = __fadd_rn(A1, red[s7 * 64 + 32 + lane]);
            }
#pragma unroll 1
            for (int i = lane; i < nb; i += 32)
                if (mm[i] != -INFINITY) Ls = __fmaf_rn(ll[i], expf(__fsub_rn(mm[i], M)), Ls);
            Ls = warp_sum(Ls);
            const float o0 = __fdiv_rn(A0, Ls), o1 = __fdiv_rn(A1, Ls);
            const int d0 = hf * 64 + lane;
            p.att[(size_t) h * 128 + d0] = o0;
            p.att[(size_t) h * 128 + d0 + 32] = o1;
            // quantise the group (q8.c:5-30) and store packed code words + scale
            const float scale = q8_scale(warp_max(fmaxf(fabsf(o0), fabsf(o1))));
            const uint32_t c0 = (uint32_t) (q8_code(o0, scale) & 0xff), c1 = (uint32_t) (q8_code(o1, scale) & 0xff);
            uint32_t w0 = c0, w1 = c1;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                w0 |= __shfl_down_sync(0xffffffffu, c0, k) << (8 * k);
                w1 |= __shfl_down_sync(0xffffffffu, c1, k) << (8 * k);
            }
            const int g = h * 2 + hf;
            uint8_t* rec = reinterpret_cast<uint8_t*>(fl + p.o_attq) + (size_t) (g >> 2) * QW_SG_BYTES;
            if ((lane & 3) == 0) {
                stf_u32(rec + (g & 3) * 64 + lane, w0);
                stf_u32(rec + (g & 3) * 64 + 32 + lane, w1);
            }
            if (lane == 0) stf_f32(reinterpret_cast<float*>(rec + 256 + (g & 3) * 4), scale);
            flush_stores(); // flush (see consume_mat)
        }
    }
}

// ---------------------------------------------------------------- consumer main
template <int KV_MUL, bool TP>
__device__ void consumer(const Shared& sh, const MegaParams& p) {
    // refill the OTHER arena with the sentinel for the next launch (nobody reads it during this one)
    {
        const unsigned per = (unsigned) ((p.flow_words / 4 + gridDim.x - 1) / gridDim.x); // 16-byte pieces per CTA
        const unsigned a = per * blockIdx.x, b = min((unsigned) (p.flow_words / 4), a + per);
        uint4* dst = reinterpret_cast<uint4*>(p.flow_other);
        const uint4 s4 = make_uint4(kSent, kSent, kSent, kSent);
#pragma unroll 1
        for (unsigned i = a + threadIdx.x; i < b; i += kConsumerThreads) dst[i] = s4;
        if (TP) __threadfence_system(); // peers write into that arena during the next launch: the refill must be ahead of them
    }
    // the residual stream starts as the dequantised embedding row (forward.c:237): every CTA
    // contributes its slice of it to the flow vector x0
    constexpr bool tp = TP; // a separate instantiation: the single-GPU kernel carries no tensor-parallel code (instruction cache)
    const int tok = p.token_dev ? *p.token_dev : p.token;
    if (!tp) {
        const uint8_t* row = p.w_emb + (size_t) tok * qw_row_bytes(p.D);
        const int c0 = (int) ((unsigned) p.D * blockIdx.x / gridDim.x), c1 = (int) ((unsigned) p.D * (blockIdx.x + 1) / gridDim.x);
#pragma unroll 1
        for (int c = c0 + threadIdx.x; c < c1; c += kConsumerThreads) {
            const uint8_t* rec = row + (size_t) (c >> 8) * QW_SG_BYTES;
            const float sc = *reinterpret_cast<const float*>(rec + 256 + ((c >> 6) & 3) * 4);
            stf_f32(p.flow_x0 + c, p.x_inject ? p.x_inject[c] : __fmul_rn((float) reinterpret_cast<const int8_t*>(rec)[c & 255], sc));
        }
        flush_stores();
    } else {
        // tensor parallel: every CTA keeps the whole residual stream in shared memory; each thread dequantises the
        // elements it owns in prologue_quant (records warp, warp + 15, warp + 30; 8 columns per lane)
        const uint8_t* row = p.w_emb + (size_t) tok * qw_row_bytes(p.D);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int recs = qw_sg_per_row(p.D), groups = p.D / 64;
#pragma unroll 1
        for (int rec = warp; rec < recs; rec += kConsumerWarps) {
            if (rec * 4 + (lane >> 3) >= groups) continue;
            const uint8_t* r = row + (size_t) rec * QW_SG_BYTES;
            const uint2 cw = *reinterpret_cast<const uint2*>(r + lane * 8);
            const float sc = *reinterpret_cast<const float*>(r + 256 + (lane >> 3) * 4);
            float* xr = sh.xres + rec * 256 + lane * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int code = (int) (int8_t) (((i < 4 ? cw.x : cw.y) >> (8 * (i & 3))) & 0xffu);
                xr[i] = p.x_inject ? p.x_inject[rec * 256 + lane * 8 + i] : __fmul_rn((float) code, sc);
            }
        }
    }
    RingPos rp{0u, 0u};
    const float* xprev = p.flow_x0; // residual stream entering the layer (tp: the last w2's per-rank partials, or null)
    if (tp) xprev = nullptr;
    const int nph = 4 * p.layers_run;
#pragma unroll 1
    for (int ph = 4 * p.l_begin; ph <= nph; ++ph) {
        const int l = ph >> 2, k = ph == nph ? 4 : (ph & 3);
        float* fl = flow_layer(p, k == 4 ? 0 : l);
        const int lp = k == 4 ? p.L : l; // profile row
        float* out;
        const float* resid = nullptr;
        stamp(p, lp, 4 * (k & 3));
        if (k == 1) { // attention block, second half (forward.c:261-298)
            consume_attn<KV_MUL>(sh, p, l, rp);
            stamp(p, lp, 13);
            combine_attn(sh, p, l);
            stamp(p, lp, 5);
            prologue_load_codes(sh, p, reinterpret_cast<const uint8_t*>(fl + p.o_attq), p.Pl);
            out = fl + p.o_xa;
            resid = xprev;
            if (tp) out = reinterpret_cast<float*>(4 * ((size_t) l * p.flow_layer_words + p.o_xa + (size_t) p.tp_rank * p.part_stride)); // kind 3: byte offset
        } else {
            const float *src, *nw;
            int n = p.D, nparts = -1;
            if (k == 0) { // forward.c:254-259
                src = xprev; nw = p.att_norm + (size_t) l * p.D; out = fl + p.o_qkv;
                if (tp) nparts = xprev ? p.tp : 0;
            } else if (k == 2) { // forward.c:303-318
                src = fl + p.o_xa; nw = p.ffn_norm + (size_t) l * p.D; out = fl + p.o_h;
                if (tp) nparts = p.tp;
            } else if (k == 3) { // forward.c:319-338
                src = fl + p.o_h; nw = nullptr; n = p.Hdl; out = fl + p.o_xb; resid = fl + p.o_xa;
                if (tp) out = reinterpret_cast<float*>(4 * ((size_t) l * p.flow_layer_words + p.o_xb + (size_t) p.tp_rank * p.part_stride)); // kind 3: byte offset
            } else { // final norm + classifier (forward.c:344-348)
                src = xprev; nw = p.out_norm; out = p.logits;
                if (tp) nparts = xprev ? p.tp : 0;
            }
            prologue_quant<TP>(sh, p, src, n, nw, nparts);
        }
        if (p.dbg_codes && blockIdx.x == 0) { // debug: what this GEMV is fed (every CTA holds the same vector)
            const int pieces = (int) (qw_row_bytes(p.mat[k].n) / 16);
            uint4* dst = reinterpret_cast<uint4*>(p.dbg_codes + (size_t) (k == 4 ? 4 * p.L : 4 * l + k) * p.dbg_stride);
            for (int i = threadIdx.x; i < pieces; i += kConsumerThreads) dst[i] = reinterpret_cast<const uint4*>(sh.xq)[i];
        }
        stamp(p, lp, 4 * (k & 3) + 2);
        consume_mat<TP>(sh, p, p.mat[k], l, rp, out, resid);
        stamp(p, lp, 4 * (k & 3) + 3);
        if (k == 3) xprev = fl + p.o_xb;
    }
}

#include "decode_pw.cuh"

// one instantiation per GQA ratio: only the attention code of the model at hand is in the kernel
template <int KV_MUL, bool TP>
__global__ void __launch_bounds__(kThreads, 1) k_decode(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int abort_flag;
    Shared sh;
    sh.ring = smem;
    sh.xq = smem + p.off_xq;
    sh.scr = reinterpret_cast<float*>(smem + p.off_scr);
    sh.xres = reinterpret_cast<float*>(smem + p.off_xres);
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
    consumer<KV_MUL, TP>(sh, p);
}

// test hook: the decode kernel's own quantiser (quant_record: reciprocal candidate + exact fallback) over a
// plain vector, one warp per 256-column record, output in SG layout
__global__ void k_quant_records(const float* __restrict__ x, int n, uint8_t* __restrict__ sg) {
    const int lane = threadIdx.x & 31, rec = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (rec >= qw_sg_per_row(n)) return;
    const bool live = rec * 4 + (lane >> 3) < n / 64;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = live ? x[rec * 256 + lane * 8 + i] : 0.0f;
    quant_record(sg, rec, lane, v, live);
}

__global__ void k_fill_u32(uint32_t* dst, size_t n, uint32_t v) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) dst[i] = v;
}

} // namespace

// ---------------------------------------------------------------- host side
static MegaState* state_of(QwenCudaCtx* c) { return reinterpret_cast<MegaState*>(c->mega); }
static const void* decode_kernel_pw(int kv_mul) {
    switch (kv_mul) {
        case 1: return (const void*) k_decode_pw<1>;
        case 2: return (const void*) k_decode_pw<2>;
        case 4: return (const void*) k_decode_pw<4>;
        default: return (const void*) k_decode_pw<8>;
    }
}
static const void* decode_kernel(int kv_mul, bool tp) {
    switch (kv_mul) {
        case 1: return tp ? (const void*) k_decode<1, true> : (const void*) k_decode<1, false>;
        case 2: return tp ? (const void*) k_decode<2, true> : (const void*) k_decode<2, false>;
        case 4: return tp ? (const void*) k_decode<4, true> : (const void*) k_decode<4, false>;
        default: return tp ? (const void*) k_decode<8, true> : (const void*) k_decode<8, false>;
    }
}

// both arenas back to "nothing written" (at create, and after an aborted launch)
int qw_mega_reset(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st || !st->grid) return 0;
    for (int s = 0; s < 2; ++s) k_fill_u32<<<592, 256, 0, c->stream>>>(reinterpret_cast<uint32_t*>(st->arena[s]), st->words, kSent);
    QW_CUDA(cudaGetLastError());
    QW_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int qw_mega_init(QwenCudaCtx* c) {
    MegaState* st = new MegaState();
    c->mega = st;
    const int kv_mul = c->KVHl > 0 ? c->Hl / c->KVHl : 0;
    if (c->tp_size > kMaxTp || !(kv_mul == 1 || kv_mul == 2 || kv_mul == 4 || kv_mul == 8) || kv_mul * c->KVHl != c->Hl) {
        c->path = 1; // unusual head ratios use the per-op path
        return 0;
    }
    const int amax = qw_pad_cols(std::max(c->D, std::max(c->Pl, c->Hdl)));
    if (qw_row_bytes(amax) > (size_t) kSlotBytes || 2 * qw_row_bytes(c->D) > (size_t) kSlotBytes) {
        qw_set_error("persistent decode kernel: a weight row (%d columns) does not fit one %d-byte ring slot", amax, kSlotBytes);
        return -1;
    }
    if (qw_sg_per_row(c->D) > kRecBatch * kConsumerWarps) {
        qw_set_error("persistent decode kernel: dim %d exceeds the fused RMSNorm prologue's %d columns", c->D,
                     kRecBatch * kConsumerWarps * 256);
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
    const int scr_b = std::max(kScrFloats, (kv_mul + 2) * 128) * 4;
    const int misc_b = 1024, bar_b = std::max(2 * kMaxSlots * 8, 3 * kConsumerWarps * 8);
    const int xres_b = c->tp_size > 1 ? qw_pad_cols(c->D) * 4 : 0; // tensor parallel: the residual stream lives in every CTA
    const int fixed = ((xq_b + 127) & ~127) + ((scr_b + 127) & ~127) + ((misc_b + 127) & ~127) + ((bar_b + 127) & ~127)
                      + ((xres_b + 127) & ~127);
    const int avail = dev_smem - fixed - 1024; // 1 KB left for static shared + driver reserve
    // Ring depth: 5 slots by default although 7 fit. Measured on B200 (decode at context 4096): 4B shape 1.877 / 1.865 /
    // 1.826 / 1.886 ms per token with 7 / 6 / 5 / 4 slots, 8B shape 2.459 (7) vs 2.426 (5): fewer bulk copies in flight per
    // SM shorten every hand-off (scripts/ubench/handoff.cu: 13 us at 6 x 28 KB in flight, 3.3 us at 2) by more than the
    // shallower prefetch costs. QWEN_MEGA_NSLOT overrides (up to what fits).
    // Single-GPU contexts run the per-warp streaming variant (decode_pw.cuh) when its 15 warp regions fit beside the
    // activation vector; QWEN_MEGA_PW=0 selects the ring variant (the kernel of tensor-parallel contexts).
    // ... and when a row of the dim-column matrices (wq|wk|wv, w1/w3, classifier: most of the bytes) is ONE chunk, so that a
    // row pair travels as one 5 KB copy; with two copies per item the issue server saturates (measured: 8B shape 388 tok/s
    // against 420 with the ring variant, 4B 570 against 535).
    st->pw = c->tp_size == 1 && avail >= kConsumerWarps * kPwRegion && qw_sg_per_row(c->D) <= kPwChunkRecs;
    if (const char* e = getenv("QWEN_MEGA_PW")) st->pw = atoi(e) != 0 && c->tp_size == 1 && avail >= kConsumerWarps * kPwRegion;
    const int fit = std::min(kMaxSlots, avail / kSlotBytes);
    st->nslot = std::min(fit, 5);
    if (const char* e = getenv("QWEN_MEGA_NSLOT")) st->nslot = std::max(2, std::min(fit, atoi(e)));
    if (const char* e = getenv("QWEN_MEGA_MODE")) st->dbg_mode = atoi(e);
    if (st->pw) st->l2_ahead = 2; // per-warp variant: HBM -> L2 prefetch two sub-phases ahead (decode_pw.cuh: pw_prefetcher)
    if (const char* e = getenv("QWEN_MEGA_L2AHEAD")) st->l2_ahead = std::max(0, atoi(e));
    if (const char* e = getenv("QWEN_MEGA_INFLIGHT")) st->inflight = std::max(0, atoi(e));
    if (getenv("QWEN_MEGA_VERBOSE")) fprintf(stderr, "[mega] nslot %d mode %d smem %d\n", st->nslot, st->dbg_mode, fixed);
    if (st->nslot < 2) {
        qw_set_error("persistent decode kernel: not enough shared memory for a 2-slot ring (%d bytes free)", avail);
        return -1;
    }
    take(st->pw ? kConsumerWarps * kPwRegion : st->nslot * kSlotBytes);
    st->off_xq = take(xq_b);
    st->off_scr = take(scr_b);
    st->off_misc = take(misc_b);
    st->off_bar = take(bar_b);
    st->off_xres = take(xres_b);
    st->smem = off;
    const void* kern = st->pw ? decode_kernel_pw(kv_mul) : decode_kernel(kv_mul, c->tp_size > 1);
    QW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) st->smem));
    int per_sm = 0;
    QW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, st->smem));
    if (per_sm < 1) {
        qw_set_error("persistent decode kernel does not fit on an SM (smem %zu)", st->smem);
        return -1;
    }
    st->grid = std::min(c->num_sms, kMaxGrid);
    if (const char* e = getenv("QWEN_MEGA_GRID")) { // experiment: fewer CTAs (small shapes are bound by the hand-offs, not by streaming)
        const int v = atoi(e);
        if (v > 0) st->grid = std::min(st->grid, v);
    }
    if (st->grid < c->KVHl) { // every kv head needs a block of its own (attn_split)
        st->grid = 0;
        c->path = 1;
        return 0;
    }
    {
        auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
        int k = std::max(1, st->grid / 3);
        while (gcd(k, st->grid) != 1) ++k;
        st->perm = k;
        if (const char* e = getenv("QWEN_MEGA_PERM")) { const int v = atoi(e); if (v > 0 && gcd(v, st->grid) == 1) st->perm = v; }
    }
    // blocks per kv head in attention: all of them on one GPU (18-19 for 8 kv heads); tensor-parallel ranks with 1-2 kv
    // heads are capped so that a combine task merges at most 2 x 21 partials
    {
        int cap = 42;
        if (const char* e = getenv("QWEN_MEGA_ATTN_CAP")) cap = std::max(1, atoi(e));
        st->attn_ga = (int) std::min<long long>(st->grid, (long long) cap * c->KVHl);
    }
    // flow arena: per layer [xa tp x D][xb tp x D][qkv P+2K][h Hd][attq row_bytes(P)/4][partials KVH x slots x kv_mul x 132]
    // (tp > 1: xa / xb hold one partial vector per rank, written by that rank over NVLink)
    {
        auto up4 = [](size_t w) { return (w + 3) & ~(size_t) 3; };
        size_t o = 0;
        st->part_stride = (int) up4(c->D);
        st->o_xa = (int) o; o += (size_t) c->tp_size * st->part_stride;
        st->o_xb = (int) o; o += (size_t) c->tp_size * st->part_stride;
        st->o_qkv = (int) o; o += up4((size_t) c->Pl + 2 * c->Kl);
        st->o_h = (int) o; o += up4(c->Hdl);
        st->o_attq = (int) o; o += up4(qw_row_bytes(c->Pl) / 4);
        st->part_slots = st->attn_ga / c->KVHl + 1; // blocks per kv head (attn_split): floor or ceil of attn_ga / KVH
        st->o_part = (int) o; o += up4((size_t) c->KVHl * st->part_slots * kv_mul * kPartStride);
        st->layer_words = o;
        st->x0_off = o * c->L; // after the layers: the embedding row
        st->words = o * c->L + up4(c->D);
        for (int s = 0; s < 2; ++s) QW_CUDA(cudaMalloc((void**) &st->arena[s], st->words * 4));
    }
    // tensor-parallel contexts run the per-op path until qw_mega_tp_connect has mapped the peers' arenas
    c->path = c->tp_size > 1 ? 1 : 0;
    for (int s = 0; s < 2; ++s) st->peer[s][c->tp_rank] = st->arena[s];
    return qw_mega_reset(c);
}

// (Cross-rank ordering of the arena reuse: see the invariant at the logits all-gather in qw_decode_mega.)
// Tensor parallelism, one process per GPU: exchange cudaIpc handles of the two flow arenas through the NCCL
// communicator and map every peer's arenas into this process (NVLink peer access). After this the persistent
// kernel's wo / w2 epilogues store their partial sums straight into every rank's arena -- the all-reduce is part
// of the kernel. Returns 0 and leaves the context on the per-op + NCCL path if peer mapping is not possible.
int qw_mega_tp_connect(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st || !st->grid || c->tp_size <= 1 || st->peers_open) return 0;
    if (getenv("QWEN_TP_NO_PEER")) return 0;
    const int tp = c->tp_size;
    struct Rec { cudaIpcMemHandle_t h[2]; int ok; int pad[31]; };
    static_assert(sizeof(Rec) == 2 * 64 + 128, "handle record");
    Rec mine;
    memset(&mine, 0, sizeof mine);
    mine.ok = cudaIpcGetMemHandle(&mine.h[0], st->arena[0]) == cudaSuccess
              && cudaIpcGetMemHandle(&mine.h[1], st->arena[1]) == cudaSuccess;
    cudaGetLastError();
    Rec* dev = nullptr;
    std::vector<Rec> all(tp);
    QW_CUDA(cudaMalloc((void**) &dev, sizeof(Rec) * (tp + 1)));
    QW_CUDA(cudaMemcpyAsync(dev + tp, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    if (qw_tp_allgather(c, reinterpret_cast<const float*>(dev + tp), reinterpret_cast<float*>(dev), sizeof(Rec) / 4)) return -1;
    QW_CUDA(cudaMemcpyAsync(all.data(), dev, sizeof(Rec) * tp, cudaMemcpyDeviceToHost, c->stream));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    bool ok = true;
    for (int q = 0; q < tp; ++q) ok = ok && all[q].ok;
    int opened = 0;
    for (int q = 0; ok && q < tp; ++q) {
        if (q == c->tp_rank) continue;
        for (int s = 0; s < 2; ++s) {
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[q].h[s], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                fprintf(stderr, "[TP] rank %d cannot map rank %d's arena: %s\n", c->tp_rank, q, cudaGetErrorString(cudaGetLastError()));
                ok = false;
                break;
            }
            st->peer[s][q] = reinterpret_cast<float*>(ptr);
            ++opened;
        }
    }
    // every rank must reach the same verdict: all-gather the outcome
    float verdict = ok ? 1.0f : 0.0f;
    QW_CUDA(cudaMemcpyAsync(reinterpret_cast<float*>(dev + tp), &verdict, 4, cudaMemcpyHostToDevice, c->stream));
    if (qw_tp_allgather(c, reinterpret_cast<const float*>(dev + tp), reinterpret_cast<float*>(dev), 1)) return -1;
    std::vector<float> verdicts(tp);
    QW_CUDA(cudaMemcpyAsync(verdicts.data(), dev, 4 * tp, cudaMemcpyDeviceToHost, c->stream));
    QW_CUDA(cudaStreamSynchronize(c->stream)); // also the barrier: every rank's arenas are initialised and mapped
    cudaFree(dev);
    for (int q = 0; q < tp; ++q) ok = ok && verdicts[q] == 1.0f;
    if (!ok) {
        for (int q = 0; q < tp; ++q)
            for (int s = 0; s < 2; ++s)
                if (q != c->tp_rank && st->peer[s][q]) {
                    cudaIpcCloseMemHandle(st->peer[s][q]);
                    st->peer[s][q] = nullptr;
                }
        cudaGetLastError();
        if (c->tp_rank == 0) fprintf(stderr, "[TP] peer mapping unavailable: per-op kernels + NCCL all-reduce\n");
        return 0;
    }
    (void) opened;
    st->peers_open = true;
    c->path = 0;
    return 0;
}
bool qw_mega_tp_ready(const QwenCudaCtx* c) {
    const MegaState* st = reinterpret_cast<const MegaState*>(c->mega);
    return st && st->grid && (c->tp_size == 1 || st->peers_open);
}

void qw_mega_free(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st) return;
    for (int q = 0; q < kMaxTp; ++q)
        for (int s = 0; s < 2; ++s)
            if (st->peers_open && q != c->tp_rank && st->peer[s][q]) cudaIpcCloseMemHandle(st->peer[s][q]);
    void* bufs[] = {st->arena[0], st->arena[1], st->prof};
    for (void* b : bufs)
        if (b) cudaFree(b);
    delete st;
    c->mega = nullptr;
}

int qw_decode_mega(QwenCudaCtx* c, int token, const int* token_dev, int pos) {
    MegaState* st = state_of(c);
    if (!st || !st->grid || (c->tp_size > 1 && !st->peers_open)) {
        qw_set_error("persistent decode kernel is not initialised for this context");
        return -4;
    }
    MegaParams p;
    memset(&p, 0, sizeof p);
    p.tp = c->tp_size; p.tp_rank = c->tp_rank;
    p.attn_ga = st->attn_ga; p.part_stride = st->part_stride; p.off_xres = st->off_xres;
    for (int q = 0; q < c->tp_size; ++q) p.peer_flow[q] = st->peer[st->launches & 1][q];
    p.D = c->D; p.Hdl = c->Hdl; p.L = c->L; p.Hl = c->Hl; p.KVHl = c->KVHl; p.Pl = c->Pl; p.Kl = c->Kl; p.Vl = c->Vl;
    p.S = c->S; p.kv_mul = c->Hl / c->KVHl;
    p.pos = pos; p.token = token; p.token_dev = token_dev;
    p.layers_run = (c->layers_run >= 0 && c->layers_run <= c->L) ? c->layers_run : c->L;
    p.l_begin = (c->layer_begin > 0 && c->layer_begin <= p.layers_run) ? c->layer_begin : 0;
    p.x_inject = c->x_inject_on ? c->x_inject : nullptr;
    p.dbg_codes = c->dbg_codes;
    p.dbg_stride = (int) c->dbg_codes_stride;
    int stage_env = 1; // QWEN_MEGA_STAGE=0: per-warp stores + fence instead of the staged bulk store
    if (const char* e = getenv("QWEN_MEGA_STAGE")) stage_env = atoi(e);
    int stage_mask = 31; // QWEN_MEGA_STAGE_MASK: bit k = matrix k (0 qkv, 1 wo, 2 w1/w3, 3 w2, 4 classifier) uses the staged store
    if (const char* e = getenv("QWEN_MEGA_STAGE_MASK")) stage_mask = atoi(e);
    int mat_index = 0;
    const int grid = st->grid;
    auto desc = [stage_env, stage_mask, &mat_index, grid](const uint8_t* base, size_t stride, int rows, int n, int kind) {
        int rt = kSlotBytes / (int) qw_row_bytes(n);
        if (rt >= 2) rt &= ~1; // whole 2-row units per tile
        if (rt < 2 && kind == 2) rt = 2;
        // staged bulk store: a CTA's slice of the output must start and end on 16 bytes (4 rows; 8 for the w1/w3 pairs)
        const int sg = kind == 2 ? 8 : 4;
        const bool stage = stage_env && ((stage_mask >> mat_index++) & 1) && rows % sg == 0 && (rows / sg / grid + 1) * sg <= (kind == 2 ? 2048 : kScrFloats); // kind 2: raw sums wait at +2048 (decode_pw.cuh)
        const int gran = stage ? sg : (kind == 2 ? 2 : 1);
        const int recs = qw_sg_per_row(n), C = (recs + kPwChunkRecs - 1) / kPwChunkRecs, cr = (recs + C - 1) / C;
        return MatDesc{base, stride, rows, n, gran, rt, stage ? 1 : 0, C, cr, kind};
    };
    p.mat[0] = desc(c->w_qkv, c->w_qkv_stride, c->Pl + 2 * c->Kl, c->D, 0);
    const int kres = c->tp_size > 1 ? 3 : 1; // wo / w2 epilogue: residual add on one GPU, partial push under tensor parallelism
    p.mat[1] = desc(c->w_o, c->w_o_stride, c->D, c->Pl, kres);
    p.mat[2] = desc(c->w_13, c->w_13_stride, 2 * c->Hdl, c->D, 2);
    p.mat[3] = desc(c->w_2, c->w_2_stride, c->D, c->Hdl, kres);
    p.mat[4] = desc(c->w_cls, 0, c->Vl, c->D, 0);
    p.w_emb = c->w_emb;
    p.att_norm = c->att_norm; p.ffn_norm = c->ffn_norm; p.out_norm = c->out_norm; p.q_norm = c->q_norm; p.k_norm = c->k_norm;
    p.rope_cos = c->rope_cos; p.rope_sin = c->rope_sin;
    p.k_cache = c->k_cache; p.v_cache = c->v_cache;
    p.flow = st->arena[st->launches & 1];
    p.flow_other = st->arena[(st->launches + 1) & 1];
    p.flow_x0 = p.flow + st->x0_off;
    p.flow_layer_words = st->layer_words; p.flow_words = st->words;
    p.o_xa = st->o_xa; p.o_xb = st->o_xb; p.o_qkv = st->o_qkv; p.o_h = st->o_h; p.o_attq = st->o_attq; p.o_part = st->o_part;
    p.part_slots = st->part_slots;
    p.att = c->att; p.logits = c->logits;
    p.err = c->err_flag;
    p.dbg_mode = st->dbg_mode;
    p.l2_ahead = st->l2_ahead;
    if (const char* e = getenv("QWEN_PW_LATE")) p.pw_late = atoi(e);
    if (const char* e = getenv("QWEN_PW_PUB")) p.pw_pub = atoi(e);
    p.l2_cls = st->pw ? 384 * 1024 : st->l2_ahead * 32768;
    if (const char* e = getenv("QWEN_MEGA_L2CLS_KB")) p.l2_cls = std::max(0, atoi(e)) * 1024;
    p.inflight = st->inflight < st->nslot ? st->inflight : 0;
    p.perm = st->perm;
    p.prof = st->prof;
    p.nslot = st->nslot; p.off_xq = st->off_xq; p.off_scr = st->off_scr;
    p.off_misc = st->off_misc; p.off_bar = st->off_bar;
    void* args[] = {&p};
    // cooperative launch: the CTAs wait for each other's results, so all of them must be resident
    if (st->pw)
        QW_CUDA(cudaLaunchCooperativeKernel(decode_kernel_pw(p.kv_mul), dim3(st->grid), dim3(kThreads), args, st->smem, c->stream));
    else
        QW_CUDA(cudaLaunchCooperativeKernel(decode_kernel(p.kv_mul, c->tp_size > 1), dim3(st->grid), dim3(kThreads), args, st->smem, c->stream));
    st->last_layers = p.layers_run;
    ++st->launches;
    // The classifier is split over vocabulary rows: gather the logits slices.
    // INVARIANT this collective also carries: rank A's launch N peer-stores into rank B's arena[N & 1], which B's launch
    // N - 1 refilled with the sentinel. Nothing inside the kernels orders A's launch N after B's launch N - 1 -- the
    // all-gather enqueued after EVERY launch does: it completes on A only once every rank has contributed the logits of
    // its launch N - 1, i.e. has finished that kernel including its refill. A step variant without a collective after the
    // launch (a per-rank argmax chain, a no-logits step) would need its own cross-rank epoch exchange before the next
    // launch. The arena parity (launches & 1) is per process: all ranks must run the same sequence of steps.
    if (c->tp_size > 1 && qw_tp_allgather(c, c->logits, c->logits_all, c->Vl)) return -1;
    return 0;
}

int qw_decode_mega_launches(const QwenCudaCtx* c) { return c->tp_size > 1 ? 2 : 1; }

// test hook behind qwen_cuda_debug_quantize_fused: x[n] (device) -> SG-layout codes + scales (device)
void qw_mega_quant_records(const float* x, int n, uint8_t* sg, cudaStream_t st) {
    const int recs = qw_sg_per_row(n);
    k_quant_records<<<(recs + 7) / 8, 256, 0, st>>>(x, n, sg);
}

// debug: where the last persistent-kernel step left the vector `what` of the last layer it ran
const float* qw_mega_debug_ptr(QwenCudaCtx* c, const char* what) {
    MegaState* st = state_of(c);
    if (!st || !st->grid || !st->launches || st->last_layers < 1) return nullptr;
    const float* fl = st->arena[(st->launches - 1) & 1] + (size_t) (st->last_layers - 1) * st->layer_words;
    if (what[0] == 'x' && !what[1]) return fl + st->o_xb;
    if (what[0] == 'h' && !what[1]) return fl + st->o_h;
    if (what[0] == 'q' && what[1] == 'k') return fl + st->o_qkv;
    return nullptr;
}

// debug: per-CTA phase timestamps of the NEXT steps; read back with qw_mega_profile_read
int qw_mega_profile_enable(QwenCudaCtx* c) {
    MegaState* st = state_of(c);
    if (!st || !st->grid) return -1;
    // + per-warp GEMV cycle counters [grid][16 warps][5 matrices][8: wait, math, epilogue, units, end barrier, bulk store, entry] (filled by -DQW_UNITPROF builds)
    const size_t n = (size_t) st->grid * (c->L + 1) * kProfSlots + (size_t) st->grid * 16 * 5 * 8;
    if (!st->prof) QW_CUDA(cudaMalloc((void**) &st->prof, n * 8));
    QW_CUDA(cudaMemset(st->prof, 0, n * 8));
    QW_CUDA(cudaDeviceSynchronize());
    return (int) n;
}
int qw_mega_profile_read(QwenCudaCtx* c, unsigned long long* host, size_t max_elems) {
    MegaState* st = state_of(c);
    if (!st || !st->prof) return -1;
    size_t n = (size_t) st->grid * (c->L + 1) * kProfSlots + (size_t) st->grid * 16 * 5 * 8;
    if (n > max_elems) n = max_elems;
    QW_CUDA(cudaStreamSynchronize(c->stream));
    QW_CUDA(cudaMemcpy(host, st->prof, n * 8, cudaMemcpyDeviceToHost));
    return st->grid;
}
