// prefill.cu -- prompt prefill: T tokens through the layer loop at once.
//
// The reference has no prefill: completion.c:57-66 calls forward() once per prompt token and throws
// the logits of all but the last away. Here the same per-token arithmetic (src/forward.c:225-350) is
// applied to a chunk of T tokens together, so every weight matrix is read once per chunk instead of
// once per token and the seven projections become dense int8 contractions on the tcgen05 tensor cores
// (prefill_gemm.cu: exact int32 group dots, the reference's ((float) dot * ws) * xs terms folded in group
// order -- bit-identical to the reference matmul of each token). Between the GEMMs sit batched versions
// of the small ops -- RMSNorm + Q8_0 quantise, per-head q/k norm + RoPE + KV-cache write, causal GQA
// attention, SwiGLU, residual add -- each the reference's formula per element. The KV cache is written in
// the decode kernel's layout, so decoding continues from the prefilled cache with the persistent kernel.
// Only the last token's logits are produced (the reference discards the others).
//
// Chunks of kChunkTokens tokens bound the activation memory; chunk c attends over the cache rows of the
// earlier chunks plus its own causal part.
//
// Per layer and chunk (round 2): k_prep_quant (norm + quantise, previous residual add fused) -> qkv GEMM -> k_qkv_post_rows ->
// k_attn_prefill_t (tiled fp32 attention in fixed key blocks) -> k_attn_merge_quant (merge of split rows + quantiser) -> wo GEMM
// -> k_prep_quant (residual add + norm + quantise) -> w1/w3 GEMM -> k_prep_quant (SwiGLU + quantise) -> w2 GEMM.
#include <algorithm>
#include <cstring>
#include <cstdlib>

#include "attn_core.cuh"
#include "common.cuh"

int qw_prefill_gemm(const uint8_t* w, const int8_t* xq, const float* xsT, float* out, int32_t* dots, int d, int n, int T,
                    int Tpad, int* err_dev, cudaStream_t st, float* ms_out);

namespace {

constexpr int kChunkTokens = 512;

struct PrefillBufs {
    int cap = 0;          // tokens the buffers hold
    int* tokens = nullptr;
    float* apart = nullptr; // partial attention rows [parts][cap][Hl][132] (k_attn_prefill_t), grown on demand
    size_t apart_floats = 0;
    float *x = nullptr, *xa = nullptr, *xb = nullptr, *qkv = nullptr, *q = nullptr, *att = nullptr, *h13 = nullptr;
    int8_t* q8 = nullptr; // [cap][maxn] activation codes of the GEMM at hand (row pitch = its n)
    float* xsT = nullptr; // [maxn / 64][cap] activation scales, transposed for the GEMM epilogue
    int* err = nullptr;
};

// dequantised embedding rows (forward.c:237): grid T
__global__ void k_embed_rows(float* __restrict__ x, const uint8_t* __restrict__ w_emb, const int* __restrict__ tokens, int D) {
    const int tok = tokens[blockIdx.x];
    const uint8_t* row = w_emb + (size_t) tok * qw_row_bytes(D);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        const uint8_t* rec = row + (size_t) (c >> 8) * QW_SG_BYTES;
        const float sc = *reinterpret_cast<const float*>(rec + 256 + ((c >> 6) & 3) * 4);
        x[(size_t) blockIdx.x * D + c] = __fmul_rn((float) reinterpret_cast<const int8_t*>(rec)[c & 255], sc);
    }
}

// One block per token: optional RMSNorm (forward.c:12-28; w == nullptr: none), optional SwiGLU over
// interleaved (w1, w3) pairs (forward.c:122-139; src then has 2n values per token), then the Q8_0
// quantiser (q8.c:5-30) with the reference's exact arithmetic. Codes go to q8[t][n], scales transposed to
// xsT[g][Tpad]. 256 threads; warp w handles groups w, w + 8, ...
// add != nullptr (with w): the residual add of the previous block (forward.c:295-298, 335-338) is done here first,
// x[t] += add[t] written back, instead of in a pass of its own.
__global__ void __launch_bounds__(256)
k_prep_quant(float* __restrict__ src, const float* __restrict__ add, const float* __restrict__ w, int8_t* __restrict__ q8,
             float* __restrict__ xsT, int n, int Tpad, int swiglu) {
    __shared__ float red[8];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* s = src + (size_t) t * (swiglu ? 2 * n : n);
    float r = 1.0f;
    if (w) {
        // sum of squares: element i goes to thread i % 256, as in round 1 -- the order fixes the last bits of the scale, and with
        // them which values sit on a rounding boundary of the quantiser (a four-per-thread order was measured: same speed
        // class, but it moved 436 of 307 200 layer-0 K/V values of the 600-token parity case past 1e-4)
        float ss = 0.0f;
        if (add) {
            const float* a = add + (size_t) t * n;
            for (int i = tid; i < n; i += 256) {
                const float v = __fadd_rn(s[i], a[i]);
                s[i] = v; // re-read by other threads after the barrier below
                ss = __fmaf_rn(v, v, ss);
            }
        } else {
            for (int i = tid; i < n; i += 256) ss = __fmaf_rn(s[i], s[i], ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot = __fadd_rn(tot, red[i]);
        r = rms_rscale(tot, n);
    }
    // quantiser: a half-warp per group of 64 (four consecutive values per lane: 16-byte loads, 4-byte code stores), warp w
    // takes the group pairs w, w + 8, ...; per value the reference's arithmetic (q8.c:5-30)
    const int groups = n / 64;
    for (int g2 = warp; 2 * g2 < groups; g2 += 8) {
        const int g = 2 * g2 + (lane >> 4);
        const bool live = g < groups;
        const int i0 = g * 64 + (lane & 15) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            if (swiglu) {
                const float4 p0 = *reinterpret_cast<const float4*>(s + 2 * i0), p1 = *reinterpret_cast<const float4*>(s + 2 * i0 + 4);
                v[0] = __fmul_rn(silu_ref(p0.x), p0.y);
                v[1] = __fmul_rn(silu_ref(p0.z), p0.w);
                v[2] = __fmul_rn(silu_ref(p1.x), p1.y);
                v[3] = __fmul_rn(silu_ref(p1.z), p1.w);
            } else {
                const float4 x4 = *reinterpret_cast<const float4*>(s + i0);
                if (w) {
                    const float4 w4 = *reinterpret_cast<const float4*>(w + i0);
                    v[0] = __fmul_rn(w4.x, __fmul_rn(r, x4.x));
                    v[1] = __fmul_rn(w4.y, __fmul_rn(r, x4.y));
                    v[2] = __fmul_rn(w4.z, __fmul_rn(r, x4.z));
                    v[3] = __fmul_rn(w4.w, __fmul_rn(r, x4.w));
                } else {
                    v[0] = x4.x; v[1] = x4.y; v[2] = x4.z; v[3] = x4.w;
                }
            }
        }
        float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); // within the half-warp
        const float scale = q8_scale(amax);
        if (live) {
            const int c0 = q8_code(v[0], scale), c1 = q8_code(v[1], scale), c2 = q8_code(v[2], scale), c3 = q8_code(v[3], scale);
            *reinterpret_cast<uint32_t*>(q8 + (size_t) t * n + i0) =
                (uint32_t) (c0 & 0xff) | ((uint32_t) (c1 & 0xff) << 8) | ((uint32_t) (c2 & 0xff) << 16) | ((uint32_t) (c3 & 0xff) << 24);
            if ((lane & 15) == 0) xsT[(size_t) g * Tpad + t] = scale;
        }
    }
}

// per (token, head): q/k RMSNorm + RoPE at position pos0 + t, K and V written into the cache
// (forward.c:261-280, 244-248). grid (T, Hl + 2 KVHl), 128 threads.
__global__ void __launch_bounds__(128)
k_qkv_post_rows(const float* __restrict__ qkv, float* __restrict__ q_out, float* __restrict__ k_layer, float* __restrict__ v_layer,
                const float* __restrict__ gq, const float* __restrict__ gk, const float* __restrict__ rope_cos,
                const float* __restrict__ rope_sin, int Hl, int KVHl, int S, int pos0) {
    __shared__ float y[128];
    __shared__ float red[4];
    const int t = blockIdx.x, b = blockIdx.y, i = threadIdx.x, pos = pos0 + t;
    const int width = (Hl + 2 * KVHl) * 128;
    const float v = qkv[(size_t) t * width + (size_t) b * 128 + i];
    if (b >= Hl + KVHl) { // V: stored raw
        v_layer[((size_t) (b - Hl - KVHl) * S + pos) * 128 + i] = v;
        return;
    }
    float ss = warp_sum(__fmul_rn(v, v));
    if ((i & 31) == 0) red[i >> 5] = ss;
    __syncthreads();
    ss = __fadd_rn(__fadd_rn(red[0], red[1]), __fadd_rn(red[2], red[3]));
    const float r = rms_rscale(ss, 128);
    const float* g = (b < Hl) ? gq : gk;
    y[i] = __fmul_rn(g[i], __fmul_rn(r, v));
    __syncthreads();
    const int j = i & 63;
    const float c = rope_cos[(size_t) pos * 64 + j], s = rope_sin[(size_t) pos * 64 + j];
    const float a = y[j], bb = y[j + 64];
    const float o = (i < 64) ? __fsub_rn(__fmul_rn(a, c), __fmul_rn(bb, s)) : __fadd_rn(__fmul_rn(a, s), __fmul_rn(bb, c));
    if (b < Hl)
        q_out[(size_t) t * Hl * 128 + (size_t) b * 128 + i] = o;
    else
        k_layer[((size_t) (b - Hl) * S + pos) * 128 + i] = o;
}

// Causal GQA attention for a chunk (forward.c:141-195 per token). grid (ceil(T / 8), KVHl, NHG), 256 threads:
// warp w = query token 8 * blockIdx.x + w with the HW query heads of head group blockIdx.z; the 8 tokens
// share K/V tiles of 32 positions staged in shared memory; each warp runs the decode kernel's online-softmax
// step (attn_core.cuh) over the positions it may see (0 .. pos0 + t).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
constexpr int kAttnTile = 32;                                   // KV positions per shared-memory tile
constexpr int kAttnSmem = 2 * 2 * kAttnTile * 128 * 4;          // double-buffered K and V tiles

template <int KV_MUL>
__global__ void __launch_bounds__(256)
k_attn_prefill(const float* __restrict__ q, const float* __restrict__ k_layer, const float* __restrict__ v_layer,
               float* __restrict__ out, int Hl, int S, int pos0, int T) {
    constexpr int HW = KV_MUL < 4 ? KV_MUL : 4;
    extern __shared__ __align__(16) float tiles[]; // [2 buffers][K 32 x 128 | V 32 x 128]
    const int kvh = blockIdx.y, hg = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x * 8 + warp;
    const bool live = t < T;
    const int last_pos = pos0 + min(T, blockIdx.x * 8 + 8) - 1; // last position any token of the block sees
    const int my_pos = pos0 + t;
    const int h0 = kvh * KV_MUL + hg * HW;
    float4 qv[HW];
#pragma unroll
    for (int j = 0; j < HW; ++j)
        qv[j] = live ? *reinterpret_cast<const float4*>(q + ((size_t) t * Hl + h0 + j) * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    AttnState<HW> st;
    attn_state_reset(st);
    const float* K = k_layer + (size_t) kvh * S * 128;
    const float* V = v_layer + (size_t) kvh * S * 128;
    // tile `it` (positions 32 it ..) is fetched with cp.async into buffer it & 1 while tile it - 1 is being used
    auto fetch = [&](int it) {
        const int p0 = it * kAttnTile;
        if (p0 <= last_pos) {
            const int rows = min(kAttnTile, last_pos + 1 - p0);
            float* kb = tiles + (size_t) (it & 1) * (2 * kAttnTile * 128);
            for (int i = tid; i < rows * 32; i += 256) {
                cp_async16(kb + i * 4, K + (size_t) p0 * 128 + (size_t) i * 4);
                cp_async16(kb + kAttnTile * 128 + i * 4, V + (size_t) p0 * 128 + (size_t) i * 4);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch(0);
    for (int it = 0; it * kAttnTile <= last_pos; ++it) {
        fetch(it + 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory"); // tile `it` has landed (this thread's part)
        __syncthreads();                                      // ... and everybody else's
        const int p0 = it * kAttnTile;
        const int rows = min(kAttnTile, last_pos + 1 - p0);
        const int cnt = live ? min(rows, my_pos + 1 - p0) : 0;
        const float* kb = tiles + (size_t) (it & 1) * (2 * kAttnTile * 128);
        if (cnt > 0) attn_rows<HW>(kb, cnt, kb, cnt, kAttnTile * 128, qv, st, lane);
        __syncthreads(); // the buffer is refilled by the fetch of the next iteration
    }
    if (live) {
#pragma unroll
        for (int j = 0; j < HW; ++j) {
            float4 o;
            o.x = __fdiv_rn(st.acc[j].x, st.l[j]);
            o.y = __fdiv_rn(st.acc[j].y, st.l[j]);
            o.z = __fdiv_rn(st.acc[j].z, st.l[j]);
            o.w = __fdiv_rn(st.acc[j].w, st.l[j]);
            *reinterpret_cast<float4*>(out + ((size_t) t * Hl + h0 + j) * 128 + lane * 4) = o;
        }
    }
}

// ---- round 2: tiled causal attention for a chunk --------------------------------------------------------------------
// k_attn_prefill above gives every query token a warp and runs the decode kernel's per-warp online-softmax step: 31 shuffles
// per 32 scores, one LDS.128 per K row per warp -- 7.7 TFLOP/s on B200 (10 % of the fp32 peak; a third of the prefill time).
// k_attn_prefill_t is the register-blocked form: one CTA = 64 query tokens of ONE head, 128 threads as 8 x 16; for every
// tile of 64 cached positions (aligned to absolute position 0, so a prompt prefilled in several calls sees the same tiles)
//   S = Q K^T : thread (ty, tx) owns rows 8 ty .. 8 ty + 7 x keys tx, tx + 16, tx + 32, tx + 48; per float4 of the head
//               dimension 8 broadcast Q loads + 4 conflict-free K loads (row pitch 132 floats) feed 128 FMAs;
//   softmax   : score / sqrtf(128), causal mask, row maximum and sum over the 16 tx lanes by xor shuffles, running
//               maximum / sum / rescale per row (forward.c:156-181 in online form);
//   O += P V  : P goes through shared memory (over the K tile); thread (ty, tx) owns the same 8 rows x dims 4 tx .. 4 tx + 3
//               and 64 + 4 tx .. : per 4 keys 8 broadcast P loads + 8 V loads feed 256 FMAs.
// fp32 throughout (the cache is fp32, as the reference's). The heaviest query tiles (latest positions) are launched first.
constexpr int kTQ = 64, kTK = 64, kTKPitch = 132, kTPPitch = 68;
constexpr int kTPartStride = 132;
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float exp2f_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
} // floats per partial row: 128 outputs, running maximum, running sum, 2 pad

// The key axis is cut into FIXED blocks of absolute positions -- 128 positions each up to 512, 512 positions beyond -- and a
// (query tile, head) item whose visible keys span more than one block is computed as one CTA per block ("part"), each
// leaving an un-normalised partial (O, m, l) per row, merged by k_attn_merge in block order. Why: a query tile's work grows
// with its position (1 .. 8 key tiles at T = 512), so whole items balance badly over 148 SMs when heads are few (1.7B
// shape: 128 items, the busiest SM had 8 tile-units of an average 3.9); parts are at most 2 units. The blocks depend on
// absolute positions only, so a row's result does not depend on how a prompt was cut into calls
// (test_prefill_in_two_calls_equals_one_call); a row whose keys all lie in block 0 merges to O_0 * 1 / (l_0 * 1): exactly
// what the single-part path gives.
__host__ __device__ inline int attn_parts(int nkt) { return nkt <= 8 ? (nkt + 1) / 2 : 4 + (nkt - 8 + 7) / 8; }
__host__ __device__ inline void attn_part_range(int p, int& k0, int& k1) {
    if (p < 4) {
        k0 = 2 * p;
        k1 = k0 + 2;
    } else {
        k0 = 8 + 8 * (p - 4);
        k1 = k0 + 8;
    }
}
struct AttnPlan {
    int nq;             // query tiles of the chunk
    int first[9];       // first[i]: first block index (in units of heads) of query tile nq - 1 - i; first[nq] = total
    int parts[8];       // parts[i]: parts of query tile nq - 1 - i
};
inline AttnPlan attn_plan(int pos0, int T) {
    AttnPlan pl{};
    pl.nq = (T + kTQ - 1) / kTQ;
    int acc = 0;
    for (int i = 0; i < pl.nq; ++i) {
        const int qt = pl.nq - 1 - i;
        const int last_pos = pos0 + std::min(T, (qt + 1) * kTQ) - 1;
        pl.parts[i] = attn_parts(last_pos / kTK + 1);
        pl.first[i] = acc;
        acc += pl.parts[i];
    }
    pl.first[pl.nq] = acc;
    return pl;
}
constexpr int kTSmem = (kTQ * 128 + kTK * kTKPitch + kTK * 128) * 4; // Q | K (later P) | V = 99 328 B: two CTAs per SM

__global__ void __launch_bounds__(128, 2)
k_attn_prefill_t(const float* __restrict__ q, const float* __restrict__ k_layer, const float* __restrict__ v_layer,
                 float* __restrict__ out, float* __restrict__ partial, int Hl, int kv_mul, int S, int pos0, int T,
                 const AttnPlan plan) {
    extern __shared__ __align__(16) float tsm[];
    float* sQ = tsm;
    float* sK = tsm + kTQ * 128;
    float* sP = sK; // the probabilities overwrite the K tile once every thread has its scores
    float* sV = sK + kTK * kTKPitch;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // block -> (query tile, part, head): latest query tiles first, heads fastest
    const int unit = (int) blockIdx.x / Hl, h = (int) blockIdx.x % Hl, kvh = h / kv_mul;
    int qi = 0;
    while (qi + 1 < plan.nq && unit >= plan.first[qi + 1]) ++qi;
    const int qt = plan.nq - 1 - qi, part = unit - plan.first[qi], nparts = plan.parts[qi];
    const int t0 = qt * kTQ;
    const int rows = min(kTQ, T - t0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // the Q tile travels with the first K tile (same cp.async group)
    for (int i = tid; i < kTQ * 32; i += 128) {
        const int r = i >> 5, c = (i & 31) * 4;
        if (r < rows) cp_async16(sQ + r * 128 + c, q + ((size_t) (t0 + r) * Hl + h) * 128 + c);
        else *reinterpret_cast<float4*>(sQ + r * 128 + c) = zero4;
    }
    const float* K = k_layer + (size_t) kvh * S * 128;
    const float* V = v_layer + (size_t) kvh * S * 128;
    const int last_pos = pos0 + t0 + rows - 1;        // last position any query of the tile sees
    const int qpos0 = pos0 + t0 + ty * 8;              // position of this thread's first row
    const float inv = sqrtf(128.0f);
    float m[8], l[8], o[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m[i] = -INFINITY;
        l[i] = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[i][e] = 0.0f;
    }
    int kt0 = 0, kt1 = 0;
    attn_part_range(part, kt0, kt1);
    const int p_end = min(last_pos, kt1 * kTK - 1); // this part's last key position
    for (int p0 = kt0 * kTK; p0 <= p_end; p0 += kTK) {
        const int nk = min(kTK, last_pos + 1 - p0);
        __syncthreads(); // the previous tile's P and V have been consumed (first pass: Q is in place)
        // the K tile and the V tile travel as two cp.async groups: the scores wait for K only, V lands behind them
        for (int i = tid; i < kTK * 32; i += 128) {
            const int r = i >> 5, c = (i & 31) * 4;
            if (r < nk) cp_async16(sK + r * kTKPitch + c, K + (size_t) (p0 + r) * 128 + c);
            else *reinterpret_cast<float4*>(sK + r * kTKPitch + c) = zero4; // rows nobody may see: zeros, not whatever the cache holds
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int i = tid; i < kTK * 32; i += 128) {
            const int r = i >> 5, c = (i & 31) * 4;
            if (r < nk) cp_async16(sV + r * 128 + c, V + (size_t) (p0 + r) * 128 + c);
            else *reinterpret_cast<float4*>(sV + r * 128 + c) = zero4;       // (0 * NaN would poison P V)
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        // ---- S = Q K^T
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.0f;
#pragma unroll 2
        for (int d = 0; d < 128; d += 4) {
            float4 kf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) kf[j] = *reinterpret_cast<const float4*>(sK + (tx + 16 * j) * kTKPitch + d);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 qf = *reinterpret_cast<const float4*>(sQ + (ty * 8 + i) * 128 + d);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = __fmaf_rn(qf.x, kf[j].x, s[i][j]);
                    s[i][j] = __fmaf_rn(qf.y, kf[j].y, s[i][j]);
                    s[i][j] = __fmaf_rn(qf.z, kf[j].z, s[i][j]);
                    s[i][j] = __fmaf_rn(qf.w, kf[j].w, s[i][j]);
                }
            }
        }
        // ---- scale, causal mask, online softmax per row (the 16 tx lanes of a row sit in one half-warp)
        const bool diag = p0 + kTK - 1 > qpos0; // some key of the tile may lie beyond some row of this thread
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = __fdiv_rn(s[i][j], inv);
                if (diag && p0 + tx + 16 * j > qpos0 + i) v = -INFINITY;
                s[i][j] = v;
                mx = fmaxf(mx, v);
            }
#pragma unroll
            for (int ofs = 8; ofs > 0; ofs >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, ofs));
            const float m_new = fmaxf(m[i], mx);
            // exp(x) as ex2.approx(x * log2 e): 2 instructions instead of expf's ~30 (40 exponentials per thread and tile were
            // a sixth of the tile's issue slots); relative error ~2e-7, far inside the op tolerance the kernel is tested to
            const float scl = (m[i] == -INFINITY) ? 0.0f : exp2f_fast(__fmul_rn(__fsub_rn(m[i], m_new), kLog2e));
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = (s[i][j] == -INFINITY) ? 0.0f : exp2f_fast(__fmul_rn(__fsub_rn(s[i][j], m_new), kLog2e));
                s[i][j] = e;
                sum = __fadd_rn(sum, e);
            }
#pragma unroll
            for (int ofs = 8; ofs > 0; ofs >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, ofs));
            m[i] = m_new;
            l[i] = __fmaf_rn(l[i], scl, sum);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[i][e] = __fmul_rn(o[i][e], scl);
        }
        __syncthreads(); // every thread is done with the K tile
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sP[(ty * 8 + i) * kTPPitch + tx + 16 * j] = s[i][j];
        asm volatile("cp.async.wait_group 0;" ::: "memory"); // this thread's part of the V tile
        __syncthreads();
        // ---- O += P V
        const int nk4 = (nk + 3) & ~3;
#pragma unroll 2
        for (int kk = 0; kk < nk4; kk += 4) {
            float4 va[4], vb[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                va[c] = *reinterpret_cast<const float4*>(sV + (kk + c) * 128 + tx * 4);
                vb[c] = *reinterpret_cast<const float4*>(sV + (kk + c) * 128 + 64 + tx * 4);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 pf = *reinterpret_cast<const float4*>(sP + (ty * 8 + i) * kTPPitch + kk);
                const float pw[4] = {pf.x, pf.y, pf.z, pf.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    o[i][0] = __fmaf_rn(pw[c], va[c].x, o[i][0]);
                    o[i][1] = __fmaf_rn(pw[c], va[c].y, o[i][1]);
                    o[i][2] = __fmaf_rn(pw[c], va[c].z, o[i][2]);
                    o[i][3] = __fmaf_rn(pw[c], va[c].w, o[i][3]);
                    o[i][4] = __fmaf_rn(pw[c], vb[c].x, o[i][4]);
                    o[i][5] = __fmaf_rn(pw[c], vb[c].y, o[i][5]);
                    o[i][6] = __fmaf_rn(pw[c], vb[c].z, o[i][6]);
                    o[i][7] = __fmaf_rn(pw[c], vb[c].w, o[i][7]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = ty * 8 + i;
        if (r < rows) {
            if (nparts == 1) {
                float* dst = out + ((size_t) (t0 + r) * Hl + h) * 128;
                *reinterpret_cast<float4*>(dst + tx * 4) =
                    make_float4(__fdiv_rn(o[i][0], l[i]), __fdiv_rn(o[i][1], l[i]), __fdiv_rn(o[i][2], l[i]), __fdiv_rn(o[i][3], l[i]));
                *reinterpret_cast<float4*>(dst + 64 + tx * 4) =
                    make_float4(__fdiv_rn(o[i][4], l[i]), __fdiv_rn(o[i][5], l[i]), __fdiv_rn(o[i][6], l[i]), __fdiv_rn(o[i][7], l[i]));
            } else { // un-normalised partial of this block of keys: k_attn_merge finishes the row
                float* dst = partial + (((size_t) part * T + t0 + r) * Hl + h) * kTPartStride;
                *reinterpret_cast<float4*>(dst + tx * 4) = make_float4(o[i][0], o[i][1], o[i][2], o[i][3]);
                *reinterpret_cast<float4*>(dst + 64 + tx * 4) = make_float4(o[i][4], o[i][5], o[i][6], o[i][7]);
                if (tx == 0) {
                    dst[128] = m[i];
                    dst[129] = l[i];
                }
            }
        }
    }
}

// rows of query tiles that were computed in parts: softmax merge of the partials in block order (as the decode kernel's
// combine does for its split-KV partials). grid (rows, Hl), 128 threads = output dims.
__global__ void __launch_bounds__(128)
k_attn_merge(const float* __restrict__ partial, float* __restrict__ out, int Hl, int pos0, int T, int t_first) {
    const int t = t_first + (int) blockIdx.x, h = blockIdx.y, d = threadIdx.x;
    const int qt = t / kTQ;
    const int last_pos = pos0 + min(T, (qt + 1) * kTQ) - 1;
    const int np = attn_parts(last_pos / kTK + 1);
    if (np == 1) return; // written directly
    const size_t pstride = (size_t) T * Hl * kTPartStride;
    const float* base = partial + ((size_t) t * Hl + h) * kTPartStride;
    float A = 0.0f, L = 0.0f;
    if (np <= 4) { // the usual case (contexts up to 512): every load issued before the first use -- one L2 round trip
        float mp[4], lp[4], op[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const bool live = p < np;
            mp[p] = live ? base[p * pstride + 128] : -INFINITY;
            lp[p] = live ? base[p * pstride + 129] : 0.0f;
            op[p] = live ? base[p * pstride + d] : 0.0f;
        }
        const float M = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (mp[p] == -INFINITY) continue; // a block this row sees nothing of
            const float w = expf(__fsub_rn(mp[p], M));
            L = __fmaf_rn(lp[p], w, L);
            A = __fmaf_rn(op[p], w, A);
        }
    } else {
        float M = -INFINITY;
        for (int p = 0; p < np; ++p) M = fmaxf(M, base[p * pstride + 128]);
        for (int p = 0; p < np; ++p) {
            const float* src = base + p * pstride;
            const float mp = src[128];
            if (mp == -INFINITY) continue;
            const float w = expf(__fsub_rn(mp, M));
            L = __fmaf_rn(src[129], w, L);
            A = __fmaf_rn(src[d], w, A);
        }
    }
    out[((size_t) t * Hl + h) * 128 + d] = __fdiv_rn(A, L);
}

// k_attn_merge and the Q8_0 quantiser of the attention output (the input of wo) in one pass: one block per token, a warp per
// head (lane = 4 output dims), rows of unsplit query tiles are read from `att`, rows of split ones merged from the partials
// with k_attn_merge's arithmetic (same order, same roundings); then the quantiser of k_prep_quant (a half-warp per group).
__global__ void __launch_bounds__(256)
k_attn_merge_quant(const float* __restrict__ partial, const float* __restrict__ att, int8_t* __restrict__ q8, float* __restrict__ xsT,
                   int Hl, int pos0, int T, int Tpad) {
    const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qt = t / kTQ;
    const int last_pos = pos0 + min(T, (qt + 1) * kTQ) - 1;
    const int np = attn_parts(last_pos / kTK + 1);
    const size_t pstride = (size_t) T * Hl * kTPartStride;
    const int n = Hl * 128;
    for (int h = warp; h < Hl; h += 8) {
        float v[4];
        if (np == 1) {
            const float4 a4 = *reinterpret_cast<const float4*>(att + ((size_t) t * Hl + h) * 128 + lane * 4);
            v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w;
        } else {
            const float* base = partial + ((size_t) t * Hl + h) * kTPartStride;
            float A[4] = {0.f, 0.f, 0.f, 0.f}, L = 0.0f;
            if (np <= 4) { // contexts up to 512: every load issued before the first use
                float mp[4], lp[4];
                float4 o4[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const bool live = p < np;
                    mp[p] = live ? base[p * pstride + 128] : -INFINITY;
                    lp[p] = live ? base[p * pstride + 129] : 0.0f;
                    o4[p] = live ? *reinterpret_cast<const float4*>(base + p * pstride + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float M = fmaxf(fmaxf(mp[0], mp[1]), fmaxf(mp[2], mp[3]));
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (mp[p] == -INFINITY) continue; // a block this row sees nothing of
                    const float w = expf(__fsub_rn(mp[p], M));
                    L = __fmaf_rn(lp[p], w, L);
                    A[0] = __fmaf_rn(o4[p].x, w, A[0]); A[1] = __fmaf_rn(o4[p].y, w, A[1]);
                    A[2] = __fmaf_rn(o4[p].z, w, A[2]); A[3] = __fmaf_rn(o4[p].w, w, A[3]);
                }
            } else {
                float M = -INFINITY;
                for (int p = 0; p < np; ++p) M = fmaxf(M, base[p * pstride + 128]);
                for (int p = 0; p < np; ++p) {
                    const float* src = base + p * pstride;
                    const float mp = src[128];
                    if (mp == -INFINITY) continue;
                    const float w = expf(__fsub_rn(mp, M));
                    const float4 o4 = *reinterpret_cast<const float4*>(src + lane * 4);
                    L = __fmaf_rn(src[129], w, L);
                    A[0] = __fmaf_rn(o4.x, w, A[0]); A[1] = __fmaf_rn(o4.y, w, A[1]);
                    A[2] = __fmaf_rn(o4.z, w, A[2]); A[3] = __fmaf_rn(o4.w, w, A[3]);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = __fdiv_rn(A[e], L);
        }
        float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); // within the half-warp = one group of 64
        const float scale = q8_scale(amax);
        const int c0 = q8_code(v[0], scale), c1 = q8_code(v[1], scale), c2 = q8_code(v[2], scale), c3 = q8_code(v[3], scale);
        *reinterpret_cast<uint32_t*>(q8 + (size_t) t * n + h * 128 + lane * 4) =
            (uint32_t) (c0 & 0xff) | ((uint32_t) (c1 & 0xff) << 8) | ((uint32_t) (c2 & 0xff) << 16) | ((uint32_t) (c3 & 0xff) << 24);
        if ((lane & 15) == 0) xsT[(size_t) (2 * h + (lane >> 4)) * Tpad + t] = scale;
    }
}

// residual add over a chunk (forward.c:295-298, 335-338)
__global__ void k_add_rows(float* __restrict__ x, const float* __restrict__ y, size_t n) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fadd_rn(x[i], y[i]);
}

int ensure_bufs(QwenCudaCtx* c, PrefillBufs*& pb, int T) {
    if (!pb) pb = new PrefillBufs();
    if (pb->cap >= T) return 0;
    const int cap = (T + 127) / 128 * 128;
    const size_t maxn = (size_t) std::max(c->D, std::max(c->Pl, c->Hdl));
    void* old[] = {pb->apart, pb->tokens, pb->x, pb->xa, pb->xb, pb->qkv, pb->q, pb->att, pb->h13, pb->q8, pb->xsT, pb->err};
    for (void* o : old)
        if (o) cudaFree(o);
    *pb = PrefillBufs();
    QW_CUDA(cudaMalloc((void**) &pb->tokens, (size_t) cap * 4));
    QW_CUDA(cudaMalloc((void**) &pb->x, (size_t) cap * c->D * 4));
    QW_CUDA(cudaMalloc((void**) &pb->xa, (size_t) cap * c->D * 4));
    QW_CUDA(cudaMalloc((void**) &pb->xb, (size_t) cap * c->D * 4));
    QW_CUDA(cudaMalloc((void**) &pb->qkv, (size_t) cap * (c->Pl + 2 * c->Kl) * 4));
    QW_CUDA(cudaMalloc((void**) &pb->q, (size_t) cap * c->Pl * 4));
    QW_CUDA(cudaMalloc((void**) &pb->att, (size_t) cap * c->Pl * 4));
    QW_CUDA(cudaMalloc((void**) &pb->h13, (size_t) cap * 2 * c->Hdl * 4));
    QW_CUDA(cudaMalloc((void**) &pb->q8, (size_t) cap * maxn));
    QW_CUDA(cudaMalloc((void**) &pb->xsT, (size_t) (maxn / 64) * cap * 4));
    QW_CUDA(cudaMalloc((void**) &pb->err, 4));
    QW_CUDA(cudaMemset(pb->err, 0, 4));
    QW_CUDA(cudaFuncSetAttribute(k_attn_prefill<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    QW_CUDA(cudaFuncSetAttribute(k_attn_prefill<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    QW_CUDA(cudaFuncSetAttribute(k_attn_prefill<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    QW_CUDA(cudaFuncSetAttribute(k_attn_prefill<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    QW_CUDA(cudaFuncSetAttribute(k_attn_prefill_t, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmem));
    pb->cap = cap;
    return 0;
}

} // namespace

static PrefillBufs*& bufs_of(QwenCudaCtx* c) { return *reinterpret_cast<PrefillBufs**>(&c->prefill); }

void qw_prefill_free(QwenCudaCtx* c) {
    PrefillBufs* pb = bufs_of(c);
    if (!pb) return;
    void* old[] = {pb->apart, pb->tokens, pb->x, pb->xa, pb->xb, pb->qkv, pb->q, pb->att, pb->h13, pb->q8, pb->xsT, pb->err};
    for (void* o : old)
        if (o) cudaFree(o);
    delete pb;
    c->prefill = nullptr;
}

// one chunk of T tokens at positions pos0 .. pos0 + T - 1; leaves the chunk's final residual rows in pb->x
static int prefill_chunk(QwenCudaCtx* c, PrefillBufs* pb, const int* tokens_host, int T, int pos0) {
    cudaStream_t st = c->stream;
    const int D = c->D, Pl = c->Pl, Kl = c->Kl, Hdl = c->Hdl, Tpad = pb->cap;
    const int kv_mul = c->Hl / c->KVHl;
    QW_CUDA(cudaMemcpyAsync(pb->tokens, tokens_host, (size_t) T * 4, cudaMemcpyHostToDevice, st));
    k_embed_rows<<<T, 256, 0, st>>>(pb->x, c->w_emb, pb->tokens, D);
    auto gemm = [&](const uint8_t* w, float* out, int d, int n) {
        return qw_prefill_gemm(w, pb->q8, pb->xsT, out, nullptr, d, n, T, Tpad, pb->err, st, nullptr);
    };
    const size_t TD = (size_t) T * D;
    for (int l = 0; l < c->L; ++l) {
        const size_t loff = (size_t) l * c->KVHl * c->S * 128;
        // attention block (forward.c:254-298)
        // the previous layer's w2 output (all-reduced under tensor parallelism) joins the residual stream inside this pass
        k_prep_quant<<<T, 256, 0, st>>>(pb->x, l > 0 ? pb->xb : nullptr, c->att_norm + (size_t) l * D, pb->q8, pb->xsT, D, Tpad, 0);
        if (gemm(c->w_qkv + l * c->w_qkv_stride, pb->qkv, Pl + 2 * Kl, D)) return -1;
        k_qkv_post_rows<<<dim3(T, c->Hl + 2 * c->KVHl), 128, 0, st>>>(pb->qkv, pb->q, c->k_cache + loff, c->v_cache + loff,
                                                                     c->q_norm + (size_t) l * 128, c->k_norm + (size_t) l * 128,
                                                                     c->rope_cos, c->rope_sin, c->Hl, c->KVHl, c->S, pos0);
        static int attn_v = -1;
        if (attn_v < 0) {
            const char* e = getenv("QWEN_ATTN_V"); // 1: the per-warp kernel of round 1; default: the tiled kernel
            attn_v = e ? atoi(e) : 2;
        }
        const dim3 ag((T + 7) / 8, c->KVHl, kv_mul > 4 ? kv_mul / 4 : 1);
        if (attn_v != 1) {
            const AttnPlan plan = attn_plan(pos0, T);
            const int maxparts = plan.parts[0];
            if (maxparts > 1) {
                const size_t need = (size_t) maxparts * T * c->Hl * kTPartStride;
                if (need > pb->apart_floats) {
                    QW_CUDA(cudaStreamSynchronize(st));
                    if (pb->apart) cudaFree(pb->apart);
                    pb->apart = nullptr;
                    pb->apart_floats = 0;
                    QW_CUDA(cudaMalloc((void**) &pb->apart, need * 4));
                    pb->apart_floats = need;
                }
            }
            k_attn_prefill_t<<<(unsigned) (plan.first[plan.nq] * c->Hl), 128, kTSmem, st>>>(pb->q, c->k_cache + loff, c->v_cache + loff, pb->att,
                                                                                          pb->apart, c->Hl, kv_mul, c->S, pos0, T, plan);
            // merge of the split rows + the quantiser of the attention output in one pass
            k_attn_merge_quant<<<T, 256, 0, st>>>(pb->apart, pb->att, pb->q8, pb->xsT, c->Hl, pos0, T, Tpad);
        } else
        switch (kv_mul) {
            case 1: k_attn_prefill<1><<<ag, 256, kAttnSmem, st>>>(pb->q, c->k_cache + loff, c->v_cache + loff, pb->att, c->Hl, c->S, pos0, T); break;
            case 2: k_attn_prefill<2><<<ag, 256, kAttnSmem, st>>>(pb->q, c->k_cache + loff, c->v_cache + loff, pb->att, c->Hl, c->S, pos0, T); break;
            case 4: k_attn_prefill<4><<<ag, 256, kAttnSmem, st>>>(pb->q, c->k_cache + loff, c->v_cache + loff, pb->att, c->Hl, c->S, pos0, T); break;
            case 8: k_attn_prefill<8><<<ag, 256, kAttnSmem, st>>>(pb->q, c->k_cache + loff, c->v_cache + loff, pb->att, c->Hl, c->S, pos0, T); break;
            default: qw_set_error("prefill: unsupported GQA ratio %d", kv_mul); return -2;
        }
        if (attn_v == 1) k_prep_quant<<<T, 256, 0, st>>>(pb->att, nullptr, nullptr, pb->q8, pb->xsT, Pl, Tpad, 0);
        if (gemm(c->w_o + l * c->w_o_stride, pb->xa, D, Pl)) return -1;
        if (qw_tp_allreduce(c, pb->xa, TD)) return -1; // wo is row-parallel under tensor parallelism
        // feed-forward block (forward.c:303-338); x += wo output inside the norm + quantise pass
        k_prep_quant<<<T, 256, 0, st>>>(pb->x, pb->xa, c->ffn_norm + (size_t) l * D, pb->q8, pb->xsT, D, Tpad, 0);
        if (gemm(c->w_13 + l * c->w_13_stride, pb->h13, 2 * Hdl, D)) return -1;
        k_prep_quant<<<T, 256, 0, st>>>(pb->h13, nullptr, nullptr, pb->q8, pb->xsT, Hdl, Tpad, 1);
        if (gemm(c->w_2 + l * c->w_2_stride, pb->xb, D, Hdl)) return -1;
        if (qw_tp_allreduce(c, pb->xb, TD)) return -1;
        if (l == c->L - 1) k_add_rows<<<(unsigned) ((TD + 255) / 256), 256, 0, st>>>(pb->x, pb->xb, TD); // the last one has no norm after it here
    }
    QW_CUDA(cudaGetLastError());
    return 0;
}

// Prefill n tokens at positions pos0 .. pos0 + n - 1 (KV cache rows written for all of them) and leave the
// logits of the LAST token in c->logits (c->logits_all under tensor parallelism), as n forward() calls would.
int qw_prefill(QwenCudaCtx* c, const int* tokens_host, int n, int pos0) {
    if (n <= 0 || pos0 < 0 || pos0 + n > c->S) {
        qw_set_error("prefill: positions %d .. %d outside [0, %d)", pos0, pos0 + n - 1, c->S);
        return -2;
    }
    for (int i = 0; i < n; ++i)
        if (tokens_host[i] < 0 || tokens_host[i] >= c->V) {
            qw_set_error("prefill: token %d outside [0, %d)", tokens_host[i], c->V);
            return -2;
        }
    if (c->D % 64 || c->Pl % 64 || c->Hdl % 64) {
        qw_set_error("prefill: dimensions must be multiples of the Q8_0 group");
        return -2;
    }
    PrefillBufs*& pb = bufs_of(c);
    if (ensure_bufs(c, pb, std::min(n, kChunkTokens))) return -1;
    int T = 0;
    for (int t0 = 0; t0 < n; t0 += T) {
        T = std::min(kChunkTokens, n - t0);
        if (int rc = prefill_chunk(c, pb, tokens_host + t0, T, pos0 + t0)) return rc;
    }
    // final norm + classifier for the last token only (forward.c:344-348)
    cudaStream_t st = c->stream;
    const int D = c->D, pad = qw_pad_cols(D);
    QW_CUDA(cudaMemcpyAsync(c->x, pb->x + (size_t) (T - 1) * D, (size_t) D * 4, cudaMemcpyDeviceToDevice, st));
    launch_rmsnorm(c->x, c->x, c->out_norm, D, st);
    if (pad != D) {
        cudaMemsetAsync(c->aq + D, 0, pad - D, st);
        cudaMemsetAsync(c->as + D / 64, 0, (size_t) (pad - D) / 64 * 4, st);
    }
    launch_quantize(c->x, c->aq, c->as, D, st);
    launch_gemv_sg(c->w_cls, c->aq, c->as, c->logits, c->Vl, D, nullptr, st);
    if (c->tp_size > 1 && qw_tp_allgather(c, c->logits, c->logits_all, c->Vl)) return -1;
    QW_CUDA(cudaGetLastError());
    int herr = 0;
    QW_CUDA(cudaMemcpyAsync(&herr, pb->err, 4, cudaMemcpyDeviceToHost, st));
    QW_CUDA(cudaStreamSynchronize(st));
    if (herr) {
        qw_set_error("prefill: GEMM pipeline wait %d timed out", herr);
        cudaMemsetAsync(pb->err, 0, 4, st);
        return -3;
    }
    return 0;
}

// Test hook: the chunk attention kernels on host data. q [T][Hl][128]; k, v [pos][KVHl * 128] (the reference's cache layout,
// positions 0 .. pos0 + T - 1); out [T][Hl][128]. variant 1 = per-warp kernel, 2 = tiled kernel. Token t sits at pos0 + t.
extern "C" int qwen_cuda_debug_attn_prefill(float* out, const float* q, const float* k, const float* v, int Hl, int KVHl, int pos0,
                                            int T, int variant) {
    if (qwen_cuda_device_count() <= 0) {
        qw_set_error("no CUDA device: this library has no CPU path");
        return -1;
    }
    const int kv_mul = KVHl > 0 ? Hl / KVHl : 0, S = pos0 + T;
    if (T <= 0 || pos0 < 0 || KVHl <= 0 || Hl % KVHl || (variant == 1 && kv_mul != 1 && kv_mul != 2 && kv_mul != 4 && kv_mul != 8)) {
        qw_set_error("debug_attn_prefill: bad shape");
        return -2;
    }
    float *dq = nullptr, *dk = nullptr, *dv = nullptr, *dout = nullptr;
    const size_t nq = (size_t) T * Hl * 128, nkv = (size_t) KVHl * S * 128;
    // device layout [kv head][pos][128]
    float* hk = (float*) malloc(nkv * 4);
    float* hv = (float*) malloc(nkv * 4);
    int rc = -1;
    do {
        if (!hk || !hv) break;
        for (int p = 0; p < S; ++p)
            for (int h = 0; h < KVHl; ++h) {
                memcpy(hk + ((size_t) h * S + p) * 128, k + ((size_t) p * KVHl + h) * 128, 512);
                memcpy(hv + ((size_t) h * S + p) * 128, v + ((size_t) p * KVHl + h) * 128, 512);
            }
        if (cudaMalloc(&dq, nq * 4) || cudaMalloc(&dout, nq * 4) || cudaMalloc(&dk, nkv * 4) || cudaMalloc(&dv, nkv * 4)) break;
        cudaMemcpy(dq, q, nq * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dk, hk, nkv * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dv, hv, nkv * 4, cudaMemcpyHostToDevice);
        cudaMemset(dout, 0xff, nq * 4);
        if (variant == 1) {
            const dim3 ag((T + 7) / 8, KVHl, kv_mul > 4 ? kv_mul / 4 : 1);
            cudaFuncSetAttribute(k_attn_prefill<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
            cudaFuncSetAttribute(k_attn_prefill<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
            cudaFuncSetAttribute(k_attn_prefill<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
            cudaFuncSetAttribute(k_attn_prefill<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
            switch (kv_mul) {
                case 1: k_attn_prefill<1><<<ag, 256, kAttnSmem>>>(dq, dk, dv, dout, Hl, S, pos0, T); break;
                case 2: k_attn_prefill<2><<<ag, 256, kAttnSmem>>>(dq, dk, dv, dout, Hl, S, pos0, T); break;
                case 4: k_attn_prefill<4><<<ag, 256, kAttnSmem>>>(dq, dk, dv, dout, Hl, S, pos0, T); break;
                default: k_attn_prefill<8><<<ag, 256, kAttnSmem>>>(dq, dk, dv, dout, Hl, S, pos0, T); break;
            }
        } else {
            cudaFuncSetAttribute(k_attn_prefill_t, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmem);
            // chunks of at most kChunkTokens tokens, as forward_prefill cuts them
            for (int c0 = 0; c0 < T; c0 += kChunkTokens) {
                const int Tc = std::min(kChunkTokens, T - c0);
                const AttnPlan plan = attn_plan(pos0 + c0, Tc);
                float* dpart = nullptr;
                if (plan.parts[0] > 1 && cudaMalloc(&dpart, (size_t) plan.parts[0] * Tc * Hl * kTPartStride * 4)) break;
                k_attn_prefill_t<<<(unsigned) (plan.first[plan.nq] * Hl), 128, kTSmem>>>(dq + (size_t) c0 * Hl * 128, dk, dv, dout + (size_t) c0 * Hl * 128,
                                                                                        dpart, Hl, kv_mul, S, pos0 + c0, Tc, plan);
                if (plan.parts[0] > 1) k_attn_merge<<<dim3((unsigned) Tc, (unsigned) Hl), 128>>>(dpart, dout + (size_t) c0 * Hl * 128, Hl, pos0 + c0, Tc, 0);
                cudaDeviceSynchronize();
                cudaFree(dpart);
            }
        }
        if (cudaDeviceSynchronize() != cudaSuccess) {
            qw_set_error("debug_attn_prefill: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaMemcpy(out, dout, nq * 4, cudaMemcpyDeviceToHost);
        rc = 0;
    } while (0);
    free(hk); free(hv);
    cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout);
    return rc;
}

// Host-only test hook (no GPU needed): the fixed key blocks of the chunk attention. For a chunk of T tokens at pos0 fills,
// per query tile qt (ascending), parts[qt] and the blocks' key-tile ranges k01[qt][part][2] (at most max_parts per tile).
// Returns the number of query tiles, or a negative value.
extern "C" int qwen_cuda_debug_attn_plan(int pos0, int T, int* parts, int* k01, int max_parts) {
    if (pos0 < 0 || T <= 0 || T > kChunkTokens || !parts || !k01 || max_parts <= 0) return -2;
    const AttnPlan pl = attn_plan(pos0, T);
    for (int i = 0; i < pl.nq; ++i) {
        const int qt = pl.nq - 1 - i;
        parts[qt] = pl.parts[i];
        if (pl.parts[i] > max_parts) return -3;
        for (int p = 0; p < pl.parts[i]; ++p) attn_part_range(p, k01[(qt * max_parts + p) * 2], k01[(qt * max_parts + p) * 2 + 1]);
    }
    return pl.nq;
}
