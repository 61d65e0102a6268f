// attn_core.cuh -- the per-warp online-softmax attention step shared by the persistent decode kernel
// (decode_mega.cu) and the prefill attention kernel (prefill.cu). Reference: src/forward.c:141-195.
#pragma once

#include "common.cuh"

// Online-softmax state of one warp for HW query heads: lane owns output dims 4*lane .. 4*lane+3 of
// every head; m and l are kept per head in every lane (identical across the warp).
template <int HW>
struct AttnState {
    float4 acc[HW];
    float m[HW], l[HW];
};
template <int HW>
__device__ __forceinline__ void attn_state_reset(AttnState<HW>& st) {
#pragma unroll
    for (int j = 0; j < HW; ++j) {
        st.acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        st.m[j] = -INFINITY;
        st.l[j] = 0.0f;
    }
}

// One warp attends over `cnt` positions (K rows 128 floats each in shared memory, V rows voff floats later)
// for HW query heads q[j] (this lane's 4 dims of each head). Scores (forward.c:156-165): every lane
// forms its 4-dim partial dot for PB = 32 / HW positions x HW heads = 32 values, and a butterfly
// transpose-reduction (31 shuffles) leaves lane i with the full dot of value i = (position i / HW,
// head i % HW). Softmax statistics per head by shuffles over the lanes of that head, then P.V with
// the probability broadcast from its lane. No shared-memory traffic besides one LDS.128 per K row
// and per V row, and no synchronisation with other warps.
template <int HW>
__device__ __forceinline__ void attn_rows(const float* Ka, int cnta, const float* Kb, int cnt, int voff, const float4 (&q)[HW], AttnState<HW>& st, int lane) {
    // rows 0 .. cnta-1 live at Ka, rows cnta .. cnt-1 at Kb (the next ring slot); V rows sit voff floats after K
    constexpr int PB = 32 / HW;
    const float inv = sqrtf(128.0f);
    Kb -= cnta * 128;
#pragma unroll 1
    for (int b0 = 0; b0 < cnt; b0 += PB) {
        const int nb = min(PB, cnt - b0);
        float v[32];
#pragma unroll
        for (int pp = 0; pp < PB; ++pp) {
            const int r = b0 + min(pp, nb - 1); // rows past the end recompute a valid row (masked below)
            const float4 kf = *reinterpret_cast<const float4*>((r < cnta ? Ka : Kb) + r * 128 + lane * 4);
#pragma unroll
            for (int j = 0; j < HW; ++j) {
                float d = __fmul_rn(q[j].x, kf.x);
                d = __fmaf_rn(q[j].y, kf.y, d);
                d = __fmaf_rn(q[j].z, kf.z, d);
                d = __fmaf_rn(q[j].w, kf.w, d);
                v[pp * HW + j] = d;
            }
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const bool hi = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
                const float send = hi ? v[i] : v[i + s];
                const float keep = hi ? v[i + s] : v[i];
                v[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, s));
            }
        }
        const int pp = lane / HW;
        const float sc = pp < nb ? __fdiv_rn(v[0], inv) : -INFINITY; // score / sqrtf(head_dim)
        float bm = sc;
#pragma unroll
        for (int o = HW; o < 32; o <<= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        // every lane learns the block max of every head (lane j holds head j's)
        float e = 0.0f;
#pragma unroll
        for (int j = 0; j < HW; ++j) {
            const float bmj = __shfl_sync(0xffffffffu, bm, j);
            const float m_new = fmaxf(st.m[j], bmj);
            const float scl = (st.m[j] == -INFINITY) ? 0.0f : expf(__fsub_rn(st.m[j], m_new));
            if ((lane % HW) == j) e = pp < nb ? expf(__fsub_rn(sc, m_new)) : 0.0f;
            st.m[j] = m_new;
            st.l[j] = __fmul_rn(st.l[j], scl);
            st.acc[j].x = __fmul_rn(st.acc[j].x, scl);
            st.acc[j].y = __fmul_rn(st.acc[j].y, scl);
            st.acc[j].z = __fmul_rn(st.acc[j].z, scl);
            st.acc[j].w = __fmul_rn(st.acc[j].w, scl);
        }
        float es = e;
#pragma unroll
        for (int o = HW; o < 32; o <<= 1) es = __fadd_rn(es, __shfl_xor_sync(0xffffffffu, es, o));
#pragma unroll
        for (int j = 0; j < HW; ++j) st.l[j] = __fadd_rn(st.l[j], __shfl_sync(0xffffffffu, es, j));
#pragma unroll 2
        for (int i = 0; i < nb; ++i) { // uniform across the warp
            const int r = b0 + i;
            const float4 vv = *reinterpret_cast<const float4*>((r < cnta ? Ka : Kb) + voff + r * 128 + lane * 4);
#pragma unroll
            for (int j = 0; j < HW; ++j) {
                const float pw = __shfl_sync(0xffffffffu, e, i * HW + j);
                st.acc[j].x = __fmaf_rn(pw, vv.x, st.acc[j].x);
                st.acc[j].y = __fmaf_rn(pw, vv.y, st.acc[j].y);
                st.acc[j].z = __fmaf_rn(pw, vv.z, st.acc[j].z);
                st.acc[j].w = __fmaf_rn(pw, vv.w, st.acc[j].w);
            }
        }
    }
}


// ONE position (K row k, V row v: this lane's 4 dims of each) for HW query heads, without the 8-position block
// machinery of attn_rows: the decode step's own position, which would otherwise cost a whole block of masked work on
// the critical path of the blocks every combine task waits for. Same operations per head as attn_rows (the xor tree
// adds the 32 lane partials in the same pairing as the butterfly).
template <int HW>
__device__ __forceinline__ void attn_one_row(const float4 kf, const float4 vv, const float4 (&q)[HW], AttnState<HW>& st) {
    const float inv = sqrtf(128.0f);
#pragma unroll
    for (int j = 0; j < HW; ++j) {
        float d = __fmul_rn(q[j].x, kf.x);
        d = __fmaf_rn(q[j].y, kf.y, d);
        d = __fmaf_rn(q[j].z, kf.z, d);
        d = __fmaf_rn(q[j].w, kf.w, d);
        d = warp_sum(d);
        const float sc = __fdiv_rn(d, inv);
        const float m_new = fmaxf(st.m[j], sc);
        const float scl = (st.m[j] == -INFINITY) ? 0.0f : expf(__fsub_rn(st.m[j], m_new));
        const float e = expf(__fsub_rn(sc, m_new));
        st.m[j] = m_new;
        st.l[j] = __fadd_rn(__fmul_rn(st.l[j], scl), e);
        st.acc[j].x = __fmaf_rn(e, vv.x, __fmul_rn(st.acc[j].x, scl));
        st.acc[j].y = __fmaf_rn(e, vv.y, __fmul_rn(st.acc[j].y, scl));
        st.acc[j].z = __fmaf_rn(e, vv.z, __fmul_rn(st.acc[j].z, scl));
        st.acc[j].w = __fmaf_rn(e, vv.w, __fmul_rn(st.acc[j].w, scl));
    }
}

// Split form of attn_rows for ONE block of up to PB = 32 / HW positions whose K rows and V rows arrive separately
// (per-warp streaming in decode_mega.cu: the K rows' shared memory is handed back before the V rows are needed).
// attn_block_scores folds the block's K rows (K: nb rows of 128 floats) into the running softmax statistics and returns
// this lane's probability e -- value i = (position i / HW, head i % HW) lives in lane i -- and attn_block_pv adds e * V.
// Same operations in the same order as one iteration of attn_rows.
template <int HW>
__device__ __forceinline__ float attn_block_scores(const float* K, int nb, const float4 (&q)[HW], AttnState<HW>& st, int lane) {
    constexpr int PB = 32 / HW;
    const float inv = sqrtf(128.0f);
    float v[32];
#pragma unroll
    for (int pp = 0; pp < PB; ++pp) {
        const int r = min(pp, nb - 1); // rows past the end recompute a valid row (masked below)
        const float4 kf = *reinterpret_cast<const float4*>(K + r * 128 + lane * 4);
#pragma unroll
        for (int j = 0; j < HW; ++j) {
            float d = __fmul_rn(q[j].x, kf.x);
            d = __fmaf_rn(q[j].y, kf.y, d);
            d = __fmaf_rn(q[j].z, kf.z, d);
            d = __fmaf_rn(q[j].w, kf.w, d);
            v[pp * HW + j] = d;
        }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool hi = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = hi ? v[i] : v[i + s];
            const float keep = hi ? v[i + s] : v[i];
            v[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, s));
        }
    }
    const int pp = lane / HW;
    const float sc = pp < nb ? __fdiv_rn(v[0], inv) : -INFINITY; // score / sqrtf(head_dim)
    float bm = sc;
#pragma unroll
    for (int o = HW; o < 32; o <<= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    float e = 0.0f;
#pragma unroll
    for (int j = 0; j < HW; ++j) {
        const float bmj = __shfl_sync(0xffffffffu, bm, j);
        const float m_new = fmaxf(st.m[j], bmj);
        const float scl = (st.m[j] == -INFINITY) ? 0.0f : expf(__fsub_rn(st.m[j], m_new));
        if ((lane % HW) == j) e = pp < nb ? expf(__fsub_rn(sc, m_new)) : 0.0f;
        st.m[j] = m_new;
        st.l[j] = __fmul_rn(st.l[j], scl);
        st.acc[j].x = __fmul_rn(st.acc[j].x, scl);
        st.acc[j].y = __fmul_rn(st.acc[j].y, scl);
        st.acc[j].z = __fmul_rn(st.acc[j].z, scl);
        st.acc[j].w = __fmul_rn(st.acc[j].w, scl);
    }
    float es = e;
#pragma unroll
    for (int o = HW; o < 32; o <<= 1) es = __fadd_rn(es, __shfl_xor_sync(0xffffffffu, es, o));
#pragma unroll
    for (int j = 0; j < HW; ++j) st.l[j] = __fadd_rn(st.l[j], __shfl_sync(0xffffffffu, es, j));
    return e;
}
template <int HW>
__device__ __forceinline__ void attn_block_pv(const float* V, int nb, float e, AttnState<HW>& st, int lane) {
#pragma unroll 2
    for (int i = 0; i < nb; ++i) { // uniform across the warp
        const float4 vv = *reinterpret_cast<const float4*>(V + i * 128 + lane * 4);
#pragma unroll
        for (int j = 0; j < HW; ++j) {
            const float pw = __shfl_sync(0xffffffffu, e, i * HW + j);
            st.acc[j].x = __fmaf_rn(pw, vv.x, st.acc[j].x);
            st.acc[j].y = __fmaf_rn(pw, vv.y, st.acc[j].y);
            st.acc[j].z = __fmaf_rn(pw, vv.z, st.acc[j].z);
            st.acc[j].w = __fmaf_rn(pw, vv.w, st.acc[j].w);
        }
    }
}
