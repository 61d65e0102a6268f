// decode_ops.cu -- the decode step as one kernel per op (debug / cross-check path).
//
// Same arithmetic as the persistent kernel (decode_mega.cu), but every op of
// reference src/forward.c:225-350 is its own launch and every intermediate lands in
// global memory where qwen_cuda_debug_read can fetch it. Selected with
// qwen_cuda_set_path(ctx, 1). Also hosts the attention kernels behind the op-level
// attention() wrapper.
#include "common.cuh"

// K1: embedding row, dequantised on the fly from the SG layout (forward.c:237, q8.c:35)
__global__ void k_embed(float* __restrict__ x, const uint8_t* __restrict__ w_emb, int token,
                        const int* __restrict__ token_dev, int D) {
    const int tok = token_dev ? *token_dev : token;
    const uint8_t* row = w_emb + (size_t) tok * qw_row_bytes(D);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) {
        const uint8_t* rec = row + (size_t) (i >> 8) * QW_SG_BYTES;
        const int within = i & 255;
        const float sc = reinterpret_cast<const float*>(rec + 256)[within >> 6];
        x[i] = __fmul_rn((float) reinterpret_cast<const int8_t*>(rec)[within], sc);
    }
}

// K5: per-head q/k RMSNorm + RoPE, K and V written into cache slot `pos`
// (forward.c:267-280, 244-248). One block of 128 threads per head.
// blocks [0,Hl): q heads; [Hl, Hl+KVHl): k heads; [Hl+KVHl, Hl+2KVHl): v heads.
__global__ void __launch_bounds__(128)
k_qkv_post(const float* __restrict__ qkv, float* __restrict__ q_out, float* __restrict__ k_layer,
           float* __restrict__ v_layer, const float* __restrict__ gq, const float* __restrict__ gk,
           const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int Hl, int KVHl, int S,
           int pos) {
    __shared__ float y[128];
    __shared__ float red[4];
    const int b = blockIdx.x, t = threadIdx.x;
    const float v = qkv[(size_t) b * 128 + t];
    if (b >= Hl + KVHl) { // V: stored raw
        const int h = b - Hl - KVHl;
        v_layer[((size_t) h * S + pos) * 128 + t] = v;
        return;
    }
    float ss = warp_sum(__fmul_rn(v, v));
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    ss = __fadd_rn(__fadd_rn(red[0], red[1]), __fadd_rn(red[2], red[3]));
    const float r = rms_rscale(ss, 128);
    const float* g = (b < Hl) ? gq : gk;
    y[t] = __fmul_rn(g[t], __fmul_rn(r, v));
    __syncthreads();
    const int i = t & 63;
    const float c = rope_cos[(size_t) pos * 64 + i], s = rope_sin[(size_t) pos * 64 + i];
    const float a = y[i], bb = y[i + 64];
    const float o = (t < 64) ? __fsub_rn(__fmul_rn(a, c), __fmul_rn(bb, s))
                             : __fadd_rn(__fmul_rn(a, s), __fmul_rn(bb, c));
    if (b < Hl)
        q_out[(size_t) b * 128 + t] = o;
    else
        k_layer[((size_t) (b - Hl) * S + pos) * 128 + t] = o;
}

// K6a: split-KV attention partials. grid (KVHl, nsplit), 256 threads.
// Serves the kv_mul query heads that share one KV head (GQA, forward.c:146,158).
__global__ void __launch_bounds__(256)
k_attn_partial(const float* __restrict__ q, const float* __restrict__ k_layer, const float* __restrict__ v_layer,
               float* __restrict__ part_m, float* __restrict__ part_l, float* __restrict__ part_acc, int kv_mul,
               int S, int pos, int chunk, int max_splits) {
    extern __shared__ float sm[];
    float* sq = sm;                 // [kv_mul][128]
    float* ss = sm + kv_mul * 128;  // [kv_mul][chunk]
    const int kvh = blockIdx.x, sp = blockIdx.y;
    const int p0 = sp * chunk;
    const int p1 = min(pos + 1, p0 + chunk);
    const int cnt = p1 - p0;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* K = k_layer + (size_t) kvh * S * 128;
    const float* V = v_layer + (size_t) kvh * S * 128;
    for (int i = t; i < kv_mul * 128; i += 256) sq[i] = q[(size_t) kvh * kv_mul * 128 + i];
    __syncthreads();
    const float inv = sqrtf(128.0f);
    for (int i = warp; i < cnt; i += 8) {
        const float4 kv = *reinterpret_cast<const float4*>(K + (size_t) (p0 + i) * 128 + lane * 4);
        for (int j = 0; j < kv_mul; ++j) {
            const float4 qv = *reinterpret_cast<const float4*>(sq + j * 128 + lane * 4);
            float d = __fmul_rn(qv.x, kv.x);
            d = __fmaf_rn(qv.y, kv.y, d);
            d = __fmaf_rn(qv.z, kv.z, d);
            d = __fmaf_rn(qv.w, kv.w, d);
            d = warp_sum(d);
            if (lane == 0) ss[j * chunk + i] = __fdiv_rn(d, inv); // score / sqrtf(head_dim), forward.c:164
        }
    }
    __syncthreads();
    for (int j = warp; j < kv_mul; j += 8) {
        float m = -INFINITY;
        for (int i = lane; i < cnt; i += 32) m = fmaxf(m, ss[j * chunk + i]);
        m = warp_max(m);
        float l = 0.0f;
        for (int i = lane; i < cnt; i += 32) {
            const float e = expf(__fsub_rn(ss[j * chunk + i], m));
            ss[j * chunk + i] = e;
            l = __fadd_rn(l, e);
        }
        l = warp_sum(l);
        if (lane == 0) {
            const int h = kvh * kv_mul + j;
            part_m[h * max_splits + sp] = m;
            part_l[h * max_splits + sp] = l;
        }
    }
    __syncthreads();
    for (int idx = t; idx < kv_mul * 128; idx += 256) {
        const int j = idx >> 7, d = idx & 127;
        float acc = 0.0f;
        for (int i = 0; i < cnt; ++i) acc = __fmaf_rn(ss[j * chunk + i], V[(size_t) (p0 + i) * 128 + d], acc);
        const int h = kvh * kv_mul + j;
        part_acc[((size_t) h * max_splits + sp) * 128 + d] = acc;
    }
}

// K6b: combine the splits of one head (online-softmax merge). grid Hl, 128 threads.
__global__ void __launch_bounds__(128)
k_attn_combine(float* __restrict__ out, const float* __restrict__ part_m, const float* __restrict__ part_l,
               const float* __restrict__ part_acc, int nsplit, int max_splits) {
    const int h = blockIdx.x, d = threadIdx.x;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_m[h * max_splits + s]);
    float L = 0.0f, acc = 0.0f;
    for (int s = 0; s < nsplit; ++s) {
        const float w = expf(__fsub_rn(part_m[h * max_splits + s], M));
        L = __fmaf_rn(part_l[h * max_splits + s], w, L);
        acc = __fmaf_rn(part_acc[((size_t) h * max_splits + s) * 128 + d], w, acc);
    }
    out[(size_t) h * 128 + d] = __fdiv_rn(acc, L);
}

// K8: residual add (forward.c:295-298, 335-338)
__global__ void k_add(float* __restrict__ x, const float* __restrict__ y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fadd_rn(x[i], y[i]);
}

// K7 on the interleaved w1/w3 output: h[i] = silu(h13[2i]) * h13[2i+1] (forward.c:319-321)
__global__ void k_swiglu_pairs(float* __restrict__ h, const float* __restrict__ h13, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) h[i] = __fmul_rn(silu_ref(h13[2 * i]), h13[2 * i + 1]);
}

static void quantize_padded(QwenCudaCtx* c, const float* src, int n) {
    const int pad = qw_pad_cols(n);
    if (pad != n) {
        cudaMemsetAsync(c->aq + n, 0, pad - n, c->stream);
        cudaMemsetAsync(c->as + n / 64, 0, (size_t) (pad - n) / 64 * 4, c->stream);
    }
    launch_quantize(src, c->aq, c->as, n, c->stream);
}

int qw_attention_device(QwenCudaCtx* c, int layer, int pos, const float* q_dev, float* out_dev) {
    const int kv_mul = c->Hl / c->KVHl;
    int nsplit = (pos + 1 + 31) / 32;
    if (nsplit > c->att_max_splits) nsplit = c->att_max_splits;
    const int chunk = (pos + 1 + nsplit - 1) / nsplit;
    nsplit = (pos + 1 + chunk - 1) / chunk;
    const size_t smem = (size_t) kv_mul * (128 + chunk) * sizeof(float);
    if (smem > 48 * 1024) {
        qw_set_error("attention: chunk too large for shared memory (pos %d)", pos);
        return -2;
    }
    const size_t loff = (size_t) layer * c->KVHl * c->S * 128;
    k_attn_partial<<<dim3(c->KVHl, nsplit), 256, smem, c->stream>>>(q_dev, c->k_cache + loff, c->v_cache + loff,
                                                                    c->att_m, c->att_l, c->att_acc, kv_mul, c->S,
                                                                    pos, chunk, c->att_max_splits);
    k_attn_combine<<<c->Hl, 128, 0, c->stream>>>(out_dev, c->att_m, c->att_l, c->att_acc, nsplit,
                                                 c->att_max_splits);
    return 0;
}

int qw_decode_ops(QwenCudaCtx* c, int token, const int* token_dev, int pos) {
    cudaStream_t st = c->stream;
    const int D = c->D, Pl = c->Pl, Kl = c->Kl, Hdl = c->Hdl;
    const int layers = (c->layers_run >= 0 && c->layers_run <= c->L) ? c->layers_run : c->L;
    const int l0 = (c->layer_begin > 0 && c->layer_begin <= layers) ? c->layer_begin : 0; // debug window (qwen_cuda_debug_set_window)
    if (c->x_inject_on) QW_CUDA(cudaMemcpyAsync(c->x, c->x_inject, (size_t) D * 4, cudaMemcpyDeviceToDevice, st));
    else k_embed<<<(D + 255) / 256, 256, 0, st>>>(c->x, c->w_emb, token, token_dev, D);
    for (int l = l0; l < layers; ++l) {
        const size_t loff = (size_t) l * c->KVHl * c->S * 128;
        launch_rmsnorm(c->xb, c->x, c->att_norm + (size_t) l * D, D, st);
        quantize_padded(c, c->xb, D);
        launch_gemv_sg(c->w_qkv + l * c->w_qkv_stride, c->aq, c->as, c->qkv, Pl + 2 * Kl, D, nullptr, st);
        k_qkv_post<<<c->Hl + 2 * c->KVHl, 128, 0, st>>>(c->qkv, c->q, c->k_cache + loff, c->v_cache + loff,
                                                       c->q_norm + (size_t) l * 128, c->k_norm + (size_t) l * 128,
                                                       c->rope_cos, c->rope_sin, c->Hl, c->KVHl, c->S, pos);
        if (qw_attention_device(c, l, pos, c->q, c->att)) return -1;
        quantize_padded(c, c->att, Pl);
        launch_gemv_sg(c->w_o + l * c->w_o_stride, c->aq, c->as, c->xb, D, Pl, nullptr, st);
        if (qw_tp_allreduce(c, c->xb, D)) return -1; // wo is row-parallel: partial sums over the head slices
        k_add<<<(D + 255) / 256, 256, 0, st>>>(c->x, c->xb, D);
        launch_rmsnorm(c->xb, c->x, c->ffn_norm + (size_t) l * D, D, st);
        quantize_padded(c, c->xb, D);
        launch_gemv_sg(c->w_13 + l * c->w_13_stride, c->aq, c->as, c->h13, 2 * Hdl, D, nullptr, st);
        k_swiglu_pairs<<<(Hdl + 255) / 256, 256, 0, st>>>(c->h, c->h13, Hdl);
        quantize_padded(c, c->h, Hdl);
        launch_gemv_sg(c->w_2 + l * c->w_2_stride, c->aq, c->as, c->xb, D, Hdl, nullptr, st);
        if (qw_tp_allreduce(c, c->xb, D)) return -1; // w2 is row-parallel over the hidden slices
        k_add<<<(D + 255) / 256, 256, 0, st>>>(c->x, c->xb, D);
    }
    launch_rmsnorm(c->x, c->x, c->out_norm, D, st);
    quantize_padded(c, c->x, D);
    launch_gemv_sg(c->w_cls, c->aq, c->as, c->logits, c->Vl, D, nullptr, st);
    if (c->tp_size > 1 && qw_tp_allgather(c, c->logits, c->logits_all, c->Vl)) return -1; // classifier is split over V
    QW_CUDA(cudaGetLastError());
    return 0;
}

// launches per token on this path (for bench.py's gpu_launches claim)
int qw_decode_ops_launches(const QwenCudaCtx* c) { return 1 + c->L * 14 + 3; } // + 2 NCCL kernels per layer when tp_size > 1
