// context.cu -- device context: upload + repack of a checkpoint, KV cache, the
// forward entry points of include/qwen_cuda.h. Stands behind model_create /
// model_free / forward (reference: src/model.c:162-282, 321-406, 491-500;
// src/forward.c:225-350).
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <thread>
#include <vector>

#include "common.cuh"

int qw_attention_device(QwenCudaCtx* c, int layer, int pos, const float* q_dev, float* out_dev);
int qw_decode_ops_launches(const QwenCudaCtx* c);
int qw_decode_mega_launches(const QwenCudaCtx* c);

namespace {

template <typename T>
int dev_alloc(T** p, size_t count, bool zero = true) {
    QW_CUDA(cudaMalloc((void**) p, std::max<size_t>(count * sizeof(T), 16)));
    if (zero) QW_CUDA(cudaMemset(*p, 0, std::max<size_t>(count * sizeof(T), 16)));
    return 0;
}

int upload_f32(float** dst, const float* src, size_t n) {
    if (dev_alloc(dst, n, false)) return -1;
    QW_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// Upload pipeline (the reference maps the file and dequantises the whole embedding table on the host,
// src/model.c:19-48, 199-206). A tensor slice travels in chunks of <= 32 MB of codes:
//   host threads GATHER it from the page-cache mapping into one of three pinned staging buffers -- only the rows and the
//   column window this rank owns, so a tensor-parallel rank reads 1/tp of every matrix instead of all of it --,
//   cudaMemcpyAsync (pinned source: truly asynchronous, unlike the pageable mapping) into a device staging buffer,
//   k_repack into the SG layout, all on the context's stream; an event per buffer says when it may be refilled.
// The gather of chunk i+1 overlaps the copy and repack of chunk i.
struct Uploader {
    static constexpr int kBufs = 3;
    static constexpr size_t kCodes = (size_t) 32 << 20;          // codes per chunk
    uint8_t* pin[kBufs] = {};
    int8_t* dq[kBufs] = {};
    float* ds[kBufs] = {};
    cudaEvent_t done[kBufs] = {};
    bool used[kBufs] = {};
    int next = 0, nthreads = 1;
    size_t bytes_read = 0;
    ~Uploader() {
        for (int i = 0; i < kBufs; ++i) {
            if (pin[i]) cudaFreeHost(pin[i]);
            if (dq[i]) cudaFree(dq[i]);
            if (ds[i]) cudaFree(ds[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
    }
    int init() {
        for (int i = 0; i < kBufs; ++i) {
            QW_CUDA(cudaHostAlloc((void**) &pin[i], kCodes + kCodes / 16, cudaHostAllocDefault));
            QW_CUDA(cudaMalloc((void**) &dq[i], kCodes));
            QW_CUDA(cudaMalloc((void**) &ds[i], kCodes / 16));
            QW_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        nthreads = (int) std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        if (const char* e = getenv("QWEN_CUDA_UPLOAD_THREADS")) nthreads = std::max(1, atoi(e));
        return 0;
    }
    // rows [row0, row0 + rows) x columns [col0, col0 + n) of a [*][src_n] tensor -> dst rows dst_row0 + r * dst_row_step
    int put(const QwenCudaQ8& t, int src_n, int row0, int rows, int col0, int n, uint8_t* dst, int dst_row0, int dst_row_step,
            cudaStream_t stream) {
        const int rows_per_chunk = (int) std::max<size_t>(1, kCodes / (size_t) n);
        for (int r = 0; r < rows; r += rows_per_chunk) {
            const int nr = std::min(rows_per_chunk, rows - r);
            const int b = next;
            next = (next + 1) % kBufs;
            if (used[b]) QW_CUDA(cudaEventSynchronize(done[b]));
            int8_t* hq = reinterpret_cast<int8_t*>(pin[b]);
            float* hs = reinterpret_cast<float*>(pin[b] + kCodes);
            const int8_t* sq = t.q + (size_t) (row0 + r) * src_n + col0;
            const float* ss = t.s + ((size_t) (row0 + r) * src_n + col0) / 64;
            auto gather = [=](int a, int e) {
                if (n == src_n) { // whole rows: one contiguous range
                    memcpy(hq + (size_t) a * n, sq + (size_t) a * src_n, (size_t) (e - a) * n);
                    memcpy(hs + (size_t) a * (n / 64), ss + (size_t) a * (src_n / 64), (size_t) (e - a) * (n / 64) * 4);
                } else {
                    for (int i = a; i < e; ++i) {
                        memcpy(hq + (size_t) i * n, sq + (size_t) i * src_n, (size_t) n);
                        memcpy(hs + (size_t) i * (n / 64), ss + (size_t) i * (src_n / 64), (size_t) (n / 64) * 4);
                    }
                }
            };
            const int T = (size_t) nr * n < ((size_t) 1 << 20) ? 1 : std::min(nthreads, nr);
            if (T <= 1) {
                gather(0, nr);
            } else {
                std::vector<std::thread> th;
                for (int k = 1; k < T; ++k) th.emplace_back(gather, (int) ((long long) nr * k / T), (int) ((long long) nr * (k + 1) / T));
                gather(0, (int) ((long long) nr / T));
                for (auto& x : th) x.join();
            }
            const size_t codes = (size_t) nr * n;
            bytes_read += codes + codes / 16;
            QW_CUDA(cudaMemcpyAsync(dq[b], hq, codes, cudaMemcpyHostToDevice, stream));
            QW_CUDA(cudaMemcpyAsync(ds[b], hs, codes / 16, cudaMemcpyHostToDevice, stream));
            launch_repack(dq[b], ds[b], n, 0, n, nr, dst, dst_row0 + r * dst_row_step, dst_row_step, stream);
            QW_CUDA(cudaGetLastError());
            QW_CUDA(cudaEventRecord(done[b], stream));
            used[b] = true;
        }
        return 0;
    }
};

} // namespace

extern "C" QwenCudaCtx* qwen_cuda_create(const QwenCudaModelDesc* m, int device, QwenCudaTp tp) {
    if (!m) {
        qw_set_error("qwen_cuda_create: null descriptor");
        return nullptr;
    }
    if (qwen_cuda_device_count() <= device || device < 0) {
        qw_set_error("qwen_cuda_create: CUDA device %d not available (%d visible); there is no CPU path", device,
                     qwen_cuda_device_count());
        return nullptr;
    }
    if (m->group_size != 64 || m->head_dim != 128) {
        qw_set_error("qwen_cuda_create: only block_size 64 / head_dim 128 checkpoints are supported (got %d / %d)",
                     m->group_size, m->head_dim);
        return nullptr;
    }
    if (tp.size < 1 || tp.rank < 0 || tp.rank >= tp.size || m->n_kv_heads % tp.size || m->n_heads % tp.size
        || m->vocab_size % tp.size || m->hidden_dim % (64 * tp.size) || m->dim % 64
        || m->n_heads % m->n_kv_heads || (m->n_heads / m->n_kv_heads) > 8) {
        qw_set_error("qwen_cuda_create: shape not divisible for tp=%d (heads %d/%d, hidden %d, vocab %d, dim %d)",
                     tp.size, m->n_heads, m->n_kv_heads, m->hidden_dim, m->vocab_size, m->dim);
        return nullptr;
    }
    QW_CUDA_NULL(cudaSetDevice(device));
    QwenCudaCtx* c = new QwenCudaCtx();
    memset(c, 0, sizeof *c);
    c->device = device;
    c->layers_run = -1;
    cudaDeviceProp prop;
    QW_CUDA_NULL(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (prop.major != 10) {
        qw_set_error("qwen_cuda_create: device is sm_%d%d, this library is built for sm_100a only", prop.major,
                     prop.minor);
        delete c;
        return nullptr;
    }
    QW_CUDA_NULL(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    QW_CUDA_NULL(cudaEventCreate(&c->ev0));
    QW_CUDA_NULL(cudaEventCreate(&c->ev1));

    c->D = m->dim; c->Hd = m->hidden_dim; c->L = m->n_layers; c->H = m->n_heads; c->KVH = m->n_kv_heads;
    c->V = m->vocab_size; c->S = m->seq_len; c->hd = m->head_dim; c->G = m->group_size;
    c->tp_rank = tp.rank; c->tp_size = tp.size;
    c->Hl = c->H / tp.size; c->KVHl = c->KVH / tp.size; c->Pl = c->Hl * 128; c->Kl = c->KVHl * 128;
    c->Hdl = c->Hd / tp.size; c->Vl = c->V / tp.size;
    const int D = c->D, L = c->L, Pl = c->Pl, Kl = c->Kl, Hdl = c->Hdl, Vl = c->Vl, r = tp.rank;
    const int P = c->H * 128, K = c->KVH * 128;

    auto fail = [&]() -> QwenCudaCtx* {
        qwen_cuda_destroy(c);
        return nullptr;
    };

    // ---- weights -> SG layout ------------------------------------------------
    c->w_qkv_stride = qw_row_bytes(D) * (size_t) (Pl + 2 * Kl);
    c->w_o_stride = qw_row_bytes(Pl) * (size_t) D;
    c->w_13_stride = qw_row_bytes(D) * (size_t) (2 * Hdl);
    c->w_2_stride = qw_row_bytes(Hdl) * (size_t) D;
    const size_t cls_bytes = qw_row_bytes(D) * (size_t) Vl, emb_bytes = qw_row_bytes(D) * (size_t) c->V;
    const bool emb_alias = m->shared_classifier && tp.size == 1;
    if (dev_alloc(&c->w_qkv, c->w_qkv_stride * L, false) || dev_alloc(&c->w_o, c->w_o_stride * L, false)
        || dev_alloc(&c->w_13, c->w_13_stride * L, false) || dev_alloc(&c->w_2, c->w_2_stride * L, false)
        || dev_alloc(&c->w_cls, cls_bytes, false))
        return fail();
    if (emb_alias) {
        c->w_emb = c->w_cls;
    } else if (dev_alloc(&c->w_emb, emb_bytes, false)) {
        return fail();
    }
    c->bytes_weights = (c->w_qkv_stride + c->w_o_stride + c->w_13_stride + c->w_2_stride) * L + cls_bytes
                       + (emb_alias ? 0 : emb_bytes);
    {
        Uploader up;
        if (up.init()) return fail();
        cudaStream_t s = c->stream;
        for (int l = 0; l < L; ++l) {
            uint8_t* qkv = c->w_qkv + l * c->w_qkv_stride;
            if (up.put(m->wq[l], D, r * Pl, Pl, 0, D, qkv, 0, 1, s) || up.put(m->wk[l], D, r * Kl, Kl, 0, D, qkv, Pl, 1, s)
                || up.put(m->wv[l], D, r * Kl, Kl, 0, D, qkv, Pl + Kl, 1, s)
                || up.put(m->wo[l], P, 0, D, r * Pl, Pl, c->w_o + l * c->w_o_stride, 0, 1, s)
                || up.put(m->w1[l], D, r * Hdl, Hdl, 0, D, c->w_13 + l * c->w_13_stride, 0, 2, s)
                || up.put(m->w3[l], D, r * Hdl, Hdl, 0, D, c->w_13 + l * c->w_13_stride, 1, 2, s)
                || up.put(m->w2[l], c->Hd, 0, D, r * Hdl, Hdl, c->w_2 + l * c->w_2_stride, 0, 1, s))
                return fail();
        }
        if (up.put(m->cls, D, r * Vl, Vl, 0, D, c->w_cls, 0, 1, s)) return fail();
        if (!emb_alias && up.put(m->emb, D, 0, c->V, 0, D, c->w_emb, 0, 1, s)) return fail();
        if (cudaStreamSynchronize(s) != cudaSuccess) {
            qw_set_error("qwen_cuda_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            return fail();
        }
        c->bytes_read = up.bytes_read;
        (void) K;
    }
    if (upload_f32(&c->att_norm, m->att_rms_norm, (size_t) L * D) || upload_f32(&c->ffn_norm, m->ffn_rms_norm, (size_t) L * D)
        || upload_f32(&c->out_norm, m->out_rms_norm, D) || upload_f32(&c->q_norm, m->q_rms_norm, (size_t) L * 128)
        || upload_f32(&c->k_norm, m->k_rms_norm, (size_t) L * 128)
        || upload_f32(&c->rope_cos, m->rope_cos, (size_t) c->S * 64) || upload_f32(&c->rope_sin, m->rope_sin, (size_t) c->S * 64))
        return fail();

    // ---- KV cache + activations ------------------------------------------------
    const size_t kv = (size_t) L * c->KVHl * c->S * 128;
    c->bytes_kv = 2 * kv * sizeof(float);
    if (dev_alloc(&c->k_cache, kv) || dev_alloc(&c->v_cache, kv)) return fail();
    const int amax = qw_pad_cols(std::max(D, std::max(Pl, Hdl)));
    c->att_max_splits = 64;
    c->argmax_cap = 4096;
    if (dev_alloc(&c->x, D) || dev_alloc(&c->xb, amax) || dev_alloc(&c->qkv, Pl + 2 * Kl) || dev_alloc(&c->q, Pl)
        || dev_alloc(&c->att, Pl) || dev_alloc(&c->h, Hdl) || dev_alloc(&c->h13, 2 * Hdl) || dev_alloc(&c->logits, Vl)
        || dev_alloc(&c->aq, amax) || dev_alloc(&c->as, amax / 64) || dev_alloc(&c->token_dev, 1)
        || dev_alloc(&c->argmax_out, c->argmax_cap) || dev_alloc(&c->att_m, (size_t) c->Hl * c->att_max_splits)
        || dev_alloc(&c->att_l, (size_t) c->Hl * c->att_max_splits)
        || dev_alloc(&c->att_acc, (size_t) c->Hl * c->att_max_splits * 128) || dev_alloc(&c->bar_counter, 512))
        return fail();
    // error flag lives in mapped pinned memory: the kernel writes it only on a barrier
    // timeout, the host reads it for free after every synchronise
    if (cudaHostAlloc((void**) &c->err_flag, 64, cudaHostAllocMapped) != cudaSuccess) {
        qw_set_error("qwen_cuda_create: pinned error flag allocation failed");
        return fail();
    }
    *c->err_flag = 0;
    if (cudaMallocHost((void**) &c->logits_pinned, (size_t) Vl * sizeof(float)) != cudaSuccess) {
        qw_set_error("qwen_cuda_create: pinned logits buffer allocation failed");
        return fail();
    }
    if (tp.size > 1 && dev_alloc(&c->logits_all, c->V)) return fail();
    if (qw_mega_init(c)) return fail();
    // dev_alloc's memsets run on the legacy default stream, which does NOT order against our
    // non-blocking stream: without this a late memset can zero buffers the first step already wrote
    if (cudaDeviceSynchronize() != cudaSuccess) {
        qw_set_error("qwen_cuda_create: %s", cudaGetErrorString(cudaGetLastError()));
        return fail();
    }
    return c;
}

extern "C" void qwen_cuda_destroy(QwenCudaCtx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    qw_mega_free(c);
    qw_prefill_free(c);
    qw_tp_free(c);
    if (c->logits_all) cudaFree(c->logits_all);
    if (c->sample_ws) cudaFree(c->sample_ws);
    if (c->x_inject) cudaFree(c->x_inject);
    if (c->dbg_codes) cudaFree(c->dbg_codes);
    if (c->w_emb != c->w_cls) cudaFree(c->w_emb);
    void* bufs[] = {c->w_qkv, c->w_o, c->w_13, c->w_2, c->w_cls, c->att_norm, c->ffn_norm, c->out_norm, c->q_norm,
                    c->k_norm, c->rope_cos, c->rope_sin, c->k_cache, c->v_cache, c->x, c->xb, c->qkv, c->q, c->att,
                    c->h, c->h13, c->logits, c->aq, c->as, c->token_dev, c->argmax_out, c->att_m, c->att_l,
                    c->att_acc, c->bar_counter};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (c->logits_pinned) cudaFreeHost(c->logits_pinned);
    if (c->err_flag) cudaFreeHost(c->err_flag);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void* qwen_cuda_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) {
        qw_set_error("pinned host allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    memset(p, 0, bytes);
    return p;
}
extern "C" void qwen_cuda_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int qw_mega_profile_enable(QwenCudaCtx* c);
int qw_mega_profile_read(QwenCudaCtx* c, unsigned long long* host, size_t max_elems);
int qw_mega_reset(QwenCudaCtx* c);
bool qw_mega_tp_ready(const QwenCudaCtx* c);
const float* qw_mega_debug_ptr(QwenCudaCtx* c, const char* what);
extern "C" int qwen_cuda_debug_profile_enable(QwenCudaCtx* c) { return c ? qw_mega_profile_enable(c) : -2; }
extern "C" int qwen_cuda_debug_profile_read(QwenCudaCtx* c, unsigned long long* host, size_t max_elems) {
    return c ? qw_mega_profile_read(c, host, max_elems) : -2;
}

void qw_mega_quant_records(const float* x, int n, uint8_t* sg, cudaStream_t st);
// Test hook: the quantiser fused into the persistent decode kernel's prologues (reciprocal candidate with an
// exact-division fallback) applied to a host vector; must equal q8_quantize (reference q8.c:5-30) bit for bit.
extern "C" int qwen_cuda_debug_quantize_fused(int8_t* q, float* s, const float* x, int n) {
    if (!q || !s || !x || n <= 0 || n % 64) return -2;
    const int recs = qw_sg_per_row(n);
    float* dx = nullptr;
    uint8_t* dsg = nullptr;
    std::vector<uint8_t> h((size_t) recs * QW_SG_BYTES);
    int rc = -1;
    do {
        if (cudaMalloc((void**) &dx, (size_t) n * 4) || cudaMalloc((void**) &dsg, h.size())) break;
        if (cudaMemcpy(dx, x, (size_t) n * 4, cudaMemcpyHostToDevice)) break;
        qw_mega_quant_records(dx, n, dsg, 0);
        if (cudaMemcpy(h.data(), dsg, h.size(), cudaMemcpyDeviceToHost)) break;
        for (int g = 0; g < n / 64; ++g) {
            const uint8_t* rec = h.data() + (size_t) (g >> 2) * QW_SG_BYTES;
            memcpy(q + (size_t) g * 64, rec + (g & 3) * 64, 64);
            memcpy(s + g, rec + 256 + (g & 3) * 4, 4);
        }
        rc = 0;
    } while (0);
    if (rc) qw_set_error("debug_quantize_fused: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(dx);
    cudaFree(dsg);
    return rc;
}

extern "C" int qwen_cuda_debug_set_layers(QwenCudaCtx* c, int n) {
    if (!c) return -2;
    c->layers_run = n;
    return 0;
}

// Debug (layer-by-layer parity with teacher forcing): the next steps run layers [l0, l1) only, then the final norm and the
// classifier; x_host (dim floats) replaces the embedding row as the residual stream entering layer l0. Both paths.
extern "C" int qwen_cuda_debug_set_window(QwenCudaCtx* c, int l0, int l1, const float* x_host) {
    if (!c || l0 < 0 || l0 > c->L || (l1 >= 0 && l1 < l0) || l1 > c->L) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    c->layer_begin = l0;
    c->layers_run = l1;
    c->x_inject_on = 0;
    if (x_host) {
        if (!c->x_inject) QW_CUDA(cudaMalloc((void**) &c->x_inject, (size_t) qw_pad_cols(c->D) * 4));
        QW_CUDA(cudaMemcpy(c->x_inject, x_host, (size_t) c->D * 4, cudaMemcpyHostToDevice));
        c->x_inject_on = 1;
    }
    return 0;
}

// Debug (flip audit): record the Q8_0 activation vector -- codes and scales -- the persistent kernel feeds to every GEMV of
// the following steps. Vector `which` = 4 * layer + {0: wq|wk|wv input, 1: wo input, 2: w1/w3 input, 3: w2 input},
// 4 * n_layers: classifier input. read unpacks n codes and n / 64 scales of the last step.
extern "C" int qwen_cuda_debug_codes_enable(QwenCudaCtx* c, int on) {
    if (!c) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    if (!on) {
        if (c->dbg_codes) cudaFree(c->dbg_codes);
        c->dbg_codes = nullptr;
        return 0;
    }
    if (!c->dbg_codes) {
        c->dbg_codes_stride = qw_row_bytes(std::max(c->D, std::max(c->Pl, c->Hdl)));
        QW_CUDA(cudaMalloc((void**) &c->dbg_codes, (size_t) (4 * c->L + 1) * c->dbg_codes_stride));
        QW_CUDA(cudaMemset(c->dbg_codes, 0, (size_t) (4 * c->L + 1) * c->dbg_codes_stride));
    }
    return 0;
}
extern "C" int qwen_cuda_debug_codes_read(QwenCudaCtx* c, int which, int8_t* q, float* s, int n) {
    if (!c || !c->dbg_codes || which < 0 || which > 4 * c->L || n <= 0 || n % 64 || qw_row_bytes(n) > c->dbg_codes_stride) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<uint8_t> h(qw_row_bytes(n));
    QW_CUDA(cudaMemcpy(h.data(), c->dbg_codes + (size_t) which * c->dbg_codes_stride, h.size(), cudaMemcpyDeviceToHost));
    for (int g = 0; g < n / 64; ++g) {
        const uint8_t* rec = h.data() + (size_t) (g >> 2) * QW_SG_BYTES;
        memcpy(q + (size_t) g * 64, rec + (g & 3) * 64, 64);
        memcpy(s + g, rec + 256 + (g & 3) * 4, 4);
    }
    return 0;
}

extern "C" int qwen_cuda_set_path(QwenCudaCtx* c, int path) {
    if (!c || path < 0 || path > 1) return -2;
    if (path == 0 && !qw_mega_tp_ready(c)) {
        qw_set_error("persistent kernel path: the peers' flow arenas are not mapped (qwen_cuda_tp_init) or the shape is unsupported");
        return -2;
    }
    c->path = path;
    return 0;
}

extern "C" int qwen_cuda_get_path(const QwenCudaCtx* c) { return c ? c->path : -2; }

// What this context (= this tensor-parallel rank) holds in HBM and what it read from the checkpoint at create.
extern "C" int qwen_cuda_memory(const QwenCudaCtx* c, size_t* weight_bytes, size_t* kv_bytes, size_t* read_bytes) {
    if (!c) return -2;
    if (weight_bytes) *weight_bytes = c->bytes_weights;
    if (kv_bytes) *kv_bytes = c->bytes_kv;
    if (read_bytes) *read_bytes = c->bytes_read;
    return 0;
}

static int qw_check_flag(QwenCudaCtx* c);
static int step(QwenCudaCtx* c, int token, const int* token_dev, int pos) {
    if (pos < 0 || pos >= c->S) {
        qw_set_error("forward: pos %d outside [0, %d)", pos, c->S);
        return -2;
    }
    if (!token_dev && (token < 0 || token >= c->V)) {
        qw_set_error("forward: token %d outside [0, %d)", token, c->V);
        return -2;
    }
    if (c->tp_size > 1 && *(volatile int*) c->err_flag) return qw_check_flag(c); // sticky (see qw_check_flag)
    return c->path == 1 ? qw_decode_ops(c, token, token_dev, pos) : qw_decode_mega(c, token, token_dev, pos);
}

static int qw_check_flag(QwenCudaCtx* c) {
    const int flag = *(volatile int*) c->err_flag;
    if (flag) {
        if (c->tp_size > 1) {
            // Tensor parallel: the peers store into this rank's arenas (and this rank into theirs) -- a rank-local refill would
            // race with them, and the ranks would disagree on which arena the next launch uses. The group stays in a STICKY
            // error state: every later step of this context fails with the same code; recovery = destroy all ranks' contexts.
            qw_set_error("decode kernel reported error %d on tensor-parallel rank %d (a wait timed out); the context is unusable", flag,
                         c->tp_rank);
            return -3;
        }
        qw_set_error("decode kernel reported error %d (a wait inside the persistent kernel timed out)", flag);
        *c->err_flag = 0;
        qw_mega_reset(c); // the flow arenas are meaningless after an abort
        return -3;
    }
    return 0;
}

extern "C" int qwen_cuda_forward_async(QwenCudaCtx* c, int token, int pos) {
    if (!c) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    return step(c, token, nullptr, pos);
}

extern "C" int qwen_cuda_sync(QwenCudaCtx* c) {
    if (!c) return -2;
    QW_CUDA(cudaStreamSynchronize(c->stream));
    return qw_check_flag(c);
}

extern "C" int qwen_cuda_logits_to_host(QwenCudaCtx* c, float* logits_host) {
    if (!c || !logits_host) return -2;
    // tensor-parallel contexts hold the all-gathered logits (the classifier is split over vocab rows)
    const float* src = c->tp_size > 1 ? c->logits_all : c->logits;
    QW_CUDA(cudaMemcpyAsync(logits_host, src, (size_t) c->V * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    return qw_check_flag(c);
}

extern "C" int qwen_cuda_forward(QwenCudaCtx* c, int token, int pos, float* logits_host) {
    if (!c) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    if (int rc = step(c, token, nullptr, pos)) return rc;
    return qwen_cuda_logits_to_host(c, logits_host);
}

// Stands behind the prompt loop of the generation code (reference: src/completion.c:57-66 calls forward()
// once per prompt token and keeps only the last logits): n tokens at positions pos0 .. pos0 + n - 1.
extern "C" int qwen_cuda_prefill(QwenCudaCtx* c, const int* tokens, int n, int pos0, float* logits_host) {
    if (!c || !tokens) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    if (int rc = qw_prefill(c, tokens, n, pos0)) return rc;
    return logits_host ? qwen_cuda_logits_to_host(c, logits_host) : 0;
}

extern "C" int qwen_cuda_decode_greedy(QwenCudaCtx* c, int first_token, int pos0, int n, int* out_tokens_host) {
    if (!c || n < 0 || n > c->argmax_cap) {
        qw_set_error("decode_greedy: bad arguments (n %d, cap %d, tp %d)", n, c ? c->argmax_cap : 0, c ? c->tp_size : 0);
        return -2;
    }
    QW_CUDA(cudaSetDevice(c->device));
    QW_CUDA(cudaMemcpyAsync(c->token_dev, &first_token, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    for (int i = 0; i < n; ++i) {
        if (int rc = step(c, 0, c->token_dev, pos0 + i)) return rc;
        launch_argmax(c->tp_size > 1 ? c->logits_all : c->logits, c->V, c->argmax_out + i, c->token_dev, c->stream);
    }
    QW_CUDA(cudaMemcpyAsync(out_tokens_host, c->argmax_out, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return qwen_cuda_sync(c);
}

extern "C" int qwen_cuda_time_decode(QwenCudaCtx* c, int token, int pos0, int steps, int warmup, float* ms_total,
                                     int* launches) {
    if (!c || steps <= 0 || warmup < 0) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    int pos = pos0;
    for (int i = 0; i < warmup; ++i)
        if (int rc = step(c, token, nullptr, pos++)) return rc;
    QW_CUDA(cudaStreamSynchronize(c->stream));
    QW_CUDA(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < steps; ++i)
        if (int rc = step(c, token, nullptr, pos++)) return rc;
    QW_CUDA(cudaEventRecord(c->ev1, c->stream));
    if (int rc = qwen_cuda_sync(c)) return rc;
    QW_CUDA(cudaEventElapsedTime(ms_total, c->ev0, c->ev1));
    if (launches) *launches = steps * (c->path == 1 ? qw_decode_ops_launches(c) : qw_decode_mega_launches(c));
    return 0;
}

// ---- KV cache access: host side is the reference's [npos][KVH*128] per layer ----
static int kv_copy(QwenCudaCtx* c, int layer, int pos0, int npos, float* k_host, float* v_host, bool to_device) {
    if (!c || layer < 0 || layer >= c->L || pos0 < 0 || npos < 0 || pos0 + npos > c->S) {
        qw_set_error("kv access out of range (layer %d pos %d+%d)", layer, pos0, npos);
        return -2;
    }
    QW_CUDA(cudaSetDevice(c->device));
    const size_t hpitch = (size_t) c->KVH * 128 * sizeof(float);
    for (int h = 0; h < c->KVHl; ++h) {
        const size_t doff = (((size_t) layer * c->KVHl + h) * c->S + pos0) * 128;
        const size_t hoff = (size_t) (c->tp_rank * c->KVHl + h) * 128;
        if (to_device) {
            QW_CUDA(cudaMemcpy2DAsync(c->k_cache + doff, 512, k_host + hoff, hpitch, 512, npos, cudaMemcpyHostToDevice, c->stream));
            QW_CUDA(cudaMemcpy2DAsync(c->v_cache + doff, 512, v_host + hoff, hpitch, 512, npos, cudaMemcpyHostToDevice, c->stream));
        } else {
            QW_CUDA(cudaMemcpy2DAsync(k_host + hoff, hpitch, c->k_cache + doff, 512, 512, npos, cudaMemcpyDeviceToHost, c->stream));
            QW_CUDA(cudaMemcpy2DAsync(v_host + hoff, hpitch, c->v_cache + doff, 512, 512, npos, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    QW_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int qwen_cuda_kv_write(QwenCudaCtx* c, int layer, int pos0, int npos, const float* k, const float* v) {
    return kv_copy(c, layer, pos0, npos, const_cast<float*>(k), const_cast<float*>(v), true);
}
extern "C" int qwen_cuda_kv_read(QwenCudaCtx* c, int layer, int pos0, int npos, float* k, float* v) {
    return kv_copy(c, layer, pos0, npos, k, v, false);
}

extern "C" int qwen_cuda_debug_read(QwenCudaCtx* c, const char* what, void* host, size_t max_bytes) {
    if (!c || !what || !host) return -2;
    const void* src = nullptr;
    size_t bytes = 0, elems = 0;
    if (!strcmp(what, "x")) { src = c->x; elems = c->D; bytes = elems * 4; }
    else if (!strcmp(what, "xb")) { src = c->xb; elems = c->D; bytes = elems * 4; }
    else if (!strcmp(what, "q")) { src = c->q; elems = c->Pl; bytes = elems * 4; }
    else if (!strcmp(what, "qkv")) { src = c->qkv; elems = c->Pl + 2 * c->Kl; bytes = elems * 4; }
    else if (!strcmp(what, "att")) { src = c->att; elems = c->Pl; bytes = elems * 4; }
    else if (!strcmp(what, "h")) { src = c->h; elems = c->Hdl; bytes = elems * 4; }
    else if (!strcmp(what, "logits")) { src = c->logits; elems = c->Vl; bytes = elems * 4; }
    else if (!strcmp(what, "aq")) { src = c->aq; elems = qw_pad_cols(std::max(c->D, std::max(c->Pl, c->Hdl))); bytes = elems; }
    else if (!strcmp(what, "as")) { src = c->as; elems = qw_pad_cols(std::max(c->D, std::max(c->Pl, c->Hdl))) / 64; bytes = elems * 4; }
    else {
        qw_set_error("debug_read: unknown buffer '%s'", what);
        return -2;
    }
    if (bytes > max_bytes) bytes = max_bytes;
    QW_CUDA(cudaSetDevice(c->device));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    if (c->path == 0) // the persistent kernel keeps x / h / qkv of every layer in its flow arena
        if (const float* f = qw_mega_debug_ptr(c, what)) src = f;
    QW_CUDA(cudaMemcpy(host, src, bytes, cudaMemcpyDeviceToHost));
    return (int) elems;
}

extern "C" int qwen_cuda_attention(QwenCudaCtx* c, int layer, int pos, const float* q_host, float* out_host) {
    if (!c || layer < 0 || layer >= c->L || pos < 0 || pos >= c->S) {
        qw_set_error("attention: layer/pos out of range");
        return -2;
    }
    QW_CUDA(cudaSetDevice(c->device));
    const size_t off = (size_t) c->tp_rank * c->Pl;
    QW_CUDA(cudaMemcpyAsync(c->q, q_host + off, (size_t) c->Pl * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = qw_attention_device(c, layer, pos, c->q, c->att)) return rc;
    QW_CUDA(cudaGetLastError());
    QW_CUDA(cudaMemcpyAsync(out_host + off, c->att, (size_t) c->Pl * 4, cudaMemcpyDeviceToHost, c->stream));
    QW_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
