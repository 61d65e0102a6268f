/* host_priv.h -- the private record behind a Model* (host C side). */
#ifndef QWEN_HOST_PRIV_H
#define QWEN_HOST_PRIV_H

#include "../../include/forward.h"
#include "../../include/model.h"
#include "../../include/q8.h"
#include "../../include/qwen_cuda.h"

typedef struct ModelPriv {
    Model pub;        /* must stay first: callers hold &pub */
    QwenCudaCtx* ctx; /* device context */
} ModelPriv;

QwenCudaCtx* model_cuda_ctx(Model* m);

#endif
