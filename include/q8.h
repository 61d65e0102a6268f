/*
 * q8.h -- Q8_0 tensor view and the two Q8_0 primitives, B200 build.
 *
 * ABI mirror of the reference header (reference: include/q8.h:10-30). The
 * struct layout and the two prototypes are the drop-in contract: unchanged
 * reference callers (src/model.c, src/forward.c, examples) compile against this
 * file and link against libqwen3.so from this repo.
 *
 * In this build both functions are host-in / host-out wrappers: the buffers the
 * caller passes are ordinary host memory, the arithmetic runs in sm_100a
 * kernels (csrc/ops.cu) behind qwen_cuda_q8_quantize / qwen_cuda_q8_dequantize.
 * There is no CPU implementation in the product library.
 */
#ifndef QWEN_Q8_H
#define QWEN_Q8_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Largest magnitude an int8 code may take (symmetric range, -128 unused). */
#define Q8_MAX 127.0f

/*
 * A Q8_0 tensor is a pair of borrowed pointers; it carries no shape.
 *   q : numel int8 codes
 *   s : numel / block_size fp32 scales, one per run of block_size codes
 * Field order matters (reference: include/q8.h:16-19).
 */
typedef struct Q8Tensor {
    float* s;
    int8_t* q;
} Q8Tensor;

/*
 * Per block of `block_size` floats: scale = absmax / 127 (1e-6f for an all-zero
 * block), code = clamp(roundf(x / scale), -127, 127). A tail of n % block_size
 * elements is ignored. Bit-exact with the reference for identical input
 * (reference: src/q8.c:5-30).
 */
void q8_quantize(Q8Tensor* qt, float* x, int n, int block_size);

/* x[i] = (float) q[i] * s[i / block_size]  (reference: src/q8.c:32-37). */
void q8_dequantize(Q8Tensor* qt, float* x, int n, int block_size);

#ifdef __cplusplus
}
#endif

#endif /* QWEN_Q8_H */
