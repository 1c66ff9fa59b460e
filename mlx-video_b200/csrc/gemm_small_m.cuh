// Small-M (weight-streaming) variant of K1, see gemm_small_m.cu; called from ltxb_gemm_bf16 (gemm.cu).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ltxb.h"

namespace ltxb {

bool gemm_small_m_supported(int M, int N, int K);
// MLX affine-quantised weights (bits 4 / 8): W = uint32 words [N, ldw], levels lowest bits first; scales / biases [N, lds]
struct WsPacked {
  const void* scales;
  const void* biases;
  long long lds;
  int group, bits, aux_f32;
};
// partials / counters: the registered split-K workspace (NULL: no split-K); want_splits > 0 forces the k-range count
int launch_gemm_small_m(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                        const ltxb_epilogue* epi, float* partials, long long partial_bytes, int* counters, int want_splits,
                        cudaStream_t stream, const WsPacked* packed = nullptr);

}  // namespace ltxb
