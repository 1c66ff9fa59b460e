// Shared between the two attention kernels (attention.cu: one query tile per CTA; attention_pair.cu: two).
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace ltxb {

struct AttnParams {
  int B, Tq, Tk, H;
  float scale_log2;  // softmax scale * log2(e)
  __nv_bfloat16* O;
  long long ldo;
  const float* kv_bias;  // [B, Tk] additive (natural-log domain) or null
  // Ulysses gather fused into the epilogue: query rows [i*rows_per_peer, (i+1)*rows_per_peer) are stored straight
  // into rank i's receive buffer over NVLink (rows_per_peer == 0: plain local output)
  int rows_per_peer;
  __nv_bfloat16* o_peer[8];
  // pair kernel only: job = (batch, head, pair of query tiles); CTAs [0, n_full) run a whole job, the remaining
  // jobs (the ragged last wave) are cut into n_split key ranges whose partial (O, m, l) go to the workspace
  int n_qp, n_full, n_split;
  float* ws_o;   // [slots][256][dh] unnormalised partial outputs
  float* ws_ml;  // [slots][256][2]  running maximum (log2 domain) and row sum
};

// where output row `row` of batch b, head h starts (local O or a peer's receive buffer)
__device__ __forceinline__ __nv_bfloat16* attn_out_row(const AttnParams& p, int b, int row, int h, int dh) {
  if (p.rows_per_peer > 0) {
    const int dst = min(row / p.rows_per_peer, 7);
    return p.o_peer[dst] + static_cast<long long>(row - dst * p.rows_per_peer) * p.ldo + h * dh;
  }
  return p.O + (static_cast<long long>(b) * p.Tq + row) * p.ldo + h * dh;
}

// attention_pair.cu
int launch_attention_pair(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                          AttnParams p, int dh, cudaStream_t stream);

}  // namespace ltxb
