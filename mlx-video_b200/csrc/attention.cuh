// Shared between the two attention kernels (attention.cu: one query tile per CTA; attention_pair.cu: two).
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"
#include "peer_sync.cuh"

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
  float* ws_o;   // [slots][part_rows][dh] unnormalised partial outputs
  float* ws_ml;  // [slots][part_rows][2]  running maximum (log2 domain) and row sum
  // partial rows of one job in the workspace (256 = two full query tiles; fewer for the compact cross-rank partials of
  // video->audio attention), and how attention_combine_kernel finds part i of job j: row j*cmb_job_stride of the arrays
  // ws_o + i*cmb_part_o and ws_ml + i*cmb_part_ml (strides in floats)
  int part_rows;
  long long cmb_job_stride, cmb_part_o, cmb_part_ml;
  // key-split merge inside attention_pair64_kernel (no separate combine launch): arrivals / departures per split job, zero
  // between launches; nullptr = attention_combine_kernel merges
  int* cmb_counters;
  // cross-GPU flag barrier before the first read of Q / K / V (they were stored by the peers): folded into the prologue of
  // attention_pair64_kernel; the launchers of the other kernels run the stand-alone barrier kernel first
  PeerSync sync;
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
// partial mode: every job (batch, head, query-tile pair) over its whole key range, unnormalised (O, m, l) to part_o / part_ml
int launch_attention_partial(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, AttnParams p,
                             int dh, float* part_o, float* part_ml, cudaStream_t stream);
// merge n_parts partial sets (part i at base + i * part_stride floats: [jobs*rows][dh] then [jobs*rows][2]) into p.O
int launch_attention_merge(AttnParams p, int dh, const float* parts, long long part_stride, int n_parts, cudaStream_t stream);
int launch_attention_pair(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                          AttnParams p, int dh, cudaStream_t stream);

}  // namespace ltxb
