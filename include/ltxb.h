/*
 * ltxb — C ABI of the B200-native LTX-2 DiT denoise-step kernels (sm_100a).
 *
 * The reference (CharafChnioune/mlx-video) has NO native / FFI layer on this path: the hot path is
 * Python over MLX library ops (ltx_core/loader/kernels.py:1-3 is an empty
 * "custom kernels" stub).  Each entry point below therefore replaces one MLX library op *call site*
 * of the reference, cited per function.  The Python host (mlx-video_b200/) binds these with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; no torch types cross this boundary;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host sync, no
 *     allocation inside the library (scratch, if any, is passed by the caller);
 *   - row-major matrices, leading dimensions (`ld*`) in ELEMENTS;
 *   - bf16 = __nv_bfloat16 storage, "f32" = float;
 *   - return value 0 = ok, negative = error (LTXB_ERR_*); ltxb_last_error() gives the message of
 *     the last failure on the calling thread.
 */
#ifndef LTXB_H_
#define LTXB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTXB_OK 0
#define LTXB_ERR_BAD_ARG (-1)     /* null pointer, misaligned pointer, negative size           */
#define LTXB_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are built for              */
#define LTXB_ERR_CUDA (-3)        /* a CUDA runtime / driver call failed, see ltxb_last_error  */
#define LTXB_ERR_NO_DEVICE (-4)   /* no sm_100 device visible                                  */

const char* ltxb_last_error(void);
/* ABI version of this header; bump on any signature change. */
int ltxb_abi_version(void);
/* 0 if a compute-capability 10.x device is current, LTXB_ERR_NO_DEVICE otherwise. */
int ltxb_device_check(void);
/* Make `device` current for the library's (statically linked) CUDA runtime on the calling thread.
 * Call once per thread before the first launch when the host framework uses a device other than 0. */
int ltxb_set_device(int32_t device);

/* Kernels launched by this library since it was loaded (one C-ABI call may launch more than one: the attention
 * key-split adds a merge kernel).  bench.py reports the delta over its timed region as `gpu_launches`. */
int64_t ltxb_kernel_launches(void);

/* ------------------------------------------------------------------------------------------------
 * K1  nn.Linear  (attention.py:91-93,100,123-126,142; feed_forward.py:31,33; ltx.py:130,292,301;
 *                 adaln.py:27,130-132; text_projection.py:18,20)
 *
 *   acc[m,n] = sum_k A[m,k] * W[n,k]         A: bf16 [M,K] (lda)   W: bf16 [N,K] (ldw), i.e. the
 *                                            (out,in) weight layout of nn.Linear
 * followed by a fused epilogue:
 *   LTXB_EPI_BIAS_BF16      out bf16 = acc + bias
 *   LTXB_EPI_GELU_BF16      out bf16 = gelu_tanh(acc + bias)     (feed_forward.py:12, text_projection.py:19)
 *   LTXB_EPI_SILU_BF16      out bf16 = silu(acc + bias)          (adaln.py:26,131)
 *   LTXB_EPI_BIAS_F32       out f32  = acc + bias
 *   LTXB_EPI_RESID_GATE_F32 out f32  = resid + (acc + bias) * g  (transformer.py:254,257-261,347)
 *        g = 1                                  when gate == NULL (attn2, un-gated)
 *        g = gate_table[n] + gate[grow, n]      otherwise, grow = gate_row_index ? gate_row_index[m]
 *                                                                                : m / gate_row_div
 * tcgen05/TMEM tensor-core kernel fed by TMA; persistent over output tiles.
 * Requirements: K % 64 == 0, N % 16 == 0, lda/ldw/ldo multiples of 8, 16-byte aligned pointers.
 * block_n: N extent of one output tile (multiple of 16, 32..256) or 0 = choose for wave efficiency.
 * cta_pair: 1 = cta_group::2 (two SMs share one 256-row tile), 0 = single CTA, -1 = choose; 2 / 3 = force the
 *   contiguous stream-K schedule (every SM an equal, contiguous share of the (tile, k-block) list; what -1 picks for
 *   M <= 640 when the tile count leaves a ragged wave) with single-CTA / pair tiles; 4 = force the few-row weight-streaming
 *   kernel (M <= 512: weight rows on the TMEM lanes, tokens as the MMA's N; what -1 picks for M <= 512), block_n then =
 *   k-range splits per weight tile (0 = choose).
 * ---------------------------------------------------------------------------------------------- */
enum {
  LTXB_EPI_BIAS_BF16 = 0,
  LTXB_EPI_GELU_BF16 = 1,
  LTXB_EPI_SILU_BF16 = 2,
  LTXB_EPI_BIAS_F32 = 3,
  LTXB_EPI_RESID_GATE_F32 = 4,
  LTXB_EPI_COUNT = 5
};

/* The flag barrier of the NVLink-fused Ulysses exchange (ltxb_peer_barrier) folded into the prologue of the kernel that
 * consumes what the peers wrote: once the stream predecessor has completed, block 0 raises this rank's flag on every
 * peer and every block waits for all peers' flags before it reads; the last block to leave advances the epoch.  Same
 * flags / epoch counter as ltxb_peer_barrier (the two can be mixed); done_counter is one more device int, zero between
 * launches.  Accepted by ltxb_gemm_bf16 / ltxb_gemm_qw_bf16 (ltxb_epilogue.peer_sync) and ltxb_attention_fwd_peers_sync. */
typedef struct ltxb_peer_sync {
  int32_t* flags[8];      /* flags[i] = rank i's flag array (n_peers ints), mapped into this process */
  int32_t n_peers;        /* 1..8 */
  int32_t my_rank;
  int32_t* epoch_counter; /* device int: barriers this rank has passed */
  int32_t* done_counter;  /* device int, zero between launches */
} ltxb_peer_sync;

/* ltxb_epilogue.flags.  LTXB_GEMM_CONST_W: W is not written by any kernel that may still be running ahead of this
 * launch in the stream (model weights): the few-row kernel then starts streaming W while its stream predecessor is
 * still finishing (programmatic dependent launch) and only the activations wait for it. */
#define LTXB_GEMM_CONST_W 1

typedef struct ltxb_epilogue {
  int32_t mode;                  /* LTXB_EPI_*                                       */
  int32_t gate_row_div;          /* >= 1; rows of `gate` = ceil(M / gate_row_div)    */
  const float* bias;             /* [N] or NULL                                      */
  const float* resid;            /* RESID_GATE: f32 [M, ldr]; may alias out          */
  int64_t ldr;
  const float* gate;             /* RESID_GATE: f32 [*, gate_ld] or NULL             */
  int64_t gate_ld;
  const int32_t* gate_row_index; /* optional [M]                                     */
  const float* gate_table;       /* [N] added to gate, or NULL                       */
  /* A-operand layout (0 = plain row-major [M, K]).  a_group_cols = g > 0: A is stored head-group-major,
   * [K / g][M][g] with rows lda apart and groups a_group_stride elements apart — what the Ulysses gather
   * all-to-all delivers — and is read in place through a 3-D TMA map (no transposing copy). */
  int32_t a_group_cols;
  int32_t flags;                 /* LTXB_GEMM_* bits                                 */
  int64_t a_group_stride;
  const ltxb_peer_sync* peer_sync; /* NULL, or: run the cross-GPU flag barrier before the first read of A (A = what the peers
                                    * stored into this rank's receive buffer: the out-projection after ltxb_attention_fwd_peers) */
} ltxb_epilogue;

int ltxb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int32_t M,
                   int32_t N, int32_t K, const ltxb_epilogue* epi, int32_t block_n, int32_t cta_pair,
                   void* stream);

/* K1 over MLX affine-quantised weights, for FEW rows (M <= 256) — nn.QuantizedLinear on the DiT path (ltx.py:641-725 loads
 * such checkpoints; mx.quantized_matmul):  acc[m,n] = sum_k A[m,k] * bf16(scales[n, k/G] * q[n,k] + biases[n, k/G]).
 * Wq: uint32 [N, ldq], 32/bits levels per word, lowest bits first (MLX packing); scales / biases [N, lds], bf16 or f32
 * (aux_f32); G = group_size (32, 64, 128); bits 4 or 8.  The packed tiles are what travels from HBM (a quarter / half of
 * the bf16 bytes — at few rows the weight stream bounds the GEMM) and are expanded on the SM, into tensor memory, with the
 * arithmetic of ltxb_dequant_affine_bf16, so the result is bit-identical to ltxb_dequant_affine_bf16 followed by ltxb_gemm_bf16
 * (cta_pair 4, same splits).  Same epilogues as ltxb_gemm_bf16; splits = k-range pieces per weight tile, 0 = choose.
 * More than 256 rows: LTXB_ERR_UNSUPPORTED (expand the weights once and call ltxb_gemm_bf16 — that regime is tensor-bound). */
int ltxb_gemm_qw_bf16(const void* A, int64_t lda, const uint32_t* Wq, int64_t ldq, const void* scales, const void* biases,
                      int64_t lds, int32_t aux_f32, int32_t group_size, int32_t bits, void* out, int64_t ldo, int32_t M,
                      int32_t N, int32_t K, const ltxb_epilogue* epi, int32_t splits, void* stream);

/* Split-K scratch for ltxb_gemm_bf16.  The library never allocates: the host hands it one device buffer of
 * ltxb_gemm_workspace_bytes() per device (kept until replaced; NULL unregisters).  With a workspace registered,
 * a problem whose tile count leaves a ragged wave on the 148 SMs (M = 1280 tokens: 80 tiles for 74 SM pairs) runs
 * the leftover tiles first, each cut into 2-4 k-ranges on different SM pairs; the pair holding a tile's first
 * k-range adds the others' parked fp32 partials in k order (bit-reproducible) and applies the epilogue.  Without
 * a workspace every GEMM runs data-parallel.  GEMMs sharing a workspace must be issued on one stream. */
int64_t ltxb_gemm_workspace_bytes(void);
int ltxb_gemm_set_workspace(void* workspace, int64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3 + AdaLN modulate   rms_norm(x) * (1 + scale) + shift   (utils.py:398-400; transformer.py:253,
 *                       257,315-325,346)   x f32 [R, D] (ldx) -> out bf16 [R, D] (ldo)
 *   scale[r,:] = table_scale[:] + mod[mrow(r), scale_off : scale_off + D]   (get_ada_values,
 *   shift[r,:] = table_shift[:] + mod[mrow(r), shift_off : shift_off + D]    transformer.py:135-177)
 *   mrow(r) = row_index ? row_index[r] : r / row_div.   mod == NULL -> plain weightless rms_norm.
 * One warp-shuffle reduction per row, 128-bit loads/stores. D % 8 == 0, D <= 16384.
 * ---------------------------------------------------------------------------------------------- */
int ltxb_rmsnorm_modulate(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t R, int32_t D, float eps,
                          const float* mod, int64_t ld_mod, int32_t scale_off, int32_t shift_off,
                          const float* table_scale, const float* table_shift, int32_t row_div,
                          const int32_t* row_index, void* stream);

/* K3r  The residual add of a projection and the NEXT sub-layer's rms_norm + AdaLN in one pass over the fp32 row
 *      (transformer.py:254 -> 256-257 and 257 -> 343-346):
 *        x[r,:] += y[r,:] * g[r,:]                      (x updated in place; y bf16 [R,D] = projection output incl. bias)
 *        out[r,:] = rms_norm(x[r,:]) * (1 + scale[r,:]) + shift[r,:]
 *      g     = table_gate[:]  + mod[mrow(r), gate_off  + :]   (either part NULL / offset < 0 -> absent; both absent -> 1)
 *      scale = table_scale[:] + mod[mrow(r), scale_off + :]   (likewise; both absent -> plain rms_norm)
 *      The out-projection GEMM then keeps its bf16 epilogue (LTXB_EPI_BIAS_BF16) instead of LTXB_EPI_RESID_GATE_F32,
 *      whose fp32 read-modify-write of the tile is exposed when there is one tile per SM pair (N = 4096). */
int ltxb_residual_rmsnorm_modulate(float* x, int64_t ldx, const void* y, int64_t ldy, void* out, int64_t ldo, int32_t R,
                                   int32_t D, float eps, const float* mod, int64_t ld_mod, int32_t gate_off,
                                   int32_t scale_off, int32_t shift_off, const float* table_gate,
                                   const float* table_scale, const float* table_shift, int32_t row_div,
                                   const int32_t* row_index, void* stream);

/* K6  nn.LayerNorm(affine=False) then x*(1+scale)+shift  (ltx.py:300,432-457)
 *   scale/shift[r,:] = table_*[:] + emb[mrow(r), :]   (emb = embedded_timestep f32 [*, D]) */
int ltxb_layernorm_modulate(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t R, int32_t D, float eps,
                            const float* emb, int64_t ld_emb, const float* table_scale,
                            const float* table_shift, int32_t row_div, const int32_t* row_index,
                            void* stream);

/* x f32 [R,D] += y bf16 [R,D] * g   (transformer.py:254,347 when the projection ran un-fused)
 *   g as in LTXB_EPI_RESID_GATE_F32 (gate NULL -> 1). */
int ltxb_gate_residual(float* x, int64_t ldx, const void* y, int64_t ldy, int32_t R, int32_t D, const float* gate,
                       int64_t gate_ld, int32_t gate_off, const float* gate_table, int32_t row_div,
                       const int32_t* row_index, void* stream);

/* K3w + K4  q_norm / k_norm (nn.RMSNorm over the FULL inner dim, learned weight; attention.py:96-97,
 *           129-130) followed by split RoPE per head (rope.py:109-172; attention.py:133-136), in place.
 *   x bf16 [B*T, D] (ldx), D = H * dh, dh in {64,128};  weight f32 [D];
 *   cos/sin f32 [B_pe, H, T, dh/2] (B_pe == 1 broadcasts) or NULL for no rotation (attn2). */
int ltxb_qknorm_rope(void* x, int64_t ldx, int32_t B, int32_t T, int32_t H, int32_t dh, const float* weight,
                     float eps, const float* cos_tab, const float* sin_tab, int32_t B_pe, void* stream);

/* The same over n_seg independent column slices of one buffer in ONE launch: slice s is x + s*seg_stride (columns)
 * with norm weight `weight + s*w_seg_stride` — q and k of the fused QKV buffer (n_seg = 2, attention.py:129-136), or
 * the text keys of all 48 blocks in the stacked K/V buffer (n_seg = 48, no rotation).  weight2 (optional, same
 * striding) is a second per-column factor multiplied in after the norm. */
int ltxb_qknorm_rope_segments(void* x, int64_t ldx, int32_t n_seg, int64_t seg_stride, int32_t B, int32_t T, int32_t H,
                              int32_t dh, const float* weight, int64_t w_seg_stride, const float* weight2, float eps,
                              const float* cos_tab, const float* sin_tab, int32_t B_pe, void* stream);

/* Same operation writing to a SEPARATE, head-grouped destination — the send buffer of the Ulysses
 * head-scatter all-to-all (new: the reference is single-device).  Head h of row r goes to
 *   out + (h / heads_per_group) * group_stride + r * ldo + (h % heads_per_group) * dh
 * so each destination rank's heads are one contiguous [rows, *] block.  weight == NULL copies without
 * norm / rotation (the V operand).  The norm statistic still spans all H heads of the row (attention.py:96-97),
 * which is why it must run BEFORE the heads are scattered. */
int ltxb_qknorm_rope_scatter(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t heads_per_group,
                             int64_t group_stride, int32_t B, int32_t T, int32_t H, int32_t dh, const float* weight,
                             float eps, const float* cos_tab, const float* sin_tab, int32_t B_pe, void* stream);

/* The same, fused with the head-scatter all-to-all: group g's heads are stored to group_bases[g] (host array of
 * n_groups device pointers — NVLink-mapped receive buffers of the destination ranks, each already offset to this
 * rank's chunk), rows ldo apart.  No send buffer and no collective call: the exchange is this kernel's stores. */
int ltxb_qknorm_rope_scatter_peers(const void* x, int64_t ldx, void* const* group_bases, int32_t n_groups, int64_t ldo,
                                   int32_t B, int32_t T, int32_t H, int32_t dh, const float* weight, float eps,
                                   const float* cos_tab, const float* sin_tab, int32_t B_pe, void* stream);

/* q | k | v of the fused QKV buffer in ONE launch (three column segments seg_stride elements apart): q and k get their
 * full-width RMSNorm (qk_weight f32 [2, H*dh]: q_norm row 0, k_norm row 1; attention.py:96-97,129-130) and split RoPE
 * (rope.py:109-172), v is copied; head group g of segment s is stored to group_bases[g] + s * slot_stride (elements),
 * rows ldo apart — the receive buffer layout [rows][q | k | v][heads of this rank].  Replaces three
 * ltxb_qknorm_rope_scatter_peers launches per block of the sequence-parallel forward. */
int ltxb_qkv_norm_rope_scatter_peers(const void* qkv, int64_t ldx, int64_t seg_stride, void* const* group_bases,
                                     int32_t n_groups, int64_t ldo, int64_t slot_stride, int32_t T, int32_t H, int32_t dh,
                                     const float* qk_weight, float eps, const float* cos_tab, const float* sin_tab,
                                     int32_t B_pe, void* stream);

/* Cross-GPU barrier on NVLink peer memory (new).  flag_ptrs: host array, flag_ptrs[i] = rank i's
 * int32[n_peers] flag array mapped into this process (zero-initialised once); epoch_counter: one int32 in LOCAL
 * device memory, zero-initialised once, advanced by the kernel (so a captured CUDA graph can replay the barrier).
 * Stream-ordered: once all earlier work of the stream has completed, raises this rank's flag on every peer to the
 * next epoch (release, system scope), then waits until every peer's flag here has reached it.  Every rank must
 * call it the same number of times. */
int ltxb_peer_barrier(int32_t* const* flag_ptrs, int32_t n_peers, int32_t my_rank, int32_t* epoch_counter, void* stream);

/* Store `bytes` (multiple of 16) from src to dst_ptrs[0..n_peers) — this rank's block of an all-gather, written
 * straight into every peer's (and its own) gather buffer over NVLink. */
int ltxb_peer_broadcast(const void* src, int64_t bytes, void* const* dst_ptrs, int32_t n_peers, void* stream);

/* a5  sinusoidal timestep features  (utils.py:486-526 with adaln.py:66: dim 256, flip_sin_to_cos,
 *     shift 0):  out bf16 [n, dim] = [cos(t*scale*f_i) | sin(t*scale*f_i)], f_i = exp(-ln(1e4) i/(dim/2)). */
int ltxb_timestep_embed(const float* t, int32_t n, float scale, int32_t dim, void* out, int64_t ldo, void* stream);

/* F7 dedupe of per-token timesteps (generate.py:604-606,653,793 build timesteps = sigma * mask, so at most
 *   F+1 distinct values exist):  values f32 [cap] = the distinct entries of t[0..n) (unused slots 0),
 *   index i32 [n] = slot of each token, count i32 [1] = slots used.  More than `cap` distinct values:
 *   count = cap+1 and all values are NaN (the forward then yields NaN; hosts check `count`). */
int ltxb_timestep_groups(const float* t, int32_t n, int32_t cap, float* values, int32_t* index, int32_t* count,
                         void* stream);

/* a4  RoPE table  (rope.py:419-529, SPLIT layout):  cos/sin f32 [B, H, T, dim/(2H)].
 *   positions f32 [B, n_axes, T, 2] ([start,end) bounds; middle = (start+end)/2 when use_middle, else start),
 *   max_pos f32 [n_axes] (host), freq f32 [nfreq] (device; theta^linspace(0,1,nfreq) * pi/2, built on
 *   the host exactly as the reference does), nfreq = dim / (2 n_axes); left pad = dim/2 - nfreq*n_axes. */
int ltxb_rope_table(const float* positions, int32_t B, int32_t n_axes, int32_t T, const float* max_pos_host,
                    const float* freq, int32_t nfreq, int32_t dim, int32_t H, int32_t use_middle, float* cos_out,
                    float* sin_out, void* stream);

/* elementwise helpers around the GEMMs */
int ltxb_silu_bf16(const void* x, void* out, int64_t n, void* stream);        /* adaln.py:26 */
int ltxb_cast_f32_to_bf16(const float* x, void* out, int64_t n, void* stream);
int ltxb_cast_bf16_to_f32(const void* x, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  scaled_dot_product_attention  (attention.py:13-53 -> mx.fast.scaled_dot_product_attention)
 *   O[b,tq,h,:] = softmax_k(Q[b,tq,h,:] . K[b,tk,h,:] * scale + kv_bias[b,tk]) V[b,tk,h,:]
 *   Q bf16 [B*Tq, *] (ldq), head h at columns [h*dh, (h+1)*dh); K, V likewise with Tk rows per batch;
 *   O bf16 [B*Tq, H*dh] (ldo).  Non-causal.  kv_bias f32 [B, Tk] additive or NULL
 *   (ltx.py:91-107 turns a 0/1 context mask into (m-1)*1e9).  dh in {64,128}.
 * FlashAttention-style online softmax; both contractions on tcgen05 with TMEM accumulators, TMA loads.
 * Tq <= 128: one query tile per CTA, P staged in shared memory.  Tq > 128: two query tiles per CTA ping-pong on
 * the tensor core, P kept in TMEM as the A operand of O += P V (attention_pair.cu).
 * ---------------------------------------------------------------------------------------------- */
int ltxb_attention_fwd(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                       int64_t ldo, int32_t B, int32_t Tq, int32_t Tk, int32_t H, int32_t dh, float scale,
                       const float* kv_bias, void* stream);

/* Scratch for the key-split of the ragged last wave of attention CTAs (Tq > 128): when the (batch, head, 256-query)
 * jobs do not divide the SM count, the jobs of the last partial wave are cut into key ranges whose partial
 * (O, max, sum) are parked here and merged by a second small kernel.  The library never allocates: the host
 * framework registers one buffer per device (bytes >= ltxb_attention_workspace_bytes()); without it (or with
 * workspace == NULL) the last wave simply runs unsplit.  The buffer must stay valid while attention calls are in
 * flight; its contents are scratch. */
int64_t ltxb_attention_workspace_bytes(void);
int ltxb_attention_set_workspace(void* workspace, int64_t bytes);

/* The same with the Ulysses sequence-gather all-to-all fused into the epilogue (B = 1): output row r is stored to
 * o_peers[r / rows_per_peer] + (r % rows_per_peer) * ldo — rank i's NVLink-mapped receive buffer, already offset
 * to this rank's head-group chunk — instead of a local O followed by a collective. */
/* ..._sync: with the flag barrier that orders the peers' q/k/v stores (ltxb_qkv_norm_rope_scatter_peers on every rank)
 * before this kernel's reads folded into its prologue (sync may be NULL = plain ltxb_attention_fwd_peers). */
int ltxb_attention_fwd_peers_sync(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                  void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, int64_t ldo, int32_t Tq,
                                  int32_t Tk, int32_t H, int32_t dh, float scale, const ltxb_peer_sync* sync, void* stream);
int ltxb_attention_fwd_peers(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                             void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, int64_t ldo, int32_t Tq,
                             int32_t Tk, int32_t H, int32_t dh, float scale, void* stream);

/* Attention over a SLICE of the keys, left un-normalised, and the merge of such slices (new: video->audio attention with
 * the video rows sharded across ranks, SURVEY.md 8e — queries = the replicated audio stream, keys / values = this rank's
 * video rows; transformer.py:326-339 computes it over all keys at once).  ltxb_attention_partial writes, for every
 * (batch, head, query row), O~ = sum_j 2^(s_j - m) v_j (f32 [dh]), the stabiliser m (log2 domain) and l = sum_j 2^(s_j - m)
 * into `part`: [B*H*Tq][dh] floats followed by [B*H*Tq][2] (ltxb_attention_partial_floats floats in all, Tq <= 256).
 * ltxb_attention_merge combines n_parts such blocks (part i at parts + i*part_stride floats — e.g. the all-gathered blocks
 * of the ranks) by log-sum-exp:  O = sum_i 2^(m_i - M) O~_i / sum_i 2^(m_i - M) l_i  -> bf16 O [B*Tq, H*dh]. */
int64_t ltxb_attention_partial_floats(int32_t B, int32_t Tq, int32_t H, int32_t dh);
int ltxb_attention_partial(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, float* part,
                           int32_t B, int32_t Tq, int32_t Tk, int32_t H, int32_t dh, float scale, void* stream);
int ltxb_attention_merge(const float* parts, int64_t part_stride, int32_t n_parts, void* O, int64_t ldo, int32_t B, int32_t Tq,
                         int32_t H, int32_t dh, void* stream);

/* N3  LoRA merged into a non-quantised bf16 weight, the path the reference takes for such checkpoints
 *     (lora.py:93-129 via generate.py:2997-3007):  w[r,c] = bf16( w[r,c] + bf16(delta[r,c] * strength) )
 *     delta f32 [R,C] = B (out, rank) . A (rank, in), produced by ltxb_gemm_bf16 (LTXB_EPI_BIAS_F32, no bias) with
 *     a = B and W = A^T, rank zero-padded to a multiple of 64.  Both roundings of the reference are kept. */
int ltxb_lora_merge_bf16(void* w, int64_t ldw, const float* delta, int64_t ldd, int64_t R, int32_t C, float strength,
                         void* stream);

/* N3  MLX affine group-quantised linears (pre-quantised 4/8-bit checkpoints: ltx.py:614-725 swaps the linears named by
 *     ".scales" tensors for nn.QuantizedLinear, whose forward is mx.quantized_matmul).  On a GPU that keeps all 25.8 GB
 *     of bf16 weights resident the packed form buys nothing at M >= 1280 rows (the GEMMs are tensor-bound), so the
 *     checkpoint is expanded ONCE at load — exactly what MLX's matrix kernels do per tile — and the bf16 GEMM runs:
 *         out[r,c] = bf16( scales[r, c/G] * q[r,c] + biases[r, c/G] )
 *     wq: uint32 [R, C*bits/32] (ldq words per row), level c of a row in bits [bits*(c % (32/bits)), ...) of word
 *     c / (32/bits);  scales / biases: bf16 (aux_f32 = 0) or f32 (aux_f32 = 1) [R, C/G] (lds elements per row);
 *     out: bf16 [R, C] (ldo, may be a row view of fused q|k|v storage).  bits in {2,4,8}, G in {32,64,128}. */
int ltxb_dequant_affine_bf16(const uint32_t* wq, int64_t ldq, const void* scales, const void* biases, int64_t lds,
                             int32_t aux_f32, void* out, int64_t ldo, int64_t R, int32_t C, int32_t group_size,
                             int32_t bits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N4  latent 2x spatial upsampler between the pipeline stages (mlx_video/models/ltx/upsampler.py).
 *     Activations: f32, channels-last [N, D, H, W, C].  Every convolution is ltxb_gemm_bf16 over the rows built here.
 * ---------------------------------------------------------------------------------------------- */
/* Conv3d / nn.Conv2d operand (upsampler.py:51-72,149,163; mx.conv3d, stride 1, "same" zero padding):
 *   out bf16 [N*D*H*W, kd*kh*kw*C], out[m, ((kz*kh + ky)*kw + kx)*C + c] = x[n, d+kz-kd/2, h+ky-kh/2, w+kx-kw/2, c]
 *   — the K order of the reference's weight layout (C_out, kd, kh, kw, C_in), which is therefore the GEMM's W as stored.
 *   kd = 1 gives the frame-by-frame 2-D convolution of SpatialRationalResampler.  C % 8 == 0, odd kernel sizes. */
int ltxb_im2col_cl(const float* x, void* out, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t kd, int32_t kh,
                   int32_t kw, void* stream);

/* GroupNorm3d (upsampler.py:85-114: f32 statistics per sample and group over S = D*H*W positions x C/G channels,
 * population variance, eps inside the sqrt) fused with what follows it in ResBlock3D / LatentUpsampler
 * (upsampler.py:189-197,258-262):  out = silu?( (x - mean_g) * rstd_g * weight[c] + bias[c] [+ resid] ).
 *   x, out, resid: f32 [N, S, C] (out may alias x);  two launches (per-chunk partial sums, then one pass).
 *   workspace: caller-owned scratch of ltxb_groupnorm_workspace_bytes(N, S, G) bytes.  C <= 1024, C % G == 0, G <= 64. */
int64_t ltxb_groupnorm_workspace_bytes(int32_t N, int64_t S, int32_t G);
int ltxb_groupnorm_silu(const float* x, float* out, int32_t N, int64_t S, int32_t C, int32_t G, float eps, const float* weight,
                        const float* bias, const float* resid, int32_t silu, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* PixelShuffle2D(2) (upsampler.py:124-139): x f32 [F, H, W, 4*Co] -> out f32 [F, 2H, 2W, Co], input channel (co*2 + rh)*2 + rw. */
int ltxb_pixel_shuffle2(const float* x, float* out, int64_t F, int32_t H, int32_t W, int32_t Co, void* stream);

/* The layout moves of LatentUpsampler.__call__ (upsampler.py:250,290) carrying upsample_latents' un- / re-normalisation
 * (upsampler.py:305-314): to_channels_last != 0: x f32 (B, C, S) -> out (B, S, C), out = x * scale[c] + shift[c];
 * else x (B, S, C) -> out (B, C, S), out = (x - shift[c]) / scale[c].  scale = shift = NULL: plain transposition. */
int ltxb_latent_layout(const float* x, float* out, const float* scale, const float* shift, int64_t B, int32_t C, int64_t S,
                       int32_t to_channels_last, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a21/a22 sampler-side elementwise (utils.py:404-440; generate.py:1255,1283,1288-1301)
 *   CFG combine + to_denoised + fp32 Euler in one pass over the latent:
 *     v  = v_pos + (cfg_scale - 1) (v_pos - v_neg)           (v_neg NULL -> v = v_pos)
 *     x0 = x - sigma_tok * v ; if mask: x0 = x0*mask + clean*(1-mask)
 *     x' = x0 + sigma_next * (x - x0) / sigma
 *   x, v_*, clean: f32 [n_tok, C];  sigma_tok f32 [n_tok] or NULL (-> sigma); mask f32 [n_tok] or NULL.
 * ---------------------------------------------------------------------------------------------- */
int ltxb_euler_step(float* x, const float* v_pos, const float* v_neg, float cfg_scale, const float* sigma_tok,
                    float sigma, float sigma_next, const float* mask, const float* clean, int64_t n_tok,
                    int32_t C, float* x0_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row N4, second half: the LTX-2 video VAE decoder (video_vae/decoder.py:94-450) and tiled decoding (tiling.py:279-520).
 *     Activations: f32, channels-last [N, D, H, W, C]; every CausalConv3d is ltxb_gemm_bf16 over the rows built here.
 * ---------------------------------------------------------------------------------------------- */
/* CausalConv3d operand (convolution.py:120-166, kernel 3, stride 1): rows [m0, m0 + rows) of the [N*D*H*W, 27*C] bf16
 * matrix whose row m holds the 27 taps of output position m — temporal padding by frame replication (causal != 0: two
 * copies of the first frame; else first + last), reflect padding in H and W (convolution.py:13-40).  pre_op != 0 applies
 * the chain in front of a ResNet convolution to every tap while it is gathered (decoder.py:140-180, 357-359, 425-440):
 *     silu( x / sqrt(mean_c(x^2) + eps) * (1 + table_scale[c] + emb_scale[n, c]) + table_shift[c] + emb_shift[n, c] )
 * (tables / embeddings may be NULL: pixel norm + SiLU only).  HBM: 27 x (4C read (L2-resident after the first tap) +
 * 2C written) bytes per position. */
int ltxb_vae_gather_rows(const float* x, void* out, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t causal,
                         int64_t m0, int64_t rows, const float* table_scale, const float* table_shift, const float* emb_scale,
                         const float* emb_shift, int64_t emb_ld, float eps, int32_t pre_op, void* stream);
/* DepthToSpaceUpsample, stride (2,2,2), residual, first frame dropped (sampling.py:143-197): y = the convolution's
 * output [N, D, H, W, 4C], x = its input [N, D, H, W, C] -> out [N, 2D-1, 2H, 2W, C/2]:
 *     out[n, 2d+st-1, 2h+sh, 2w+sw, c] = y[n,d,h,w, ((c*2+st)*2+sh)*2+sw] + x[n,d,h,w, (((c % (C/8))*2+st)*2+sh)*2+sw] */
int ltxb_vae_depth_to_space(const float* y, const float* x, float* out, int64_t N, int32_t D, int32_t H, int32_t W, int32_t C,
                            void* stream);
/* decoder.py:380-384: channels-first latents (N, C, S) [+ noise] -> channels-last (N, S, C):
 *     out = (noise * noise_scale + (1 - noise_scale) * sample) * std[c] + mean[c]        (noise NULL = zeros) */
int ltxb_vae_prepare_latent(const float* sample, const float* noise, float noise_scale, const float* stdv, const float* mean,
                            float* out, int64_t N, int32_t C, int64_t S, void* stream);
/* ops.py:47-80 (patch 4): z channels-last [N, F, H, W, 48] -> video channels-first (N, 3, F, 4H, 4W),
 *     video[n, c, f, 4h+pq, 4w+pr] = z[n, f, h, w, (c*4+pr)*4+pq] */
int ltxb_vae_unpatchify(const float* z, float* video, int64_t N, int32_t F, int32_t H, int32_t W, void* stream);
/* Tiled decoding (tiling.py:404-470): the (at, ah, aw) corner of a decoded tile (N, 3, Ft, Ht, Wt) is accumulated into
 * output (N, 3, F, H, W) at (t0, h0, w0) with the separable trapezoid mask mt[t] * mh[h] * mw[w]; weights (N, 1, F, H, W)
 * accumulates the mask.  ltxb_vae_blend_normalize: output /= max(weights, 1e-8) (tiling.py:506-508), plane = F*H*W. */
int ltxb_vae_blend_tile(const float* tile, int64_t N, int32_t Ft, int32_t Ht, int32_t Wt, int32_t at, int32_t ah, int32_t aw,
                        const float* mt, const float* mh, const float* mw, float* output, float* weights, int32_t F, int32_t H,
                        int32_t W, int32_t t0, int32_t h0, int32_t w0, void* stream);
int ltxb_vae_blend_normalize(float* output, const float* weights, int64_t N, int64_t plane, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTXB_H_ */
