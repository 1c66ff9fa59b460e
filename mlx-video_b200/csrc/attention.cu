// K2: non-causal multi-head attention, FlashAttention-style online softmax, both contractions on
// tcgen05 tensor cores with TMEM accumulators, operands staged by TMA.
//
// Replaces mlx_video/models/ltx/attention.py:13-53 (reshape to heads + mx.fast.scaled_dot_product_attention):
// video self-attention, text cross-attention and the audio<->video cross-attentions of
// transformer.py:247-339 all come through here (dh = 128 video, 64 audio).
//
// One CTA = one 128-row query tile of one (batch, head); it streams 128-row K/V tiles.
//   warp 0      TMA producer: Q once, then K and V tiles into 2-deep rings
//   warp 1      MMA issuer:   S_j = Q K_j^T  (K-major x K-major)      -> TMEM S[j&1]   (128 fp32 cols)
//                             O  += P_j V_j  (K-major x MN-major V)   -> TMEM O        (dh fp32 cols)
//   warps 2..9  softmax: TWO threads per query row (TMEM lane), one per half of the 128 score columns —
//               two warps per SM sub-partition, so one warp's dependent ALU / MUFU latencies hide behind the
//               other's (with a single warp per sub-partition the kernel was softmax-issue bound).  The halves
//               exchange their row maxima through smem (one named barrier per KV tile), keep partial row sums,
//               write their half of P_j to smem as bf16 in the canonical 128B-swizzled K-major layout and
//               rescale / store their half of the O columns.  O is rescaled in TMEM only when a maximum moved.
// S_{j+1} is issued before P_j V_j, so the tensor core computes the next scores while the softmax
// warps work on the current ones.
#include <cstdlib>

#include "attention.cuh"
#include "ptx.cuh"

namespace ltxb {

constexpr int kSplit = 2;  // softmax threads per query row
constexpr int kAttnThreads = 64 + 128 * kSplit;
constexpr int kTileQ = 128;
constexpr int kTileKV = 128;
constexpr int kAttnHeader = 4096;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO = 256;

struct AttnSmemHeader {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2];
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full;
  uint64_t pv_done;
  uint32_t tmem_base;
  float row_stat[2][kSplit][128];  // per-tile partial row maxima (double-buffered); partial row sums at the end
};
static_assert(sizeof(AttnSmemHeader) <= kAttnHeader, "header overflow");

template <int kDh>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  constexpr int kBlocks = kDh / 64;                  // 64-column (128 B) swizzle blocks per row
  constexpr uint32_t kBlockBytes = 128 * 128;        // 128 rows x 128 B
  constexpr uint32_t kTileBytes = kBlocks * kBlockBytes;  // one Q / K / V tile
  constexpr uint32_t kPBytes = 2 * kBlockBytes;      // P: 128 x 128 bf16

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  AttnSmemHeader* hdr = reinterpret_cast<AttnSmemHeader*>(smem);
  uint8_t* sQ = smem + kAttnHeader;
  uint8_t* sK = sQ + kTileBytes;       // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;   // 2 stages
  uint8_t* sP = sV + 2 * kTileBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kTileQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.Tk + kTileKV - 1) / kTileKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(&hdr->q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&hdr->k_full[i], 1);
        mbar_init(&hdr->k_empty[i], 1);
        mbar_init(&hdr->v_full[i], 1);
        mbar_init(&hdr->v_empty[i], 1);
        mbar_init(&hdr->s_full[i], 1);
        mbar_init(&hdr->s_empty[i], 128 * kSplit);
      }
      mbar_init(&hdr->p_full, 128 * kSplit);
      mbar_init(&hdr->pv_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(&hdr->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&hdr->tmem_base);
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&hdr->q_full, kTileBytes);
#pragma unroll
      for (int j = 0; j < kBlocks; ++j) tma_load_3d(sQ + j * kBlockBytes, &tmap_q, &hdr->q_full, h * kDh + 64 * j, q0, b);
      for (int it = 0; it < n_kv; ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&hdr->k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&hdr->k_full[st], kTileBytes);
#pragma unroll
        for (int j = 0; j < kBlocks; ++j)
          tma_load_3d(sK + st * kTileBytes + j * kBlockBytes, &tmap_k, &hdr->k_full[st], h * kDh + 64 * j, it * kTileKV, b);
        mbar_wait(&hdr->v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&hdr->v_full[st], kTileBytes);
#pragma unroll
        for (int j = 0; j < kBlocks; ++j)
          tma_load_3d(sV + st * kTileBytes + j * kBlockBytes, &tmap_v, &hdr->v_full[st], h * kDh + 64 * j, it * kTileKV, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(kTileQ, kTileKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(kTileQ, kDh, 0, 1);  // B = V is MN-major (dh contiguous)
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&hdr->s_empty[st], ph ^ 1);
        mbar_wait(&hdr->k_full[st], ph);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + st * kTileBytes);
        const uint32_t d = tmem_base + (st ? kColS1 : kColS0);
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBlockBytes + (kk & 3) * 32;
          umma_bf16_ss<1>(d, make_smem_desc_sw128(q_addr + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                          idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(&hdr->k_empty[st]);
        umma_commit(&hdr->s_full[st]);
      };
      mbar_wait(&hdr->q_full, 0);
      issue_s(0);
      for (int it = 0; it < n_kv; ++it) {
        if (it + 1 < n_kv) issue_s(it + 1);
        const int st = it & 1;
        mbar_wait(&hdr->p_full, it & 1);
        mbar_wait(&hdr->v_full[st], (it >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + st * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kTileKV / 16; ++kk) {
          const uint32_t p_off = (kk >> 2) * kBlockBytes + (kk & 3) * 32;
          // V tile: [kv rows][64-col block] with 128 B rows; 16 kv rows = 2048 B; dh blocks 16 KB apart
          umma_bf16_ss<1>(tmem_base + kColO, make_smem_desc_sw128(p_addr + p_off, 16, 1024),
                          make_smem_desc_sw128(v_addr + kk * 2048, kBlockBytes, 1024), idesc_o, (it | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&hdr->v_empty[st]);
        umma_commit(&hdr->pv_done);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (kSplit threads per query row) =========
    constexpr int kCols = 128 / kSplit;   // score columns per thread
    constexpr int kOCols = kDh / kSplit;  // O columns per thread
    static_assert(kCols % 32 == 0 && kOCols % 32 == 0, "column split must keep 32-column TMEM accesses");
    const int quarter = warp & 3;
    const int part = (warp - 2) >> 2;    // which column slice of the row this thread owns
    const int r = quarter * 32 + lane;   // row inside the tile == TMEM lane
    const int c0 = part * kCols;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    auto softmax_bar = [] { asm volatile("bar.sync 1, %0;" ::"n"(128 * kSplit) : "memory"); };
    float m = -INFINITY, l = 0.f;
    constexpr float kLog2e = 1.4426950408889634f;
    for (int it = 0; it < n_kv; ++it) {
      const int st = it & 1;
      mbar_wait(&hdr->s_full[st], (it >> 1) & 1);
      tc_fence_after_sync();
      uint32_t sr[kCols];
      {
        const uint32_t t_s = t_lane + (st ? kColS1 : kColS0) + c0;
#pragma unroll
        for (int c = 0; c < kCols; c += 32) tmem_ld_x32(t_s + c, *reinterpret_cast<uint32_t(*)[32]>(&sr[c]));
        tmem_wait_ld();
      }
      tc_fence_before_sync();
      mbar_arrive(&hdr->s_empty[st]);

      const int kv_valid = p.Tk - it * kTileKV - c0;  // valid columns of this thread's slice (may be <= 0)
      float mt = -INFINITY;
      if (p.kv_bias != nullptr) {
        const float* bias = p.kv_bias + static_cast<long long>(b) * p.Tk + it * kTileKV + c0;
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
          float s = __uint_as_float(sr[c]) * p.scale_log2;
          if (c < kv_valid) s += __ldg(bias + c) * kLog2e; else s = -INFINITY;
          sr[c] = __float_as_uint(s);
          mt = fmaxf(mt, s);
        }
      } else if (kv_valid < kCols) {
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
          float s = __uint_as_float(sr[c]) * p.scale_log2;
          if (c >= kv_valid) s = -INFINITY;
          sr[c] = __float_as_uint(s);
          mt = fmaxf(mt, s);
        }
      } else {
#pragma unroll
        for (int c = 0; c < kCols; ++c) {
          const float s = __uint_as_float(sr[c]) * p.scale_log2;
          sr[c] = __float_as_uint(s);
          mt = fmaxf(mt, s);
        }
      }
      // the row's maximum over all slices (double-buffered exchange: one barrier per tile is enough)
      hdr->row_stat[st][part][r] = mt;
      softmax_bar();
#pragma unroll
      for (int q = 0; q < kSplit; ++q) mt = fmaxf(mt, hdr->row_stat[st][q][r]);
      const float m_new = fmaxf(m, mt);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = fast_exp2(m - m_use);  // m = -inf -> 0
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < kCols; ++c) {
        const float e = fast_exp2(__uint_as_float(sr[c]) - m_use);
        rs += e;
        sr[c] = __float_as_uint(e);
      }
      l = l * alpha + rs;
      m = m_new;

      // P smem and the O accumulator are free once the previous P.V has completed
      if (it > 0) {
        mbar_wait(&hdr->pv_done, (it - 1) & 1);
        tc_fence_after_sync();
      }
      // P_j -> smem, bf16, K-major, 128-byte swizzle: 16-byte chunk index XOR (row % 8)
      {
        uint8_t* prow = sP + r * 128;
#pragma unroll
        for (int j8 = 0; j8 < kCols / 8; ++j8) {
          const int c = c0 + j8 * 8;         // first of 8 score columns
          const int blk = c >> 6, j = (c & 63) >> 3;
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(sr[j8 * 8 + 0]), __uint_as_float(sr[j8 * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(sr[j8 * 8 + 2]), __uint_as_float(sr[j8 * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(sr[j8 * 8 + 4]), __uint_as_float(sr[j8 * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(sr[j8 * 8 + 6]), __uint_as_float(sr[j8 * 8 + 7]));
          *reinterpret_cast<uint4*>(prow + blk * kBlockBytes + ((j ^ (r & 7)) << 4)) = w;
        }
      }
      // rescale this thread's slice of the running O (TMEM) when some row of this warp moved its maximum
      if (it > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
        for (int c = 0; c < kOCols; c += 32) {
          uint32_t o[32];
          tmem_ld_x32(t_lane + kColO + part * kOCols + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x32(t_lane + kColO + part * kOCols + c, o);
        }
        tmem_wait_st();
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&hdr->p_full);
    }
    // ---- row sum over all slices, then epilogue: O / l -> global bf16 ----
    softmax_bar();  // every thread is done reading the last tile's maxima
    hdr->row_stat[0][part][r] = l;
    softmax_bar();
    l = 0.f;
#pragma unroll
    for (int q = 0; q < kSplit; ++q) l += hdr->row_stat[0][q][r];
    mbar_wait(&hdr->pv_done, (n_kv - 1) & 1);
    tc_fence_after_sync();
    const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
    const int row = q0 + r;
    __nv_bfloat16* orow = attn_out_row(p, b, row, h, kDh) + part * kOCols;
#pragma unroll
    for (int c = 0; c < kOCols; c += 32) {
      uint32_t o[32];
      tmem_ld_x32(t_lane + kColO + part * kOCols + c, o);
      tmem_wait_ld();
      if (row < p.Tq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c + 8 * i) = w;
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int kDh>
static int launch_attention(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                            const AttnParams& p, cudaStream_t stream) {
  constexpr size_t kTileBytes = static_cast<size_t>(kDh / 64) * 128 * 128;
  constexpr size_t smem = 1024 + kAttnHeader + 5 * kTileBytes + 2 * 128 * 128;
  auto kernel = attention_kernel<kDh>;
  static PerDeviceOnce configured;  // per instantiation and device
  if (configured.first()) {
    LTXB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  CUtensorMap tq, tk, tv;
  const uint32_t box[3] = {64, 128, 1};
  auto enc = [&](CUtensorMap* m, const void* base, long long ld, int T) {
    const uint64_t dims[3] = {static_cast<uint64_t>(p.H) * kDh, static_cast<uint64_t>(T), static_cast<uint64_t>(p.B)};
    const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * static_cast<uint64_t>(T)};
    return encode_tmap_bf16(m, base, 3, dims, strides, box);
  };
  int rc;
  if ((rc = enc(&tq, Q, ldq, p.Tq))) return rc;
  if ((rc = enc(&tk, K, ldk, p.Tk))) return rc;
  if ((rc = enc(&tv, V, ldv, p.Tk))) return rc;
  dim3 grid((p.Tq + kTileQ - 1) / kTileQ, p.H, p.B);
  LTXB_CUDA(launch_kernel(kernel, grid, dim3(kAttnThreads), smem, stream, 1, tq, tk, tv, p));
  return LTXB_OK;
}

}  // namespace ltxb

using namespace ltxb;

static int attention_impl(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                          int64_t ldo, int32_t B, int32_t Tq, int32_t Tk, int32_t H, int32_t dh, float scale,
                          const float* kv_bias, void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, void* stream,
                          const ltxb_peer_sync* sync = nullptr);

extern "C" int ltxb_attention_fwd(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                  void* O, int64_t ldo, int32_t B, int32_t Tq, int32_t Tk, int32_t H, int32_t dh,
                                  float scale, const float* kv_bias, void* stream) {
  return attention_impl(Q, ldq, K, ldk, V, ldv, O, ldo, B, Tq, Tk, H, dh, scale, kv_bias, nullptr, 0, 0, stream);
}

extern "C" int ltxb_attention_fwd_peers_sync(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                             void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, int64_t ldo, int32_t Tq,
                                             int32_t Tk, int32_t H, int32_t dh, float scale, const ltxb_peer_sync* sync, void* stream) {
  LTXB_CHECK_ARG(o_peers && n_peers >= 1 && n_peers <= 8 && rows_per_peer > 0 && n_peers * rows_per_peer >= Tq,
                 "ltxb_attention_fwd_peers_sync: %d peers x %d rows do not cover Tq=%d", n_peers, rows_per_peer, Tq);
  return attention_impl(Q, ldq, K, ldk, V, ldv, o_peers[0], ldo, 1, Tq, Tk, H, dh, scale, nullptr, o_peers, n_peers, rows_per_peer, stream, sync);
}

extern "C" int ltxb_attention_fwd_peers(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                        void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, int64_t ldo,
                                        int32_t Tq, int32_t Tk, int32_t H, int32_t dh, float scale, void* stream) {
  LTXB_CHECK_ARG(o_peers && n_peers >= 1 && n_peers <= 8 && rows_per_peer > 0 && n_peers * rows_per_peer >= Tq,
                 "ltxb_attention_fwd_peers: %d peers x %d rows do not cover Tq=%d", n_peers, rows_per_peer, Tq);
  return attention_impl(Q, ldq, K, ldk, V, ldv, o_peers[0], ldo, 1, Tq, Tk, H, dh, scale, nullptr, o_peers, n_peers, rows_per_peer, stream);
}

extern "C" int64_t ltxb_attention_partial_floats(int32_t B, int32_t Tq, int32_t H, int32_t dh) {
  const long long n_qp = (Tq + 255) / 256;
  const long long rows = static_cast<long long>(B) * H * n_qp * (n_qp == 1 ? Tq : 256);
  return ((rows * (dh + 2) + 3) / 4) * 4;
}

extern "C" int ltxb_attention_partial(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, float* part,
                                      int32_t B, int32_t Tq, int32_t Tk, int32_t H, int32_t dh, float scale, void* stream) {
  LTXB_CHECK_ARG(Q && K && V && part, "ltxb_attention_partial: null pointer");
  LTXB_CHECK_ARG(B > 0 && Tq > 0 && Tk > 0 && H > 0, "ltxb_attention_partial: bad shape B=%d Tq=%d Tk=%d H=%d", B, Tq, Tk, H);
  LTXB_CHECK_SUPPORTED(dh == 64 || dh == 128, "ltxb_attention_partial: head dim %d not in {64,128}", dh);
  LTXB_CHECK_ARG(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(part), "ltxb_attention_partial: pointers must be 16-byte aligned");
  LTXB_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldq >= H * dh && ldk >= H * dh && ldv >= H * dh,
                 "ltxb_attention_partial: leading dims must be multiples of 8 covering H*dh");
  AttnParams p{};
  p.B = B, p.Tq = Tq, p.Tk = Tk, p.H = H;
  p.scale_log2 = scale * 1.4426950408889634f;
  const long long n_qp = (Tq + 255) / 256;
  const long long rows = static_cast<long long>(B) * H * n_qp * (n_qp == 1 ? Tq : 256);
  LTXB_CHECK_SUPPORTED((rows * dh) % 4 == 0, "ltxb_attention_partial: partial O block must stay 16-byte aligned");
  return launch_attention_partial(Q, ldq, K, ldk, V, ldv, p, dh, part, part + rows * dh, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ltxb_attention_merge(const float* parts, int64_t part_stride, int32_t n_parts, void* O, int64_t ldo, int32_t B, int32_t Tq,
                                    int32_t H, int32_t dh, void* stream) {
  LTXB_CHECK_ARG(parts && O && n_parts >= 1 && B > 0 && Tq > 0 && H > 0, "ltxb_attention_merge: bad argument");
  LTXB_CHECK_SUPPORTED(dh == 64 || dh == 128, "ltxb_attention_merge: head dim %d not in {64,128}", dh);
  LTXB_CHECK_ARG(aligned16(parts) && aligned16(O) && ldo % 8 == 0 && ldo >= H * dh, "ltxb_attention_merge: misaligned operands");
  AttnParams p{};
  p.B = B, p.Tq = Tq, p.Tk = 0, p.H = H;
  p.O = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  return launch_attention_merge(p, dh, parts, part_stride, n_parts, reinterpret_cast<cudaStream_t>(stream));
}

static int attention_impl(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                          int64_t ldo, int32_t B, int32_t Tq, int32_t Tk, int32_t H, int32_t dh, float scale,
                          const float* kv_bias, void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, void* stream,
                          const ltxb_peer_sync* sync) {
  LTXB_CHECK_ARG(Q && K && V && O, "ltxb_attention_fwd: null pointer");
  if (B == 0 || Tq == 0) return LTXB_OK;
  LTXB_CHECK_ARG(B > 0 && Tq > 0 && Tk > 0 && H > 0, "ltxb_attention_fwd: bad shape B=%d Tq=%d Tk=%d H=%d", B, Tq, Tk, H);
  LTXB_CHECK_SUPPORTED(dh == 64 || dh == 128, "ltxb_attention_fwd: head dim %d not in {64,128}", dh);
  LTXB_CHECK_SUPPORTED(B <= 65535 && H <= 65535, "ltxb_attention_fwd: B and H must fit a grid dimension");
  LTXB_CHECK_ARG(aligned16(Q) && aligned16(K) && aligned16(V) && aligned16(O), "ltxb_attention_fwd: pointers must be 16-byte aligned");
  LTXB_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "ltxb_attention_fwd: leading dims must be multiples of 8");
  LTXB_CHECK_ARG(ldq >= H * dh && ldk >= H * dh && ldv >= H * dh && ldo >= H * dh, "ltxb_attention_fwd: leading dims must cover H*dh");
  AttnParams p{};
  p.B = B, p.Tq = Tq, p.Tk = Tk, p.H = H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.kv_bias = kv_bias;
  p.rows_per_peer = rows_per_peer;
  for (int i = 0; i < 8; ++i) {
    void* base = (i < n_peers) ? o_peers[i] : (n_peers > 0 ? o_peers[n_peers - 1] : nullptr);
    if (i < n_peers) LTXB_CHECK_ARG(base && aligned16(base), "ltxb_attention_fwd_peers: null / misaligned peer base %d", i);
    p.o_peer[i] = reinterpret_cast<__nv_bfloat16*>(base);
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = peer_sync_from_abi(sync, &p.sync, "ltxb_attention_fwd_peers_sync")) return rc;
  // more than one query tile: two tiles per CTA ping-pong on the tensor core (attention_pair.cu)
  static const bool force_single = [] { const char* e = getenv("LTXB_ATTN_SINGLE"); return e != nullptr && atoi(e) != 0; }();
  if (Tq > kTileQ && !force_single) return launch_attention_pair(Q, ldq, K, ldk, V, ldv, p, dh, s);
  if (p.sync.n_peers > 0) {  // the one-tile kernel has no folded barrier: run it as a kernel of its own
    if (int rc = launch_peer_barrier(p.sync, s)) return rc;
    p.sync = PeerSync{};
  }
  if (dh == 128) return launch_attention<128>(Q, ldq, K, ldk, V, ldv, p, s);
  return launch_attention<64>(Q, ldq, K, ldk, V, ldv, p, s);
}
