// K2, more than one query tile: two 128-row query tiles per CTA share every K/V tile and take turns on the
// tensor core, so one tile's softmax (MUFU / FMA pipes) runs under the other tile's contractions.
//
// Replaces mlx_video/models/ltx/attention.py:13-53 (mx.fast.scaled_dot_product_attention) for the video
// self-attention, the text cross-attention and a2v of transformer.py:247-339.
//
//   warp 0       TMA producer: Q0, Q1 once; K and V tiles into 2-deep rings
//   warp 1       MMA issuer (one thread):   S_t = Q_t K_j^T  (smem x smem)            -> TMEM S_t   (128 fp32 columns)
//                                           O_t += P_t V_j   (P_t read from TMEM)     -> TMEM O_t   (dh fp32 columns)
//   warps 2,3    idle (keep the softmax warps aligned to their TMEM lane quarters)
//   warps 4..7   softmax of tile 0, one thread per query row: the whole 128-column score row lives in registers,
//   warps 8..11  softmax of tile 1  so there is no cross-thread exchange; P_t is written back over S_t in TMEM as
//                packed bf16 (no shared-memory round trip) and consumed by the next tcgen05.mma as its A operand.
// Issue order per K/V tile j:  PV_0(j), S_0(j+1), PV_1(j), S_1(j+1)  — tile 0's softmax of j+1 overlaps tile 1's
// MMAs and vice versa.  The running maximum is only moved when a row's tile maximum exceeds it by more than
// 2^kTau (exponents stay in fp32 / bf16 range), so the O accumulator is almost never rescaled after tile 0.
//
// Ragged last wave: jobs = B*H*ceil(Tq/256) rarely divide the SM count.  The jobs of the last partial wave are
// cut into n_split key ranges, one CTA each; those CTAs park unnormalised (O, m, l) in the workspace and
// attention_combine_kernel merges them (log-sum-exp) into the output.
#include <algorithm>
#include <cstdlib>

#include "attention.cuh"
#include "ptx.cuh"

namespace ltxb {

constexpr int kPairThreads = 384;
constexpr int kPairHeader = 1024;
constexpr float kTau = 8.0f;  // log2 units
#ifndef LTXB_ATTN_EMU
#define LTXB_ATTN_EMU 2
#endif
constexpr float kMasked = -1.0e9f;  // score of an out-of-range key: finite, so the polynomial exp2 never sees -inf
#ifndef LTXB_ATTN_CHUNK
#define LTXB_ATTN_CHUNK 1
#endif
constexpr bool kChunkP = LTXB_ATTN_CHUNK != 0;  // signal P_t in two key halves so PV_t starts under the second half's exps
constexpr int kEmu = LTXB_ATTN_EMU;  // of every 8 column pairs, how many take the FMA-pipe exp2
constexpr int kRegsIssue = 104, kRegsSoftmax = 200;  // 128 * 104 + 256 * 200 <= 64 K registers

// LTXB_ATTN_TRACE: CTA 0 records clock64() at the phase boundaries of its first 16 key tiles into the workspace
// (long long [4 roles][16 tiles][8 events]; roles: softmax 0, softmax 1, MMA issuer, CTA life cycle) — scripts/attn_trace.py prints it.
#ifdef LTXB_ATTN_TRACE
#define TRACE(role, it, ev)                                                                                  \
  do {                                                                                                       \
    if (blockIdx.x == 0 && (it) < 16 && p.ws_o != nullptr)                                                   \
      reinterpret_cast<long long*>(p.ws_o)[((role) * 16 + (it)) * 8 + (ev)] = clock64();                      \
  } while (0)
#else
#define TRACE(role, it, ev) do {} while (0)
#endif

struct PairSmemHeader {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2];
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full[2];   // S_t(j) complete              (MMA -> softmax t)
  uint64_t p_full[2][2];  // P_t(j) keys [0,64) / [64,128) written, O_t rescaled (softmax t -> MMA); one arrival per warp
  uint64_t pv_done[2];  // O_t += P_t(j) V_j complete   (MMA -> softmax t)
  uint32_t tmem_base;
};
static_assert(sizeof(PairSmemHeader) <= kPairHeader, "header overflow");

struct PairJob {
  int b, h, q0, kv_lo, kv_hi, slot;  // slot < 0: whole job, normalised bf16 output
};

__device__ __forceinline__ PairJob decode_pair_job(const AttnParams& p, int cta) {
  PairJob j;
  const int n_kv = (p.Tk + 127) / 128;
  int job;
  if (cta < p.n_full) {
    job = cta, j.slot = -1, j.kv_lo = 0, j.kv_hi = n_kv;
  } else {
    const int idx = cta - p.n_full;
    job = p.n_full + idx / p.n_split;
    const int part = idx % p.n_split;
    j.slot = idx;
    j.kv_lo = static_cast<int>(static_cast<long long>(part) * n_kv / p.n_split);
    j.kv_hi = static_cast<int>(static_cast<long long>(part + 1) * n_kv / p.n_split);
  }
  const int bh = job / p.n_qp;
  j.q0 = (job - bh * p.n_qp) * 256;
  j.b = bh / p.H;
  j.h = bh - j.b * p.H;
  return j;
}

template <int kDh>
__global__ void __launch_bounds__(kPairThreads, 1)
attention_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  constexpr int kBlocks = kDh / 64;                       // 64-column (128 B) swizzle blocks per row
  constexpr uint32_t kBlockBytes = 128 * 128;             // 128 rows x 128 B
  constexpr uint32_t kTileBytes = kBlocks * kBlockBytes;  // one Q / K / V tile
  constexpr uint32_t kColO = 256;                         // O_t at kColO + t*kDh; S_t (and P_t) at t*128

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PairSmemHeader* hdr = reinterpret_cast<PairSmemHeader*>(smem);
  uint8_t* sQ = smem + kPairHeader;   // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;  // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;  // 2 stages

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TRACE(3, 0, 0);  // kernel entry
  const PairJob job = decode_pair_job(p, blockIdx.x);
  const int n_it = job.kv_hi - job.kv_lo;

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
        mbar_init(&hdr->p_full[i][0], 4);
        mbar_init(&hdr->p_full[i][1], 4);
        mbar_init(&hdr->pv_done[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(&hdr->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&hdr->tmem_base);
  if (threadIdx.x == 0) TRACE(3, 0, 1);  // barriers + TMEM ready
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) TRACE(3, 0, 2);  // predecessor kernel finished

  // the softmax threads hold a whole 128-column score row: they take registers from the producer / issuer warpgroup
  if (warp < 4) {
    reg_dealloc<kRegsIssue>();
    if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(&hdr->q_full, 2 * kTileBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int j = 0; j < kBlocks; ++j)
          tma_load_3d(sQ + t * kTileBytes + j * kBlockBytes, &tmap_q, &hdr->q_full, job.h * kDh + 64 * j, job.q0 + 128 * t, job.b);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int row = (job.kv_lo + it) * 128;
        mbar_wait(&hdr->k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&hdr->k_full[st], kTileBytes);
#pragma unroll
        for (int j = 0; j < kBlocks; ++j)
          tma_load_3d(sK + st * kTileBytes + j * kBlockBytes, &tmap_k, &hdr->k_full[st], job.h * kDh + 64 * j, row, job.b);
        mbar_wait(&hdr->v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&hdr->v_full[st], kTileBytes);
#pragma unroll
        for (int j = 0; j < kBlocks; ++j)
          tma_load_3d(sV + st * kTileBytes + j * kBlockBytes, &tmap_v, &hdr->v_full[st], job.h * kDh + 64 * j, row, job.b);
      }
    }
    } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, kDh, 0, 1);  // B = V is MN-major (dh contiguous)
      auto issue_s = [&](int t, int it) {                        // K stage it&1 is full
        const uint32_t q_addr = smem_u32(sQ + t * kTileBytes);
        const uint32_t k_addr = smem_u32(sK + (it & 1) * kTileBytes);
        const uint32_t d = tmem_base + t * 128;
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBlockBytes + (kk & 3) * 32;
          umma_bf16_ss<1>(d, make_smem_desc_sw128(q_addr + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024),
                          idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit(&hdr->s_full[t]);
      };
      // O_t += P_t(it) V: keys [0,64) as soon as the first half of P_t is in TMEM, the rest under the second half's exps
      auto issue_pv = [&](int t, int it) {  // V stage it&1 is full
        const uint32_t v_addr = smem_u32(sV + (it & 1) * kTileBytes);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (kChunkP || hf == 1) {
            mbar_wait(&hdr->p_full[t][hf], it & 1);
            tc_fence_after_sync();
          }
          TRACE(2, it, t * 4 + hf);  // P_t half hf seen by the issuer
#pragma unroll
          for (int kk = 4 * hf; kk < 4 * hf + 4; ++kk)  // 16 keys per MMA = 8 packed TMEM columns of P; V: 16 kv rows = 2048 B
            umma_bf16_ts(tmem_base + kColO + t * kDh, tmem_base + t * 128 + kk * 8,
                         make_smem_desc_sw128(v_addr + kk * 2048, kBlockBytes, 1024), idesc_o, (it | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&hdr->pv_done[t]);
        TRACE(2, it, t * 4 + 2);  // PV_t issued
      };
      mbar_wait(&hdr->q_full, 0);
      mbar_wait(&hdr->k_full[0], 0);
      tc_fence_after_sync();
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(&hdr->k_empty[0]);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const bool more = it + 1 < n_it;
        mbar_wait(&hdr->v_full[st], (it >> 1) & 1);
        issue_pv(0, it);
        if (more) {  // S_0(it+1) overwrites P_0(it): the tensor pipe runs MMAs in issue order
          mbar_wait(&hdr->k_full[st ^ 1], ((it + 1) >> 1) & 1);
          tc_fence_after_sync();
          issue_s(0, it + 1);
          TRACE(2, it, 3);  // S_0(it+1) issued
        }
        issue_pv(1, it);
        umma_commit(&hdr->v_empty[st]);
        if (more) {
          issue_s(1, it + 1);
          TRACE(2, it, 7);  // S_1(it+1) issued
          umma_commit(&hdr->k_empty[st ^ 1]);
        }
      }
    }
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row =====================
    reg_alloc<kRegsSoftmax>();
    const int t = (warp - 4) >> 2;      // query tile of this warpgroup
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_s = t_lane + t * 128;
    const uint32_t t_o = t_lane + kColO + t * kDh;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    for (int it = 0; it < n_it; ++it) {
      if (lane == 0 && quarter == 0) TRACE(t, it, 0);  // start waiting for S_t(it)
      mbar_wait(&hdr->s_full[t], it & 1);
      tc_fence_after_sync();
      if (lane == 0 && quarter == 0) TRACE(t, it, 1);  // S_t(it) ready
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 128; c += 32) tmem_ld_x32(t_s + c, *reinterpret_cast<uint32_t(*)[32]>(&sr[c]));
      tmem_wait_ld();
      if (lane == 0 && quarter == 0) TRACE(t, it, 2);  // scores in registers

      const int kv0 = (job.kv_lo + it) * 128;
      const int kv_valid = p.Tk - kv0;
      float sc = p.scale_log2;
      if (p.kv_bias != nullptr) {  // rare (context masks): fold scale and bias into the scores first
        const float* bias = p.kv_bias + static_cast<long long>(job.b) * p.Tk + kv0;
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const float bv = (c < kv_valid) ? __ldg(bias + c) * kLog2e : 0.f;
          sr[c] = __float_as_uint(fmaf(__uint_as_float(sr[c]), sc, bv));
        }
        sc = 1.0f;
      }
      if (kv_valid < 128) {
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (c >= kv_valid) sr[c] = __float_as_uint(kMasked);
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(sr[c + u]));
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;  // sc > 0
      // move the running maximum only when it is exceeded by more than 2^kTau (or on the first valid tile); keep it an
      // INTEGER (any stabiliser near the maximum will do) so that 2^(m - m') and the range reduction below are exact
      const float m_new = (m_tile > m + kTau) ? fmaxf(ceilf(m_tile), -1048576.0f) : m;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_new == m) ? 1.0f : fast_exp2(m - m_use);  // m = -inf -> 0
      // rescale the running O when some row of this warp moved its maximum (O is free once PV_t(it-1) is done)
      if (it > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
        mbar_wait(&hdr->pv_done[t], (it - 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < kDh; c += 32) {
          uint32_t o[32];
          tmem_ld_x32(t_o + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x32(t_o + c, o);
        }
      }
      // P = 2^(s*sc - m).  MUFU.EX2 (16 / clk / SM) is the scarcest unit of this kernel, so kEmu of every 8 column
      // pairs are computed on the FMA pipe instead: n = round(x) through the 1.5*2^23 trick, f = x - n in
      // [-0.5, 0.5], 2^f by a degree-3 minimax polynomial (7.5e-5 relative, far below the bf16 rounding of P),
      // 2^n by an add into the exponent field.  Packed f32x2 arithmetic throughout.
      if (lane == 0 && quarter == 0) TRACE(t, it, 3);  // maximum known
      constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
      const uint64_t sc2 = pack_f32x2(sc, sc);
      const uint64_t negm2 = pack_f32x2(-m_use, -m_use);
      const uint64_t c2 = pack_f32x2(kMagic - m_use, kMagic - m_use);  // exact: m_use is an integer below 2^22
      const uint64_t neg1 = pack_f32x2(-1.0f, -1.0f);
      const uint64_t k0 = pack_f32x2(0.99992807f, 0.99992807f), k1 = pack_f32x2(0.69326099f, 0.69326099f);
      const uint64_t k2 = pack_f32x2(0.24261114f, 0.24261114f), k3 = pack_f32x2(0.05517167f, 0.05517167f);
      uint64_t rs2[2] = {0ull, 0ull};
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[32];
#pragma unroll
        for (int pi = 0; pi < 32; ++pi) {
          const int c = half * 64 + 2 * pi;
          const uint64_t s2 = pack_f32x2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1]));
          float e0, e1;
          if ((pi & 7) < kEmu) {
            float x0, x1, p0, p1;
            unpack_f32x2(fma_f32x2(s2, sc2, c2), x0, x1);  // kMagic + round(x)
            x0 = fmaxf(x0, kMagic - 126.0f), x1 = fmaxf(x1, kMagic - 126.0f);
            const uint64_t f = fma_f32x2(s2, sc2, fma_f32x2(pack_f32x2(x0, x1), neg1, c2));  // x - round(x)
            unpack_f32x2(fma_f32x2(fma_f32x2(fma_f32x2(k3, f, k2), f, k1), f, k0), p0, p1);
            e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(x0) << 23));
            e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(x1) << 23));
          } else {
            float x0, x1;
            unpack_f32x2(fma_f32x2(s2, sc2, negm2), x0, x1);
            e0 = fast_exp2(x0), e1 = fast_exp2(x1);
          }
          rs2[pi & 1] = add_f32x2(rs2[pi & 1], pack_f32x2(e0, e1));
          pk[pi] = pack_bf16x2(e0, e1);
        }
        tmem_st_x32(t_s + half * 32, pk);  // P_t over the first 64 columns of S_t
        if (kChunkP || half == 1) {
          tmem_wait_st();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&hdr->p_full[t][half]);
        }
        if (lane == 0 && quarter == 0) TRACE(t, it, 4 + half);  // P half signalled
      }
      float r0, r1, r2, r3;
      unpack_f32x2(rs2[0], r0, r1);
      unpack_f32x2(rs2[1], r2, r3);
      l = l * alpha + ((r0 + r1) + (r2 + r3));
      m = m_new;
    }
    // ---- epilogue ----
    if (lane == 0 && quarter == 0) TRACE(3, 0, 3 + t);  // softmax of the last key tile done
    if (n_it > 0) {
      mbar_wait(&hdr->pv_done[t], (n_it - 1) & 1);
      tc_fence_after_sync();
    }
    if (lane == 0 && quarter == 0 && t == 0) TRACE(3, 0, 5);  // last PV done
    const int row = job.q0 + t * 128 + r;
    if (job.slot < 0) {
      const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
      __nv_bfloat16* orow = attn_out_row(p, job.b, row, job.h, kDh);
#pragma unroll
      for (int c = 0; c < kDh; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(t_o + c, o);
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
    } else {
      const long long prow = static_cast<long long>(job.slot) * p.part_rows + t * 128 + r;
      float* wo = p.ws_o + prow * kDh;
      if (row < p.Tq) *reinterpret_cast<float2*>(p.ws_ml + prow * 2) = make_float2(m, l);
#pragma unroll
      for (int c = 0; c < kDh; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(t_o + c, o);
        tmem_wait_ld();
        if (row < p.Tq) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(wo + c + 4 * i) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(3, 0, 6);  // output stored by every warp
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// ================================================================================================================
// attention_pair64_kernel: the same job layout (two 128-row query tiles per CTA, key-split ragged wave), but the
// scores are produced in 64-KEY STEPS into DOUBLE-BUFFERED TMEM tiles, so S_t(s+1) is on the tensor pipe a whole step
// ahead of its softmax.
//
// Why (profiles/r1e/attn_ncu_5184.md + the source-level samples of that report): in attention_pair_kernel S_t(j+1)
// overwrites P_t(j) in TMEM, so it can only be issued after PV_t(j); 52 % of the softmax warps' samples (31.5 % of all)
// sat on the wait for that S tile while no pipe was saturated (tensor 52 %, MUFU 40 %).  Here the chain per query tile is
// softmax(s) -> P(s) -> PV(s); S(s+1) is already in the other buffer when softmax(s) ends.
//
// TMEM (512 columns, all used at dh = 128): score buffer (tile t, buffer b) at (t*2 + b)*64 — 64 fp32 columns = 64 keys;
// P(s) is written back over its first 32 columns as packed bf16 — and O_t at 256 + t*dh.
// K / V stay 128-key TMA tiles in 2-deep smem rings; step s uses key rows [64*(s&1), +64) of tile s>>1.
//   warp 0      TMA producer (Q0, Q1 once; K, V tiles)
//   warp 1      MMA issuer.  Per step s:  S_0(s+1), S_1(s+1), then PV_0(s) / PV_1(s) in the order their P arrives
//   warps 2,3   idle (keep the softmax warps aligned to their TMEM lane quarters)
//   warps 4..7  softmax of tile 0, warps 8..11 of tile 1: one thread per query row, 64 scores per step in registers
// ================================================================================================================
struct Pair64Header {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2];
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full[2][2];  // S_t(s) complete in buffer s&1          (MMA -> softmax t)
  uint64_t p_full[2][2];  // P_t(s) written over buffer s&1         (softmax t -> MMA); one arrival per warp
  // O_t += P_t(s) V(s) complete (MMA -> softmax t), one barrier per score buffer: a softmax thread only knows that
  // PV_t(s-2) has completed when it holds S_t(s), so with a single barrier per tile a parity wait could not tell
  // "PV(s-2) done" from "PV(s) done" (two phases apart); per buffer the uncertainty is one phase
  uint64_t pv_done[2][2];
  uint32_t tmem_base;
};
static_assert(sizeof(Pair64Header) <= kPairHeader, "header overflow");

constexpr int kCmbJobs = 160;  // split jobs a launch can have (<= SM count): arrival counters, then departure counters
template <int kDh>
__device__ __forceinline__ void attention_combine_row(const AttnParams& p, int jl, int rr, int lane);

template <int kDh>
__global__ void __launch_bounds__(kPairThreads, 1)
attention_pair64_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  constexpr int kBlocks = kDh / 64;                       // 64-column (128 B) swizzle blocks per row
  constexpr uint32_t kBlockBytes = 128 * 128;             // 128 rows x 128 B
  constexpr uint32_t kTileBytes = kBlocks * kBlockBytes;  // one Q / K / V tile
  constexpr uint32_t kColO = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Pair64Header* hdr = reinterpret_cast<Pair64Header*>(smem);
  uint8_t* sQ = smem + kPairHeader;   // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;  // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;  // 2 stages

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TRACE(3, 0, 0);  // kernel entry
  const PairJob job = decode_pair_job(p, blockIdx.x);
  const int n_kt = job.kv_hi - job.kv_lo;                               // 128-key K / V tiles of this CTA
  const int s_lo = 2 * job.kv_lo;                                       // first 64-key step (global index)
  const int n_steps = min(2 * job.kv_hi, (p.Tk + 63) / 64) - s_lo;      // the last tile may hold a single step

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
        mbar_init(&hdr->v_empty[i], 2);  // one commit from each PV issuer
        for (int j = 0; j < 2; ++j) {
          mbar_init(&hdr->pv_done[i][j], 1);
          mbar_init(&hdr->s_full[i][j], 1);
          mbar_init(&hdr->p_full[i][j], 4);
        }
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(&hdr->tmem_base, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&hdr->tmem_base);
  if (threadIdx.x == 0) TRACE(3, 0, 1);  // barriers + TMEM ready
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) TRACE(3, 0, 2);  // predecessor kernel finished
  const int sync_epoch = peer_sync_enter(p.sync);  // Q / K / V stored by the peers: flag barrier before the first load

  if (warp < 4) {
    reg_dealloc<kRegsIssue>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(&hdr->q_full, 2 * kTileBytes);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            tma_load_3d(sQ + t * kTileBytes + j * kBlockBytes, &tmap_q, &hdr->q_full, job.h * kDh + 64 * j, job.q0 + 128 * t, job.b);
        for (int lt = 0; lt < n_kt; ++lt) {
          const int st = lt & 1;
          const uint32_t ph = (lt >> 1) & 1;
          const int row = (job.kv_lo + lt) * 128;
          mbar_wait(&hdr->k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->k_full[st], kTileBytes);
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            tma_load_3d(sK + st * kTileBytes + j * kBlockBytes, &tmap_k, &hdr->k_full[st], job.h * kDh + 64 * j, row, job.b);
          mbar_wait(&hdr->v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->v_full[st], kTileBytes);
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            tma_load_3d(sV + st * kTileBytes + j * kBlockBytes, &tmap_v, &hdr->v_full[st], job.h * kDh + 64 * j, row, job.b);
        }
      }
    } else {
      // ===================== MMA issuers: warp 1 = scores of both tiles, warp 2 / 3 = PV of tile 0 / 1 =====================
      // Three issuing warps on three different schedulers instead of one: a single issuing thread needed ~2500 cycles for
      // the ~350 instructions of one 64-key step (waits, polls, 24 MMAs) next to two busy softmax warps on its scheduler
      // — twice the 1280 cycles the tensor pipe needs for them (scripts/umma_bench.cu: 48 cycles per 128x64x16 SS MMA,
      // 64 per 128x128x16 TS MMA, with or without softmax-like background load) — so the pipe starved and the softmax
      // waited ~1000 cycles per step for its scores (profiles/r2/attention_s64_notes.md).  Each warp runs its control
      // flow warp-uniformly (descriptor arithmetic in uniform registers, ~3 issue slots per MMA) and one elected lane
      // issues.  Descriptors are {lo, hi} words: only the 14-bit start-address field of lo changes between MMAs.
      // Ordering that the single in-order issuer used to give for free: S_t(i) overwrites the buffer of P_t(i-2), so the
      // score warp waits for PV_t(i-2) to COMPLETE (pv_done) before it issues S(i); that is a full softmax ahead of need.
      const bool issuer = elect_one();
      constexpr uint32_t kHiK = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, 128B swizzle
      constexpr uint32_t kTile16 = kTileBytes >> 4, kBlock16 = kBlockBytes >> 4;
      if (warp == 1) {
        const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t q_lo = ((smem_u32(sQ) & 0x3FFFFu) >> 4) | (1u << 16);  // K-major operands: LBO field = 1 (unused)
        const uint32_t k_lo = ((smem_u32(sK) & 0x3FFFFu) >> 4) | (1u << 16);
        mbar_wait(&hdr->q_full, 0);
        for (int i = 0; i < n_steps; ++i) {
          const int lt = i >> 1;
          if ((i & 1) == 0) {  // step i opens key tile lt
            mbar_wait(&hdr->k_full[lt & 1], (lt >> 1) & 1);
          }
          // S_t(i) = Q_t . K[64 key rows of step i]^T into buffer i&1, tile after tile, each gated by ITS OWN PV_t(i-2)
          // only (the buffer still holds P_t(i-2) until that PV has read it): the two tiles' softmax phases — row maximum
          // first, MUFU idle; then the exponentials — drift out of phase instead of colliding on the MUFU pipe (+4-8 %
          // against issuing both tiles' MMAs interleaved behind a common wait, profiles/r2/attention_s64_notes.md)
          const uint32_t ka = k_lo + (lt & 1) * kTile16 + (i & 1) * ((64 * 128) >> 4);  // rows 64.. of every block
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (i >= 2) mbar_wait(&hdr->pv_done[t][i & 1], ((i - 2) >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t d = tmem_base + (t * 2 + (i & 1)) * 64;
            if (issuer) {
#pragma unroll
              for (int kk = 0; kk < kDh / 16; ++kk) {
                const uint32_t off = (kk >> 2) * kBlock16 + (kk & 3) * 2;
                umma_bf16_ss<1>(d, desc_from_words(q_lo + t * kTile16 + off, kHiK), desc_from_words(ka + off, kHiK), idesc_s, kk != 0 ? 1u : 0u);
              }
              umma_commit(&hdr->s_full[t][i & 1]);
            }
            __syncwarp();
          }
          if (issuer && ((i & 1) == 1 || i + 1 == n_steps)) umma_commit(&hdr->k_empty[lt & 1]);  // K tile lt fully contracted
          __syncwarp();
          if (lane == 0) TRACE(2, i, 3);
        }
      } else {
        const int t = warp - 2;  // query tile of this PV issuer
        const uint32_t idesc_o = make_idesc_bf16(128, kDh, 0, 1);  // B = V is MN-major (dh contiguous)
        const uint32_t v_lo = ((smem_u32(sV) & 0x3FFFFu) >> 4) | ((kBlockBytes >> 4) << 16);  // MN-major: LBO = next 64-column block
        const uint32_t t_o = tmem_base + kColO + t * kDh;
        for (int i = 0; i < n_steps; ++i) {
          const int lt = i >> 1;
          if ((i & 1) == 0) mbar_wait(&hdr->v_full[lt & 1], (lt >> 1) & 1);  // first step of a key tile: its V must have landed
          mbar_wait(&hdr->p_full[t][i & 1], (i >> 1) & 1);
          tc_fence_after_sync();
          // O_t += P_t(i) . V[64 key rows of step i]  (P read from TMEM buffer i&1)
          const uint32_t va = v_lo + (lt & 1) * kTile16 + (i & 1) * (4 * 2048 >> 4);
          const uint32_t p_tmem = tmem_base + (t * 2 + (i & 1)) * 64;
          if (issuer) {
#pragma unroll
            for (int kq = 0; kq < 4; ++kq)  // 16 keys per MMA = 8 packed TMEM columns of P; V: 16 key rows = 2048 B
              umma_bf16_ts(t_o, p_tmem + kq * 8, desc_from_words(va + kq * (2048 >> 4), kHiK), idesc_o, (i | kq) != 0 ? 1u : 0u);
            umma_commit(&hdr->pv_done[t][i & 1]);
            if ((i & 1) == 1 || i + 1 == n_steps) umma_commit(&hdr->v_empty[lt & 1]);  // this tile's share of V tile lt is done
          }
          __syncwarp();
          if (lane == 0) TRACE(2, i, t * 4);
        }
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row, 64 keys per step ==========
    reg_alloc<kRegsSoftmax>();
    const int t = (warp - 4) >> 2;      // query tile of this warpgroup
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_lane + kColO + t * kDh;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < n_steps; ++i) {
      const uint32_t t_s = t_lane + (t * 2 + (i & 1)) * 64;
      if (lane == 0 && quarter == 0) TRACE(t, i, 0);  // start waiting for S_t(i)
      mbar_wait(&hdr->s_full[t][i & 1], (i >> 1) & 1);
      tc_fence_after_sync();
      if (lane == 0 && quarter == 0) TRACE(t, i, 1);  // S_t(i) ready
      uint32_t sr[64];
#pragma unroll
      for (int c = 0; c < 64; c += 32) tmem_ld_x32(t_s + c, *reinterpret_cast<uint32_t(*)[32]>(&sr[c]));
      tmem_wait_ld();
      if (lane == 0 && quarter == 0) TRACE(t, i, 2);  // scores in registers

      const int kv0 = (s_lo + i) * 64;
      const int kv_valid = p.Tk - kv0;
      float sc = p.scale_log2;
      if (p.kv_bias != nullptr) {  // rare (context masks): fold scale and bias into the scores first
        const float* bias = p.kv_bias + static_cast<long long>(job.b) * p.Tk + kv0;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const float bv = (c < kv_valid) ? __ldg(bias + c) * kLog2e : 0.f;
          sr[c] = __float_as_uint(fmaf(__uint_as_float(sr[c]), sc, bv));
        }
        sc = 1.0f;
      }
      if (kv_valid < 64) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= kv_valid) sr[c] = __float_as_uint(kMasked);
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(sr[c + u]));
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;  // sc > 0
      const float m_new = (m_tile > m + kTau) ? fmaxf(ceilf(m_tile), -1048576.0f) : m;  // integer stabiliser, moved lazily
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_new == m) ? 1.0f : fast_exp2(m - m_use);  // m = -inf -> 0
      if (i > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {  // O is free once PV_t(i-1) is done
        mbar_wait(&hdr->pv_done[t][(i - 1) & 1], ((i - 1) >> 1) & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int c = 0; c < kDh; c += 32) {
          uint32_t o[32];
          tmem_ld_x32(t_o + c, o);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
          tmem_st_x32(t_o + c, o);
        }
      }
      if (lane == 0 && quarter == 0) TRACE(t, i, 3);  // maximum known
      // P = 2^(s*sc - m): kEmu of every 8 column pairs on the FMA pipe (degree-3 polynomial), the rest on MUFU
      constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
      const uint64_t sc2 = pack_f32x2(sc, sc);
      const uint64_t negm2 = pack_f32x2(-m_use, -m_use);
      const uint64_t c2 = pack_f32x2(kMagic - m_use, kMagic - m_use);  // exact: m_use is an integer below 2^22
      const uint64_t neg1 = pack_f32x2(-1.0f, -1.0f);
      const uint64_t k0 = pack_f32x2(0.99992807f, 0.99992807f), k1 = pack_f32x2(0.69326099f, 0.69326099f);
      const uint64_t k2 = pack_f32x2(0.24261114f, 0.24261114f), k3 = pack_f32x2(0.05517167f, 0.05517167f);
      uint64_t rs2[2] = {0ull, 0ull};
      uint32_t pk[32];
#pragma unroll
      for (int pi = 0; pi < 32; ++pi) {
        const int c = 2 * pi;
        const uint64_t s2 = pack_f32x2(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1]));
        float e0, e1;
        if ((pi & 7) < kEmu) {
          float x0, x1, p0, p1;
          unpack_f32x2(fma_f32x2(s2, sc2, c2), x0, x1);  // kMagic + round(x)
          x0 = fmaxf(x0, kMagic - 126.0f), x1 = fmaxf(x1, kMagic - 126.0f);
          const uint64_t f = fma_f32x2(s2, sc2, fma_f32x2(pack_f32x2(x0, x1), neg1, c2));  // x - round(x)
          unpack_f32x2(fma_f32x2(fma_f32x2(fma_f32x2(k3, f, k2), f, k1), f, k0), p0, p1);
          e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(x0) << 23));
          e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(x1) << 23));
        } else {
          float x0, x1;
          unpack_f32x2(fma_f32x2(s2, sc2, negm2), x0, x1);
          e0 = fast_exp2(x0), e1 = fast_exp2(x1);
        }
        rs2[pi & 1] = add_f32x2(rs2[pi & 1], pack_f32x2(e0, e1));
        pk[pi] = pack_bf16x2(e0, e1);
      }
      tmem_st_x32(t_s, pk);  // P_t(i): packed bf16 over the first 32 columns of this buffer
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->p_full[t][i & 1]);
      if (lane == 0 && quarter == 0) TRACE(t, i, 4);  // P signalled
      float r0, r1, r2, r3;
      unpack_f32x2(rs2[0], r0, r1);
      unpack_f32x2(rs2[1], r2, r3);
      l = l * alpha + ((r0 + r1) + (r2 + r3));
      m = m_new;
    }
    // ---- epilogue ----
    if (lane == 0 && quarter == 0) TRACE(3, 0, 3 + t);  // softmax of the last step done
    if (n_steps > 0) {
      mbar_wait(&hdr->pv_done[t][(n_steps - 1) & 1], ((n_steps - 1) >> 1) & 1);  // MMAs complete in issue order: all PV_t are done
      tc_fence_after_sync();
    }
    if (lane == 0 && quarter == 0 && t == 0) TRACE(3, 0, 5);  // last PV done
    const int row = job.q0 + t * 128 + r;
    if (job.slot < 0) {
      const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
      __nv_bfloat16* orow = attn_out_row(p, job.b, row, job.h, kDh);
#pragma unroll
      for (int c = 0; c < kDh; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(t_o + c, o);
        tmem_wait_ld();
        if (row < p.Tq) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * k + 0]) * inv_l, __uint_as_float(o[8 * k + 1]) * inv_l);
            w.y = pack_bf16x2(__uint_as_float(o[8 * k + 2]) * inv_l, __uint_as_float(o[8 * k + 3]) * inv_l);
            w.z = pack_bf16x2(__uint_as_float(o[8 * k + 4]) * inv_l, __uint_as_float(o[8 * k + 5]) * inv_l);
            w.w = pack_bf16x2(__uint_as_float(o[8 * k + 6]) * inv_l, __uint_as_float(o[8 * k + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + 8 * k) = w;
          }
        }
      }
    } else {
      const long long prow = static_cast<long long>(job.slot) * p.part_rows + t * 128 + r;
      float* wo = p.ws_o + prow * kDh;
      if (row < p.Tq) *reinterpret_cast<float2*>(p.ws_ml + prow * 2) = make_float2(m, l);
#pragma unroll
      for (int c = 0; c < kDh; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(t_o + c, o);
        tmem_wait_ld();
        if (row < p.Tq) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<uint4*>(wo + c + 4 * k) = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
        }
      }
      if (p.cmb_counters != nullptr) {
        // ---- merge inside the kernel (no combine launch): the key ranges of a split job meet at a counter, then each
        // merges ITS share of the job's 256 rows by log-sum-exp (attention_combine_row) and writes the bf16 output.  The
        // split CTAs are the last n_left * n_split <= SM count blocks of the grid, so they are all resident together.
        const int jl = job.slot / p.n_split, part = job.slot - jl * p.n_split;
        int* arrive = p.cmb_counters + jl;
        int* depart = p.cmb_counters + kCmbJobs + jl;
        asm volatile("bar.sync 2, 256;" ::: "memory");  // all eight softmax warps have parked their rows
        if (warp == 4 && lane == 0) {
          __threadfence();
          atomicAdd(arrive, 1);
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(arrive) < p.n_split) {
            if (clock64() - t0 > LTXB_WATCHDOG_CYCLES) {
              printf("ltxb: attention key-split watchdog: job %d part %d sees %d of %d arrivals\n", jl, part, *reinterpret_cast<volatile int*>(arrive), p.n_split);
              __trap();
            }
          }
          __threadfence();
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const int rpp = (256 + p.n_split - 1) / p.n_split;
        const int r_end = min(256, (part + 1) * rpp);
        for (int rr = part * rpp + (warp - 4); rr < r_end; rr += 8) attention_combine_row<kDh>(p, jl, rr, lane);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (warp == 4 && lane == 0 && atomicAdd(depart, 1) == p.n_split - 1) {  // the last range to leave re-arms the counters
          *arrive = 0;
          *depart = 0;
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(3, 0, 6);  // output stored by every warp
  if (threadIdx.x == 0) peer_sync_exit(p.sync, sync_epoch);
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// One warp per query row of a split job: O = sum_i 2^(m_i - M) O_i / sum_i 2^(m_i - M) l_i.  (jl: index of the job among the
// split ones, rr: row inside the job's partial block.)  Partials are read through L2 (__ldcg): inside the attention kernel
// they were written by other SMs moments ago.
template <int kDh>
__device__ __forceinline__ void attention_combine_row(const AttnParams& p, int jl, int rr, int lane) {
  constexpr int kPer = kDh / 32;  // columns per lane
  const int jb = p.n_full + jl;
  const int bh = jb / p.n_qp;
  const int row = (jb - bh * p.n_qp) * 256 + rr;
  if (row >= p.Tq) return;
  const int b = bh / p.H, h = bh - b * p.H;
  const long long prow0 = static_cast<long long>(jl) * p.cmb_job_stride + rr;
  float M = -INFINITY;
  for (int i = 0; i < p.n_split; ++i) M = fmaxf(M, __ldcg(p.ws_ml + i * p.cmb_part_ml + prow0 * 2));
  float acc[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) acc[u] = 0.f;
  float L = 0.f;
  for (int i = 0; i < p.n_split; ++i) {
    const float2 ml = __ldcg(reinterpret_cast<const float2*>(p.ws_ml + i * p.cmb_part_ml + prow0 * 2));
    const float w = (ml.x == -INFINITY) ? 0.f : fast_exp2(ml.x - M);
    L += w * ml.y;
    const float* src = p.ws_o + i * p.cmb_part_o + prow0 * kDh + lane * kPer;
    if constexpr (kPer == 4) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(src));
      acc[0] += w * v.x, acc[1] += w * v.y, acc[2] += w * v.z, acc[3] += w * v.w;
    } else {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(src));
      acc[0] += w * v.x, acc[1] += w * v.y;
    }
  }
  const float inv = (L > 0.f) ? 1.0f / L : 0.f;
  __nv_bfloat16* dst = attn_out_row(p, b, row, h, kDh) + lane * kPer;
  if constexpr (kPer == 4) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv));
  } else {
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(acc[0] * inv, acc[1] * inv);
  }
}

template <int kDh>
__global__ void __launch_bounds__(256) attention_combine_kernel(const AttnParams p, int n_left) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int jl = gw / p.part_rows, rr = gw - jl * p.part_rows;
  if (jl >= n_left) return;
  attention_combine_row<kDh>(p, jl, rr, lane);
}

// ---- split-KV workspace: registered by the host framework (the library never allocates) -----------
constexpr int kMaxDevices = 16;
constexpr int kMaxSlots = 160;  // <= SM count
struct AttnWorkspace {
  float* base = nullptr;
  long long bytes = 0;
};
static AttnWorkspace g_attn_ws[kMaxDevices];
constexpr long long kCmbCounterBytes = 2048;  // 2 x kCmbJobs ints in front of the partials, zeroed when the workspace is registered
static long long attn_ws_bytes_for(int slots, int dh) { return kCmbCounterBytes + static_cast<long long>(slots) * 256 * (dh + 2) * sizeof(float); }

template <int kDh>
static int launch_pair(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, AttnParams p,
                       cudaStream_t stream) {
  constexpr size_t kTileBytes = static_cast<size_t>(kDh / 64) * 128 * 128;
  constexpr size_t smem = 1024 + kPairHeader + 6 * kTileBytes;
  // LTXB_ATTN_S64=0: the single-buffered 128-key-step kernel (kept for A/B runs)
  static const bool s64 = [] { const char* e = getenv("LTXB_ATTN_S64"); return e == nullptr || atoi(e) != 0; }();
  auto kernel = s64 ? attention_pair64_kernel<kDh> : attention_pair_kernel<kDh>;
  if (p.sync.n_peers > 0 && !s64) {  // only the 64-key kernel has the folded barrier
    if (int rc = launch_peer_barrier(p.sync, stream)) return rc;
    p.sync = PeerSync{};
  }
  static PerDeviceOnce configured;  // per instantiation and device
  if (configured.first()) {
    LTXB_CUDA(cudaFuncSetAttribute(attention_pair64_kernel<kDh>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    LTXB_CUDA(cudaFuncSetAttribute(attention_pair_kernel<kDh>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const int sms = num_sms();
  LTXB_CHECK_SUPPORTED(sms > 0, "ltxb_attention_fwd: no device");
  const int n_kv = (p.Tk + 127) / 128;
  p.n_qp = (p.Tq + 255) / 256;
  const long long jobs = static_cast<long long>(p.B) * p.H * p.n_qp;
  LTXB_CHECK_SUPPORTED(jobs < (1ll << 30), "ltxb_attention_fwd: too many (batch, head, query tile) jobs");
  // ragged last wave -> split its jobs over the keys so that one CTA per SM finishes it in 1/n_split of a wave
  static const bool split_on = [] { const char* e = getenv("LTXB_ATTN_SPLIT"); return e == nullptr || atoi(e) != 0; }();
  int dev = 0;
  LTXB_CUDA(cudaGetDevice(&dev));
  const int n_left = static_cast<int>(jobs % sms);
  int n_split = 1;
  if (split_on && n_left > 0 && dev < kMaxDevices && g_attn_ws[dev].base != nullptr) {
    n_split = std::min(n_kv, sms / n_left);
    if (n_split * n_left > kMaxSlots || attn_ws_bytes_for(n_split * n_left, kDh) > g_attn_ws[dev].bytes) n_split = 1;
  }
  if (n_split > 1) {
    p.n_full = static_cast<int>(jobs - n_left);
    p.n_split = n_split;
    p.ws_o = g_attn_ws[dev].base + kCmbCounterBytes / sizeof(float);
    p.ws_ml = p.ws_o + static_cast<long long>(n_split) * n_left * 256 * kDh;
    p.part_rows = 256;
    p.cmb_job_stride = static_cast<long long>(n_split) * 256;
    p.cmb_part_o = 256ll * kDh;
    p.cmb_part_ml = 512;
    // LTXB_ATTN_FUSED_COMBINE=1: the 64-key kernel merges inside (every key range its share of the rows, after meeting the
    // others at a counter) instead of leaving it to attention_combine_kernel.  Measured slower — 50.0 against 42.6 us at
    // 1280 x 1280, 33.5 against 33.0 ms per step: the ranges wait for the slowest one and merge with eight warps each, where
    // the second launch (its prologue already overlapped by PDL) uses the whole GPU — so the separate launch is the default.
    static const bool fused = [] { const char* e = getenv("LTXB_ATTN_FUSED_COMBINE"); return e != nullptr && atoi(e) != 0; }();
    if (fused && s64 && n_left <= kCmbJobs) p.cmb_counters = reinterpret_cast<int*>(g_attn_ws[dev].base);
  } else {
    p.n_full = static_cast<int>(jobs);
    p.n_split = 1;
#ifdef LTXB_ATTN_TRACE
    if (dev < kMaxDevices) p.ws_o = g_attn_ws[dev].base;
#endif
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
  const int grid = p.n_full + (n_split > 1 ? n_left * n_split : 0);
  LTXB_CUDA(launch_kernel(kernel, dim3(grid), dim3(kPairThreads), smem, stream, 1, tq, tk, tv, p));
  if (n_split > 1 && p.cmb_counters == nullptr)
    LTXB_CUDA(launch_kernel(attention_combine_kernel<kDh>, dim3(n_left * 256 / 8), dim3(256), 0, stream, 1, p, n_left));
  return LTXB_OK;
}

template <int kDh>
static int launch_partial(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, AttnParams p,
                          float* part_o, float* part_ml, cudaStream_t stream) {
  constexpr size_t kTileBytes = static_cast<size_t>(kDh / 64) * 128 * 128;
  constexpr size_t smem = 1024 + kPairHeader + 6 * kTileBytes;
  auto kernel = attention_pair64_kernel<kDh>;
  static PerDeviceOnce configured;
  if (configured.first()) LTXB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  p.n_qp = (p.Tq + 255) / 256;
  const long long jobs = static_cast<long long>(p.B) * p.H * p.n_qp;
  LTXB_CHECK_SUPPORTED(jobs < (1ll << 30), "ltxb_attention_partial: too many (batch, head, query tile) jobs");
  p.n_full = 0;  // every CTA is a "piece" (slot = job) covering the whole key range: unnormalised output
  p.n_split = 1;
  p.ws_o = part_o;
  p.ws_ml = part_ml;
  p.part_rows = p.n_qp == 1 ? p.Tq : 256;  // compact when a job holds fewer than 256 query rows
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
  LTXB_CUDA(launch_kernel(kernel, dim3(static_cast<unsigned>(jobs)), dim3(kPairThreads), smem, stream, 1, tq, tk, tv, p));
  return LTXB_OK;
}

int launch_attention_partial(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, AttnParams p,
                             int dh, float* part_o, float* part_ml, cudaStream_t stream) {
  if (dh == 128) return launch_partial<128>(Q, ldq, K, ldk, V, ldv, p, part_o, part_ml, stream);
  return launch_partial<64>(Q, ldq, K, ldk, V, ldv, p, part_o, part_ml, stream);
}

int launch_attention_merge(AttnParams p, int dh, const float* parts, long long part_stride, int n_parts, cudaStream_t stream) {
  p.n_qp = (p.Tq + 255) / 256;
  const long long jobs = static_cast<long long>(p.B) * p.H * p.n_qp;
  p.n_full = 0;
  p.n_split = n_parts;
  p.part_rows = p.n_qp == 1 ? p.Tq : 256;
  const long long rows = jobs * p.part_rows;
  // part i: [rows][dh] floats at parts + i*part_stride, then [rows][2]
  LTXB_CHECK_ARG(part_stride >= rows * (dh + 2) && part_stride % 4 == 0, "ltxb_attention_merge: part stride %lld too small / misaligned for %lld rows",
                 part_stride, rows);
  p.ws_o = const_cast<float*>(parts);
  p.ws_ml = const_cast<float*>(parts) + rows * dh;
  p.cmb_job_stride = p.part_rows;
  p.cmb_part_o = part_stride;
  p.cmb_part_ml = part_stride;
  const int warps = static_cast<int>(jobs * p.part_rows);
  if (dh == 128) LTXB_CUDA(launch_kernel(attention_combine_kernel<128>, dim3((warps + 7) / 8), dim3(256), 0, stream, 1, p, static_cast<int>(jobs)));
  else LTXB_CUDA(launch_kernel(attention_combine_kernel<64>, dim3((warps + 7) / 8), dim3(256), 0, stream, 1, p, static_cast<int>(jobs)));
  return LTXB_OK;
}

int launch_attention_pair(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                          AttnParams p, int dh, cudaStream_t stream) {
  if (dh == 128) return launch_pair<128>(Q, ldq, K, ldk, V, ldv, p, stream);
  return launch_pair<64>(Q, ldq, K, ldk, V, ldv, p, stream);
}

}  // namespace ltxb

using namespace ltxb;

extern "C" int64_t ltxb_attention_workspace_bytes(void) { return attn_ws_bytes_for(kMaxSlots, 128); }

extern "C" int ltxb_attention_set_workspace(void* workspace, int64_t bytes) {
  int dev = 0;
  LTXB_CUDA(cudaGetDevice(&dev));
  LTXB_CHECK_SUPPORTED(dev >= 0 && dev < kMaxDevices, "ltxb_attention_set_workspace: device %d out of range", dev);
  if (workspace == nullptr) {
    g_attn_ws[dev] = AttnWorkspace{};
    return LTXB_OK;
  }
  LTXB_CHECK_ARG(aligned16(workspace) && bytes > 0, "ltxb_attention_set_workspace: need a 16-byte aligned, non-empty buffer");
  LTXB_CHECK_ARG(bytes > kCmbCounterBytes, "ltxb_attention_set_workspace: buffer smaller than the counter block");
  LTXB_CUDA(cudaMemset(workspace, 0, kCmbCounterBytes));  // arrival / departure counters of the in-kernel key-split merge
  g_attn_ws[dev].base = reinterpret_cast<float*>(workspace);
  g_attn_ws[dev].bytes = bytes;
  return LTXB_OK;
}
