// PROPOSED, NOT YET RUN ON HARDWARE (written at the end of round 1 with the GPU budget spent; compiles for sm_100a).
// Kept under profiles/ — not in the product build — until it has passed tests/test_gpu_kernels.py::test_attention.
// To try it: copy to mlx-video_b200/csrc/, add it to SRCS in the Makefile, declare launch_attention_pair64 in
// attention.cuh and call it from launch_attention_pair (attention_pair.cu) when an env switch is set.
//
// K2 variant: two 128-row query tiles per CTA as in attention_pair.cu, but the scores are produced in 64-key steps
// into DOUBLE-BUFFERED TMEM tiles, so that S_t(s+1) is issued a whole step ahead of its softmax.
//
// Why (profiles/r1e/attn_ncu_5184.md): in attention_pair.cu S_t(j+1) overwrites P_t(j) in TMEM, so it can only be issued
// after PV_t(j); 31.5 % of all warp samples are softmax warps waiting for that S tile (~1.65 k cycles of P-ready -> PV ->
// S turnaround per key tile) while no pipe is saturated (tensor 52 %, MUFU 40 %).  Here the chain per query tile is
// softmax(s) -> P(s) -> PV(s) only; S(s+1) is already in the other buffer when softmax(s) ends.
//
// TMEM (512 columns, all used): S buffer (tile t, buffer b) at (t*2 + b)*64 (64 fp32 columns = 64 keys; P(s) is written
// back over its first 32 columns as packed bf16), O_t at 256 + t*dh.
// K / V stay 128-key TMA tiles in 2-deep smem rings; step s uses key rows [64*(s&1), +64) of tile s>>1.
//   warp 0      TMA producer (Q0, Q1 once; K, V tiles)
//   warp 1      MMA issuer.  Per step s:  S_0(s+1), PV_0(s), S_1(s+1), PV_1(s)
//   warps 2,3   idle (keep the softmax warps aligned to their TMEM lane quarters)
//   warps 4..7  softmax of tile 0, warps 8..11 of tile 1: one thread per query row, 64 scores per step in registers
// Whole jobs only (no key split of the ragged last wave yet): grid = B * H * ceil(Tq / 256).
#include <algorithm>
#include <cstdlib>

#include "attention.cuh"
#include "ptx.cuh"

namespace ltxb {
namespace {

constexpr int kThreads64 = 384;
constexpr int kHeader64 = 1024;
constexpr float kTau64 = 8.0f;        // log2 units: the running maximum moves only when exceeded by more than 2^kTau
constexpr float kMasked64 = -1.0e9f;  // score of an out-of-range key
constexpr int kEmu64 = 2;             // of every 8 column pairs, how many take the FMA-pipe exp2
constexpr int kRegsIssue64 = 104, kRegsSoftmax64 = 200;  // 128 * 104 + 256 * 200 <= 64 K registers

struct Pair64Header {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2];
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full[2][2];  // S_t(s) complete in buffer s&1          (MMA -> softmax t)
  uint64_t p_full[2][2];  // P_t(s) written over buffer s&1         (softmax t -> MMA); one arrival per warp
  uint64_t pv_done[2];    // O_t += P_t(s) V(s) complete, per step  (MMA -> softmax t)
  uint32_t tmem_base;
};
static_assert(sizeof(Pair64Header) <= kHeader64, "header overflow");

template <int kDh>
__global__ void __launch_bounds__(kThreads64, 1)
attention_pair64_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  constexpr int kBlocks = kDh / 64;                       // 64-column (128 B) swizzle blocks per row
  constexpr uint32_t kBlockBytes = 128 * 128;             // 128 rows x 128 B
  constexpr uint32_t kTileBytes = kBlocks * kBlockBytes;  // one Q / K / V tile
  constexpr uint32_t kColO = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Pair64Header* hdr = reinterpret_cast<Pair64Header*>(smem);
  uint8_t* sQ = smem + kHeader64;      // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;   // 2 stages
  uint8_t* sV = sK + 2 * kTileBytes;   // 2 stages

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // job = (batch, head, pair of query tiles)
  const int job = blockIdx.x;
  const int bh = job / p.n_qp;
  const int q0 = (job - bh * p.n_qp) * 256;
  const int b = bh / p.H, h = bh - b * p.H;
  const int n_kt = (p.Tk + 127) / 128;    // 128-key K / V tiles
  const int n_steps = (p.Tk + 63) / 64;   // 64-key score steps (the last tile may hold a single one)

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
        mbar_init(&hdr->pv_done[i], 1);
        for (int j = 0; j < 2; ++j) {
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
  pdl_launch_dependents();
  pdl_wait();

  if (warp < 4) {
    reg_dealloc<kRegsIssue64>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(&hdr->q_full, 2 * kTileBytes);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            tma_load_3d(sQ + t * kTileBytes + j * kBlockBytes, &tmap_q, &hdr->q_full, h * kDh + 64 * j, q0 + 128 * t, b);
        for (int kt = 0; kt < n_kt; ++kt) {
          const int st = kt & 1;
          const uint32_t ph = (kt >> 1) & 1;
          mbar_wait(&hdr->k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->k_full[st], kTileBytes);
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            tma_load_3d(sK + st * kTileBytes + j * kBlockBytes, &tmap_k, &hdr->k_full[st], h * kDh + 64 * j, kt * 128, b);
          mbar_wait(&hdr->v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&hdr->v_full[st], kTileBytes);
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            tma_load_3d(sV + st * kTileBytes + j * kBlockBytes, &tmap_v, &hdr->v_full[st], h * kDh + 64 * j, kt * 128, b);
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t idesc_o = make_idesc_bf16(128, kDh, 0, 1);  // B = V is MN-major (dh contiguous)
        // S_t(s) = Q_t . K[64 key rows of step s]^T into buffer s&1 (the K stage of tile s>>1 must be full)
        auto issue_s = [&](int t, int s) {
          const uint32_t q_addr = smem_u32(sQ + t * kTileBytes);
          const uint32_t k_addr = smem_u32(sK + ((s >> 1) & 1) * kTileBytes) + (s & 1) * (64 * 128);  // rows 64.. of every block
          const uint32_t d = tmem_base + (t * 2 + (s & 1)) * 64;
#pragma unroll
          for (int kk = 0; kk < kDh / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kBlockBytes + (kk & 3) * 32;
            umma_bf16_ss<1>(d, make_smem_desc_sw128(q_addr + off, 16, 1024), make_smem_desc_sw128(k_addr + off, 16, 1024), idesc_s,
                            kk != 0 ? 1u : 0u);
          }
          umma_commit(&hdr->s_full[t][s & 1]);
        };
        // O_t += P_t(s) . V[64 key rows of step s]  (P read from TMEM buffer s&1; the V stage of tile s>>1 must be full)
        auto issue_pv = [&](int t, int s) {
          mbar_wait(&hdr->p_full[t][s & 1], (s >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t v_addr = smem_u32(sV + ((s >> 1) & 1) * kTileBytes);
          const uint32_t p_tmem = tmem_base + (t * 2 + (s & 1)) * 64;
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) {  // 16 keys per MMA = 8 packed TMEM columns of P; V: 16 key rows = 2048 B
            const int kk = 4 * (s & 1) + kq;
            umma_bf16_ts(tmem_base + kColO + t * kDh, p_tmem + kq * 8, make_smem_desc_sw128(v_addr + kk * 2048, kBlockBytes, 1024),
                         idesc_o, (s | kq) != 0 ? 1u : 0u);
          }
          umma_commit(&hdr->pv_done[t]);
        };
        mbar_wait(&hdr->q_full, 0);
        mbar_wait(&hdr->k_full[0], 0);
        tc_fence_after_sync();
        issue_s(0, 0);
        issue_s(1, 0);
        for (int s = 0; s < n_steps; ++s) {
          const int kt = s >> 1;
          const bool more = s + 1 < n_steps;
          if ((s & 1) == 0) {  // first step of a key tile: its V must have landed before the first PV
            mbar_wait(&hdr->v_full[kt & 1], (kt >> 1) & 1);
          } else if (more) {   // step s+1 opens the next key tile
            mbar_wait(&hdr->k_full[(kt + 1) & 1], ((kt + 1) >> 1) & 1);
          }
          tc_fence_after_sync();
          // S_t(s+1) goes into the buffer P_t(s-1) lived in: PV_t(s-1) was issued one iteration ago and the tensor
          // pipe runs MMAs in issue order, so no wait is needed — the scores are a step ahead of their softmax
          if (more) issue_s(0, s + 1);
          issue_pv(0, s);
          if (more) issue_s(1, s + 1);
          issue_pv(1, s);
          if (more && (s & 1) == 0) umma_commit(&hdr->k_empty[kt & 1]);   // both halves of K tile kt have been contracted
          if (!more || (s & 1) == 1) umma_commit(&hdr->v_empty[kt & 1]);  // V tile kt is done
        }
      }
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row, 64 keys per step ==========
    reg_alloc<kRegsSoftmax64>();
    const int t = (warp - 4) >> 2;      // query tile of this warpgroup
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t t_o = t_lane + kColO + t * kDh;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    for (int s = 0; s < n_steps; ++s) {
      const uint32_t t_s = t_lane + (t * 2 + (s & 1)) * 64;
      mbar_wait(&hdr->s_full[t][s & 1], (s >> 1) & 1);
      tc_fence_after_sync();
      uint32_t sr[64];
#pragma unroll
      for (int c = 0; c < 64; c += 32) tmem_ld_x32(t_s + c, *reinterpret_cast<uint32_t(*)[32]>(&sr[c]));
      tmem_wait_ld();

      const int kv0 = s * 64;
      const int kv_valid = p.Tk - kv0;
      float sc = p.scale_log2;
      if (p.kv_bias != nullptr) {  // rare (context masks): fold scale and bias into the scores first
        const float* bias = p.kv_bias + static_cast<long long>(b) * p.Tk + kv0;
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
          if (c >= kv_valid) sr[c] = __float_as_uint(kMasked64);
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 64; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(sr[c + u]));
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sc;  // sc > 0
      const float m_new = (m_tile > m + kTau64) ? fmaxf(ceilf(m_tile), -1048576.0f) : m;  // integer stabiliser, moved lazily
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_new == m) ? 1.0f : fast_exp2(m - m_use);  // m = -inf -> 0
      if (s > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {  // O is free once PV_t(s-1) is done
        mbar_wait(&hdr->pv_done[t], (s - 1) & 1);
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
      // P = 2^(s*sc - m): kEmu of every 8 column pairs on the FMA pipe (degree-3 polynomial), the rest on MUFU
      constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
      const uint64_t sc2 = pack_f32x2(sc, sc);
      const uint64_t negm2 = pack_f32x2(-m_use, -m_use);
      const uint64_t c2 = pack_f32x2(kMagic - m_use, kMagic - m_use);
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
        if ((pi & 7) < kEmu64) {
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
      tmem_st_x32(t_s, pk);  // P_t(s): packed bf16 over the first 32 columns of this buffer
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hdr->p_full[t][s & 1]);
      float r0, r1, r2, r3;
      unpack_f32x2(rs2[0], r0, r1);
      unpack_f32x2(rs2[1], r2, r3);
      l = l * alpha + ((r0 + r1) + (r2 + r3));
      m = m_new;
    }
    // ---- epilogue ----
    if (n_steps > 0) {
      mbar_wait(&hdr->pv_done[t], (n_steps - 1) & 1);
      tc_fence_after_sync();
    }
    const int row = q0 + t * 128 + r;
    const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
    __nv_bfloat16* orow = attn_out_row(p, b, row, h, kDh);
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
int launch_pair64(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, AttnParams p,
                  cudaStream_t stream) {
  constexpr size_t kTileBytes = static_cast<size_t>(kDh / 64) * 128 * 128;
  constexpr size_t smem = 1024 + kHeader64 + 6 * kTileBytes;
  auto kernel = attention_pair64_kernel<kDh>;
  static bool configured = false;
  if (!configured) {
    LTXB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  p.n_qp = (p.Tq + 255) / 256;
  const long long jobs = static_cast<long long>(p.B) * p.H * p.n_qp;
  LTXB_CHECK_SUPPORTED(jobs < (1ll << 30), "ltxb_attention_fwd: too many (batch, head, query tile) jobs");
  p.n_full = static_cast<int>(jobs);
  p.n_split = 1;
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
  LTXB_CUDA(launch_kernel(kernel, dim3(static_cast<unsigned>(jobs)), dim3(kThreads64), smem, stream, 1, tq, tk, tv, p));
  return LTXB_OK;
}

}  // namespace

int launch_attention_pair64(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                            AttnParams p, int dh, cudaStream_t stream) {
  if (dh == 128) return launch_pair64<128>(Q, ldq, K, ldk, V, ldv, p, stream);
  return launch_pair64<64>(Q, ldq, K, ldk, V, ldv, p, stream);
}

}  // namespace ltxb
