// K1 for FEW rows: out = epilogue(A[M,K] . W[N,K]^T) with M <= 512 — the weight-streaming regime.
//
// Where it runs: the sequence-parallel shards of the LTX-2 DiT (1280 tokens on 8 GPUs = 160 rows per rank against the
// full 4096 / 16384-wide weights), the audio stream (68 tokens per sample) and the AdaLN tables (one row per sample):
// the same nn.Linear call sites as gemm.cu (attention.py:91-93,100,123-126,142; feed_forward.py:31,33; adaln.py:27,130-132).
// A tile of the big kernel is 256 token rows x 256 weight rows: at M = 160 it spends the tensor time of 256 rows, holds only
// 64 KB of weights in flight per SM and pays an 8 us prologue / fix-up / teardown (profiles/r2/gemm_small_m.md).
//
// Design (B200 / sm_100a) — the operands are SWAPPED:
//   * the WEIGHT rows are the MMA's M dimension (TMEM lanes): one CTA pair (cta_group::2) owns 256 weight rows, 128 per
//     CTA; the TOKENS are the MMA's N dimension (16 .. 256 columns, one or two MMAs per k-slice), so the tensor work and
//     the TMEM footprint scale with the token count and nothing is padded to 128 rows;
//   * each CTA of the pair stages its own 128 weight rows (16 KB per k-block, from HBM) and HALF of the token rows
//     (from L2) — the pair MMA reads the B operand from both CTAs' shared memory, which halves the L2 -> SM traffic of
//     re-reading the activations per weight tile;
//   * one (weight tile, k-range) per CTA pair, no persistence: grid = tiles x splits pairs, all co-resident; one CTA per SM
//     with all of its shared memory as stages (8 x 26 KB at 160 tokens) when that still fills the machine, else two per SM
//     (the accumulators of both fit 2 x 256 TMEM columns) for twice the pair slots, i.e. more k-range splits;
//   * split-K is a reduce-scatter through L2: every pair parks the 8-token chunks it does not own, all pairs of a tile
//     meet at a counter, and each adds the others' parts to the chunks it owns, in split order (run-to-run
//     reproducible), then applies the fused epilogue;
//   * in the epilogue (eight warps) a thread owns one OUTPUT COLUMN (its TMEM lane) and walks the tokens; the finished
//     8-token chunks are staged in the operand stages (idle by then) and handed to the TMA: a plain tensor store, or for
//     the residual epilogue an element-wise ADD into the residual stream (out += (acc + bias) * gate, one add per element);
//   * weights the host marks constant (LTXB_GEMM_CONST_W) are requested before the programmatic-launch wait, so the first
//     stages fill under the stream predecessor's tail; the MMA issuer is a warp-uniform loop with one elected lane.
// gemm_small_m_packed_kernel (further down) is the same product over MLX affine-quantised weights kept packed in HBM.
#include "common.cuh"
#include "gemm_small_m.cuh"
#include "peer_sync.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <type_traits>
#include <cstdlib>

namespace ltxb {

constexpr int kWsTileRows = 128;  // weight rows (= output columns) per CTA: the TMEM lanes
constexpr int kWsBlockK = 64;     // 64 bf16 = one 128-byte swizzle span
constexpr int kWsThreads = 320;   // warp 0 TMA producer, warp 1 MMA issuer + TMEM, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kWsMaxStages = 8;
constexpr int kWsChunk = 8;  // tokens per reduce-scatter chunk
constexpr int kWsHeader = 1024;
constexpr uint32_t kWsWBytes = kWsTileRows * kWsBlockK * 2;
constexpr int kWsDepartOffset = 1 << 15;  // counters: arrivals at [0, 32768), departures behind them

struct WsParams {
  int M, N, K;
  int m_pad;      // token columns of the accumulator (multiple of 16; of 32 when n_mma == 2)
  int n_mma;      // MMAs per k-slice: 1 (m_pad <= 256) or 2
  int tmem_cols;  // power of two >= m_pad
  int num_stages;
  int splits;  // k-range pieces per weight tile
  const float* bias;
  void* out;
  long long ldo;
  const float* resid;
  long long ldr;
  const float* gate;
  long long gate_ld;
  int gate_row_div;
  const int* gate_row_index;
  const float* gate_table;
  int a_group_cols;
  // packed weights (kWBits 4 / 8): MLX affine groups, out = bf16(scales * q + biases) (ltxb_dequant_affine_bf16's arithmetic)
  const void* scales;
  const void* biases;
  long long lds;
  int group;    // columns per (scale, bias) pair: 32, 64 or 128
  int aux_f32;  // scales / biases are f32 (else bf16)
  uint32_t magic;  // 0x4B000000 (bits of 2^23): see the expanding warps
  int const_w;      // LTXB_GEMM_CONST_W: the first stages' weight tiles are requested before the PDL wait
  int out_tma;      // 0: per-thread global stores; 1: output chunks staged in shared memory, written by TMA; 2: staged, TMA add
                    // into `out` (RESID_GATE with out == resid: out += (acc + bias) * g, one add per element)
  float* partials;  // [grid][m_pad / 8][2][128] float4: parked chunks
  int* counters;
  PeerSync sync;    // cross-GPU flag barrier before the first read of the token rows (n_peers 0: none)
  long long* trace;  // LTXB_WS_DEBUG builds, debug & 8: [grid][4] globaltimer at PDL wait / accumulator / met / done (tail of the workspace)
  int debug;  // LTXB_WS_DEBUG builds only: 1 = no token loads after the first k-block, 2 = no epilogue, 4 = no weight loads
              // after the first k-block, 8 = every CTA records its life-cycle times (timing experiments; results are wrong)
};
#ifdef LTXB_WS_DEBUG
__device__ __forceinline__ long long ws_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define WS_DBG(bit) ((p.debug & (bit)) != 0)
// role profiling (debug & 64): cycles a role's lane 0 spent in a barrier wait, accumulated per role, printed by CTA 0
#define WS_TIMED_WAIT(acc, call)            \
  do {                                      \
    const long long t0__ = clock64();       \
    call;                                   \
    (acc) += clock64() - t0__;              \
  } while (0)
#else
#define WS_TIMED_WAIT(acc, call) call
#define WS_DBG(bit) false
#endif

struct WsSmemHeader {
  uint64_t full[kWsMaxStages];
  uint64_t empty[kWsMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};
static_assert(sizeof(WsSmemHeader) <= kWsHeader, "header overflow");

__device__ __forceinline__ void ws_epilogue_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Fused epilogue for the 8 tokens [m0, m0 + 8) of output column n (one thread): the warp's 32 lanes hold 32 consecutive
// columns, so every load / store below is one contiguous 64- or 128-byte piece per token.
template <int kEpi>
__device__ __forceinline__ void ws_finish_chunk(const WsParams& p, const float* acc, int m0, long long n, float bias_n,
                                                float table_n) {
  if (WS_DBG(16)) {  // timing experiment: no output stores (the values still have to be produced)
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) t += acc[i];
    if (t == 123.456f) reinterpret_cast<float*>(p.out)[0] = t;
    return;
  }
  float v[kWsChunk];
#pragma unroll
  for (int i = 0; i < kWsChunk; ++i) v[i] = acc[i] + bias_n;
  if constexpr (kEpi == LTXB_EPI_GELU_BF16) {
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) v[i] = gelu_tanh(v[i]);
  } else if constexpr (kEpi == LTXB_EPI_SILU_BF16) {
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) v[i] = silu(v[i]);
  }
  if constexpr (kEpi == LTXB_EPI_BIAS_BF16 || kEpi == LTXB_EPI_GELU_BF16 || kEpi == LTXB_EPI_SILU_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + n;
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i)
      if (m0 + i < p.M) o[static_cast<long long>(m0 + i) * p.ldo] = __float2bfloat16_rn(v[i]);
  } else if constexpr (kEpi == LTXB_EPI_BIAS_F32) {
    float* o = reinterpret_cast<float*>(p.out) + n;
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i)
      if (m0 + i < p.M) o[static_cast<long long>(m0 + i) * p.ldo] = v[i];
  } else {  // LTXB_EPI_RESID_GATE_F32; `out` normally IS `resid`: read the whole chunk before the first store
    float r[kWsChunk], g[kWsChunk];
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) {
      const int m = m0 + i;
      r[i] = 0.f, g[i] = 1.f;
      if (m < p.M) {
        r[i] = p.resid[static_cast<long long>(m) * p.ldr + n];
        if (p.gate != nullptr) {
          const long long grow = p.gate_row_index != nullptr ? __ldg(p.gate_row_index + m) : m / p.gate_row_div;
          g[i] = __ldg(p.gate + grow * p.gate_ld + n) + table_n;
        }
      }
    }
    float* o = reinterpret_cast<float*>(p.out) + n;
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i)
      if (m0 + i < p.M) o[static_cast<long long>(m0 + i) * p.ldo] = fmaf(v[i], g[i], r[i]);
  }
}

// The same epilogue into a shared-memory chunk tile [8 tokens][128 columns] (row-major, what the TMA store reads):
// `dst` = this thread's column in the tile.  RESID_GATE stages the gated update only, the TMA adds it into the residual.
template <int kEpi>
__device__ __forceinline__ void ws_stage_chunk(const WsParams& p, const float* acc, int m0, long long n, bool n_ok, float bias_n,
                                               float table_n, uint8_t* dst) {
  float v[kWsChunk];
#pragma unroll
  for (int i = 0; i < kWsChunk; ++i) v[i] = acc[i] + bias_n;
  if constexpr (kEpi == LTXB_EPI_GELU_BF16) {
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) v[i] = gelu_tanh(v[i]);
  } else if constexpr (kEpi == LTXB_EPI_SILU_BF16) {
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) v[i] = silu(v[i]);
  }
  if constexpr (kEpi == LTXB_EPI_BIAS_BF16 || kEpi == LTXB_EPI_GELU_BF16 || kEpi == LTXB_EPI_SILU_BF16) {
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) *reinterpret_cast<__nv_bfloat16*>(dst + i * (kWsTileRows * 2)) = __float2bfloat16_rn(v[i]);
  } else {
    if constexpr (kEpi == LTXB_EPI_RESID_GATE_F32) {
      if (p.gate != nullptr) {
        float g[kWsChunk];
#pragma unroll
        for (int i = 0; i < kWsChunk; ++i) {
          const int m = m0 + i;
          g[i] = 0.f;
          if (m < p.M && n_ok) {
            const long long grow = p.gate_row_index != nullptr ? __ldg(p.gate_row_index + m) : m / p.gate_row_div;
            g[i] = __ldg(p.gate + grow * p.gate_ld + n) + table_n;
          }
        }
#pragma unroll
        for (int i = 0; i < kWsChunk; ++i) v[i] *= g[i];
      }
    }
#pragma unroll
    for (int i = 0; i < kWsChunk; ++i) *reinterpret_cast<float*>(dst + i * (kWsTileRows * 4)) = v[i];
  }
}

// The epilogue shared by the bf16 and the packed-weight kernels (warps 2..9 of a CTA): accumulator [128 weight rows (lanes)]
// [m_pad tokens (columns)] at `tmem_acc` -> reduce-scatter over the k-range splits -> fused epilogue -> global memory.
// `stage_area`: the operand stages, idle once the accumulator is complete, reused for the TMA output tiles.
template <int kEpi, int kOthers>  // kOthers: parts of how many other splits a reduce pass requests together (registers!)
__device__ __forceinline__ void ws_epilogue(const WsParams& p, const CUtensorMap* tmap_out, uint64_t* tmem_full, uint8_t* stage_area,
                                            uint32_t stage_area_bytes, uint32_t tmem_acc, int tile, int split, uint32_t cta_rank,
                                            int warp, int lane, long long* dbg) {
  // ===================== epilogue: thread = output column, walks the tokens =====================
  // Eight warps, two per TMEM lane quarter: the epilogue is a long run of per-token address / convert / store
  // instructions per thread, so it is paced by how many warps the four schedulers can interleave.
  const int quarter = warp & 3;          // TMEM lanes [32 * quarter, +32) belong to this warp
  const int half = (warp - 2) >> 2;      // which half of the token chunks this warp walks
  const int lane_row = quarter * 32 + lane;
  const long long n = static_cast<long long>(tile) * (2 * kWsTileRows) + cta_rank * kWsTileRows + lane_row;
  const bool n_ok = n < p.N;
  const float bias_n = (p.bias != nullptr && n_ok) ? __ldg(p.bias + n) : 0.f;
  const float table_n = (p.gate_table != nullptr && n_ok) ? __ldg(p.gate_table + n) : 0.f;
  const bool epi_leader = (warp == 2 && lane == 0);
  const int chunks = p.m_pad / kWsChunk;
  const uint32_t t_row = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16);
  // ---- output staging (p.out_tma): the four warps of a half fill chunk tiles in the (now idle) operand stages, one
  // thread hands them to the TMA.  A round = as many consecutive chunks as the half's share of the stages holds.
  constexpr int kEsize = (kEpi == LTXB_EPI_BIAS_F32 || kEpi == LTXB_EPI_RESID_GATE_F32) ? 4 : 2;
  constexpr int kChunkBytes = kWsChunk * kWsTileRows * kEsize;
  const uint32_t half_bytes = (stage_area_bytes / 2) & ~1023u;
  uint8_t* half_stage = stage_area + half * half_bytes;
  const int stage_cap = static_cast<int>(half_bytes / kChunkBytes) & ~1;
  const bool staged = p.out_tma != 0;
  const bool half_issuer = (((warp - 2) & 3) == 0) && lane == 0;
  const int n0_cta = tile * (2 * kWsTileRows) + static_cast<int>(cta_rank) * kWsTileRows;
  int staged_k = 0, staged_c0 = 0;
  auto half_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory"); };
  auto flush = [&](bool last) {  // warp-uniform and identical in the four warps of the half
    if (staged_k == 0) return;
    fence_proxy_async_smem();
    half_bar();
    if (half_issuer) {
      for (int i = 0; i < staged_k; ++i) {
        if (p.out_tma == 2) tma_reduce_add_2d(tmap_out, half_stage + i * kChunkBytes, n0_cta, (staged_c0 + i) * kWsChunk);
        else tma_store_2d(tmap_out, half_stage + i * kChunkBytes, n0_cta, (staged_c0 + i) * kWsChunk);
      }
      tma_store_commit();
      if (!last) tma_store_wait_read();
    }
    if (!last) half_bar();
    staged_k = 0;
  };
  auto emit = [&](const float* vals, int c) {  // chunk c: consecutive within a round
    if (!staged) {
      if (n_ok) ws_finish_chunk<kEpi>(p, vals, c * kWsChunk, n, bias_n, table_n);
      return;
    }
    if (staged_k == 0) staged_c0 = c;
    ws_stage_chunk<kEpi>(p, vals, c * kWsChunk, n, n_ok, bias_n, table_n, half_stage + staged_k * kChunkBytes + lane_row * kEsize);
    if (++staged_k == stage_cap) flush(false);
  };
  mbar_wait(tmem_full, 0);
  tc_fence_after_sync();
#ifdef LTXB_WS_DEBUG
  dbg[0] = ws_globaltimer();
#endif
  if (WS_DBG(2)) {
  } else if (p.splits == 1) {
    const int groups = chunks / 2;  // 16-token groups (m_pad is a multiple of 16)
    const int g0 = half == 0 ? 0 : (groups + 1) / 2, g1 = half == 0 ? (groups + 1) / 2 : groups;
    for (int g = g0; g < g1; ++g) {
      uint32_t r[16];
      tmem_ld_x16(t_row + g * 16, r);
      tmem_wait_ld();
      emit(reinterpret_cast<const float*>(r), 2 * g);
      emit(reinterpret_cast<const float*>(r) + kWsChunk, 2 * g + 1);
    }
    flush(true);
  } else {
    // ---- reduce-scatter over the `splits` pairs of this weight tile: split s owns chunks [own0, own1)
    const int own0 = (chunks * split) / p.splits, own1 = (chunks * (split + 1)) / p.splits;
    const size_t slot_f4 = static_cast<size_t>(p.m_pad / 4) * kWsTileRows;  // float4 per CTA slot
    float4* my_slot = reinterpret_cast<float4*>(p.partials) + static_cast<size_t>(blockIdx.x) * slot_f4 + lane_row;
    for (int c = 2 * half; c < chunks; c += 4) {  // 16-token groups, alternating between the two warps of a quarter
      if (c >= own0 && c + 2 <= own1) continue;
      uint32_t r[16];
      tmem_ld_x16(t_row + c * kWsChunk, r);
      tmem_wait_ld();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = c + h;
        if (cc >= own0 && cc < own1) continue;
        // slot layout [chunk][half 0..1][128 lanes] float4: a warp's 32 lanes touch 512 contiguous bytes
        __stcg(my_slot + (cc * 2 + 0) * kWsTileRows, make_float4(__uint_as_float(r[8 * h]), __uint_as_float(r[8 * h + 1]), __uint_as_float(r[8 * h + 2]), __uint_as_float(r[8 * h + 3])));
        __stcg(my_slot + (cc * 2 + 1) * kWsTileRows, make_float4(__uint_as_float(r[8 * h + 4]), __uint_as_float(r[8 * h + 5]), __uint_as_float(r[8 * h + 6]), __uint_as_float(r[8 * h + 7])));
      }
    }
    // release / acquire through the leader: the CTA barrier orders every warp's parked chunks before the leader's fence
    // (cumulativity), and the readers' loads after its acquire
    ws_epilogue_bar();
    int* arrive = p.counters + tile * 2 + static_cast<int>(cta_rank);
    if (epi_leader) {
      red_release_gpu_add(arrive, 1);
      const long long t0 = clock64();
      while (ld_acquire_gpu(arrive) < p.splits) {
        if (clock64() - t0 > LTXB_WATCHDOG_CYCLES) {
          printf("ltxb: small-M split-K watchdog: tile %d split %d sees %d of %d arrivals\n", tile, split, ld_acquire_gpu(arrive), p.splits);
          __trap();
        }
      }
    }
    ws_epilogue_bar();
#ifdef LTXB_WS_DEBUG
    if (epi_leader) dbg[1] = ws_globaltimer();
#endif
    const float4* tile_slots = reinterpret_cast<const float4*>(p.partials) + (static_cast<size_t>(tile) * p.splits * 2 + cta_rank) * slot_f4 + lane_row;
    const int mid = own0 + (own1 - own0 + 1) / 2;
    const int c_begin = half == 0 ? own0 : mid, c_end = half == 0 ? mid : own1;
    // two chunks per pass, the parts of two other splits loaded together; parts are added in split order with the own
    // accumulator at its own position, so the sum does not depend on arrival order
    for (int c = c_begin; c < c_end; c += 2) {
      const bool two = (c + 1 < c_end);
      uint32_t r[2][8];
      tmem_ld_x8(t_row + c * kWsChunk, r[0]);
      if (two) tmem_ld_x8(t_row + (c + 1) * kWsChunk, r[1]);
      float v[2][kWsChunk];
#pragma unroll
      for (int i = 0; i < kWsChunk; ++i) v[0][i] = 0.f, v[1][i] = 0.f;
      bool own_pending = true;
      int o = 0;
      while (o < p.splits || own_pending) {
        // the next kOthers other splits in order: their parts of both chunks are requested together — one L2 round trip per
        // pass (three with the registers of the one-CTA-per-SM build: one pass per chunk pair at the usual four splits;
        // two under the 96-register cap of the two-per-SM build, where three spill: 14.5 -> 17.3 us)
        int oo[kOthers];
        int q = o;
#pragma unroll
        for (int b = 0; b < kOthers; ++b) {
          if (q == split) ++q;
          oo[b] = q++;
        }
        float4 ld[kOthers][2][2];
#pragma unroll
        for (int b = 0; b < kOthers; ++b) {
          if (oo[b] < p.splits) {
            const float4* src = tile_slots + static_cast<size_t>(oo[b]) * 2 * slot_f4 + (c * 2) * kWsTileRows;
            ld[b][0][0] = __ldcg(src), ld[b][0][1] = __ldcg(src + kWsTileRows);
            if (two) ld[b][1][0] = __ldcg(src + 2 * kWsTileRows), ld[b][1][1] = __ldcg(src + 3 * kWsTileRows);
          }
        }
        auto add_own = [&]() {
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < kWsChunk; ++i) v[0][i] += __uint_as_float(r[0][i]), v[1][i] += two ? __uint_as_float(r[1][i]) : 0.f;
          own_pending = false;
        };
#pragma unroll
        for (int b = 0; b < kOthers; ++b) {
          const bool has = oo[b] < p.splits;
          if (own_pending && (!has || split < oo[b])) add_own();
          if (has) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !two) break;
              v[h][0] += ld[b][h][0].x, v[h][1] += ld[b][h][0].y, v[h][2] += ld[b][h][0].z, v[h][3] += ld[b][h][0].w;
              v[h][4] += ld[b][h][1].x, v[h][5] += ld[b][h][1].y, v[h][6] += ld[b][h][1].z, v[h][7] += ld[b][h][1].w;
            }
          }
        }
        o = oo[kOthers - 1] + 1;
      }
      emit(v[0], c);
      if (two) emit(v[1], c + 1);
    }
    flush(true);
    // the last pair to leave re-arms the counters for the next launch
    ws_epilogue_bar();
    if (epi_leader) {
      int* depart = p.counters + kWsDepartOffset + tile * 2 + static_cast<int>(cta_rank);
      if (atomicAdd(depart, 1) == p.splits - 1) {
        *arrive = 0;
        *depart = 0;
      }
    }
  }
}

// kPerSm: CTAs per SM the launch is sized for — 2 caps a thread at 96 registers, 1 (what most shapes run, see
// launch_gemm_small_m) leaves the epilogue its registers (no spills, three parts per reduce pass)
template <int kEpi, int kPerSm>
__global__ void __launch_bounds__(kWsThreads, kPerSm)
gemm_small_m_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_out, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  WsSmemHeader* hdr = reinterpret_cast<WsSmemHeader*>(smem);
  uint8_t* tiles = smem + kWsHeader;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = (cta_rank == 0);
  const int pair_id = blockIdx.x >> 1;
  const int tile = pair_id / p.splits;
  const int split = pair_id - tile * p.splits;
  const int num_kb = p.K / kWsBlockK;
  const int kb0 = (num_kb * split) / p.splits;
  const int kb1 = (num_kb * (split + 1)) / p.splits;
  const int mma_n = p.m_pad / p.n_mma;  // tokens per MMA
  const int box_rows = mma_n / 2;       // token rows this CTA stages per MMA
  const uint32_t x_box_bytes = box_rows * kWsBlockK * 2;
  const uint32_t stage_bytes = kWsWBytes + p.n_mma * x_box_bytes;  // stage = [W tile][token boxes]
  const int num_stages = p.num_stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (p.out_tma != 0) tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < num_stages; ++s) {
        mbar_init(&hdr->full[s], 1);
        mbar_init(&hdr->empty[s], 1);
      }
      mbar_init(&hdr->tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<2>(&hdr->tmem_base, p.tmem_cols);
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&hdr->tmem_base);
  pdl_launch_dependents();
  // Weights no running kernel writes (LTXB_GEMM_CONST_W) start streaming NOW, under the stream predecessor's tail: the
  // producer fills every stage's weight half, the token rows follow after the wait.
  const int n0_w = tile * (2 * kWsTileRows) + static_cast<int>(cta_rank) * kWsTileRows;
  int prefetched = 0;
  if (p.const_w) {
    prefetched = min(num_stages, kb1 - kb0);
    if (warp == 0 && lane == 0) {
      for (int s = 0; s < prefetched; ++s) {
        if (is_leader) mbar_arrive_expect_tx(&hdr->full[s], stage_bytes * 2);
        tma_load_2d_pair(tiles + static_cast<size_t>(s) * stage_bytes, &tmap_w, mapa_u32(smem_u32(&hdr->full[s]), 0), (kb0 + s) * kWsBlockK, n0_w);
      }
    }
  }
  pdl_wait();  // everything above overlapped the previous kernel's tail; activations / outputs are touched only below
  const int sync_epoch = peer_sync_enter(p.sync);  // token rows written by the peers: flag barrier before the first load
  long long dbg[2] = {0, 0};  // LTXB_WS_DEBUG builds: accumulator complete / split partners met
#ifdef LTXB_WS_DEBUG
  const long long t_pdl = ws_globaltimer();
#endif

  if (warp == 0) {
    // ===================== TMA producer, weight stream: this CTA's 128 weight rows per k-block =====================
    // Two producer warps — this one for the weights, warp 2 (an epilogue warp, idle during the main loop) for the token
    // rows: one lone producer spent ~600 cycles per k-block on its own instruction stream (barrier poll, expect_tx, two TMA
    // issues; LTXB_WS_DEBUG=64 role profile), which paced the kernel.  Warp-uniform loops, one elected lane issues, every
    // operand a 32-bit shared-space address or a running counter.
    {
      const bool issuer = elect_one();
      const uint32_t tiles_u32 = smem_u32(tiles), full_u32 = smem_u32(&hdr->full[0]);
      const uint32_t full_leader = mapa_u32(full_u32, 0);  // both CTAs report their bytes on the leader's barrier
      uint32_t stage = 0, phase = 0;
      int ka = kb0 * kWsBlockK;
      for (int kb = kb0; kb < kb1; ++kb) {
        const bool w_done = (kb - kb0) < prefetched;  // this stage's weight tile went out before the wait
        if (!w_done) {
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          if (issuer) {
            if (is_leader) mbar_arrive_expect_tx_u32(full_u32 + stage * 8, stage_bytes * 2);
            tma_load_2d_pair_u32(tiles_u32 + stage * stage_bytes, &tmap_w, full_leader + stage * 8, ka, n0_w);
          }
          __syncwarp();
        }
        ka += kWsBlockK;
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
      // neither CTA of the pair may retire while commit arrivals for its barriers are still in flight
      for (int s = 0; s < num_stages; ++s) {
        mbar_wait(&hdr->empty[stage], phase ^ 1);
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA): D[256 weight rows, tokens] += W_tile . X^T =====================
    // The whole warp runs the loop (waits, fences) and one elected lane issues: control flow stays warp-uniform and the
    // descriptors are two 32-bit words with compile-time k offsets — a lone `lane == 0` thread needed ~100 cycles per
    // MMA (profiles/r2/attention_s64_notes.md), which at four MMAs per k-block would pace this kernel.
    if (is_leader) {
      const bool issuer = elect_one();
      const uint32_t idesc = make_idesc_bf16(2 * kWsTileRows, mma_n, 0, 0);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, descriptor version 1, 128-byte swizzle
      const uint32_t tiles_lo = ((smem_u32(tiles) & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&hdr->full[stage], phase);
        tc_fence_after_sync();
        if (issuer) {
          const uint32_t w_lo = tiles_lo + ((stage * stage_bytes) >> 4);
          for (int j = 0; j < p.n_mma; ++j) {
            const uint32_t x_lo = w_lo + ((kWsWBytes + j * x_box_bytes) >> 4);
#pragma unroll
            for (int k = 0; k < kWsBlockK / 16; ++k)
              umma_bf16_ss<2>(tmem_base + j * mma_n, desc_from_words(w_lo + 2 * k, kDescHi), desc_from_words(x_lo + 2 * k, kDescHi), idesc,
                              (kb != kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit_pair(&hdr->empty[stage], 3);
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
      if (issuer) umma_commit_pair(&hdr->tmem_full, 3);
      __syncwarp();
    }
  } else {
    if (warp == 2) {
      // ===================== TMA producer, token-row stream: this CTA's half of the token rows per k-block =====================
      // (their bytes are counted on the leader's barrier by the weight stream's expect_tx; a complete_tx that lands before
      // that arrive is fine — the phase cannot complete without it)
      const bool issuer = elect_one();
      const uint32_t tiles_u32 = smem_u32(tiles), full_leader = mapa_u32(smem_u32(&hdr->full[0]), 0);
      uint32_t stage = 0, phase = 0;
      int ka = kb0 * kWsBlockK, gcol = 0, ggrp = 0;  // head-group-major token rows: column inside the group, group index
      if (p.a_group_cols > 0) gcol = ka % p.a_group_cols, ggrp = ka / p.a_group_cols;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&hdr->empty[stage], phase ^ 1);
        if (issuer) {
          const uint32_t sx = tiles_u32 + stage * stage_bytes + kWsWBytes, bar = full_leader + stage * 8;
          for (int j = 0; j < p.n_mma; ++j) {
            const int row0 = j * mma_n + static_cast<int>(cta_rank) * box_rows;
            if (p.a_group_cols > 0) tma_load_3d_pair_u32(sx + j * x_box_bytes, &tmap_x, bar, gcol, row0, ggrp);
            else tma_load_2d_pair_u32(sx + j * x_box_bytes, &tmap_x, bar, ka, row0);
          }
        }
        __syncwarp();
        ka += kWsBlockK;
        if (p.a_group_cols > 0 && (gcol += kWsBlockK) == p.a_group_cols) gcol = 0, ++ggrp;
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
    }
    ws_epilogue<kEpi, kPerSm == 1 ? 3 : 2>(p, &tmap_out, &hdr->tmem_full, tiles, static_cast<uint32_t>(num_stages) * stage_bytes, tmem_base,
                                           tile, split, cta_rank, warp, lane, dbg);
  }

  if (p.out_tma != 0 && warp >= 2 && ((warp - 2) & 3) == 0 && lane == 0) tma_store_wait_read();  // the staged tiles stay valid until read
#ifdef LTXB_WS_DEBUG
  if (WS_DBG(8) && threadIdx.x == 64 && p.trace != nullptr) {  // warp 2 lane 0 = the epilogue leader
    long long* t = p.trace + static_cast<size_t>(blockIdx.x) * 4;
    t[0] = t_pdl, t[1] = dbg[0], t[2] = dbg[1], t[3] = ws_globaltimer();
  }
#endif
  // teardown: nobody may leave while the peer can still read this CTA's shared memory / TMEM
  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();
  if (threadIdx.x == 0) peer_sync_exit(p.sync, sync_epoch);
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<2>(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Packed weights (MLX affine groups, 4 / 8 bits): the same GEMM with W streamed PACKED from HBM (a quarter / half of the
// bf16 bytes — at few rows the weight stream is what bounds the product).
//   * a stage holds the packed tile (4 / 8 KB) and this CTA's half of the token rows (10 KB at 160 tokens): up to 20
//     stages in flight per SM.  That depth is the point: with a bf16 copy of the tile in every stage (first cut) only 7
//     stages fit, and at a ~3 us round trip per stage the pipeline, not HBM, paced the kernel (measured: 49 us against the
//     bf16 kernel's 29 us at 160 x 16384 x 4096, profiles/r2/gemm_small_m.md);
//   * sixteen expanding warps — four groups of four (one warp per TMEM lane quarter), group g owning every fourth k-block —
//     expand the tiles in the arithmetic of ltxb_dequant_affine_bf16 (bit-identical to dequantise-then-GEMM) and write them
//     with tcgen05.st into a ring of A-operand slots in TENSOR MEMORY (one thread = one weight row = one TMEM lane: 32
//     columns per k-block, up to 14 slots behind the accumulator's columns); the MMA reads A from TMEM (the form the
//     attention kernel uses for P.V) and only the token rows from shared memory.  Groups on different k-blocks because a
//     trip (barrier polls, shared-memory read, arithmetic, tcgen05.st + wait, fences, remote arrive) has ~600 cycles of fixed
//     latency that more warps on the SAME k-block do not shorten;
//   * two TMA producer warps: warp 0 streams the packed tiles, warp 18 the token rows.
// ------------------------------------------------------------------------------------------------
constexpr int kPkMaxSlots = 14;  // ring slots: as many as fit behind the accumulator (512 - m_pad columns, 32 each)
constexpr int kPkSlotCols = kWsBlockK / 2;  // 64 bf16 of a weight row = 32 TMEM columns
constexpr int kPkAccCols = 256;             // accumulator columns in front of the slot ring (m_pad <= 256)
constexpr int kPkMaxStages = 20;
constexpr int kPkGroups = 4;                           // expanding warp groups (four warps each), one k-block in flight per group
constexpr int kPkTokenWarp = 2 + 4 * kPkGroups;        // warp 18: the token-row TMA stream
constexpr int kPkThreads = (kPkTokenWarp + 1) * 32;    // warp 0 packed-weight TMA stream, warp 1 MMA issuer, warps 2..17 expand (2..9 also run the epilogue), warp 18
struct PkSmemHeader {
  uint64_t full[kPkMaxStages];      // (leader's) the token rows of both CTAs have landed
  uint64_t raw_full[kPkMaxStages];  // this CTA's packed tile has landed
  uint64_t empty[kPkMaxStages];     // the MMAs that read the stage have completed (both CTAs; the tile was expanded before them)
  uint64_t a_full[kPkMaxSlots];     // (leader's) the owning group's four warps of both CTAs have filled the slot: 8 arrivals
  uint64_t a_empty[kPkMaxSlots];    // the MMAs that read the slot have completed (both CTAs)
  uint64_t tmem_full;
  uint32_t tmem_base;
};
static_assert(sizeof(PkSmemHeader) <= kWsHeader, "header overflow");

template <int kEpi, int kWBits>
__global__ void __launch_bounds__(kPkThreads, 1)
gemm_small_m_packed_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                           const __grid_constant__ CUtensorMap tmap_out, const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PkSmemHeader* hdr = reinterpret_cast<PkSmemHeader*>(smem);
  uint8_t* tiles = smem + kWsHeader;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = (cta_rank == 0);
  const int pair_id = blockIdx.x >> 1;
  const int tile = pair_id / p.splits;
  const int split = pair_id - tile * p.splits;
  const int num_kb = p.K / kWsBlockK;
  const int kb0 = (num_kb * split) / p.splits;
  const int kb1 = (num_kb * (split + 1)) / p.splits;
  const int box_rows = p.m_pad / 2;  // token rows this CTA stages (one MMA: m_pad <= 256)
  constexpr uint32_t kRawRowBytes = kWsBlockK * kWBits / 8;  // packed bytes of one weight row per k-block
  constexpr uint32_t kRawBytes = kWsTileRows * kRawRowBytes;
  const uint32_t x_bytes = box_rows * kWsBlockK * 2;
  const uint32_t stage_bytes = x_bytes + kRawBytes;  // stage = [token rows (swizzled, 1024-aligned)][packed tile]
  const int num_stages = p.num_stages;
  const int n0_w = tile * (2 * kWsTileRows) + static_cast<int>(cta_rank) * kWsTileRows;
  const uint32_t ring_col = (static_cast<uint32_t>(p.m_pad) + 31u) & ~31u;  // A-operand slots start behind the accumulator
  const int num_slots = min(kPkMaxSlots, static_cast<int>((512u - ring_col) / kPkSlotCols));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (p.out_tma != 0) tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < num_stages; ++s) {
        mbar_init(&hdr->full[s], 1);
        mbar_init(&hdr->raw_full[s], 1);
        mbar_init(&hdr->empty[s], 1);
      }
      for (int s = 0; s < num_slots; ++s) {
        mbar_init(&hdr->a_full[s], 2 * 4);  // the four warps of the owning group, in both CTAs
        mbar_init(&hdr->a_empty[s], 1);
      }
      mbar_init(&hdr->tmem_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<2>(&hdr->tmem_base, 512);
  }
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&hdr->tmem_base);
  pdl_launch_dependents();
  int prefetched = 0;
  if (p.const_w) {  // packed tiles of weights no running kernel writes start streaming under the predecessor's tail
    prefetched = min(num_stages, kb1 - kb0);
    if (warp == 0 && lane == 0) {
      for (int s = 0; s < prefetched; ++s) {
        mbar_arrive_expect_tx(&hdr->raw_full[s], kRawBytes);
        tma_load_2d(tiles + static_cast<size_t>(s) * stage_bytes + x_bytes, &tmap_w, &hdr->raw_full[s], (kb0 + s) * static_cast<int>(kRawRowBytes), n0_w);
      }
    }
  }
  pdl_wait();
  const int sync_epoch = peer_sync_enter(p.sync);
  long long dbg[2] = {0, 0};

  if (warp == 0) {
    // ===================== TMA producer, packed-weight stream (warp-uniform loop, one elected lane issues) =====================
    {
      const bool issuer = elect_one();
      const uint32_t tiles_u32 = smem_u32(tiles), raw_full_u32 = smem_u32(&hdr->raw_full[0]);
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0, t_loop = clock64();
      for (int kb = kb0; kb < kb1; ++kb) {
        const bool w_done = (kb - kb0) < prefetched;
        if (!w_done) {
          WS_TIMED_WAIT(w_empty, mbar_wait(&hdr->empty[stage], phase ^ 1));
          if (issuer) {
            mbar_arrive_expect_tx_u32(raw_full_u32 + stage * 8, kRawBytes);
            tma_load_2d_u32(tiles_u32 + stage * stage_bytes + x_bytes, &tmap_w, raw_full_u32 + stage * 8, kb * static_cast<int>(kRawRowBytes), n0_w);
          }
          __syncwarp();
        }
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
#ifdef LTXB_WS_DEBUG
      if (WS_DBG(64) && blockIdx.x == 0 && lane == 0) printf("ws roles: weight producer loop %lld cycles, %lld waiting for empty stages (%d k-blocks)\n", clock64() - t_loop, w_empty, kb1 - kb0);
#endif
      (void)t_loop, (void)w_empty;
      for (int s = 0; s < num_stages; ++s) {  // no CTA retires while commit arrivals for its barriers are in flight
        mbar_wait(&hdr->empty[stage], phase ^ 1);
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == kPkTokenWarp) {
    // ===================== TMA producer, token-row stream (a warp of its own: one producer for both streams spent ~600
    // cycles per k-block on its instruction stream and paced the kernel) =====================
    {
      const bool issuer = elect_one();
      const uint32_t tiles_u32 = smem_u32(tiles), full_u32 = smem_u32(&hdr->full[0]);
      const uint32_t full_leader = mapa_u32(full_u32, 0);  // both CTAs report their token-row bytes on the leader's barrier
      const int row0 = static_cast<int>(cta_rank) * box_rows;
      uint32_t stage = 0, phase = 0;
      int ka = kb0 * kWsBlockK, gcol = 0, ggrp = 0;
      if (p.a_group_cols > 0) gcol = ka % p.a_group_cols, ggrp = ka / p.a_group_cols;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&hdr->empty[stage], phase ^ 1);
        if (issuer) {
          const uint32_t sx = tiles_u32 + stage * stage_bytes;
          const bool skip_x = WS_DBG(128) && kb != kb0;  // (timing experiment: token rows loaded once)
          if (is_leader) mbar_arrive_expect_tx_u32(full_u32 + stage * 8, skip_x ? 0u : x_bytes * 2);
          if (skip_x) {
          } else if (p.a_group_cols > 0) tma_load_3d_pair_u32(sx, &tmap_x, full_leader + stage * 8, gcol, row0, ggrp);
          else tma_load_2d_pair_u32(sx, &tmap_x, full_leader + stage * 8, ka, row0);
        }
        __syncwarp();
        ka += kWsBlockK;
        if (p.a_group_cols > 0 && (gcol += kWsBlockK) == p.a_group_cols) gcol = 0, ++ggrp;
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader): D += A (TMEM slot) . X^T (shared memory) =====================
    if (is_leader) {  // warp-uniform loop, one elected lane issues (see the bf16 kernel)
      const bool issuer = elect_one();
      const uint32_t idesc = make_idesc_bf16(2 * kWsTileRows, p.m_pad, 0, 0);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t tiles_lo = ((smem_u32(tiles) & 0x3FFFFu) >> 4) | (1u << 16);
      uint32_t stage = 0, phase = 0, slot = 0, sphase = 0;
      long long w_full = 0, w_afull = 0, t_loop = clock64();
      for (int kb = kb0; kb < kb1; ++kb) {
        WS_TIMED_WAIT(w_full, mbar_wait(&hdr->full[stage], phase));
        WS_TIMED_WAIT(w_afull, mbar_wait(&hdr->a_full[slot], sphase));
        tc_fence_after_sync();
        if (issuer) {
          const uint32_t x_lo = tiles_lo + ((stage * stage_bytes) >> 4);
          const uint32_t a_tmem = tmem_base + ring_col + slot * kPkSlotCols;
#pragma unroll
          for (int k = 0; k < kWsBlockK / 16; ++k)
            umma_bf16_ts_pair(tmem_base, a_tmem + 8 * k, desc_from_words(x_lo + 2 * k, kDescHi), idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit_pair(&hdr->empty[stage], 3);
          umma_commit_pair(&hdr->a_empty[slot], 3);
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
        if (++slot == static_cast<uint32_t>(num_slots)) slot = 0, sphase ^= 1;
      }
      if (issuer) umma_commit_pair(&hdr->tmem_full, 3);
      __syncwarp();
#ifdef LTXB_WS_DEBUG
      if (WS_DBG(64) && blockIdx.x == 0 && lane == 0) printf("ws roles: MMA loop %lld cycles, %lld waiting for token rows, %lld for expanded tiles\n", clock64() - t_loop, w_full, w_afull);
#endif
      (void)t_loop, (void)w_full, (void)w_afull;
    }
  } else {
    // ===================== expanding warps: packed tile -> bf16 A-operand slot in TMEM =====================
    // Four GROUPS of four warps (one per TMEM lane quarter); group g expands every fourth k-block, whole weight rows (64
    // levels per thread), into that k-block's slot.  A trip is a chain of barrier polls, a shared-memory read, the
    // arithmetic, tcgen05.st + wait, fences and a remote arrive: ~1 000 cycles per k-block however few levels a thread
    // expands (measured with 8 and with 16 warps sharing ONE k-block: same time) — four k-blocks in flight overlap it.
    const int quarter = warp & 3, grp = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;  // weight row of the tile = TMEM lane
    const long long row_abs = static_cast<long long>(n0_w) + row;
    const bool row_ok = row_abs < p.N;
    // (scale, bias) of the row: 16 bytes (8 bf16 / 4 f32 groups) at a time; the chunk of the group's NEXT k-block is requested
    // one trip ahead (a load issued under a saturated HBM stream takes microseconds to come back)
    const int esz = p.aux_f32 ? 4 : 2;
    const int chunk_groups = 16 / esz;
    const int groups_per_row = p.K / p.group;
    const bool chunked = (groups_per_row % chunk_groups == 0) && ((p.lds * esz) % 16 == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.scales) | reinterpret_cast<uintptr_t>(p.biases)) % 16 == 0);
    const int group_shift = p.group == 32 ? 5 : (p.group == 64 ? 6 : 7);
    auto load_chunk = [&](int c, uint4& S, uint4& B) {
      S = make_uint4(0, 0, 0, 0), B = make_uint4(0, 0, 0, 0);
      if (!row_ok || c * chunk_groups >= groups_per_row) return;
      const size_t off = (static_cast<size_t>(row_abs) * p.lds + static_cast<size_t>(c) * chunk_groups) * esz;
      S = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.scales) + off));
      B = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.biases) + off));
    };
    auto pick = [&](const uint4& V, int i) -> float {  // element i of a chunk without indexing registers dynamically
      if (p.aux_f32) {
        const uint32_t lo = (i & 1) ? V.y : V.x, hi = (i & 1) ? V.w : V.z;
        return __uint_as_float((i & 2) ? hi : lo);
      }
      const uint32_t a = (i & 2) ? V.y : V.x, b = (i & 2) ? V.w : V.z;
      const uint32_t wd = (i & 4) ? b : a;
      return __uint_as_float((i & 1) ? (wd & 0xffff0000u) : (wd << 16));
    };
    auto load_scalar = [&](int g, float& sc, float& bi) {  // layouts the 16-byte chunks do not fit
      sc = 0.f, bi = 0.f;
      if (!row_ok) return;
      const long long a = row_abs * p.lds + g;
      if (p.aux_f32) {
        sc = __ldg(static_cast<const float*>(p.scales) + a);
        bi = __ldg(static_cast<const float*>(p.biases) + a);
      } else {
        sc = __bfloat162float(static_cast<const __nv_bfloat16*>(p.scales)[a]);
        bi = __bfloat162float(static_cast<const __nv_bfloat16*>(p.biases)[a]);
      }
    };
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t raw_thread_off = x_bytes + row * kRawRowBytes;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ring_col;
    const uint32_t magic = p.magic;  // 0x4B000000 from a kernel parameter, i.e. in a REGISTER: PRMT then takes its selector as the
                                     // immediate (one instruction per level; a literal makes ptxas re-materialise the selectors)
    const int nkb = kb1 - kb0;
    uint4 s_cur, b_cur, s_next, b_next;
    int cur_chunk = -1, next_chunk = -1;
    if (chunked && grp < nkb) {
      cur_chunk = ((kb0 + grp) * kWsBlockK >> group_shift) / chunk_groups;
      load_chunk(cur_chunk, s_cur, b_cur);
    }
    long long w_raw = 0, w_aempty = 0, w_st = 0, t_loop = clock64();
    for (int i = grp; i < nkb; i += kPkGroups) {
      const int kb = kb0 + i;
      const uint32_t stage = i % num_stages, phase = (i / num_stages) & 1, slot = i % num_slots, sphase = (i / num_slots) & 1;
      const int ga = (kb * kWsBlockK) >> group_shift, gb = (kb * kWsBlockK + 32) >> group_shift;  // groups of columns 0-31 / 32-63
      float sca, bia, scb, bib;
      if (chunked) {
        const int c = ga / chunk_groups;
        if (c != cur_chunk) s_cur = s_next, b_cur = b_next, cur_chunk = c;  // (requested one trip ago)
        if (i + kPkGroups < nkb) {
          const int cn = (((kb + kPkGroups) * kWsBlockK) >> group_shift) / chunk_groups;
          if (cn != cur_chunk && cn != next_chunk) {
            load_chunk(cn, s_next, b_next);
            next_chunk = cn;
          }
        }
        sca = pick(s_cur, ga - c * chunk_groups), bia = pick(b_cur, ga - c * chunk_groups);
        scb = pick(s_cur, gb - c * chunk_groups), bib = pick(b_cur, gb - c * chunk_groups);
      } else {
        load_scalar(ga, sca, bia);
        load_scalar(gb, scb, bib);
      }
      WS_TIMED_WAIT(w_raw, mbar_wait(&hdr->raw_full[stage], phase));
      constexpr int kWords = kWBits * 2;  // packed words of one weight row per k-block: 8 (4-bit) or 16 (8-bit)
      uint32_t w[kWords], packed[32];
      const uint32_t st = tiles_u32 + stage * stage_bytes + raw_thread_off;
#pragma unroll
      for (int v = 0; v < kWords / 4; ++v) lds128(st + 16 * v, w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
#ifdef LTXB_WS_DEBUG
      if (WS_DBG(32)) {  // timing experiment: no expansion arithmetic
#pragma unroll
        for (int v = 0; v < 32; ++v) packed[v] = w[v % kWords];
      } else
#endif
      {
        // level -> float through the 2^23 trick (bits 0x4B0000qq = 8388608 + q, exact), then scales * q + biases with the
        // roundings of ltxb_dequant_affine_bf16: the product of a bf16 scale and an 8-bit level is exact in f32, so one fused
        // multiply-add rounds like its multiply + add; f32 scales keep the two instructions
        auto expand = [&](auto fused) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 8 levels per j
            const float sc = j < 4 ? sca : scb, bi = j < 4 ? bia : bib;
            uint32_t f[8];  // float bit patterns 2^23 + level
            if constexpr (kWBits == 4) {
              const uint32_t lo = w[j] & 0x0F0F0F0Fu, hi = (w[j] >> 4) & 0x0F0F0F0Fu;  // even / odd levels, one per byte
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __byte_perm((e & 1) ? hi : lo, magic, 0x7540u | (e >> 1));
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __byte_perm(w[2 * j + (e >> 2)], magic, 0x7540u | (e & 3));
            }
            if constexpr (decltype(fused)::value) {
              // bf16 scales: two levels per instruction (add / fma .f32x2 round each half exactly like the scalar forms)
              const uint64_t sc2 = pack_f32x2(sc, sc), bi2 = pack_f32x2(bi, bi), neg2 = pack_f32x2(-8388608.0f, -8388608.0f);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint64_t q2 = add_f32x2(pack_f32x2(__uint_as_float(f[2 * e]), __uint_as_float(f[2 * e + 1])), neg2);
                float v0, v1;
                unpack_f32x2(fma_f32x2(sc2, q2, bi2), v0, v1);
                packed[4 * j + e] = pack_bf16x2(v0, v1);
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float q0 = __uint_as_float(f[2 * e]) - 8388608.0f, q1 = __uint_as_float(f[2 * e + 1]) - 8388608.0f;
                packed[4 * j + e] = pack_bf16x2(__fadd_rn(__fmul_rn(sc, q0), bi), __fadd_rn(__fmul_rn(sc, q1), bi));
              }
            }
          }
        };
        if (p.aux_f32) expand(std::false_type{}); else expand(std::true_type{});
      }
      // the MMAs that read this slot one ring turn ago have completed
      WS_TIMED_WAIT(w_aempty, mbar_wait(&hdr->a_empty[slot], sphase ^ 1));
      tc_fence_after_sync();
      WS_TIMED_WAIT(w_st, tmem_st_x32(t_lane + slot * kPkSlotCols, packed); tmem_wait_st(); tc_fence_before_sync(); __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&hdr->a_full[slot], 0));
    }
#ifdef LTXB_WS_DEBUG
    if (WS_DBG(64) && blockIdx.x == 0 && threadIdx.x == 64) printf("ws roles: expanding warp loop %lld cycles, %lld waiting for packed tiles, %lld for free slots, %lld in tcgen05.st + arrive\n", clock64() - t_loop, w_raw, w_aempty, w_st);
#endif
    (void)t_loop, (void)w_raw, (void)w_aempty, (void)w_st;
    if (warp < 10)  // the first two groups are the eight epilogue warps (two per TMEM lane quarter)
      ws_epilogue<kEpi, 2>(p, &tmap_out, &hdr->tmem_full, tiles, static_cast<uint32_t>(num_stages) * stage_bytes, tmem_base, tile, split,
                           cta_rank, warp, lane, dbg);
  }

  if (p.out_tma != 0 && warp >= 2 && ((warp - 2) & 3) == 0 && lane == 0) tma_store_wait_read();
  __syncwarp();
  tc_fence_before_sync();
  cluster_sync_all();
  if (threadIdx.x == 0) peer_sync_exit(p.sync, sync_epoch);
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int kEpi, int kWBits>
static int launch_ws(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& to, const WsParams& p, int grid, size_t smem,
                     cudaStream_t stream, int per_sm) {
  static PerDeviceOnce configured;
  if constexpr (kWBits == 16) {
    if (configured.first()) {
      LTXB_CUDA(cudaFuncSetAttribute(gemm_small_m_kernel<kEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      LTXB_CUDA(cudaFuncSetAttribute(gemm_small_m_kernel<kEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    if (per_sm == 1) LTXB_CUDA(launch_kernel(gemm_small_m_kernel<kEpi, 1>, dim3(grid), dim3(kWsThreads), smem, stream, 2, tx, tw, to, p));
    else LTXB_CUDA(launch_kernel(gemm_small_m_kernel<kEpi, 2>, dim3(grid), dim3(kWsThreads), smem, stream, 2, tx, tw, to, p));
  } else {
    auto kernel = gemm_small_m_packed_kernel<kEpi, kWBits>;
    if (configured.first()) LTXB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    LTXB_CUDA(launch_kernel(kernel, dim3(grid), dim3(kPkThreads), smem, stream, 2, tx, tw, to, p));
  }
  return LTXB_OK;
}

bool gemm_small_m_supported(int M, int N, int K) { return M >= 1 && M <= 512 && N >= 16 && K >= kWsBlockK && K % kWsBlockK == 0; }

int launch_gemm_small_m(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int M, int N, int K,
                        const ltxb_epilogue* epi, float* partials, long long partial_bytes, int* counters, int want_splits,
                        cudaStream_t stream, const WsPacked* packed) {
  const int w_bits = packed != nullptr ? packed->bits : 16;
  const int sms = num_sms();
  if (sms <= 0) return set_error(LTXB_ERR_NO_DEVICE, "ltxb_gemm_bf16: no CUDA device");
  WsParams p{};
  p.M = M, p.N = N, p.K = K;
  p.n_mma = M > 256 ? 2 : 1;
  p.m_pad = p.n_mma == 1 ? ((M + 15) / 16) * 16 : ((M + 31) / 32) * 32;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.m_pad) p.tmem_cols *= 2;
  // packed weights: a stage is the packed tile + the token rows (the bf16 tile lives in tensor memory), one CTA per SM
  const size_t stage_bytes = w_bits != 16 ? static_cast<size_t>(p.m_pad / 2) * kWsBlockK * 2 + kWsTileRows * kWsBlockK * w_bits / 8
                                          : kWsWBytes + static_cast<size_t>(p.m_pad / 2) * kWsBlockK * 2;
  if (w_bits != 16) {
    LTXB_CHECK_SUPPORTED(M <= kPkAccCols, "ltxb_gemm_qw_bf16: M=%d: the packed-weight kernel keeps the expanded tiles in tensor memory behind a %d-column accumulator (expand the weights with ltxb_dequant_affine_bf16 for more rows)", M, kPkAccCols);
    p.tmem_cols = 512;
  }
  const int tiles = (N + 2 * kWsTileRows - 1) / (2 * kWsTileRows);
  const int num_kb = K / kWsBlockK;
  static const int env_min_kb = [] { const char* e = getenv("LTXB_GEMM_SMALL_M_MIN_KB"); return e ? atoi(e) : 4; }();
  static const int env_max_splits = [] { const char* e = getenv("LTXB_GEMM_SMALL_M_MAX_SPLITS"); return e ? atoi(e) : 16; }();
  static const int env_per_sm = [] { const char* e = getenv("LTXB_GEMM_SMALL_M_PER_SM"); return e ? atoi(e) : 0; }();
  const bool can_split = partials != nullptr && counters != nullptr && tiles * 2 <= kWsDepartOffset;
  auto splits_for = [&](int slots) {
    if (!can_split) return 1;
    int s = std::max(1, std::min({slots / tiles, num_kb / std::max(1, env_min_kb), env_max_splits}));
    if (want_splits > 0) s = std::min({want_splits, num_kb, std::max(1, slots / tiles)});
    return s;
  };
  // One CTA per SM with all of its shared memory as stages (8 x 26 KB at 160 tokens), or two per SM with half each (when
  // both accumulators fit 256 TMEM columns).  Two per SM doubles the pair slots, i.e. allows more k-range splits, but every
  // split adds reduce-scatter traffic and a meeting point, and the deeper pipeline streams faster: measured at 68 / 160
  // tokens (profiles/r2/gemm/few_row_kernel_per_sm_sweep.txt) one per SM wins or ties on every shape whose weight tiles fit
  // the SM pairs in one go — even the qkv shape, 48 tiles on 74 pairs without any split: 25.4 us against 27.6 with three
  // splits on two CTAs per SM — and two per SM only when there are more tiles than pairs (AdaLN, 144 tiles: 48.7 against 51.9).
  const int pairs = sms / 2;
  int per_sm = 1;
  if (p.tmem_cols <= 256 && (113 * 1024 - 1024 - kWsHeader) / stage_bytes >= 3) {
    per_sm = (env_per_sm == 1 || env_per_sm == 2) ? env_per_sm : (tiles <= pairs ? 1 : 2);
  }
  const size_t budget = (per_sm == 2 ? 113 * 1024 : 232448) - 1024 - kWsHeader;
  p.num_stages = static_cast<int>(std::min<size_t>(w_bits != 16 ? kPkMaxStages : kWsMaxStages, budget / stage_bytes));
  const size_t smem = 1024 + kWsHeader + p.num_stages * stage_bytes;
  const int slots = pairs * per_sm;  // co-resident CTA pairs
  int splits = splits_for(slots);
  // every CTA of the grid parks into its own slot
  while (splits > 1 && static_cast<long long>(tiles) * splits * 2 * p.m_pad * kWsTileRows * 4 > partial_bytes) --splits;
  p.splits = splits;
  p.bias = epi->bias;
  p.out = out;
  p.ldo = ldo;
  p.resid = epi->resid;
  p.ldr = epi->ldr;
  p.gate = epi->gate;
  p.gate_ld = epi->gate_ld;
  p.gate_row_div = epi->gate_row_div > 0 ? epi->gate_row_div : 1;
  p.gate_row_index = epi->gate_row_index;
  p.gate_table = epi->gate_table;
  p.partials = partials;
  p.counters = counters;
  if (int rc = peer_sync_from_abi(epi->peer_sync, &p.sync, "ltxb_gemm")) return rc;
  if (packed != nullptr) {
    p.magic = 0x4B000000u;
    p.scales = packed->scales, p.biases = packed->biases, p.lds = packed->lds, p.group = packed->group, p.aux_f32 = packed->aux_f32;
  }
  static const int env_const_w = [] { const char* e = getenv("LTXB_GEMM_CONST_W"); return e ? atoi(e) : 1; }();
  p.const_w = (env_const_w && (epi->flags & LTXB_GEMM_CONST_W)) ? 1 : 0;
#ifdef LTXB_WS_DEBUG
  {
    const char* e = getenv("LTXB_WS_DEBUG");
    p.debug = e ? atoi(e) : 0;
    if ((p.debug & 8) && partials != nullptr) p.trace = reinterpret_cast<long long*>(reinterpret_cast<char*>(partials) + partial_bytes) - 296 * 4 * 4;
  }
#endif

  const uint32_t box_rows = static_cast<uint32_t>(p.m_pad / p.n_mma / 2);
  CUtensorMap tx, tw;
  if (epi->a_group_cols > 0) {
    LTXB_CHECK_SUPPORTED(epi->a_group_cols % kWsBlockK == 0 && K % epi->a_group_cols == 0 && epi->a_group_stride % 8 == 0,
                         "ltxb_gemm_bf16: a_group_cols=%d must divide K=%d and be a multiple of %d", epi->a_group_cols, K, kWsBlockK);
    p.a_group_cols = epi->a_group_cols;
    const uint64_t dims[3] = {static_cast<uint64_t>(epi->a_group_cols), static_cast<uint64_t>(M), static_cast<uint64_t>(K / epi->a_group_cols)};
    const uint64_t strides[2] = {static_cast<uint64_t>(lda) * 2, static_cast<uint64_t>(epi->a_group_stride) * 2};
    const uint32_t box[3] = {kWsBlockK, box_rows, 1};
    int rc = encode_tmap_bf16(&tx, A, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {kWsBlockK, box_rows};
    int rc = encode_tmap_bf16(&tx, A, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (packed != nullptr) {  // the packed tile as bytes: 64 levels = 32 / 64 bytes per row and k-block, 128 rows
    const uint64_t dims[2] = {static_cast<uint64_t>(K) * w_bits / 8, static_cast<uint64_t>(N)};
    const uint32_t box[2] = {static_cast<uint32_t>(kWsBlockK * w_bits / 8), kWsTileRows};
    int rc = encode_tmap_plain_2d(&tw, W, 1, dims, static_cast<uint64_t>(ldw) * 4, box);
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kWsBlockK, kWsTileRows};
    int rc = encode_tmap_bf16(&tw, W, 2, dims, strides, box);
    if (rc) return rc;
  }
  // Output through shared memory + TMA (16x fewer store instructions than a 2-byte store per thread and token, measured:
  // 2.7 us of a 3.6 us epilogue at 160 tokens were the stores).  RESID_GATE is an in-place update of the residual stream
  // (out == resid) on the model path: the TMA adds the gated update into it; other aliasing keeps the per-thread path.
  static const int env_tma = [] { const char* e = getenv("LTXB_GEMM_SMALL_M_TMA_OUT"); return e ? atoi(e) : 1; }();
  CUtensorMap to{};
  const bool f32_out = (epi->mode == LTXB_EPI_BIAS_F32 || epi->mode == LTXB_EPI_RESID_GATE_F32);
  const int esize = f32_out ? 4 : 2;
  p.out_tma = 0;
  if (env_tma && aligned16(out) && (static_cast<uint64_t>(ldo) * esize) % 16 == 0 && N % (16 / esize) == 0) {
    if (epi->mode != LTXB_EPI_RESID_GATE_F32) p.out_tma = 1;
    else if (epi->resid == out && epi->ldr == ldo) p.out_tma = 2;
  }
  if (p.out_tma != 0) {
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    const uint32_t box[2] = {kWsTileRows, kWsChunk};
    int rc = encode_tmap_plain_2d(&to, out, esize, dims, static_cast<uint64_t>(ldo) * esize, box);
    if (rc) return rc;
  }
  const int grid = tiles * splits * 2;
#define LTXB_WS_DISPATCH(BITS)                                                                                         \
  switch (epi->mode) {                                                                                                 \
    case LTXB_EPI_BIAS_BF16: return launch_ws<LTXB_EPI_BIAS_BF16, BITS>(tx, tw, to, p, grid, smem, stream, per_sm);             \
    case LTXB_EPI_GELU_BF16: return launch_ws<LTXB_EPI_GELU_BF16, BITS>(tx, tw, to, p, grid, smem, stream, per_sm);             \
    case LTXB_EPI_SILU_BF16: return launch_ws<LTXB_EPI_SILU_BF16, BITS>(tx, tw, to, p, grid, smem, stream, per_sm);             \
    case LTXB_EPI_BIAS_F32: return launch_ws<LTXB_EPI_BIAS_F32, BITS>(tx, tw, to, p, grid, smem, stream, per_sm);               \
    case LTXB_EPI_RESID_GATE_F32: return launch_ws<LTXB_EPI_RESID_GATE_F32, BITS>(tx, tw, to, p, grid, smem, stream, per_sm);   \
    default: return set_error(LTXB_ERR_BAD_ARG, "ltxb_gemm: unknown epilogue mode %d", epi->mode);                     \
  }
  if (w_bits == 4) { LTXB_WS_DISPATCH(4) }
  if (w_bits == 8) { LTXB_WS_DISPATCH(8) }
  LTXB_WS_DISPATCH(16)
#undef LTXB_WS_DISPATCH
}

}  // namespace ltxb
