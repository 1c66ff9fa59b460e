// K1: out = epilogue(A[M,K] . W[N,K]^T) — bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear call site on the LTX-2 DiT path of the reference (mlx_video/models/ltx/
// attention.py:91-93,100,123-126,142; feed_forward.py:31,33; ltx.py:130,292,301; adaln.py:27,130-132;
// text_projection.py:18,20), with the elementwise op that follows it fused into the epilogue.
//
// Design (B200 / sm_100a):
//   * persistent kernel, one CTA (or one CTA pair, cta_group::2) per SM (pair), static tile striding,
//     M-fastest tile order so concurrently resident tiles share the same W panels in L2;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM alloc),
//     warps 4..11 = epilogue (TMEM -> registers -> global), two per TMEM lane quarter;
//   * smem ring of `num_stages` {A 128x64, W (BN/ctas)x64} bf16 tiles, 128-byte swizzle, filled by
//     cp.async.bulk.tensor, consumed by tcgen05.mma straight from smem descriptors;
//   * two TMEM accumulators (2 x BN fp32 columns) so the epilogue of tile i overlaps the main loop of
//     tile i+1;
//   * BN is a runtime value (multiple of 16, <= 256): the host picks it per problem so that the tile
//     count fills whole waves of 148 SMs (e.g. M=1280, N=4096: BN=144 in pair mode -> 145 tiles on 74
//     pairs = 96 % wave efficiency instead of 54 % at BN=256).
#include "common.cuh"
#include "gemm_small_m.cuh"
#include "peer_sync.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace ltxb {

constexpr int kBlockM = 128;  // rows of the output tile owned by ONE CTA (TMEM lanes)
constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle span
constexpr int kUmmaK = 16;
constexpr int kMaxStages = 8;
constexpr int kGemmThreads = 384;  // warp 0 TMA, warp 1 MMA, warps 2,3 idle, warps 4..11 epilogue (two per TMEM lane quarter)
// 10 warps would cap every thread at 168 registers (3 warps on one scheduler) and the residual epilogue spilled;
// with three full warpgroups the epilogue warps take what the producer / issuer warpgroup does not need
// (setmaxnreg): 128 * 104 + 256 * 200 <= 64 K
constexpr int kRegsIssue = 104, kRegsEpilogue = 200;
constexpr int kEpiThreads = 256;
constexpr int kSmemHeader = 1024;  // barriers + tmem pointer live in front of the tile ring
constexpr int kTmemCols = 512;
constexpr uint32_t kAccStride = 256;  // TMEM column stride between the two accumulators

// LTXB_GEMM_TRACE: CTA 0 records clock64() at its life-cycle events into the first bytes of the split-K partial
// workspace (long long [16]); scripts/gemm_trace.py prints them.
#ifdef LTXB_GEMM_TRACE
#define GTRACE(ev)                                                                                           \
  do {                                                                                                       \
    if (blockIdx.x == 0 && p.sk_partials != nullptr) reinterpret_cast<long long*>(p.sk_partials)[(ev)] = clock64(); \
  } while (0)
#else
#define GTRACE(ev) do {} while (0)
#endif

struct GemmParams {
  int M, N, K;
  int block_n;
  int num_stages;
  int num_m_tiles;  // in units of kBlockM * ctas rows
  int num_n_tiles;
  // epilogue
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
  int a_group_cols;  // > 0: A is [groups][M][a_group_cols] (a 3-D tensor map), K index = group * a_group_cols + col
  // split-K of the ragged wave (sk_splits <= 1: plain data-parallel tile striding)
  int sk_rem;          // number of tiles in the ragged wave (tile indices [0, sk_rem))
  int sk_splits;       // k-range pieces per such tile
  float* sk_partials;  // [clusters][ctas][kBlockM * block_n] fp32 parked partial accumulators
  int* sk_counters;    // [sk_rem][ctas] arrivals per cut tile; zero outside a launch
  // contiguous stream-K (small M, weight-streaming regime): cluster c owns k-block units [c*U/C, (c+1)*U/C) of the
  // tile-major unit list (U = tiles * k-blocks), boundaries rounded to sk_gran k-blocks
  int sk_contig;
  int sk_gran;
  PeerSync sync;  // cross-GPU flag barrier before the first read of A (n_peers 0: none)
};

// Tiles are dealt to clusters round-robin, all clusters marching through K in lockstep (tiles that share an A
// or W panel then hit it in L2 while it is hot).  When the tile count leaves a ragged wave — M = 1280 gives 80
// tiles for 74 SM pairs — the `sk_rem` leftover tiles are run FIRST, each cut into `sk_splits` k-ranges on
// different clusters (still in lockstep), so the wave costs 1/sk_splits of a tile instead of a whole one.  The
// cluster holding a cut tile's first k-range owns its epilogue: the others park their fp32 partial in L2 and
// signal; the owner adds them in k order (bit-reproducible) while its next tile's main loop is already running.
struct WorkItem {
  int tile, kb0, kb1;
};
struct WorkIter {
  int num_kb, num_tiles, stride, tile, split_tile, split_kb0, split_kb1;
  long long u, u_end;  // contiguous mode: next unit / end of this cluster's range (unit = tile * num_kb + kb)
  bool contig;
  // start of cluster c's unit range in contiguous mode
  static __device__ __forceinline__ long long range_start(long long units, int clusters, int c, int gran) {
    if (c >= clusters) return units;
    const long long b = units * c / clusters;
    return b - b % gran;
  }
  __device__ WorkIter(const GemmParams& p, int cluster_id, int num_clusters, int num_kb_)
      : num_kb(num_kb_), num_tiles(p.num_m_tiles * p.num_n_tiles), stride(num_clusters), split_tile(-1), u(0), u_end(0),
        contig(p.sk_contig != 0) {
    if (contig) {
      const long long units = static_cast<long long>(num_tiles) * num_kb_;
      u = range_start(units, num_clusters, cluster_id, p.sk_gran);
      u_end = range_start(units, num_clusters, cluster_id + 1, p.sk_gran);
      tile = 0;
      return;
    }
    int first = 0;
    if (p.sk_splits > 1) {
      first = p.sk_rem;
      if (cluster_id < p.sk_rem * p.sk_splits) {
        const int piece = cluster_id / p.sk_rem;  // consecutive clusters take consecutive tiles of the same k-range
        split_tile = cluster_id % p.sk_rem;
        split_kb0 = (num_kb_ * piece) / p.sk_splits;
        split_kb1 = (num_kb_ * (piece + 1)) / p.sk_splits;
      }
    }
    tile = first + cluster_id;
  }
  __device__ __forceinline__ bool next(WorkItem& w) {
    if (contig) {
      if (u >= u_end) return false;
      w.tile = static_cast<int>(u / num_kb);
      w.kb0 = static_cast<int>(u - static_cast<long long>(w.tile) * num_kb);
      w.kb1 = static_cast<int>(min(static_cast<long long>(num_kb), w.kb0 + (u_end - u)));
      u += w.kb1 - w.kb0;
      return true;
    }
    if (split_tile >= 0) {
      w.tile = split_tile, w.kb0 = split_kb0, w.kb1 = split_kb1;
      split_tile = -1;
      return true;
    }
    if (tile >= num_tiles) return false;
    w.tile = tile, w.kb0 = 0, w.kb1 = num_kb;
    tile += stride;
    return true;
  }
};

struct GemmSmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(GemmSmemHeader) <= kSmemHeader, "header overflow");

__device__ __forceinline__ void epilogue_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Apply the fused epilogue to `n` (16 or 32) consecutive accumulator columns of one output row.
template <int kEpi, int kCols>
__device__ __forceinline__ void epilogue_store(const GemmParams& p, const uint32_t* acc, long long row, int col,
                                               long long grow) {
  float v[kCols];
#pragma unroll
  for (int i = 0; i < kCols; ++i) v[i] = __uint_as_float(acc[i]);
  if (p.bias != nullptr) {
    // (tried: the tile's bias slice staged in shared memory once per tile instead of one global round trip per chunk —
    // the epilogue of a single tile got shorter, 12.3 k -> 8.9 k cycles, but the two extra named barriers per tile cost
    // more in the multi-tile kernels: +0.35 ms per step in a same-box A/B, profiles/r2/gemm_small_m.md)
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int i = 0; i < kCols / 4; ++i) {
      const float4 b = __ldg(b4 + i);
      v[4 * i + 0] += b.x;
      v[4 * i + 1] += b.y;
      v[4 * i + 2] += b.z;
      v[4 * i + 3] += b.w;
    }
  }
  if constexpr (kEpi == LTXB_EPI_GELU_BF16) {
#pragma unroll
    for (int i = 0; i < kCols; ++i) v[i] = gelu_tanh(v[i]);
  } else if constexpr (kEpi == LTXB_EPI_SILU_BF16) {
#pragma unroll
    for (int i = 0; i < kCols; ++i) v[i] = silu(v[i]);
  }
  if constexpr (kEpi == LTXB_EPI_BIAS_BF16 || kEpi == LTXB_EPI_GELU_BF16 || kEpi == LTXB_EPI_SILU_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col;
    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
    for (int i = 0; i < kCols / 8; ++i) {
      uint4 w;
      w.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      w.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      w.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      w.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      o4[i] = w;
    }
  } else if constexpr (kEpi == LTXB_EPI_BIAS_F32) {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + col);
#pragma unroll
    for (int i = 0; i < kCols / 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {  // LTXB_EPI_RESID_GATE_F32
    const float4* r4 = reinterpret_cast<const float4*>(p.resid + row * p.ldr + col);
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + col);
    // `out` normally IS `resid` (the residual stream is updated in place), so the compiler may not move a residual
    // load above an earlier store: read the whole chunk first, or the chunk becomes kCols/4 serialised L2 round trips
    // (measured: 79.8 -> 61.4 us for M=1280 N=4096 K=4096 together with the register re-balancing below)
    float4 rr[kCols / 4];
#pragma unroll
    for (int i = 0; i < kCols / 4; ++i) rr[i] = r4[i];
    if (p.gate != nullptr) {
      const float4* g4 = reinterpret_cast<const float4*>(p.gate + grow * p.gate_ld + col);
      const float4* t4 = reinterpret_cast<const float4*>(p.gate_table != nullptr ? p.gate_table + col : nullptr);
#pragma unroll
      for (int i = 0; i < kCols / 4; ++i) {
        const float4 r = rr[i];
        float4 g = __ldg(g4 + i);
        if (p.gate_table != nullptr) {
          const float4 t = __ldg(t4 + i);
          g.x += t.x, g.y += t.y, g.z += t.z, g.w += t.w;
        }
        o4[i] = make_float4(fmaf(v[4 * i], g.x, r.x), fmaf(v[4 * i + 1], g.y, r.y), fmaf(v[4 * i + 2], g.z, r.z),
                            fmaf(v[4 * i + 3], g.w, r.w));
      }
    } else {
#pragma unroll
      for (int i = 0; i < kCols / 4; ++i) {
        const float4 r = rr[i];
        o4[i] = make_float4(v[4 * i] + r.x, v[4 * i + 1] + r.y, v[4 * i + 2] + r.z, v[4 * i + 3] + r.w);
      }
    }
  }
}

template <int kCtas, int kEpi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles; the dynamic smem base is the same in every CTA.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmSmemHeader* hdr = reinterpret_cast<GemmSmemHeader*>(smem);
  uint8_t* tiles = smem + kSmemHeader;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) GTRACE(0);  // kernel entry
  const uint32_t cta_rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);
  const int num_clusters = gridDim.x / kCtas;
  const int cluster_id = blockIdx.x / kCtas;

  const int bn = p.block_n;
  const int bn_load = bn / kCtas;  // W rows this CTA stages per k-block
  const uint32_t a_bytes = kBlockM * kBlockK * 2;
  const uint32_t w_bytes = bn_load * kBlockK * 2;
  const uint32_t stage_bytes = a_bytes + w_bytes;
  const int num_stages = p.num_stages;
  const int num_kb = p.K / kBlockK;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < num_stages; ++s) {
        mbar_init(&hdr->full[s], 1);
        mbar_init(&hdr->empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&hdr->tmem_full[a], 1);
        mbar_init(&hdr->tmem_empty[a], 8 * kCtas);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kCtas>(&hdr->tmem_base, kTmemCols);
  }
  tc_fence_before_sync();
  if constexpr (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&hdr->tmem_base);
  if (threadIdx.x == 0) GTRACE(1);  // barriers + TMEM + cluster sync done
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) GTRACE(2);  // predecessor complete
  const int sync_epoch = peer_sync_enter(p.sync);  // A was written by the peers: flag barrier before the first load

  if (warp < 4) {
  reg_dealloc<kRegsIssue>();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      WorkIter it(p, cluster_id, num_clusters, num_kb);
      WorkItem w;
      while (it.next(w)) {
        const int m_tile = w.tile % p.num_m_tiles;
        const int n_tile = w.tile / p.num_m_tiles;
        const int m0 = (m_tile * kCtas + static_cast<int>(cta_rank)) * kBlockM;
        const int n0 = n_tile * bn + static_cast<int>(cta_rank) * bn_load;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&hdr->empty[stage], phase ^ 1);
          uint8_t* sa = tiles + static_cast<size_t>(stage) * stage_bytes;
          uint8_t* sw = sa + a_bytes;
          const int ka = kb * kBlockK;
          if constexpr (kCtas == 1) {
            mbar_arrive_expect_tx(&hdr->full[stage], stage_bytes);
            if (p.a_group_cols > 0) tma_load_3d(sa, &tmap_a, &hdr->full[stage], ka % p.a_group_cols, m0, ka / p.a_group_cols);
            else tma_load_2d(sa, &tmap_a, &hdr->full[stage], ka, m0);
            tma_load_2d(sw, &tmap_w, &hdr->full[stage], ka, n0);
          } else {
            // both CTAs of the pair report their bytes on the LEADER's barrier
            if (is_leader) mbar_arrive_expect_tx(&hdr->full[stage], stage_bytes * 2);
            const uint32_t bar = mapa_u32(smem_u32(&hdr->full[stage]), 0);
            if (p.a_group_cols > 0) tma_load_3d_pair(sa, &tmap_a, bar, ka % p.a_group_cols, m0, ka / p.a_group_cols);
            else tma_load_2d_pair(sa, &tmap_a, bar, ka, m0);
            tma_load_2d_pair(sw, &tmap_w, bar, ka, n0);
          }
          if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
        }
      }
      // tail: wait until every fill has been consumed, so neither CTA of a pair retires while
      // commit arrivals for its barriers are still in flight
      for (int s = 0; s < num_stages; ++s) {
        mbar_wait(&hdr->empty[stage], phase ^ 1);
        if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: leader CTA only) =====================
    if (lane == 0 && is_leader) {
      const uint32_t idesc = make_idesc_bf16(kBlockM * kCtas, bn, 0, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      WorkIter it(p, cluster_id, num_clusters, num_kb);
      WorkItem w;
      while (it.next(w)) {
        mbar_wait(&hdr->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&hdr->full[stage], phase);
          if (kb == w.kb0) GTRACE(3);  // first operands of this item landed
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(tiles + static_cast<size_t>(stage) * stage_bytes);
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t wdesc = make_smem_desc_sw128(sa + a_bytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advancing 16 elements (32 B) along K inside the 128-B swizzle span = +2 in the address field
            umma_bf16_ss<kCtas>(d_tmem, adesc + 2 * k, wdesc + 2 * k, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
          }
          if constexpr (kCtas == 1) umma_commit(&hdr->empty[stage]); else umma_commit_pair(&hdr->empty[stage], 3);
          if (++stage == static_cast<uint32_t>(num_stages)) stage = 0, phase ^= 1;
        }
        if constexpr (kCtas == 1) umma_commit(&hdr->tmem_full[acc]); else umma_commit_pair(&hdr->tmem_full[acc], 3);
        GTRACE(4);  // all MMAs of this item issued
        if (++acc == 2) acc = 0, acc_phase ^= 1;
      }
    }
  }
  } else {
    // ===================== epilogue warps =====================
    reg_alloc<kRegsEpilogue>();
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
    // two warps share each lane quarter and split the tile's 32-column chunks between them: the epilogue is a
    // chain of dependent memory round trips per chunk, so doubling the warps halves its (exposed) latency
    const int col_half = (warp - 4) >> 2;
    const int n_chunks = (bn + 31) / 32;
    const int c_begin = (col_half == 0 ? 0 : (n_chunks + 1) / 2) * 32;
    const int c_end = min(bn, (col_half == 0 ? (n_chunks + 1) / 2 : n_chunks) * 32);
    const bool epi_leader = (warp == 4 && lane == 0);
    const int row_in_cta = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    auto release_acc = [&]() {  // all of this warp's TMEM reads of the accumulator have completed
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtas == 1) mbar_arrive(&hdr->tmem_empty[acc]); else mbar_arrive_remote(&hdr->tmem_empty[acc], 0);
      }
    };
    WorkIter it(p, cluster_id, num_clusters, num_kb);
    WorkItem w;
    while (it.next(w)) {
      const int m_tile = w.tile % p.num_m_tiles;
      const int n_tile = w.tile / p.num_m_tiles;
      const long long row = static_cast<long long>(m_tile * kCtas + static_cast<int>(cta_rank)) * kBlockM + row_in_cta;
      const int n0 = n_tile * bn;
      const bool row_ok = row < p.M;
      long long grow = 0;
      if constexpr (kEpi == LTXB_EPI_RESID_GATE_F32) {
        if (row_ok && p.gate != nullptr)
          grow = p.gate_row_index != nullptr ? p.gate_row_index[row] : row / p.gate_row_div;
      }
      const bool partial = (w.kb0 > 0) || (w.kb1 < num_kb);
      // ---- stream-K roles for a cut tile.  The cluster holding the tile's FIRST k-range (kb0 == 0) owns its
      // epilogue and reaches it as the LAST item of its range; every other holder (kb0 > 0) reaches its piece as
      // the FIRST item of its range, parks the fp32 partial in its slot and signals.  By the time the owner gets
      // there the partials have been sitting in L2 for most of a tile, so its wait is a formality and it can pull
      // them in underneath its own main loop.
      const bool owner = partial && w.kb0 == 0;
      const size_t slot_elems = static_cast<size_t>(kBlockM) * bn;
      // slot layout [32-col chunk][float4 index 0..7][row 0..127]: for a fixed float4 index the 32 lanes of a warp
      // touch 512 contiguous bytes, so both the dump and the reload are fully coalesced
      constexpr int kF4Stride = kBlockM;
      auto chunk_of = [&](int cluster, int c) {
        float* slot = p.sk_partials + (static_cast<size_t>(cluster) * kCtas + cta_rank) * slot_elems;
        return reinterpret_cast<float4*>(slot + static_cast<size_t>(c / 32) * (kBlockM * 32)) + row_in_cta;
      };
      int* counter = p.sk_counters + w.tile * kCtas + static_cast<int>(cta_rank);
      int others = 0;  // owner: number of parked partials to add, held by clusters contrib(0) .. contrib(others-1) in k order
      // lockstep split: piece o+1 of tile t sits on cluster t + (o+1)*sk_rem; contiguous: on the next clusters in line
      const int contrib_stride = p.sk_contig ? 1 : p.sk_rem;
      if (owner) {
        if (p.sk_contig) {
          const long long units = static_cast<long long>(num_tiles) * num_kb;
          const long long tile_end = static_cast<long long>(w.tile + 1) * num_kb;
          for (int c = cluster_id + 1; c < num_clusters && WorkIter::range_start(units, num_clusters, c, p.sk_gran) < tile_end; ++c) ++others;
        } else {
          others = p.sk_splits - 1;
        }
        if (epi_leader) {
          const long long t0 = clock64();
          while (*reinterpret_cast<volatile int*>(counter) < others) {
            if (clock64() - t0 > LTXB_WATCHDOG_CYCLES) {
              printf("ltxb: split-K watchdog: tile %d waits for %d partials\n", w.tile, others);
              __trap();
            }
          }
          *counter = 0;  // every contributor has arrived; ready for the next launch
          GTRACE(9);  // owner: all partials parked
        }
        epilogue_bar_sync();
        __threadfence();
      }
      mbar_wait(&hdr->tmem_full[acc], acc_phase);
      if (epi_leader) GTRACE(5);  // accumulator complete
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccStride;
      auto finish_chunk = [&](const float* v, int c, int width) {
        const int col = n0 + c;
        if (!row_ok) return;
        const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
        if (width == 32 && col + 32 <= p.N) {
          epilogue_store<kEpi, 32>(p, r, row, col, grow);
        } else if (col + 16 <= p.N) {
          epilogue_store<kEpi, 16>(p, r, row, col, grow);
        }
      };
      if (c_begin >= c_end && !(partial && !owner)) release_acc();  // nothing to read for this warp (narrow tile)
      if (!partial) {
        for (int c = c_begin; c < c_end; c += 32) {
          const bool last = (c + 32 >= c_end);
          const int col = n0 + c;
          if (c + 32 <= bn) {
            uint32_t r[32];
            tmem_ld_x32(t_row + c, r);
            tmem_wait_ld();
            if (last) release_acc();
            if (row_ok) {
              if (col + 32 <= p.N) {
                epilogue_store<kEpi, 32>(p, r, row, col, grow);
              } else if (col + 16 <= p.N) {
                epilogue_store<kEpi, 16>(p, r, row, col, grow);
              }
            }
          } else {  // 16-column tail of a BN that is not a multiple of 32
            uint32_t r[16];
            tmem_ld_x16(t_row + c, r);
            tmem_wait_ld();
            release_acc();
            if (row_ok && col + 16 <= p.N) epilogue_store<kEpi, 16>(p, r, row, col, grow);
          }
        }
      } else if (!owner) {
        // ---- contributor: park the partial, publish it, move on
        for (int c = c_begin; c < c_end; c += 32) {
          float4* dst = chunk_of(cluster_id, c);
          if (c + 32 <= bn) {
            uint32_t r[32];
            tmem_ld_x32(t_row + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              __stcg(dst + i * kF4Stride, make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
          } else {
            uint32_t r[16];
            tmem_ld_x16(t_row + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 4; ++i)
              __stcg(dst + i * kF4Stride, make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
          }
        }
        release_acc();
        __threadfence();
        epilogue_bar_sync();
        if (epi_leader) atomicAdd(counter, 1);
      } else {
        // ---- owner: own accumulator (TMEM) + the parked partials of pieces 1.. in k order.  The loads of up to
        // kFixBatch partials for a 32-column chunk are issued together (the fix-up is a chain of L2 round trips; measured
        // with LTXB_GEMM_TRACE at M = 160, N = K = 4096: 20 k cycles of fix-up behind a 19 k-cycle main loop.  Batching four
        // instead of two did not shorten it — the row-per-lane stores, ~1 k LSU wavefronts per chunk and CTA, pace it).
        for (int c = c_begin; c < c_end; c += 32) {
          const int width = (c + 32 <= bn) ? 32 : 16;
          constexpr int kFixBatch = 2;
          float4 ld[kFixBatch][8];
          const int batch = min(others, kFixBatch);
#pragma unroll
          for (int o = 0; o < kFixBatch; ++o) {
            if (o < batch) {
              const float4* src = chunk_of(cluster_id + (o + 1) * contrib_stride, c);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i * 4 < width) ld[o][i] = __ldcg(src + i * kF4Stride);
            }
          }
          float v[32];
          if (width == 32) {
            uint32_t r[32];
            tmem_ld_x32(t_row + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          } else {
            uint32_t r[16];
            tmem_ld_x16(t_row + c, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]), v[16 + i] = 0.f;
          }
          if (c + 32 >= c_end) release_acc();
#pragma unroll
          for (int o = 0; o < kFixBatch; ++o) {
            if (o < batch) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i * 4 < width) v[4 * i] += ld[o][i].x, v[4 * i + 1] += ld[o][i].y, v[4 * i + 2] += ld[o][i].z, v[4 * i + 3] += ld[o][i].w;
            }
          }
          for (int o = kFixBatch; o < others; ++o) {
            const float4* src = chunk_of(cluster_id + (o + 1) * contrib_stride, c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i * 4 < width) {
                const float4 t = __ldcg(src + i * kF4Stride);
                v[4 * i] += t.x, v[4 * i + 1] += t.y, v[4 * i + 2] += t.z, v[4 * i + 3] += t.w;
              }
            }
          }
          finish_chunk(v, c, width);
          if (epi_leader && c == c_begin) GTRACE(10);  // owner: first chunk reduced and stored
        }
      }
      if (epi_leader) GTRACE(6);  // epilogue of this item done (stores issued)
      if (++acc == 2) acc = 0, acc_phase ^= 1;
    }
  }

  // teardown: nobody may leave while the peer can still read this CTA's smem / TMEM
  __syncwarp();
  tc_fence_before_sync();
  if constexpr (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (threadIdx.x == 0) peer_sync_exit(p.sync, sync_epoch);
  if (threadIdx.x == 0) GTRACE(7);  // teardown sync passed
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<kCtas>(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 32) GTRACE(8);  // TMEM released
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TileChoice {
  int block_n;
  int ctas;
  double cost;
};

// Wave-quantisation model: every SM (pair) runs `waves` tiles back to back; a tile costs ~BN MMA
// columns plus a fixed prologue/drain. Pick the (BN, ctas) with the fewest total column-steps.
static TileChoice choose_tile(int M, int N, int sms, int want_bn, int want_pair) {
  TileChoice best{0, 0, 1e30};
  for (int ctas = 1; ctas <= 2; ++ctas) {
    if (want_pair == 0 && ctas == 2) continue;
    if (want_pair == 1 && ctas == 1) continue;
    if (want_pair < 0 && ctas == 2 && M <= kBlockM) continue;  // a pair would be half idle
    const int slots = sms / ctas;
    const int m_tiles = (M + kBlockM * ctas - 1) / (kBlockM * ctas);
    for (int bn = 256; bn >= 32; bn -= 16) {
      if (want_bn > 0 && bn != want_bn) continue;
      if (want_bn <= 0 && bn < 128 && N >= 128) continue;  // narrow tiles starve on smem/L2 bandwidth
      if (bn > ((N + 15) / 16) * 16 && bn != 32) continue;
      const int n_tiles = (N + bn - 1) / bn;
      const long long tiles = 1ll * m_tiles * n_tiles;
      const long long waves = (tiles + slots - 1) / slots;
      // measured (scripts/gemm_sweep.py, profiles/): a tile's main-loop time barely depends on BN between 128 and
      // 256 — the SS-mode operand fetch paces the MMA — so a narrower tile only pays when it removes a whole wave
      double cost = static_cast<double>(waves) * (std::max(bn, 224) + 12.0) - 0.01 * bn;
      if (ctas == 1) cost *= 1.03;  // pairs halve the W traffic per SM: prefer them on ties
      if (cost < best.cost) best = TileChoice{bn, ctas, cost};
    }
  }
  return best;
}

// ---- stream-K workspace: registered by the host framework (the library never allocates) -----------
constexpr int kMaxDevices = 16;
constexpr long long kSkCounterInts = 1 << 16;                                  // arrivals: [tiles][ctas]
constexpr long long kSkPartialBytes = 296ll * kBlockM * 256 * sizeof(float);  // [<=148 CTAs][128 x 256] fp32 parked partials (big tiles); [<=296 CTAs][128 x <=256] or [<=148][128 x <=512] (small-M kernel)
struct SkWorkspace {
  int* counters = nullptr;
  float* partials = nullptr;
};
static SkWorkspace g_sk_ws[kMaxDevices];

struct SkChoice {
  bool use;
  int block_n, ctas, rem, splits;
};
// Split the ragged wave in K when that beats running it as a whole wave.  Costs in k-blocks of main loop per
// cluster; parking + adding a partial costs the owner roughly `kFixup` k-blocks per contributor (hidden behind
// its next tile when one follows).
static SkChoice choose_split_k(int M, int N, int K, int sms, int want_bn, int want_pair, const TileChoice& dp) {
  SkChoice none{false, 0, 0, 0, 0};
  if (want_bn > 0) return none;  // an explicit tile request means "run exactly this" (tests, sweeps)
  const int ctas = (want_pair == 0 || (want_pair < 0 && M <= kBlockM)) ? 1 : 2;
  const int bn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
  if (bn < 32) return none;
  const int slots = sms / ctas;
  const int num_kb = K / kBlockK;
  const long long tiles = 1ll * ((M + kBlockM * ctas - 1) / (kBlockM * ctas)) * ((N + bn - 1) / bn);
  const long long waves = tiles / slots, rem = tiles % slots;
  if (rem == 0 || rem * ctas > kSkCounterInts) return none;
  static const int env_max = [] { const char* e = getenv("LTXB_SK_MAX_SPLITS"); return e ? atoi(e) : 8; }();
  static const double env_fixup = [] { const char* e = getenv("LTXB_SK_FIXUP"); return e ? atof(e) : 6.0; }();
  int splits = static_cast<int>(std::min<long long>(slots / rem, env_max));
  splits = std::min(splits, num_kb / 16);  // pieces of at least 16 k-blocks (measured: 8 is past the optimum)
  if (splits < 2) return none;
  const double fixup = env_fixup * (splits - 1) * (waves >= 1 ? 0.5 : 1.0);
  const double t_sk = static_cast<double>(waves) * num_kb + static_cast<double>((num_kb + splits - 1) / splits) + fixup;
  const int dp_slots = sms / dp.ctas;
  const long long dp_tiles = 1ll * ((M + kBlockM * dp.ctas - 1) / (kBlockM * dp.ctas)) * ((N + dp.block_n - 1) / dp.block_n);
  const double t_dp = static_cast<double>((dp_tiles + dp_slots - 1) / dp_slots) * num_kb;
  if (t_sk * 1.03 >= t_dp) return none;
  return SkChoice{true, bn, ctas, static_cast<int>(rem), splits};
}

template <int kCtas, int kEpi>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int grid, size_t smem,
                       cudaStream_t stream) {
  auto kernel = gemm_bf16_kernel<kCtas, kEpi>;
  static PerDeviceOnce configured;  // per instantiation and device
  if (configured.first()) {
    LTXB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  LTXB_CUDA(launch_kernel(kernel, dim3(grid), dim3(kGemmThreads), smem, stream, kCtas, ta, tw, p));
  return LTXB_OK;
}

template <int kCtas>
static int dispatch_epi(int mode, const CUtensorMap& ta, const CUtensorMap& tw, const GemmParams& p, int grid,
                        size_t smem, cudaStream_t stream) {
  switch (mode) {
    case LTXB_EPI_BIAS_BF16: return launch_gemm<kCtas, LTXB_EPI_BIAS_BF16>(ta, tw, p, grid, smem, stream);
    case LTXB_EPI_GELU_BF16: return launch_gemm<kCtas, LTXB_EPI_GELU_BF16>(ta, tw, p, grid, smem, stream);
    case LTXB_EPI_SILU_BF16: return launch_gemm<kCtas, LTXB_EPI_SILU_BF16>(ta, tw, p, grid, smem, stream);
    case LTXB_EPI_BIAS_F32: return launch_gemm<kCtas, LTXB_EPI_BIAS_F32>(ta, tw, p, grid, smem, stream);
    case LTXB_EPI_RESID_GATE_F32: return launch_gemm<kCtas, LTXB_EPI_RESID_GATE_F32>(ta, tw, p, grid, smem, stream);
    default: return set_error(LTXB_ERR_BAD_ARG, "ltxb_gemm_bf16: unknown epilogue mode %d", mode);
  }
}

}  // namespace ltxb

using namespace ltxb;

extern "C" int ltxb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                              int32_t M, int32_t N, int32_t K, const ltxb_epilogue* epi, int32_t block_n,
                              int32_t cta_pair, void* stream) {
  LTXB_CHECK_ARG(A && W && out && epi, "ltxb_gemm_bf16: null pointer");
  LTXB_CHECK_ARG(M > 0 && N > 0 && K > 0, "ltxb_gemm_bf16: non-positive shape M=%d N=%d K=%d", M, N, K);
  LTXB_CHECK_ARG(aligned16(A) && aligned16(W) && aligned16(out), "ltxb_gemm_bf16: pointers must be 16-byte aligned");
  LTXB_CHECK_SUPPORTED(K % kBlockK == 0, "ltxb_gemm_bf16: K=%d must be a multiple of %d", K, kBlockK);
  LTXB_CHECK_SUPPORTED(N % 16 == 0, "ltxb_gemm_bf16: N=%d must be a multiple of 16", N);
  LTXB_CHECK_SUPPORTED(lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && (lda >= K || epi->a_group_cols > 0) && ldw >= K && ldo >= N,
                       "ltxb_gemm_bf16: leading dimensions must be multiples of 8 and cover the row");
  LTXB_CHECK_ARG(epi->mode >= 0 && epi->mode < LTXB_EPI_COUNT, "ltxb_gemm_bf16: bad epilogue mode %d", epi->mode);
  LTXB_CHECK_ARG(cta_pair >= -1 && cta_pair <= 4, "ltxb_gemm_bf16: cta_pair=%d not in [-1, 4]", cta_pair);
  LTXB_CHECK_ARG(cta_pair == 4 || block_n == 0 || (block_n >= 32 && block_n <= 256 && block_n % 16 == 0),
                 "ltxb_gemm_bf16: block_n=%d must be 0 or a multiple of 16 in [32,256]", block_n);
  if (epi->bias) LTXB_CHECK_ARG(aligned16(epi->bias), "ltxb_gemm_bf16: bias must be 16-byte aligned");
  if (epi->mode == LTXB_EPI_RESID_GATE_F32) {
    LTXB_CHECK_ARG(epi->resid && aligned16(epi->resid) && epi->ldr % 4 == 0 && epi->ldr >= N,
                   "ltxb_gemm_bf16: residual epilogue needs a 16-byte aligned f32 resid with ldr %% 4 == 0");
    if (epi->gate) {
      LTXB_CHECK_ARG(aligned16(epi->gate) && epi->gate_ld % 4 == 0, "ltxb_gemm_bf16: gate must be 16-byte aligned");
      LTXB_CHECK_ARG(epi->gate_row_index || epi->gate_row_div >= 1, "ltxb_gemm_bf16: gate_row_div must be >= 1");
      if (epi->gate_table) LTXB_CHECK_ARG(aligned16(epi->gate_table), "ltxb_gemm_bf16: gate_table misaligned");
    }
  }
  if (epi->mode == LTXB_EPI_BIAS_F32 || epi->mode == LTXB_EPI_RESID_GATE_F32)
    LTXB_CHECK_SUPPORTED(ldo % 4 == 0, "ltxb_gemm_bf16: f32 output needs ldo %% 4 == 0");

  const int sms = num_sms();
  if (sms <= 0) return set_error(LTXB_ERR_NO_DEVICE, "ltxb_gemm_bf16: no CUDA device");
  static const int env_pair = [] { const char* e = getenv("LTXB_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  static const int env_bn = [] { const char* e = getenv("LTXB_GEMM_BN"); return e ? atoi(e) : 0; }();
  // Few rows (a sequence-parallel shard, the audio stream, AdaLN rows): the weight-streaming kernel with swapped operands
  // (gemm_small_m.cu).  cta_pair 4 forces it (block_n then = number of k-range splits, 0 = choose); LTXB_GEMM_SMALL_M is the
  // largest M the library sends there on its own (0 disables).
  {
    // measured (profiles/r2/gemm_small_m.md): at 160 rows the few-row kernel wins everywhere, at 320 rows (two MMAs per
    // k-slice) on three of the four LTX-2 shapes (-20 %) and loses 4 % on the fourth, beyond 512 it does not apply
    static const int env_small_m = [] { const char* e = getenv("LTXB_GEMM_SMALL_M"); return e ? atoi(e) : 512; }();
    const bool forced = (cta_pair == 4);
    if (forced) LTXB_CHECK_SUPPORTED(gemm_small_m_supported(M, N, K), "ltxb_gemm_bf16: the small-M kernel needs M <= 512 (M=%d N=%d K=%d)", M, N, K);
    if (forced || (cta_pair < 0 && block_n == 0 && M <= env_small_m && N >= 128 && K >= 256 && gemm_small_m_supported(M, N, K) && env_pair < 0 && env_bn == 0)) {
      int dev = 0;
      cudaGetDevice(&dev);
      const SkWorkspace ws = (dev >= 0 && dev < kMaxDevices) ? g_sk_ws[dev] : SkWorkspace{};
      return launch_gemm_small_m(A, lda, W, ldw, out, ldo, M, N, K, epi, ws.partials, ws.partials ? kSkPartialBytes : 0, ws.counters,
                                 forced ? block_n : 0, reinterpret_cast<cudaStream_t>(stream));
    }
  }
  if (cta_pair < 0) cta_pair = env_pair;
  if (block_n == 0) block_n = env_bn;
  // cta_pair 2 / 3: force the contiguous stream-K schedule with single-CTA / pair tiles (tests, sweeps)
  int force_contig = 0;
  if (cta_pair == 2 || cta_pair == 3) {
    force_contig = 1;
    cta_pair -= 2;
  }
  const TileChoice tc = choose_tile(M, N, sms, block_n, cta_pair);
  if (tc.block_n == 0) return set_error(LTXB_ERR_UNSUPPORTED, "ltxb_gemm_bf16: no tile for M=%d N=%d", M, N);
  static const int env_sk = [] { const char* e = getenv("LTXB_GEMM_STREAMK"); return e ? atoi(e) : 1; }();
  // measured (profiles/r2/gemm_small_m.md): at M = 160 the contiguous schedule is 2-8 us SLOWER per launch than the lockstep
  // split (every cluster pays a contributor park AND an owner fix-up), so it is opt-in: LTXB_GEMM_SK_CONTIG=-1 lets the
  // library pick it for M <= 640, 1 forces it, cta_pair 2 / 3 select it per call
  static const int env_contig = [] { const char* e = getenv("LTXB_GEMM_SK_CONTIG"); return e ? atoi(e) : 0; }();
  int dev = 0;
  cudaGetDevice(&dev);
  const SkWorkspace ws = (dev >= 0 && dev < kMaxDevices) ? g_sk_ws[dev] : SkWorkspace{};
  SkChoice sk{false, 0, 0, 0, 0};
  // Contiguous stream-K for SMALL M (the weight-streaming regime of a sequence-parallel shard: M = 160 rows per rank at
  // 1280 tokens on 8 GPUs): one m-tile, so no tile shares a W panel with another and the only thing that matters is that
  // EVERY SM streams an equal slice of W.  The unit list (tile-major, k-minor) is cut into one contiguous range per
  // cluster; at most the first item of a range is a contributor piece and the last an owner piece (see the kernel).
  bool contig = false;
  int c_ctas = 0, c_bn = 0;
  if (ws.partials != nullptr && env_sk && (force_contig || (env_contig != 0 && block_n == 0 && cta_pair < 0))) {
    c_ctas = force_contig ? (cta_pair == 1 ? 2 : 1) : (M <= kBlockM ? 1 : 2);
    c_bn = block_n > 0 ? block_n : (N >= 256 ? 256 : ((N + 15) / 16) * 16);
    const int slots = sms / c_ctas;
    const long long tiles = 1ll * ((M + kBlockM * c_ctas - 1) / (kBlockM * c_ctas)) * ((N + c_bn - 1) / c_bn);
    const long long waves = (tiles + slots - 1) / slots;
    const int num_kb = K / kBlockK;
    const double fill = static_cast<double>(tiles) / static_cast<double>(waves * slots);
    const bool small_m = M <= 2 * kBlockM * 2 + kBlockM;  // <= 640 rows: at most 3 pair m-tiles
    contig = force_contig || env_contig == 1 || (small_m && fill < 0.93 && num_kb >= 16 && tiles * c_ctas <= kSkCounterInts);
    if (c_bn < 32 || (c_ctas == 2 && (c_bn % 16 != 0 || (c_bn / 2) % 8 != 0)) || tiles * num_kb < 8) contig = false;
  }
  if (!contig && env_sk && ws.partials != nullptr) sk = choose_split_k(M, N, K, sms, block_n, cta_pair, tc);
  const int ctas = contig ? c_ctas : (sk.use ? sk.ctas : tc.ctas);
  const int bn = contig ? c_bn : (sk.use ? sk.block_n : tc.block_n);
  if (ctas == 2) LTXB_CHECK_SUPPORTED(bn % 16 == 0 && (bn / 2) % 8 == 0, "pair mode needs block_n %% 16 == 0");

  GemmParams p{};
  p.M = M, p.N = N, p.K = K;
  p.block_n = bn;
  p.num_m_tiles = (M + kBlockM * ctas - 1) / (kBlockM * ctas);
  p.num_n_tiles = (N + bn - 1) / bn;
  const size_t stage_bytes = static_cast<size_t>(kBlockM + bn / ctas) * kBlockK * 2;
  const size_t budget = 232448 - 1024 /*alignment slack*/ - kSmemHeader;
  p.num_stages = static_cast<int>(std::min<size_t>(kMaxStages, budget / stage_bytes));
  const size_t smem = 1024 + kSmemHeader + p.num_stages * stage_bytes;
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
  if (int rc = peer_sync_from_abi(epi->peer_sync, &p.sync, "ltxb_gemm_bf16")) return rc;

  CUtensorMap ta, tw;
  if (epi->a_group_cols > 0) {
    // A arrives head-group-major from the Ulysses gather all-to-all: [K / gc][M][gc]; read it in place
    LTXB_CHECK_SUPPORTED(epi->a_group_cols % kBlockK == 0 && K % epi->a_group_cols == 0 && epi->a_group_stride % 8 == 0,
                         "ltxb_gemm_bf16: a_group_cols=%d must divide K=%d and be a multiple of %d", epi->a_group_cols, K, kBlockK);
    p.a_group_cols = epi->a_group_cols;
    const uint64_t dims[3] = {static_cast<uint64_t>(epi->a_group_cols), static_cast<uint64_t>(M), static_cast<uint64_t>(K / epi->a_group_cols)};
    const uint64_t strides[2] = {static_cast<uint64_t>(lda) * 2, static_cast<uint64_t>(epi->a_group_stride) * 2};
    const uint32_t box[3] = {kBlockK, kBlockM, 1};
    int rc = encode_tmap_bf16(&ta, A, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {kBlockK, kBlockM};
    int rc = encode_tmap_bf16(&ta, A, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {kBlockK, static_cast<uint32_t>(bn / ctas)};
    int rc = encode_tmap_bf16(&tw, W, 2, dims, strides, box);
    if (rc) return rc;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  int clusters = std::min(num_tiles, sms / ctas);
  if (contig) {
    p.sk_contig = 1;
    p.sk_gran = (K / kBlockK) % 4 == 0 ? 4 : 1;  // range boundaries on multiples of 4 k-blocks: no 1-block pieces
    // every cluster must own a NON-EMPTY range (the owner of a cut tile counts the clusters that follow it as its
    // contributors): at least sk_gran units each
    const long long units = static_cast<long long>(num_tiles) * (K / kBlockK);
    clusters = static_cast<int>(std::min<long long>(sms / ctas, units / p.sk_gran));
    p.sk_partials = ws.partials;
    p.sk_counters = ws.counters;
  } else if (sk.use) {
    clusters = std::min(sms / ctas, std::max(num_tiles, sk.rem * sk.splits));
    p.sk_rem = sk.rem;
    p.sk_splits = sk.splits;
    p.sk_partials = ws.partials;
    p.sk_counters = ws.counters;
  }
  const int grid = clusters * ctas;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ctas == 2) return dispatch_epi<2>(epi->mode, ta, tw, p, grid, smem, s);
  return dispatch_epi<1>(epi->mode, ta, tw, p, grid, smem, s);
}

extern "C" int ltxb_gemm_qw_bf16(const void* A, int64_t lda, const uint32_t* Wq, int64_t ldq, const void* scales, const void* biases,
                                 int64_t lds, int32_t aux_f32, int32_t group_size, int32_t bits, void* out, int64_t ldo, int32_t M,
                                 int32_t N, int32_t K, const ltxb_epilogue* epi, int32_t splits, void* stream) {
  LTXB_CHECK_ARG(A && Wq && scales && biases && out && epi, "ltxb_gemm_qw_bf16: null pointer");
  LTXB_CHECK_ARG(M > 0 && N > 0 && K > 0, "ltxb_gemm_qw_bf16: non-positive shape M=%d N=%d K=%d", M, N, K);
  LTXB_CHECK_SUPPORTED(bits == 4 || bits == 8, "ltxb_gemm_qw_bf16: bits=%d (4 and 8 are built; expand other widths with ltxb_dequant_affine_bf16)", bits);
  LTXB_CHECK_SUPPORTED(group_size == 32 || group_size == 64 || group_size == 128, "ltxb_gemm_qw_bf16: group_size=%d (32, 64 and 128 are built)", group_size);
  LTXB_CHECK_SUPPORTED(gemm_small_m_supported(M, N, K), "ltxb_gemm_qw_bf16: the packed-weight kernel is the few-row one: M=%d must be <= 512 and K=%d a multiple of 64 (expand the weights with ltxb_dequant_affine_bf16 for more rows)", M, K);
  LTXB_CHECK_SUPPORTED(N % 16 == 0 && K % group_size == 0, "ltxb_gemm_qw_bf16: N=%d must be a multiple of 16 and K=%d of the group size %d", N, K, group_size);
  LTXB_CHECK_ARG(aligned16(A) && aligned16(Wq) && aligned16(out), "ltxb_gemm_qw_bf16: pointers must be 16-byte aligned");
  LTXB_CHECK_SUPPORTED(lda % 8 == 0 && ldo % 8 == 0 && (lda >= K || epi->a_group_cols > 0) && ldo >= N && ldq % 4 == 0 &&
                           ldq >= static_cast<int64_t>(K) * bits / 32 && lds >= K / group_size,
                       "ltxb_gemm_qw_bf16: leading dimensions must cover the row (lda, ldo multiples of 8, ldq of 4 words)");
  LTXB_CHECK_ARG(epi->mode >= 0 && epi->mode < LTXB_EPI_COUNT, "ltxb_gemm_qw_bf16: bad epilogue mode %d", epi->mode);
  LTXB_CHECK_ARG(splits >= 0, "ltxb_gemm_qw_bf16: splits=%d", splits);
  if (epi->bias) LTXB_CHECK_ARG(aligned16(epi->bias), "ltxb_gemm_qw_bf16: bias must be 16-byte aligned");
  if (epi->mode == LTXB_EPI_RESID_GATE_F32) {
    LTXB_CHECK_ARG(epi->resid && aligned16(epi->resid) && epi->ldr % 4 == 0 && epi->ldr >= N,
                   "ltxb_gemm_qw_bf16: residual epilogue needs a 16-byte aligned f32 resid with ldr %% 4 == 0");
    if (epi->gate) LTXB_CHECK_ARG(epi->gate_row_index || epi->gate_row_div >= 1, "ltxb_gemm_qw_bf16: gate_row_div must be >= 1");
  }
  if (epi->mode == LTXB_EPI_BIAS_F32 || epi->mode == LTXB_EPI_RESID_GATE_F32)
    LTXB_CHECK_SUPPORTED(ldo % 4 == 0, "ltxb_gemm_qw_bf16: f32 output needs ldo %% 4 == 0");
  int dev = 0;
  cudaGetDevice(&dev);
  const SkWorkspace ws = (dev >= 0 && dev < kMaxDevices) ? g_sk_ws[dev] : SkWorkspace{};
  const WsPacked packed{scales, biases, lds, group_size, bits, aux_f32 != 0};
  return launch_gemm_small_m(A, lda, Wq, ldq, out, ldo, M, N, K, epi, ws.partials, ws.partials ? kSkPartialBytes : 0, ws.counters, splits,
                             reinterpret_cast<cudaStream_t>(stream), &packed);
}

extern "C" int64_t ltxb_gemm_workspace_bytes(void) { return kSkCounterInts * sizeof(int) + kSkPartialBytes; }

extern "C" int ltxb_gemm_set_workspace(void* workspace, int64_t bytes, void* stream) {
  int dev = 0;
  LTXB_CUDA(cudaGetDevice(&dev));
  LTXB_CHECK_SUPPORTED(dev >= 0 && dev < kMaxDevices, "ltxb_gemm_set_workspace: device %d out of range", dev);
  if (workspace == nullptr) {
    g_sk_ws[dev] = SkWorkspace{};
    return LTXB_OK;
  }
  LTXB_CHECK_ARG(aligned16(workspace) && bytes >= ltxb_gemm_workspace_bytes(),
                 "ltxb_gemm_set_workspace: need a 16-byte aligned buffer of at least %lld bytes",
                 static_cast<long long>(ltxb_gemm_workspace_bytes()));
  LTXB_CUDA(cudaMemsetAsync(workspace, 0, kSkCounterInts * sizeof(int), reinterpret_cast<cudaStream_t>(stream)));
  g_sk_ws[dev].counters = reinterpret_cast<int*>(workspace);
  g_sk_ws[dev].partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kSkCounterInts * sizeof(int));
  return LTXB_OK;
}
