// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) for the operand shapes the attention
// kernel issues — N = 64 / 128 / 256, A from shared memory (SS) or from TMEM (TS), one accumulator chain or two
// interleaved ones.  One elected thread issues `n` MMAs back to back, commits to an mbarrier and waits; the clock brackets
// issue + completion.  Operands are whatever the shared memory holds (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mlx-video_b200/csrc scripts/umma_bench.cu -o /tmp/umma_bench
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

using namespace ltxb;

struct Case {
  int n_cols;      // MMA N
  int a_tmem;      // 1: A operand from TMEM
  int chains;      // independent accumulators used round-robin (1, 2, 4)
  int a_slices;    // distinct A k-slices cycled through (8 = like Q with dh 128; 1 = same slice every time)
  int b_slices;    // distinct B k-slices
  int n_mma;       // MMAs issued
  int d_stride;    // TMEM column stride between the accumulators of different chains
};

template <int N_COLS, int A_TMEM, int CHAINS, int A_SLICES, int D_STRIDE>
__global__ void __launch_bounds__(384, 1) umma_bench_kernel(Case c, long long* out, int noise) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  if (warp == 0) tmem_alloc<1>(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  // A: 128 rows x 128 columns bf16 = two 16 KB swizzle blocks at smem + 0; B: up to 256 rows x 128 columns at smem + 32 KB
  const uint32_t a_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo = ((smem_u32(smem + 32768) & 0x3FFFFu) >> 4) | (1u << 16);
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t idesc = make_idesc_bf16(128, N_COLS, 0, 0);
  if (warp == 0) {
    const bool issuer = elect_one();
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      if (issuer) {
        for (int o = 0; o < c.n_mma / 32; ++o) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {  // compile-time offsets: the issue loop must not be instruction-bound
            const int ka = i % A_SLICES, kb = i % 8;
            const uint32_t aoff = (ka >> 2) * 1024 + (ka & 3) * 2;
            const uint32_t boff = (kb >> 2) * 2048 + (kb & 3) * 2;  // B blocks are 256 rows x 128 B = 32 KB apart
            const uint32_t d = tmem_base + (i % CHAINS) * D_STRIDE;
            if (A_TMEM)
              umma_bf16_ts(d, tmem_base + 448 + (ka & 7) * 8, desc_from_words(b_lo + boff, kHi), idesc, (o | (i >= CHAINS)) ? 1u : 0u);
            else
              umma_bf16_ss<1>(d, desc_from_words(a_lo + aoff, kHi), desc_from_words(b_lo + boff, kHi), idesc, (o | (i >= CHAINS)) ? 1u : 0u);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      const long long t1 = clock64();  // all MMAs issued
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();  // all MMAs complete
      if (issuer && blockIdx.x == 0) {
        out[rep * 2] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
    stop = 1;
  } else if (warp >= 4 && noise != 0) {
    // "softmax-like" background on the other eight warps (two per scheduler, like the attention kernel):
    // noise 1 = ALU / MUFU work only, noise 2 = + TMEM loads and stores of a 64-column score tile per round,
    // noise 3 = TMEM traffic only
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >= 8) ? 384u : 320u);
    float acc = static_cast<float>(threadIdx.x);
    while (!stop) {
      uint32_t r[32];
      if (noise >= 2) {
        tmem_ld_x32(t_lane, r);
        tmem_wait_ld();
        tmem_ld_x32(t_lane + 32, r);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(acc + i);
      }
      if (noise != 3) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(r[i]), 0.001f, -1.0f)));
      }
      if (noise >= 2) {
        tmem_st_x32(t_lane, r);
        tmem_wait_st();
      }
      acc += __uint_as_float(r[7]);
    }
    if (acc == 12345.678f) out[7] = 1;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int N_COLS, int A_TMEM, int CHAINS, int A_SLICES, int D_STRIDE>
static void run(long long* out) {
  const Case c{N_COLS, A_TMEM, CHAINS, A_SLICES, 8, 256, D_STRIDE};
  const size_t smem = 1024 + 32768 + 65536 + 1024;
  auto k = umma_bench_kernel<N_COLS, A_TMEM, CHAINS, A_SLICES, D_STRIDE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  for (int noise : {0, 1, 2, 3}) {
    k<<<148, 384, smem>>>(c, out, noise);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("error: %s\n", cudaGetErrorString(e));
      exit(1);
    }
    printf("noise %d  N=%3d %s chains=%d a_slices=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %d)\n", noise, N_COLS,
           A_TMEM ? "TS" : "SS", CHAINS, A_SLICES, double(out[4]) / c.n_mma, double(out[5]) / c.n_mma, N_COLS / 2);
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  run<64, 0, 1, 8, 64>(out);    // S of the 64-key kernel: one chain
  run<64, 0, 2, 8, 64>(out);    // two interleaved chains
  run<128, 0, 1, 8, 128>(out);  // S of the 128-key kernel
  run<256, 0, 1, 8, 256>(out);  // GEMM-like
  run<128, 1, 1, 8, 128>(out);  // PV: A from TMEM, B = V
  run<64, 1, 1, 8, 64>(out);
  return 0;
}
