// Microbenchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256 over a CTA pair, K = 16) against N — the MMA the
// few-row GEMM issues (weight rows on the lanes, N = tokens).  One thread of the leader CTA issues `n` MMAs back to back
// (A and B from shared memory, 128-byte swizzle, operands are whatever the memory holds), commits to both CTAs and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mlx-video_b200/csrc scripts/umma_pair_bench.cu -o scripts/umma_pair_bench.bin
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

using namespace ltxb;

template <int N_COLS, int CTAS>
__global__ void __launch_bounds__(128, 1) umma_pair_kernel(long long* out, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<CTAS>(&tmem_slot, 512);
  tc_fence_before_sync();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  // A: 128 rows x 64 columns bf16 (one 16 KB swizzle block) at smem + 0; B: (N / CTAS) rows x 64 columns at smem + 16 KB
  const uint32_t a_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_lo = ((smem_u32(smem + 16384) & 0x3FFFFu) >> 4) | (1u << 16);
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t idesc = make_idesc_bf16(128 * CTAS, N_COLS, 0, 0);
  if (warp == 0) {
    const bool issuer = elect_one() && rank == 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      if (issuer) {
        for (int o = 0; o < n_mma / 4; ++o) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // the four k-slices of one 64-column block, like one k-block of the GEMM
            umma_bf16_ss<CTAS>(tmem_base, desc_from_words(a_lo + 2 * k, kHi), desc_from_words(b_lo + 2 * k, kHi), idesc, (o | k) ? 1u : 0u);
        }
        if (CTAS == 2) umma_commit_pair(&bar, 3); else umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (rank == 0 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[rep] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<CTAS>(tmem_base, 512);
  }
}

template <int N_COLS, int CTAS>
static void run(long long* out) {
  const int n_mma = 512;
  const size_t smem = 1024 + 16384 + 32768 + 1024;
  auto k = umma_pair_kernel<N_COLS, CTAS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148 / CTAS * CTAS);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, k, out, n_mma);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(e));
    exit(1);
  }
  printf("cta_group::%d  M=%3d N=%3d: %.1f cycles per MMA (math floor: %d)\n", CTAS, 128 * CTAS, N_COLS, double(out[2]) / n_mma, N_COLS / 2);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  run<16, 2>(out);
  run<32, 2>(out);
  run<64, 2>(out);
  run<80, 2>(out);
  run<96, 2>(out);
  run<128, 2>(out);
  run<160, 2>(out);
  run<192, 2>(out);
  run<256, 2>(out);
  run<64, 1>(out);
  run<80, 1>(out);
  run<128, 1>(out);
  run<160, 1>(out);
  run<256, 1>(out);
  return 0;
}
