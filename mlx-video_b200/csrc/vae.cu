// Row N4 (second half): the LTX-2 video VAE decoder that follows the last denoise loop
// (mlx_video/models/ltx/video_vae/decoder.py:94-450, convolution.py:13-166, sampling.py:106-197, ops.py:47-80,
// tiling.py:279-520).  As in the latent upsampler (conv.cu) every 3x3x3 convolution is ONE tcgen05 GEMM over bf16 rows
// gathered from channels-last fp32 activations [N, D, H, W, C]; what differs is the operand fetch:
//   * temporal padding by frame REPLICATION (two copies of the first frame when causal, first + last otherwise) and
//     REFLECT padding in H / W (convolution.py:120-166) instead of zeros;
//   * the elementwise chain in front of every ResNet convolution — pixel norm over the channels, AdaLN
//     (1 + scale) x + shift from the block's scale-shift table + timestep embedding, SiLU (decoder.py:140-180) — is applied
//     while the taps are gathered (a warp holds the whole channel run of a tap, so the norm is one warp reduction), so no
//     normalised / activated copy of the activations is ever written;
//   * rows are produced in chunks [m0, m0 + rows) so the materialised operand stays bounded at the 128-channel levels.
// Around it: depth-to-space upsampling with its tiled-channel residual and dropped first frame, latent de-normalisation
// with the decode noise, un-patchify into channels-first video, and the trapezoid blend of tiled decoding.
// All memory-bound; algorithmic bytes are stated per entry point in include/ltxb.h.
#include "common.cuh"
#include "ptx.cuh"

namespace ltxb {

static int vae_grid_for(long long work_items, int threads) {
  const long long blocks = (work_items + threads - 1) / threads;
  const long long cap = 148ll * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

__device__ __forceinline__ float vae_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct VaePreOp {
  const float* table_scale;  // [C] rows of the block's scale_shift_table (or null: no modulation)
  const float* table_shift;
  const float* emb_scale;    // [N, emb_ld] timestep-embedding slices added to the table rows (or null)
  const float* emb_shift;
  long long emb_ld;
  float eps;
  int enabled;               // 0: plain gather (conv_in, the upsampler convolutions)
};

// out[m - m0, ((kz*3 + ky)*3 + kx)*C + c] = f(x[n, td(d, kz), rh(h + ky - 1), rw(w + kx - 1), c]),  m = ((n*D + d)*H + h)*W + w
//   td: causal -> max(d + kz - 2, 0); else clamp(d + kz - 1, 0, D - 1)         (frame replication)
//   rh / rw: reflection without repeating the border pixel (-1 -> 1, H -> H - 2)
//   f: identity, or silu(pixel_norm(x) * (1 + scale[c]) + shift[c])
// One warp per output position; every tap is a contiguous C-channel run (kMaxPerLane * 32 >= C / 8 eight-float groups).
template <int kGroupsPerLane>
__global__ void __launch_bounds__(256)
vae_gather_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int D, int H, int W, int C, int causal,
                  long long m0, long long rows, const VaePreOp pre) {
  pdl_launch_dependents();
  pdl_wait();
  const int c8 = C / 8;
  const int lane = threadIdx.x & 31;
  for (long long r = blockIdx.x * 8ll + (threadIdx.x >> 5); r < rows; r += 8ll * gridDim.x) {
    const long long m = m0 + r;
    const int w = static_cast<int>(m % W);
    long long q = m / W;
    const int h = static_cast<int>(q % H);
    q /= H;
    const int d = static_cast<int>(q % D);
    const long long n = q / D;
    __nv_bfloat16* orow = out + r * 27ll * C;
    // modulation of this sample (same for all 27 taps): (1 + scale), shift per channel group held by this lane
    float sc[kGroupsPerLane][8], sh[kGroupsPerLane][8];
    if (pre.enabled && pre.table_scale != nullptr) {
#pragma unroll
      for (int g = 0; g < kGroupsPerLane; ++g) {
        const int cc = lane + g * 32;
        if (cc < c8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = cc * 8 + i;
            float s = pre.table_scale[c], t = pre.table_shift[c];
            if (pre.emb_scale != nullptr) s += pre.emb_scale[n * pre.emb_ld + c], t += pre.emb_shift[n * pre.emb_ld + c];
            sc[g][i] = 1.0f + s, sh[g][i] = t;
          }
        }
      }
    }
    int tap = 0;
    for (int kz = 0; kz < 3; ++kz) {
      const int sd = causal ? max(d + kz - 2, 0) : min(max(d + kz - 1, 0), D - 1);
      for (int ky = 0; ky < 3; ++ky) {
        int sy = h + ky - 1;
        sy = sy < 0 ? -sy : (sy >= H ? 2 * (H - 1) - sy : sy);
        for (int kx = 0; kx < 3; ++kx, ++tap) {
          int sx = w + kx - 1;
          sx = sx < 0 ? -sx : (sx >= W ? 2 * (W - 1) - sx : sx);
          const float* src = x + (((n * D + sd) * H + sy) * static_cast<long long>(W) + sx) * C;
          __nv_bfloat16* dst = orow + static_cast<long long>(tap) * C;
          float v[kGroupsPerLane][8];
          float ss = 0.f;
#pragma unroll
          for (int g = 0; g < kGroupsPerLane; ++g) {
            const int cc = lane + g * 32;
            if (cc < c8) {
              const float4 a = *reinterpret_cast<const float4*>(src + cc * 8);
              const float4 b = *reinterpret_cast<const float4*>(src + cc * 8 + 4);
              v[g][0] = a.x, v[g][1] = a.y, v[g][2] = a.z, v[g][3] = a.w, v[g][4] = b.x, v[g][5] = b.y, v[g][6] = b.z, v[g][7] = b.w;
#pragma unroll
              for (int i = 0; i < 8; ++i) ss = fmaf(v[g][i], v[g][i], ss);
            }
          }
          if (pre.enabled) {
            ss = vae_warp_sum(ss);
            const float rstd = rsqrtf(ss / static_cast<float>(C) + pre.eps);
#pragma unroll
            for (int g = 0; g < kGroupsPerLane; ++g) {
              if (lane + g * 32 < c8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float y = v[g][i] * rstd;
                  if (pre.table_scale != nullptr) y = fmaf(y, sc[g][i], sh[g][i]);
                  v[g][i] = silu(y);
                }
              }
            }
          }
#pragma unroll
          for (int g = 0; g < kGroupsPerLane; ++g) {
            const int cc = lane + g * 32;
            if (cc < c8) {
              uint4 o;
              o.x = pack_bf16x2(v[g][0], v[g][1]), o.y = pack_bf16x2(v[g][2], v[g][3]);
              o.z = pack_bf16x2(v[g][4], v[g][5]), o.w = pack_bf16x2(v[g][6], v[g][7]);
              *reinterpret_cast<uint4*>(dst + cc * 8) = o;
            }
          }
        }
      }
    }
  }
}

// DepthToSpaceUpsample, stride (2,2,2), residual, first frame dropped (sampling.py:143-197):
//   out[n, 2d + st - 1, 2h + sh, 2w + sw, c] = y[n, d, h, w, ((c*2 + st)*2 + sh)*2 + sw] + x[n, d, h, w, (((c % (C/8))*2 + st)*2 + sh)*2 + sw]
// y = the convolution's output (4C channels), x = its input (C channels), out has C/2 channels and 2D - 1 frames.
__global__ void __launch_bounds__(256)
vae_depth_to_space_kernel(const float* __restrict__ y, const float* __restrict__ x, float* __restrict__ out, long long N, int D, int H,
                          int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Co = C / 2, Cr = C / 8;
  const int Do = 2 * D - 1, Ho = 2 * H, Wo = 2 * W;
  const long long total = N * Do * Ho * Wo * Co;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int c = static_cast<int>(i % Co);
    long long q = i / Co;
    const int wo = static_cast<int>(q % Wo);
    q /= Wo;
    const int ho = static_cast<int>(q % Ho);
    q /= Ho;
    const int fo = static_cast<int>(q % Do);
    const long long n = q / Do;
    const int f = fo + 1;  // frame of the un-dropped sequence
    const int d = f >> 1, st = f & 1, h = ho >> 1, sh = ho & 1, w = wo >> 1, sw = wo & 1;
    const long long pos = ((n * D + d) * H + h) * static_cast<long long>(W) + w;
    const int sub = (st * 2 + sh) * 2 + sw;
    out[i] = y[pos * (4ll * C) + c * 8 + sub] + x[pos * C + (c % Cr) * 8 + sub];
  }
}

// decoder.py:380-384: x = (noise * ns + (1 - ns) * sample) * std[c] + mean[c]; channels-first (N, C, S) -> channels-last (N, S, C)
__global__ void __launch_bounds__(256)
vae_prepare_latent_kernel(const float* __restrict__ sample, const float* __restrict__ noise, float noise_scale,
                          const float* __restrict__ stdv, const float* __restrict__ mean, float* __restrict__ out, long long N, int C,
                          long long S) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = N * S * C;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int c = static_cast<int>(i % C);
    const long long q = i / C;
    const long long s = q % S, n = q / S;
    const long long src = (n * C + c) * S + s;
    float v = sample[src];
    if (noise != nullptr) v = noise[src] * noise_scale + (1.0f - noise_scale) * v;
    else v = (1.0f - noise_scale) * v;
    out[i] = v * stdv[c] + mean[c];
  }
}

// ops.py:47-80 (patch 4, patch_size_t 1): video[n, c, f, 4h + pq, 4w + pr] = z[n, f, h, w, (c*4 + pr)*4 + pq]; z channels-last, 48 channels
__global__ void __launch_bounds__(256)
vae_unpatchify_kernel(const float* __restrict__ z, float* __restrict__ video, long long N, int F, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = 4 * H, Wo = 4 * W;
  const long long total = N * 3 * F * Ho * Wo;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int xo = static_cast<int>(i % Wo);
    long long q = i / Wo;
    const int yo = static_cast<int>(q % Ho);
    q /= Ho;
    const int f = static_cast<int>(q % F);
    q /= F;
    const int c = static_cast<int>(q % 3);
    const long long n = q / 3;
    const int h = yo >> 2, pq = yo & 3, w = xo >> 2, pr = xo & 3;
    video[i] = z[(((n * F + f) * H + h) * static_cast<long long>(W) + w) * 48 + (c * 4 + pr) * 4 + pq];
  }
}

// tiling.py:404-470: output[.., t0+t, h0+h, w0+w] += tile[.., t, h, w] * mt[t] * mh[h] * mw[w]; weights[...] += mt*mh*mw
// (tile (N, 3, Ft, Ht, Wt) channels-first, cropped to (at, ah, aw); output (N, 3, F, H, W), weights (N, 1, F, H, W))
__global__ void __launch_bounds__(256)
vae_blend_kernel(const float* __restrict__ tile, long long N, int Ft, int Ht, int Wt, int at, int ah, int aw,
                 const float* __restrict__ mt, const float* __restrict__ mh, const float* __restrict__ mw, float* __restrict__ output,
                 float* __restrict__ weights, int F, int H, int W, int t0, int h0, int w0) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = N * at * ah * aw;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int x = static_cast<int>(i % aw);
    long long q = i / aw;
    const int y = static_cast<int>(q % ah);
    q /= ah;
    const int t = static_cast<int>(q % at);
    const long long n = q / at;
    const float m = mt[t] * mh[y] * mw[x];
    const long long o1 = ((n * F + t0 + t) * H + h0 + y) * static_cast<long long>(W) + w0 + x;  // index into one channel plane set
    weights[o1] += m;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long src = (((n * 3 + c) * Ft + t) * Ht + y) * static_cast<long long>(Wt) + x;
      const long long dst = (((n * 3 + c) * F + t0 + t) * H + h0 + y) * static_cast<long long>(W) + w0 + x;
      output[dst] += tile[src] * m;
    }
  }
}

// tiling.py:506-508: output /= max(weights, 1e-8)
__global__ void __launch_bounds__(256)
vae_blend_normalize_kernel(float* __restrict__ output, const float* __restrict__ weights, long long N, long long plane) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = N * 3 * plane;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long n = i / (3 * plane), p = i % plane;
    output[i] = output[i] / fmaxf(weights[n * plane + p], 1e-8f);
  }
}

}  // namespace ltxb

using namespace ltxb;

extern "C" int ltxb_vae_gather_rows(const float* x, void* out, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t causal,
                                    int64_t m0, int64_t rows, const float* table_scale, const float* table_shift,
                                    const float* emb_scale, const float* emb_shift, int64_t emb_ld, float eps, int32_t pre_op,
                                    void* stream) {
  LTXB_CHECK_ARG(x && out, "ltxb_vae_gather_rows: null pointer");
  if (rows == 0) return LTXB_OK;
  LTXB_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && rows > 0 && m0 >= 0 &&
                     m0 + rows <= static_cast<int64_t>(N) * D * H * W,
                 "ltxb_vae_gather_rows: bad shape N=%d D=%d H=%d W=%d C=%d rows [%lld, +%lld)", N, D, H, W, C,
                 static_cast<long long>(m0), static_cast<long long>(rows));
  LTXB_CHECK_SUPPORTED(C % 8 == 0 && C <= 1024, "ltxb_vae_gather_rows: C=%d must be a multiple of 8, <= 1024", C);
  LTXB_CHECK_SUPPORTED(H >= 2 && W >= 2, "ltxb_vae_gather_rows: reflect padding needs H, W >= 2 (got %d x %d)", H, W);
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out), "ltxb_vae_gather_rows: misaligned x / out");
  LTXB_CHECK_ARG((table_scale == nullptr) == (table_shift == nullptr) && (emb_scale == nullptr) == (emb_shift == nullptr) &&
                     (emb_scale == nullptr || table_scale != nullptr),
                 "ltxb_vae_gather_rows: scale / shift come in pairs, embeddings only on top of a table");
  VaePreOp pre{table_scale, table_shift, emb_scale, emb_shift, emb_ld, eps, pre_op ? 1 : 0};
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  const int grid = vae_grid_for(rows, 8);
  const int groups = (C / 8 + 31) / 32;
  if (groups <= 1) LTXB_CUDA(launch_kernel(vae_gather_kernel<1>, dim3(grid), dim3(256), 0, s, 1, x, o, N, D, H, W, C, causal, static_cast<long long>(m0), static_cast<long long>(rows), pre));
  else if (groups <= 2) LTXB_CUDA(launch_kernel(vae_gather_kernel<2>, dim3(grid), dim3(256), 0, s, 1, x, o, N, D, H, W, C, causal, static_cast<long long>(m0), static_cast<long long>(rows), pre));
  else LTXB_CUDA(launch_kernel(vae_gather_kernel<4>, dim3(grid), dim3(256), 0, s, 1, x, o, N, D, H, W, C, causal, static_cast<long long>(m0), static_cast<long long>(rows), pre));
  return LTXB_OK;
}

extern "C" int ltxb_vae_depth_to_space(const float* y, const float* x, float* out, int64_t N, int32_t D, int32_t H, int32_t W,
                                       int32_t C, void* stream) {
  LTXB_CHECK_ARG(y && x && out && N > 0 && D > 0 && H > 0 && W > 0, "ltxb_vae_depth_to_space: bad argument");
  LTXB_CHECK_SUPPORTED(C % 8 == 0, "ltxb_vae_depth_to_space: C=%d must be a multiple of 8", C);
  const long long total = N * (2ll * D - 1) * (2 * H) * (2 * W) * (C / 2);
  if (total == 0) return LTXB_OK;
  LTXB_CUDA(launch_kernel(vae_depth_to_space_kernel, dim3(vae_grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          y, x, out, static_cast<long long>(N), D, H, W, C));
  return LTXB_OK;
}

extern "C" int ltxb_vae_prepare_latent(const float* sample, const float* noise, float noise_scale, const float* stdv,
                                       const float* mean, float* out, int64_t N, int32_t C, int64_t S, void* stream) {
  LTXB_CHECK_ARG(sample && stdv && mean && out && N > 0 && C > 0 && S > 0, "ltxb_vae_prepare_latent: bad argument");
  LTXB_CUDA(launch_kernel(vae_prepare_latent_kernel, dim3(vae_grid_for(N * S * C, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          sample, noise, noise_scale, stdv, mean, out, static_cast<long long>(N), C, static_cast<long long>(S)));
  return LTXB_OK;
}

extern "C" int ltxb_vae_unpatchify(const float* z, float* video, int64_t N, int32_t F, int32_t H, int32_t W, void* stream) {
  LTXB_CHECK_ARG(z && video && N > 0 && F > 0 && H > 0 && W > 0, "ltxb_vae_unpatchify: bad argument");
  LTXB_CUDA(launch_kernel(vae_unpatchify_kernel, dim3(vae_grid_for(N * 3 * F * 16ll * H * W, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          z, video, static_cast<long long>(N), F, H, W));
  return LTXB_OK;
}

extern "C" int ltxb_vae_blend_tile(const float* tile, int64_t N, int32_t Ft, int32_t Ht, int32_t Wt, int32_t at, int32_t ah, int32_t aw,
                                   const float* mt, const float* mh, const float* mw, float* output, float* weights, int32_t F,
                                   int32_t H, int32_t W, int32_t t0, int32_t h0, int32_t w0, void* stream) {
  LTXB_CHECK_ARG(tile && mt && mh && mw && output && weights, "ltxb_vae_blend_tile: null pointer");
  LTXB_CHECK_ARG(N > 0 && at > 0 && ah > 0 && aw > 0 && at <= Ft && ah <= Ht && aw <= Wt && t0 >= 0 && h0 >= 0 && w0 >= 0 &&
                     t0 + at <= F && h0 + ah <= H && w0 + aw <= W,
                 "ltxb_vae_blend_tile: tile region out of range");
  LTXB_CUDA(launch_kernel(vae_blend_kernel, dim3(vae_grid_for(N * at * ah * aw, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          tile, static_cast<long long>(N), Ft, Ht, Wt, at, ah, aw, mt, mh, mw, output, weights, F, H, W, t0, h0, w0));
  return LTXB_OK;
}

extern "C" int ltxb_vae_blend_normalize(float* output, const float* weights, int64_t N, int64_t plane, void* stream) {
  LTXB_CHECK_ARG(output && weights && N > 0 && plane > 0, "ltxb_vae_blend_normalize: bad argument");
  LTXB_CUDA(launch_kernel(vae_blend_normalize_kernel, dim3(vae_grid_for(N * 3 * plane, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          output, weights, static_cast<long long>(N), static_cast<long long>(plane)));
  return LTXB_OK;
}
