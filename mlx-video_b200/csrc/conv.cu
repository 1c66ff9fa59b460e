// Row N4 (first half): the latent 2x spatial upsampler between the two stages of the LTX-2 pipelines
// (mlx_video/models/ltx/upsampler.py).  Its 17 3x3x3 convolutions and one 3x3 convolution are GEMMs on the tensor
// cores: activations stay channels-last fp32 [N, D, H, W, C]; `im2col_cl_kernel` gathers the (zero-padded) taps of every
// output position into a bf16 row [kd*kh*kw*C] — the reference's weight layout (C_out, kd, kh, kw, C_in) IS the GEMM's
// W [N, K] operand, so no weight shuffle exists — and ltxb_gemm_bf16 (bias epilogue, fp32 out) does the contraction.
// Around it: GroupNorm (per-chunk partial sums, a fold into mean / rstd, then normalise + affine [+ residual] + SiLU in one pass),
// pixel shuffle, and the channels-first <-> channels-last moves that also carry the VAE un-/re-normalisation.
// All memory-bound, 128-bit vectorised where the layout allows; sizes are tiny next to the DiT (one call per video).
#include "common.cuh"
#include "ptx.cuh"

namespace ltxb {

static int conv_grid_for(long long work_items, int threads) {
  const long long blocks = (work_items + threads - 1) / threads;
  const long long cap = 148ll * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// out[m, ((kz*kh + ky)*kw + kx)*C + c] = x[n, d+kz-pd, h+ky-ph, w+kx-pw, c] (0 outside), m = ((n*D + d)*H + h)*W + w.
// One warp per output position: the position is decoded once, every tap is a contiguous C-channel run (fp32 in,
// bf16 out, 32 B read / 16 B written per lane and step) — the stores of a warp cover whole 512-byte spans of the row.
__global__ void __launch_bounds__(256)
im2col_cl_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int D, int H, int W, int C, int kd,
                 int kh, int kw) {
  pdl_launch_dependents();
  pdl_wait();
  const int c8 = C / 8, taps = kd * kh * kw;
  const int pd = kd / 2, ph = kh / 2, pw = kw / 2;
  const int lane = threadIdx.x & 31;
  const long long M = static_cast<long long>(N) * D * H * W;
  for (long long m = blockIdx.x * 8ll + (threadIdx.x >> 5); m < M; m += 8ll * gridDim.x) {
    const int w = static_cast<int>(m % W);
    long long q = m / W;
    const int h = static_cast<int>(q % H);
    q /= H;
    const int d = static_cast<int>(q % D);
    const long long n = q / D;
    __nv_bfloat16* orow = out + m * taps * static_cast<long long>(C);
    int tap = 0;
    for (int kz = 0; kz < kd; ++kz) {
      const int sd = d + kz - pd;
      for (int ky = 0; ky < kh; ++ky) {
        const int sh = h + ky - ph;
        for (int kx = 0; kx < kw; ++kx, ++tap) {
          const int sw = w + kx - pw;
          const bool inside = sd >= 0 && sd < D && sh >= 0 && sh < H && sw >= 0 && sw < W;
          const float* src = x + (((n * D + sd) * H + sh) * static_cast<long long>(W) + sw) * C;  // only dereferenced when inside
          __nv_bfloat16* dst = orow + static_cast<long long>(tap) * C;
          for (int cc = lane; cc < c8; cc += 32) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (inside) {
              const float4 a = *reinterpret_cast<const float4*>(src + cc * 8);
              const float4 b = *reinterpret_cast<const float4*>(src + cc * 8 + 4);
              v.x = pack_bf16x2(a.x, a.y), v.y = pack_bf16x2(a.z, a.w), v.z = pack_bf16x2(b.x, b.y), v.w = pack_bf16x2(b.z, b.w);
            }
            *reinterpret_cast<uint4*>(dst + cc * 8) = v;
          }
        }
      }
    }
  }
}

// GroupNorm3d, pass 1 (upsampler.py:85-101): partial[n][chunk][g] = (sum, sum of squares) of group g over the rows
// [chunk*rows_per_chunk, ...) of sample n.  One CTA per (chunk, n); thread t walks channels t, t+256, ...
constexpr int kGnMaxC = 1024;
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const float* __restrict__ x, float2* __restrict__ partial, int S, int C, int G, int rows_per_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_sum[kGnMaxC], s_sq[kGnMaxC];
  const int chunk = blockIdx.x, n = blockIdx.y, chunks = gridDim.x;
  const int r0 = chunk * rows_per_chunk, r1 = min(S, r0 + rows_per_chunk);
  const float* base = x + static_cast<long long>(n) * S * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, b = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float v = base[static_cast<long long>(r) * C + c];
      a += v, b = fmaf(v, v, b);
    }
    s_sum[c] = a, s_sq[c] = b;
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += 256) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < cpg; ++j) a += s_sum[g * cpg + j], b += s_sq[g * cpg + j];
    partial[(static_cast<long long>(n) * chunks + chunk) * G + g] = make_float2(a, b);
  }
}

// pass 2: the chunk partials of every (sample, group) folded into (mean, rstd) — in double, fixed order, 8 threads per
// group — so that the apply pass reads 2 floats per group instead of walking all partials in every CTA
__global__ void __launch_bounds__(256)
groupnorm_finalize_kernel(const float2* __restrict__ partial, float2* __restrict__ stats, int chunks, int S, int C, int G, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double s_a[256], s_b[256];
  const int n = blockIdx.x;
  const int g = threadIdx.x >> 3, j = threadIdx.x & 7;  // 32 groups x 8 lanes per pass
  for (int g0 = 0; g0 < G; g0 += 32) {
    double a = 0.0, b = 0.0;
    if (g0 + g < G) {
      for (int k = j; k < chunks; k += 8) {
        const float2 p = partial[(static_cast<long long>(n) * chunks + k) * G + g0 + g];
        a += p.x, b += p.y;
      }
    }
    s_a[threadIdx.x] = a, s_b[threadIdx.x] = b;
    __syncthreads();
    if (j == 0 && g0 + g < G) {
      for (int t = 1; t < 8; ++t) a += s_a[threadIdx.x + t], b += s_b[threadIdx.x + t];
      const double cnt = static_cast<double>(S) * (C / G);
      const double mean = a / cnt;
      const double var = fmax(b / cnt - mean * mean, 0.0);
      stats[static_cast<long long>(n) * G + g0 + g] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
    }
    __syncthreads();
  }
}

// pass 3: y = silu?( (x - mean_g) * rstd_g * weight[c] + bias[c] [+ resid] ), fp32 in place or out of place.
template <bool kResidual, bool kSilu>
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const float* __restrict__ x, float* __restrict__ out, const float2* __restrict__ stats,
                       const float* __restrict__ weight, const float* __restrict__ bias, const float* __restrict__ resid,
                       int S, int C, int G, int rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_mean[64], s_rstd[64];
  const int n = blockIdx.y;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += 256) {
    const float2 st = stats[static_cast<long long>(n) * G + g];
    s_mean[g] = st.x, s_rstd[g] = st.y;
  }
  __syncthreads();
  const int c4 = C / 4;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(S, r0 + rows_per_cta);
  const long long base = static_cast<long long>(n) * S * C;
  for (long long i = static_cast<long long>(r0) * c4 + threadIdx.x; i < static_cast<long long>(r1) * c4; i += 256) {
    const int c = static_cast<int>(i % c4) * 4;
    const long long off = base + (i / c4) * C + c;
    const float4 v = *reinterpret_cast<const float4*>(x + off);
    const float4 w = *reinterpret_cast<const float4*>(weight + c);
    const float4 b = *reinterpret_cast<const float4*>(bias + c);
    float y[4] = {v.x, v.y, v.z, v.w};
    const float ww[4] = {w.x, w.y, w.z, w.w}, bb[4] = {b.x, b.y, b.z, b.w};
    float rr[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (kResidual) {
      const float4 r = *reinterpret_cast<const float4*>(resid + off);
      rr[0] = r.x, rr[1] = r.y, rr[2] = r.z, rr[3] = r.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;
      float t = (y[j] - s_mean[g]) * s_rstd[g] * ww[j] + bb[j] + rr[j];
      if constexpr (kSilu) t = t / (1.0f + __expf(-t));
      y[j] = t;
    }
    *reinterpret_cast<float4*>(out + off) = make_float4(y[0], y[1], y[2], y[3]);
  }
}

// PixelShuffle2D (upsampler.py:124-139): in [F, H, W, 4*Co] -> out [F, 2H, 2W, Co], in channel = (co*2 + rh)*2 + rw
__global__ void __launch_bounds__(256)
pixel_shuffle2_kernel(const float* __restrict__ x, float* __restrict__ out, long long F, int H, int W, int Co) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = F * H * 2 * W * 2 * Co;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int co = static_cast<int>(i % Co);
    long long q = i / Co;
    const int ow = static_cast<int>(q % (2 * W));
    q /= 2 * W;
    const int oh = static_cast<int>(q % (2 * H));
    const long long f = q / (2 * H);
    const int h = oh >> 1, rh = oh & 1, w = ow >> 1, rw = ow & 1;
    out[i] = x[((f * H + h) * W + w) * (4ll * Co) + (co * 2 + rh) * 2 + rw];
  }
}

// (B, C, S) channels-first -> (B, S, C) channels-last with y = x * scale[c] + shift[c]   (S = F*H*W), and back with
// y = (x - shift[c]) / scale[c]: the layout moves of upsampler.py:250,290 carrying upsample_latents' (un)normalisation
template <bool kToChannelsLast>
__global__ void __launch_bounds__(256)
latent_layout_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ scale,
                     const float* __restrict__ shift, long long B, int C, long long S) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = B * C * S;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    // i indexes the OUTPUT, so the stores are coalesced
    if constexpr (kToChannelsLast) {
      const int c = static_cast<int>(i % C);
      const long long s = (i / C) % S, b = i / (static_cast<long long>(C) * S);
      const float v = x[(b * C + c) * S + s];
      out[i] = scale != nullptr ? fmaf(v, scale[c], shift[c]) : v;
    } else {
      const long long s = i % S;
      const int c = static_cast<int>((i / S) % C);
      const long long b = i / (static_cast<long long>(C) * S);
      const float v = x[(b * S + s) * C + c];
      out[i] = scale != nullptr ? (v - shift[c]) / scale[c] : v;
    }
  }
}

}  // namespace ltxb

using namespace ltxb;

extern "C" int ltxb_im2col_cl(const float* x, void* out, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t kd,
                              int32_t kh, int32_t kw, void* stream) {
  LTXB_CHECK_ARG(x && out, "ltxb_im2col_cl: null pointer");
  LTXB_CHECK_ARG(N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "ltxb_im2col_cl: bad shape (C must be a multiple of 8)");
  LTXB_CHECK_ARG(kd >= 1 && kh >= 1 && kw >= 1 && (kd & 1) && (kh & 1) && (kw & 1), "ltxb_im2col_cl: odd kernel sizes only (same padding)");
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out), "ltxb_im2col_cl: misaligned");
  const long long rows = static_cast<long long>(N) * D * H * W;  // one warp per output position, 8 per CTA
  LTXB_CUDA(launch_kernel(im2col_cl_kernel, dim3(conv_grid_for(rows * 32, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, x,
                          reinterpret_cast<__nv_bfloat16*>(out), N, D, H, W, C, kd, kh, kw));
  return LTXB_OK;
}

constexpr int kGnRows = 16;  // rows per CTA: 320 CTAs per sample at 5 x 32 x 32 positions
extern "C" int64_t ltxb_groupnorm_workspace_bytes(int32_t N, int64_t S, int32_t G) {
  const long long chunks = (S + kGnRows - 1) / kGnRows;
  return static_cast<int64_t>(N) * (chunks + 1) * G * sizeof(float2);  // per-chunk partials + the (mean, rstd) table
}

extern "C" int ltxb_groupnorm_silu(const float* x, float* out, int32_t N, int64_t S, int32_t C, int32_t G, float eps,
                                   const float* weight, const float* bias, const float* resid, int32_t silu, void* workspace,
                                   int64_t workspace_bytes, void* stream) {
  LTXB_CHECK_ARG(x && out && weight && bias && workspace, "ltxb_groupnorm_silu: null pointer");
  LTXB_CHECK_ARG(N > 0 && N <= 65535 && S > 0 && S < (1ll << 31) && C > 0 && G > 0 && G <= 64 && C % G == 0 && C % 4 == 0 && C <= kGnMaxC,
                 "ltxb_groupnorm_silu: bad shape N=%d S=%lld C=%d G=%d (C <= %d, C %% G == 0, C %% 4 == 0, G <= 64)", N,
                 static_cast<long long>(S), C, G, kGnMaxC);
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out) && aligned16(weight) && aligned16(bias) && aligned16(workspace) &&
                     (resid == nullptr || aligned16(resid)), "ltxb_groupnorm_silu: misaligned");
  LTXB_CHECK_ARG(workspace_bytes >= ltxb_groupnorm_workspace_bytes(N, S, G), "ltxb_groupnorm_silu: workspace too small");
  const int chunks = static_cast<int>((S + kGnRows - 1) / kGnRows);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float2* partial = reinterpret_cast<float2*>(workspace);
  float2* stats = partial + static_cast<long long>(N) * chunks * G;
  LTXB_CUDA(launch_kernel(groupnorm_stats_kernel, dim3(chunks, N), dim3(256), 0, s, 1, x, partial, static_cast<int>(S), C, G, kGnRows));
  LTXB_CUDA(launch_kernel(groupnorm_finalize_kernel, dim3(N), dim3(256), 0, s, 1, static_cast<const float2*>(partial), stats, chunks,
                          static_cast<int>(S), C, G, eps));
#define LTXB_GN_APPLY(R, A)                                                                                               \
  LTXB_CUDA(launch_kernel(groupnorm_apply_kernel<R, A>, dim3(chunks, N), dim3(256), 0, s, 1, x, out, static_cast<const float2*>(stats), \
                          weight, bias, resid, static_cast<int>(S), C, G, kGnRows))
  if (resid != nullptr) {
    if (silu) LTXB_GN_APPLY(true, true); else LTXB_GN_APPLY(true, false);
  } else {
    if (silu) LTXB_GN_APPLY(false, true); else LTXB_GN_APPLY(false, false);
  }
#undef LTXB_GN_APPLY
  return LTXB_OK;
}

extern "C" int ltxb_pixel_shuffle2(const float* x, float* out, int64_t F, int32_t H, int32_t W, int32_t Co, void* stream) {
  LTXB_CHECK_ARG(x && out && F > 0 && H > 0 && W > 0 && Co > 0, "ltxb_pixel_shuffle2: bad argument");
  const long long work = F * H * 2 * W * 2 * Co;
  LTXB_CUDA(launch_kernel(pixel_shuffle2_kernel, dim3(conv_grid_for(work, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, x,
                          out, static_cast<long long>(F), H, W, Co));
  return LTXB_OK;
}

extern "C" int ltxb_latent_layout(const float* x, float* out, const float* scale, const float* shift, int64_t B, int32_t C,
                                  int64_t S, int32_t to_channels_last, void* stream) {
  LTXB_CHECK_ARG(x && out && B > 0 && C > 0 && S > 0, "ltxb_latent_layout: bad argument");
  LTXB_CHECK_ARG((scale == nullptr) == (shift == nullptr), "ltxb_latent_layout: scale and shift come together");
  const long long work = B * C * S;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (to_channels_last)
    LTXB_CUDA(launch_kernel(latent_layout_kernel<true>, dim3(conv_grid_for(work, 256)), dim3(256), 0, s, 1, x, out, scale, shift,
                            static_cast<long long>(B), C, static_cast<long long>(S)));
  else
    LTXB_CUDA(launch_kernel(latent_layout_kernel<false>, dim3(conv_grid_for(work, 256)), dim3(256), 0, s, 1, x, out, scale, shift,
                            static_cast<long long>(B), C, static_cast<long long>(S)));
  return LTXB_OK;
}
