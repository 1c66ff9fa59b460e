// Memory-bound pieces of the LTX-2 DiT block: AdaLN modulate, norms, gate+residual, QK-norm+RoPE,
// timestep features, RoPE tables, sampler step.  All are 128-bit vectorised, one row per CTA where a
// row statistic is needed (warp-shuffle + one smem hop), grid-stride otherwise.  HBM roofline kernels:
// algorithmic bytes per element are stated next to each entry point in DESIGN.md.
#include "common.cuh"
#include <algorithm>
#include "peer_sync.cuh"
#include "ptx.cuh"

namespace ltxb {

constexpr int kRowThreads = 128;  // 16 rows resident per SM: 1280-token problems fit ONE wave of 148 x 16 CTAs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the whole block (blockDim.x multiple of 32, <= 1024). `red` is 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` against the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nwarps) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void load8_ldg(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 w;
  w.x = pack_bf16x2(v[0], v[1]);
  w.y = pack_bf16x2(v[2], v[3]);
  w.z = pack_bf16x2(v[4], v[5]);
  w.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------
// rms_norm / layer_norm + AdaLN modulate.  kLayerNorm=false: x*rsqrt(mean(x^2)+eps);  true: LayerNorm.
// mod_scale / mod_shift point at the first column of the scale / shift slice of the modulation row.
// ------------------------------------------------------------------------------------------------
// kResidual: x += y * g first (y bf16 = the projection that precedes this norm, g = gate_table + gate_mod row, or 1),
// the updated x is stored back and normalised in the same pass — the residual add of an out-projection and the
// next sub-layer's norm read the same fp32 row, so the GEMM keeps its cheap bf16 epilogue and the row is read once.
struct ResidualIn {
  const __nv_bfloat16* y;
  long long ldy;
  const float* gate_mod;    // first column of the gate slice of the modulation rows, or null
  const float* gate_table;  // or null
};
template <int kChunks, bool kLayerNorm, bool kResidual>
__global__ void __launch_bounds__(kRowThreads, kChunks <= 4 ? 9 : 1)  // <= 56 registers: 9 rows per SM, 1280 tokens in ONE wave
norm_modulate_kernel(float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long ldo,
                     int D, float eps, const float* __restrict__ mod_scale, const float* __restrict__ mod_shift,
                     long long ld_mod, const float* __restrict__ table_scale, const float* __restrict__ table_shift,
                     int row_div, const int* __restrict__ row_index, const ResidualIn res) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const long long row = blockIdx.x;
  float* xr = x + row * ldx;
  // narrow rows keep the whole modulation in registers too, so its loads are in flight together with x's
  // instead of after the reduction (the kernel is a chain of L2 round trips, not a bandwidth problem)
  constexpr bool kPrefetch = kChunks <= 1;  // wider rows would push the register count past one-wave occupancy
  const bool has_mod = (mod_scale != nullptr) || (table_scale != nullptr);
  long long mrow = 0;
  if (mod_scale != nullptr || (kResidual && res.gate_mod != nullptr)) mrow = row_index != nullptr ? row_index[row] : row / row_div;
  auto load_one = [&](const float* table, const float* mod, int c, float (&o)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    if (table != nullptr) load8_ldg(table + c, o);
    if (mod != nullptr) {
      float m[8];
      load8_ldg(mod + mrow * ld_mod + c, m);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += m[i];
    }
  };
  // RMS norm is linear in x, so (1 + scale) is folded into x BEFORE the reduction: half of the modulation loads are
  // then in flight together with x's, and only the shift is fetched after the row statistic is known
  constexpr bool kFold = !kLayerNorm && !kPrefetch;
  float v[kChunks][8];
  float psc[kPrefetch ? kChunks : 1][8], psh[kPrefetch ? kChunks : 1][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kChunks; ++j) {
    const int c = (threadIdx.x + j * kRowThreads) * 8;
    if (c < D) {
      load8(xr + c, v[j]);
      if constexpr (kResidual) {
        float y[8];
        load8_bf16(res.y + row * res.ldy + c, y);
        if (res.gate_mod != nullptr || res.gate_table != nullptr) {
          float g[8];
          load_one(res.gate_table, res.gate_mod, c, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j][i] = fmaf(y[i], g[i], v[j][i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j][i] += y[i];
        }
        store8(xr + c, v[j]);
      }
      if constexpr (kPrefetch) {
        if (has_mod) {
          load_one(table_scale, mod_scale, c, psc[j]);
          load_one(table_shift, mod_shift, c, psh[j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += kLayerNorm ? v[j][i] : v[j][i] * v[j][i];
      if constexpr (kFold) {
        if (has_mod) {
          float sc[8];
          load_one(table_scale, mod_scale, c, sc);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[j][i] *= 1.0f + sc[i];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j][i] = 0.f;
    }
  }
  s = block_sum(s, red);
  float mean = 0.f, rstd;
  if constexpr (kLayerNorm) {
    mean = s / static_cast<float>(D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kChunks; ++j) {
      const int c = (threadIdx.x + j * kRowThreads) * 8;
      if (c < D) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[j][i] - mean;
          q += d * d;
        }
      }
    }
    q = block_sum(q, red);
    rstd = rsqrtf(q / static_cast<float>(D) + eps);
  } else {
    rstd = rsqrtf(s / static_cast<float>(D) + eps);
  }
  __nv_bfloat16* orow = out + row * ldo;
#pragma unroll
  for (int j = 0; j < kChunks; ++j) {
    const int c = (threadIdx.x + j * kRowThreads) * 8;
    if (c >= D) continue;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (v[j][i] - mean) * rstd;
    if (has_mod) {
      if constexpr (kPrefetch) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaf(y[i], 1.0f + psc[j][i], psh[j][i]);
      } else if constexpr (kFold) {
        float sh[8];
        load_one(table_shift, mod_shift, c, sh);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] += sh[i];
      } else {
        float sc[8], sh[8];
        load_one(table_scale, mod_scale, c, sc);
        load_one(table_shift, mod_shift, c, sh);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaf(y[i], 1.0f + sc[i], sh[i]);
      }
    }
    store8_bf16(orow + c, y);
  }
}

template <bool kLayerNorm, bool kResidual = false>
static int launch_norm_modulate(const float* x_in, long long ldx, void* out, long long ldo, int R, int D, float eps,
                                const float* mod_scale, const float* mod_shift, long long ld_mod,
                                const float* table_scale, const float* table_shift, int row_div,
                                const int* row_index, cudaStream_t s, ResidualIn res = ResidualIn{nullptr, 0, nullptr, nullptr}) {
  const int chunks = (D / 8 + kRowThreads - 1) / kRowThreads;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  float* x = const_cast<float*>(x_in);  // only the kResidual instantiation writes through it
#define LTXB_LAUNCH_NM(C)                                                                                        \
  LTXB_CUDA(launch_kernel(norm_modulate_kernel<C, kLayerNorm, kResidual>, dim3(R), dim3(kRowThreads), 0, s, 1, x, ldx, o, ldo, D, eps, mod_scale, mod_shift, ld_mod, \
                                                                 table_scale, table_shift, row_div, row_index, res))
  if (chunks <= 1) LTXB_LAUNCH_NM(1);
  else if (chunks <= 2) LTXB_LAUNCH_NM(2);
  else if (chunks <= 4) LTXB_LAUNCH_NM(4);
  else if (chunks <= 8) LTXB_LAUNCH_NM(8);
  else LTXB_LAUNCH_NM(16);
#undef LTXB_LAUNCH_NM
  LTXB_CUDA(cudaGetLastError());
  return LTXB_OK;
}

// ------------------------------------------------------------------------------------------------
// x += y * g
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_residual_kernel(float* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ y, long long ldy,
                     long long R, int D, const float* __restrict__ gate, long long gate_ld,
                     const float* __restrict__ gate_table, int row_div, const int* __restrict__ row_index) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpr = D / 8;
  const long long total = R * cpr;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / cpr;
    const int c = static_cast<int>(i - row * cpr) * 8;
    float xv[8], yv[8];
    load8(x + row * ldx + c, xv);
    load8_bf16(y + row * ldy + c, yv);
    if (gate != nullptr) {
      const long long grow = row_index != nullptr ? row_index[row] : row / row_div;
      float g[8];
      load8_ldg(gate + grow * gate_ld + c, g);
      if (gate_table != nullptr) {
        float t[8];
        load8_ldg(gate_table + c, t);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] += t[k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) xv[k] = fmaf(yv[k], g[k], xv[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) xv[k] += yv[k];
    }
    store8(x + row * ldx + c, xv);
  }
}

// ------------------------------------------------------------------------------------------------
// full-width RMSNorm with weight, then split RoPE per head, in place on bf16.
// thread j owns elements [8j', 8j'+8) of the first half and the matching 8 of the second half of a head.
// ------------------------------------------------------------------------------------------------
struct PeerBases {
  void* p[8];
};

template <int kPer>
__global__ void __launch_bounds__(128, kPer <= 2 ? 9 : 1)  // <= 56 registers: 9 rows per SM, 1280 tokens in ONE wave
qknorm_rope_kernel(const __nv_bfloat16* x, long long ldx, __nv_bfloat16* out, long long ldo,  // may alias (in place)
                   int heads_per_group, long long group_stride, const PeerBases peers, int T, int H, int dh,
                   const float* __restrict__ weight,
                   float eps, const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, int B_pe,
                   long long seg_x_stride, long long seg_w_stride, const float* __restrict__ weight2, int n_norm_seg,
                   long long seg_o_stride) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const long long row = blockIdx.x;
  // segment blockIdx.y: an independent [rows, H*dh] column slice of the same buffer with its own norm weight
  // (q and k of the fused QKV buffer; the text K of every block in the stacked K/V buffer).  Segments >= n_norm_seg
  // are copied without norm / rotation (V on its way into the peers' receive buffers, same launch as q and k).
  x += blockIdx.y * seg_x_stride;
  if (out != nullptr) out += blockIdx.y * seg_x_stride;
  if (static_cast<int>(blockIdx.y) >= n_norm_seg) weight = nullptr;
  if (weight != nullptr) weight += blockIdx.y * seg_w_stride;
  if (weight2 != nullptr) weight2 += blockIdx.y * seg_w_stride;
  const int half = dh / 2;
  const int tph = half / 8;  // work items per head: one item = 8 elements of each half of a head
  const int items = H * tph;
  const long long bb = (B_pe == 1) ? 0 : row / T;
  const long long t = row % T;
  float a[kPer][8], b[kPer][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int item = threadIdx.x + j * blockDim.x;
    if (item < items) {
      const int h = item / tph, off = (item % tph) * 8;
      const __nv_bfloat16* p1 = x + row * ldx + h * dh + off;
      load8_bf16(p1, a[j]);
      load8_bf16(p1 + half, b[j]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += a[j][i] * a[j][i] + b[j][i] * b[j][i];
    }
  }
  float rstd = 1.f;
  if (weight != nullptr) {  // weight == nullptr: plain (re-)layout copy, e.g. V on its way to the all-to-all
    s = block_sum(s, red);
    rstd = rsqrtf(s / static_cast<float>(H * dh) + eps);
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const int item = threadIdx.x + j * blockDim.x;
    if (item >= items) continue;
    const int h = item / tph, off = (item % tph) * 8;
    if (weight != nullptr) {
      float w1[8], w2[8];
      load8_ldg(weight + h * dh + off, w1);
      load8_ldg(weight + h * dh + half + off, w2);
      if (weight2 != nullptr) {  // a second per-column factor folded in (the partner's norm weight)
        float u1[8], u2[8];
        load8_ldg(weight2 + h * dh + off, u1);
        load8_ldg(weight2 + h * dh + half + off, u2);
#pragma unroll
        for (int i = 0; i < 8; ++i) w1[i] *= u1[i], w2[i] *= u2[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) a[j][i] *= rstd * w1[i], b[j][i] *= rstd * w2[i];
      if (cos_tab != nullptr) {
        const long long tab = ((bb * H + h) * T + t) * half + off;
        float c[8], sn[8];
        load8_ldg(cos_tab + tab, c);
        load8_ldg(sin_tab + tab, sn);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float f = a[j][i] * c[i] - sn[i] * b[j][i];
          const float g = b[j][i] * c[i] + sn[i] * a[j][i];
          a[j][i] = f, b[j][i] = g;
        }
      }
    }
    // head h lands in group h / heads_per_group (one group per destination rank of the Ulysses all-to-all)
    // ... and with peer bases the group IS the destination GPU: the store goes straight over NVLink into that
    // rank's receive buffer (no send buffer, no separate all-to-all)
    const int grp = h / heads_per_group;
    __nv_bfloat16* gbase = (out == nullptr) ? reinterpret_cast<__nv_bfloat16*>(peers.p[grp]) + blockIdx.y * seg_o_stride : out + grp * group_stride;
    __nv_bfloat16* o1 = gbase + row * ldo + (h % heads_per_group) * dh + off;
    store8_bf16(o1, a[j]);
    store8_bf16(o1 + half, b[j]);
  }
}

// ------------------------------------------------------------------------------------------------
// sinusoidal timestep features (flip_sin_to_cos=True, downscale_freq_shift=0, max_period=1e4)
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embed_kernel(const float* __restrict__ t, int n, float scale, int dim,
                                      __nv_bfloat16* __restrict__ out, long long ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = dim / 2;
  const long long total = static_cast<long long>(n) * half;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(blockDim.x) * gridDim.x) {
    const long long r = i / half;
    const int k = static_cast<int>(i - r * half);
    const float ex = __fdiv_rn(__fmul_rn(-9.210340371976184f, static_cast<float>(k)), static_cast<float>(half));
    const float arg = __fmul_rn(__fmul_rn(t[r], scale), expf(ex));
    float sv, cv;
    sincosf(arg, &sv, &cv);
    out[r * ldo + k] = __float2bfloat16_rn(cv);
    out[r * ldo + half + k] = __float2bfloat16_rn(sv);
  }
}

// ------------------------------------------------------------------------------------------------
// RoPE cos/sin table, SPLIT layout (B, H, T, dim/2/H), left-padded with cos=1 / sin=0.
// ------------------------------------------------------------------------------------------------
struct RopeAxes {
  float max_pos[4];
};
__global__ void rope_table_kernel(const float* __restrict__ pos, int n_axes, int T, RopeAxes ax,
                                  const float* __restrict__ freq, int nfreq, int dim, int H, int use_middle,
                                  float* __restrict__ cos_out, float* __restrict__ sin_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = dim / 2;
  const int hd2 = half / H;
  const int pad = half - nfreq * n_axes;
  const long long bt = blockIdx.x;  // b*T + t
  const long long b = bt / T, t = bt % T;
  __shared__ float s_axis[4];
  if (threadIdx.x < n_axes) {
    const float* pp = pos + ((b * n_axes + threadIdx.x) * T + t) * 2;
    const float start = pp[0], end = pp[1];
    const float mid = use_middle ? __fdiv_rn(__fadd_rn(start, end), 2.0f) : start;
    const float frac = __fdiv_rn(mid, ax.max_pos[threadIdx.x]);
    s_axis[threadIdx.x] = __fadd_rn(__fmul_rn(frac, 2.0f), -1.0f);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    float c = 1.f, s = 0.f;
    if (j >= pad) {
      const int jj = j - pad;
      const int fi = jj / n_axes, axis = jj - fi * n_axes;
      const float ang = __fmul_rn(s_axis[axis], freq[fi]);
      sincosf(ang, &s, &c);
    }
    const int h = j / hd2, w = j - h * hd2;
    const long long o = ((b * H + h) * T + t) * hd2 + w;
    cos_out[o] = c;
    sin_out[o] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// small elementwise helpers
// ------------------------------------------------------------------------------------------------
__global__ void silu_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n8,
                                 long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(blockDim.x) * gridDim.x) {
    float v[8];
    load8_bf16(x + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = v[k] / (1.0f + __expf(-v[k]));
    store8_bf16(out + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = n8 * 8; i < n; ++i) {
      const float v = __bfloat162float(x[i]);
      out[i] = __float2bfloat16_rn(v / (1.0f + __expf(-v)));
    }
  }
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n8,
                                     long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(blockDim.x) * gridDim.x) {
    float v[8];
    load8(x + i * 8, v);
    store8_bf16(out + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n8 * 8; i < n; ++i) out[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long n8,
                                     long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(blockDim.x) * gridDim.x) {
    float v[8];
    load8_bf16(x + i * 8, v);
    store8(out + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n8 * 8; i < n; ++i) out[i] = __bfloat162float(x[i]);
}

// LoRA merge into a bf16 weight (lora.py:119-127): w = bf16(w + bf16(delta * strength)) — the reference's two roundings
__global__ void __launch_bounds__(256)
lora_merge_kernel(__nv_bfloat16* __restrict__ w, long long ldw, const float* __restrict__ delta, long long ldd,
                  long long R, int C, float strength) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpr = C / 8;
  const long long total = R * cpr;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / cpr;
    const int c = static_cast<int>(i - row * cpr) * 8;
    float wv[8], dv[8];
    load8_bf16(w + row * ldw + c, wv);
    load8(delta + row * ldd + c, dv);
#pragma unroll
    for (int k = 0; k < 8; ++k) wv[k] += __bfloat162float(__float2bfloat16_rn(dv[k] * strength));
    store8_bf16(w + row * ldw + c, wv);
  }
}

// MLX affine group quantisation -> bf16 (the checkpoints ltx.py:641-725 loads into nn.QuantizedLinear; mx.dequantize):
//   out[r, c] = bf16( scales[r, c/G] * q[r, c] + biases[r, c/G] ),  q = BITS-wide level c of row r, 32/BITS levels per
//   uint32 word, lowest bits first.  One thread = 8 consecutive columns (one 128-bit store; they share one group).
template <int BITS, bool F32_AUX>
__global__ void __launch_bounds__(256)
dequant_affine_kernel(const uint32_t* __restrict__ wq, long long ldq, const void* __restrict__ scales,
                      const void* __restrict__ biases, long long lds, __nv_bfloat16* __restrict__ out, long long ldo,
                      long long R, int C, int group) {
  pdl_launch_dependents();
  pdl_wait();
  const int cpr = C / 8;
  const long long total = R * cpr;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / cpr;
    const int c = static_cast<int>(i - row * cpr) * 8;
    const long long aux = row * lds + c / group;
    float s, b;
    if constexpr (F32_AUX) {
      s = static_cast<const float*>(scales)[aux];
      b = static_cast<const float*>(biases)[aux];
    } else {
      s = __bfloat162float(static_cast<const __nv_bfloat16*>(scales)[aux]);
      b = __bfloat162float(static_cast<const __nv_bfloat16*>(biases)[aux]);
    }
    const uint32_t* wrow = wq + row * ldq;
    unsigned long long bits64;  // the 8 levels of this thread, lowest first
    if constexpr (BITS == 8) {
      const uint2 w = *reinterpret_cast<const uint2*>(wrow + c / 4);
      bits64 = (static_cast<unsigned long long>(w.y) << 32) | w.x;
    } else if constexpr (BITS == 4) {
      bits64 = wrow[c / 8];
    } else {
      bits64 = (wrow[c / 16] >> ((c & 8) * 2)) & 0xffffu;
    }
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float q = static_cast<float>(static_cast<unsigned>(bits64 >> (k * BITS)) & ((1u << BITS) - 1u));
      v[k] = __fadd_rn(__fmul_rn(s, q), b);  // two roundings like scales * q + biases (exact product for bf16 scales)
    }
    store8_bf16(out + row * ldo + c, v);
  }
}

// CFG combine + to_denoised + mask blend + fp32 Euler (utils.py:404-440; generate.py:1255-1301)
__global__ void euler_step_kernel(float* __restrict__ x, const float* __restrict__ v_pos,
                                  const float* __restrict__ v_neg, float cfg_scale,
                                  const float* __restrict__ sigma_tok, float sigma, float sigma_next,
                                  const float* __restrict__ mask, const float* __restrict__ clean, long long n_tok,
                                  int C, float* __restrict__ x0_out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = n_tok * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(blockDim.x) * gridDim.x) {
    const long long tok = i / C;
    float v = v_pos[i];
    if (v_neg != nullptr) v = v + (cfg_scale - 1.0f) * (v - v_neg[i]);
    const float xs = x[i];
    const float st = sigma_tok != nullptr ? sigma_tok[tok] : sigma;
    float x0 = xs - st * v;
    if (mask != nullptr) {
      const float m = mask[tok];
      x0 = x0 * m + clean[i] * (1.0f - m);
    }
    if (x0_out != nullptr) x0_out[i] = x0;
    x[i] = x0 + sigma_next * (xs - x0) / sigma;
  }
}

// ------------------------------------------------------------------------------------------------
// Group equal per-token timesteps (at most F+1 distinct values over B*T tokens), so that
// AdaLN-single (adaln.py:29-47) runs on `cap` rows instead of B*T.  Single CTA; slot table in smem,
// claimed with atomicCAS on the float bit pattern.  Slot ORDER depends on thread timing, which is
// harmless: every token's row is computed from that token's own value.  More than `cap` distinct
// values: count = cap + 1 and every value is poisoned with NaN so the forward fails loudly.
// ------------------------------------------------------------------------------------------------
constexpr unsigned kEmptySlot = 0xFFFFFFFFu;  // a NaN pattern no real timestep carries
__global__ void __launch_bounds__(1024)
timestep_groups_kernel(const float* __restrict__ t, int n, int cap, float* __restrict__ values,
                       int* __restrict__ index, int* __restrict__ count) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ unsigned slots[];
  __shared__ int overflow;
  for (int i = threadIdx.x; i < cap; i += blockDim.x) slots[i] = kEmptySlot;
  if (threadIdx.x == 0) overflow = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = t[i];
    if (v == 0.0f) v = 0.0f;  // -0 and +0 share a slot
    const unsigned bits = __float_as_uint(v);
    int slot = -1;
    for (int s = 0; s < cap; ++s) {
      unsigned cur = reinterpret_cast<volatile unsigned*>(slots)[s];
      if (cur == kEmptySlot) cur = atomicCAS(&slots[s], kEmptySlot, bits);
      if (cur == kEmptySlot || cur == bits) {
        slot = s;
        break;
      }
    }
    if (slot < 0) overflow = 1, slot = 0;
    index[i] = slot;
  }
  __syncthreads();
  int used = 0;
  for (int i = threadIdx.x; i < cap; i += blockDim.x) {
    const unsigned b = slots[i];
    values[i] = overflow ? __uint_as_float(0x7FC00000u) : (b == kEmptySlot ? 0.0f : __uint_as_float(b));
    used += (b != kEmptySlot);
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  atomicAdd(&total, used);
  __syncthreads();
  if (threadIdx.x == 0) *count = overflow ? cap + 1 : total;
}

// ------------------------------------------------------------------------------------------------
// Cross-GPU barrier over NVLink peer memory: every rank raises its flag on every peer to `epoch`, then waits
// until all peers have raised theirs here.  Runs stream-ordered between the kernel that WROTE peer memory and the
// kernel that READS what the peers wrote, so the exchange needs no host round trip and no collective library.
// ------------------------------------------------------------------------------------------------
struct PeerFlags {
  int* p[8];  // p[i] = rank i's flag array (n_peers ints), mapped into this process
};
__global__ void peer_barrier_kernel(const PeerFlags flags, int n_peers, int my_rank, int* epoch_counter) {
  pdl_launch_dependents();
  pdl_wait();  // everything this rank wrote to its peers before the barrier has completed
  // the epoch lives in device memory and advances by one per barrier, identically on every rank — so a captured
  // CUDA graph can replay the barrier (a host-supplied epoch would be frozen into the graph)
  int epoch = 0;
  if (threadIdx.x == 0) {
    epoch = *epoch_counter + 1;
    *epoch_counter = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const int i = threadIdx.x;
  if (i < n_peers) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flags.p[i] + my_rank), "r"(epoch) : "memory");
    const int* mine = flags.p[my_rank] + i;
    const long long t0 = clock64();
    int seen;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
      if (seen - epoch < 0 && clock64() - t0 > LTXB_WATCHDOG_CYCLES) {
        printf("ltxb: peer barrier watchdog: rank %d waits for rank %d (epoch %d, seen %d)\n", my_rank, i, epoch, seen);
        __trap();
      }
    } while (seen - epoch < 0);
    __threadfence_system();
  }
}

// one source block stored to the same offset of every peer's buffer (the all-gather of a sequence-parallel result)
struct PeerDsts {
  uint4* p[8];
};
__global__ void __launch_bounds__(256) peer_broadcast_kernel(const uint4* __restrict__ src, long long n16, const PeerDsts dst, int n_peers) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n16; i += 256ll * gridDim.x) {
    const uint4 v = src[i];
    for (int r = 0; r < n_peers; ++r) dst.p[r][i] = v;
  }
}

static int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = 148ll * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace ltxb

using namespace ltxb;

extern "C" int ltxb_rmsnorm_modulate(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t R, int32_t D,
                                     float eps, const float* mod, int64_t ld_mod, int32_t scale_off,
                                     int32_t shift_off, const float* table_scale, const float* table_shift,
                                     int32_t row_div, const int32_t* row_index, void* stream) {
  LTXB_CHECK_ARG(x && out, "ltxb_rmsnorm_modulate: null pointer");
  if (R == 0) return LTXB_OK;
  LTXB_CHECK_ARG(R > 0 && D > 0, "ltxb_rmsnorm_modulate: bad shape R=%d D=%d", R, D);
  LTXB_CHECK_SUPPORTED(D % 8 == 0 && D <= 16 * 8 * kRowThreads, "ltxb_rmsnorm_modulate: D=%d must be a multiple of 8, <= 16384", D);
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out) && ldx % 4 == 0 && ldo % 8 == 0, "ltxb_rmsnorm_modulate: misaligned x/out");
  LTXB_CHECK_ARG((table_scale == nullptr) == (table_shift == nullptr), "ltxb_rmsnorm_modulate: tables come in pairs");
  if (mod) {
    LTXB_CHECK_ARG(aligned16(mod) && ld_mod % 4 == 0 && scale_off % 4 == 0 && shift_off % 4 == 0,
                   "ltxb_rmsnorm_modulate: modulation rows must be 16-byte aligned");
    LTXB_CHECK_ARG(row_index || row_div >= 1, "ltxb_rmsnorm_modulate: row_div must be >= 1");
  }
  return launch_norm_modulate<false>(x, ldx, out, ldo, R, D, eps, mod ? mod + scale_off : nullptr,
                                     mod ? mod + shift_off : nullptr, ld_mod, table_scale, table_shift,
                                     row_div > 0 ? row_div : 1, row_index, reinterpret_cast<cudaStream_t>(stream));
}

static int residual_norm_impl(float* x, int64_t ldx, const void* y, int64_t ldy, void* out, int64_t ldo,
                              int32_t R, int32_t D, float eps, const float* mod, int64_t ld_mod, int32_t gate_off,
                              int32_t scale_off, int32_t shift_off, const float* table_gate, const float* table_scale,
                              const float* table_shift, int32_t row_div, const int32_t* row_index, void* stream) {
  LTXB_CHECK_ARG(x && y && out, "ltxb_residual_norm_modulate: null pointer");
  if (R == 0) return LTXB_OK;
  LTXB_CHECK_ARG(R > 0 && D > 0, "ltxb_residual_norm_modulate: bad shape R=%d D=%d", R, D);
  LTXB_CHECK_SUPPORTED(D % 8 == 0 && D <= 16 * 8 * kRowThreads, "ltxb_residual_norm_modulate: D=%d must be a multiple of 8, <= 16384", D);
  LTXB_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(out) && ldx % 4 == 0 && ldy % 8 == 0 && ldo % 8 == 0,
                 "ltxb_residual_norm_modulate: misaligned x / y / out");
  LTXB_CHECK_ARG((table_scale == nullptr) == (table_shift == nullptr), "ltxb_residual_norm_modulate: scale / shift tables come in pairs");
  LTXB_CHECK_ARG((scale_off >= 0) == (shift_off >= 0), "ltxb_residual_norm_modulate: scale / shift offsets come in pairs");
  const bool use_mod = mod != nullptr && (gate_off >= 0 || scale_off >= 0);
  if (use_mod) {
    LTXB_CHECK_ARG(aligned16(mod) && ld_mod % 4 == 0 && (gate_off < 0 || gate_off % 4 == 0) && (scale_off < 0 || (scale_off % 4 == 0 && shift_off % 4 == 0)),
                   "ltxb_residual_norm_modulate: modulation rows must be 16-byte aligned");
    LTXB_CHECK_ARG(row_index || row_div >= 1, "ltxb_residual_norm_modulate: row_div must be >= 1");
  }
  if (table_gate) LTXB_CHECK_ARG(aligned16(table_gate), "ltxb_residual_norm_modulate: gate table misaligned");
  ResidualIn res{reinterpret_cast<const __nv_bfloat16*>(y), ldy, (use_mod && gate_off >= 0) ? mod + gate_off : nullptr, table_gate};
  const float* ms = (use_mod && scale_off >= 0) ? mod + scale_off : nullptr;
  const float* mh = (use_mod && scale_off >= 0) ? mod + shift_off : nullptr;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return launch_norm_modulate<false, true>(x, ldx, out, ldo, R, D, eps, ms, mh, ld_mod, table_scale, table_shift, row_div > 0 ? row_div : 1, row_index, s, res);
}

extern "C" int ltxb_residual_rmsnorm_modulate(float* x, int64_t ldx, const void* y, int64_t ldy, void* out, int64_t ldo,
                                              int32_t R, int32_t D, float eps, const float* mod, int64_t ld_mod,
                                              int32_t gate_off, int32_t scale_off, int32_t shift_off,
                                              const float* table_gate, const float* table_scale, const float* table_shift,
                                              int32_t row_div, const int32_t* row_index, void* stream) {
  return residual_norm_impl(x, ldx, y, ldy, out, ldo, R, D, eps, mod, ld_mod, gate_off, scale_off, shift_off, table_gate,
                            table_scale, table_shift, row_div, row_index, stream);
}

extern "C" int ltxb_layernorm_modulate(const float* x, int64_t ldx, void* out, int64_t ldo, int32_t R, int32_t D,
                                       float eps, const float* emb, int64_t ld_emb, const float* table_scale,
                                       const float* table_shift, int32_t row_div, const int32_t* row_index,
                                       void* stream) {
  LTXB_CHECK_ARG(x && out, "ltxb_layernorm_modulate: null pointer");
  if (R == 0) return LTXB_OK;
  LTXB_CHECK_ARG(R > 0 && D > 0, "ltxb_layernorm_modulate: bad shape R=%d D=%d", R, D);
  LTXB_CHECK_SUPPORTED(D % 8 == 0 && D <= 16 * 8 * kRowThreads, "ltxb_layernorm_modulate: D=%d must be a multiple of 8, <= 16384", D);
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out) && ldx % 4 == 0 && ldo % 8 == 0, "ltxb_layernorm_modulate: misaligned x/out");
  LTXB_CHECK_ARG((table_scale == nullptr) == (table_shift == nullptr), "ltxb_layernorm_modulate: tables come in pairs");
  if (emb) {
    LTXB_CHECK_ARG(aligned16(emb) && ld_emb % 4 == 0, "ltxb_layernorm_modulate: emb rows must be 16-byte aligned");
    LTXB_CHECK_ARG(row_index || row_div >= 1, "ltxb_layernorm_modulate: row_div must be >= 1");
  }
  // scale and shift both add the SAME embedded_timestep row (ltx.py:443-451)
  return launch_norm_modulate<true>(x, ldx, out, ldo, R, D, eps, emb, emb, ld_emb, table_scale, table_shift,
                                    row_div > 0 ? row_div : 1, row_index, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ltxb_gate_residual(float* x, int64_t ldx, const void* y, int64_t ldy, int32_t R, int32_t D,
                                  const float* gate, int64_t gate_ld, int32_t gate_off, const float* gate_table,
                                  int32_t row_div, const int32_t* row_index, void* stream) {
  LTXB_CHECK_ARG(x && y, "ltxb_gate_residual: null pointer");
  if (R == 0) return LTXB_OK;
  LTXB_CHECK_ARG(R > 0 && D > 0 && D % 8 == 0, "ltxb_gate_residual: bad shape R=%d D=%d (D %% 8 == 0)", R, D);
  LTXB_CHECK_ARG(aligned16(x) && aligned16(y) && ldx % 4 == 0 && ldy % 8 == 0, "ltxb_gate_residual: misaligned x/y");
  if (gate) {
    LTXB_CHECK_ARG(aligned16(gate) && gate_ld % 4 == 0 && gate_off % 4 == 0, "ltxb_gate_residual: misaligned gate");
    LTXB_CHECK_ARG(row_index || row_div >= 1, "ltxb_gate_residual: row_div must be >= 1");
  }
  const long long work = static_cast<long long>(R) * (D / 8);
  LTXB_CUDA(launch_kernel(gate_residual_kernel, dim3(grid_for(work, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, 
      x, ldx, reinterpret_cast<const __nv_bfloat16*>(y), ldy, R, D, gate ? gate + gate_off : nullptr, gate_ld,
      gate_table, row_div > 0 ? row_div : 1, row_index));
  return LTXB_OK;
}

static int qknorm_rope_launch(const void* x, int64_t ldx, void* out, const PeerBases& peers, int64_t ldo,
                              int32_t heads_per_group, int64_t group_stride, int32_t B, int32_t T, int32_t H, int32_t dh,
                              const float* weight, float eps, const float* cos_tab, const float* sin_tab, int32_t B_pe,
                              void* stream, int32_t n_seg = 1, int64_t seg_x_stride = 0, int64_t seg_w_stride = 0,
                              const float* weight2 = nullptr, int32_t n_norm_seg = 1 << 30, int64_t seg_o_stride = 0);

extern "C" int ltxb_qknorm_rope_scatter(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t heads_per_group,
                                        int64_t group_stride, int32_t B, int32_t T, int32_t H, int32_t dh,
                                        const float* weight, float eps, const float* cos_tab, const float* sin_tab,
                                        int32_t B_pe, void* stream) {
  LTXB_CHECK_ARG(out && aligned16(out) && group_stride % 8 == 0, "ltxb_qknorm_rope: null / misaligned out");
  return qknorm_rope_launch(x, ldx, out, PeerBases{}, ldo, heads_per_group, group_stride, B, T, H, dh, weight, eps, cos_tab,
                            sin_tab, B_pe, stream);
}

extern "C" int ltxb_qknorm_rope_scatter_peers(const void* x, int64_t ldx, void* const* group_bases, int32_t n_groups,
                                              int64_t ldo, int32_t B, int32_t T, int32_t H, int32_t dh,
                                              const float* weight, float eps, const float* cos_tab,
                                              const float* sin_tab, int32_t B_pe, void* stream) {
  LTXB_CHECK_ARG(group_bases && n_groups >= 1 && n_groups <= 8 && H % n_groups == 0,
                 "ltxb_qknorm_rope_scatter_peers: 1..8 groups dividing H=%d", H);
  PeerBases peers{};
  for (int i = 0; i < n_groups; ++i) {
    LTXB_CHECK_ARG(group_bases[i] && aligned16(group_bases[i]), "ltxb_qknorm_rope_scatter_peers: null / misaligned base %d", i);
    peers.p[i] = group_bases[i];
  }
  return qknorm_rope_launch(x, ldx, nullptr, peers, ldo, H / n_groups, 0, B, T, H, dh, weight, eps, cos_tab, sin_tab, B_pe, stream);
}

extern "C" int ltxb_qkv_norm_rope_scatter_peers(const void* qkv, int64_t ldx, int64_t seg_stride, void* const* group_bases,
                                                int32_t n_groups, int64_t ldo, int64_t slot_stride, int32_t T, int32_t H, int32_t dh,
                                                const float* qk_weight, float eps, const float* cos_tab, const float* sin_tab,
                                                int32_t B_pe, void* stream) {
  LTXB_CHECK_ARG(group_bases && n_groups >= 1 && n_groups <= 8 && H % n_groups == 0,
                 "ltxb_qkv_norm_rope_scatter_peers: 1..8 groups dividing H=%d", H);
  LTXB_CHECK_ARG(qk_weight && seg_stride % 8 == 0 && slot_stride % 8 == 0, "ltxb_qkv_norm_rope_scatter_peers: null weight / misaligned strides");
  PeerBases peers{};
  for (int i = 0; i < n_groups; ++i) {
    LTXB_CHECK_ARG(group_bases[i] && aligned16(group_bases[i]), "ltxb_qkv_norm_rope_scatter_peers: null / misaligned base %d", i);
    peers.p[i] = group_bases[i];
  }
  // three column segments of the fused QKV buffer in ONE launch: q and k normalised (weights [2, H*dh]) and rotated, v copied
  return qknorm_rope_launch(qkv, ldx, nullptr, peers, ldo, H / n_groups, 0, 1, T, H, dh, qk_weight, eps, cos_tab, sin_tab, B_pe, stream, 3,
                            seg_stride, static_cast<int64_t>(H) * dh, nullptr, 2, slot_stride);
}

static int qknorm_rope_launch(const void* x, int64_t ldx, void* out, const PeerBases& peers, int64_t ldo,
                              int32_t heads_per_group, int64_t group_stride, int32_t B, int32_t T, int32_t H, int32_t dh,
                              const float* weight, float eps, const float* cos_tab, const float* sin_tab, int32_t B_pe,
                              void* stream, int32_t n_seg, int64_t seg_x_stride, int64_t seg_w_stride, const float* weight2,
                              int32_t n_norm_seg, int64_t seg_o_stride) {
  LTXB_CHECK_ARG(x, "ltxb_qknorm_rope: null pointer");
  LTXB_CHECK_ARG(n_seg >= 1 && n_seg <= 65535 && seg_x_stride % 8 == 0 && seg_w_stride % 4 == 0 && (weight2 == nullptr || aligned16(weight2)),
                 "ltxb_qknorm_rope: bad segments n=%d", n_seg);
  if (B == 0 || T == 0) return LTXB_OK;
  LTXB_CHECK_ARG(B > 0 && T > 0 && H > 0, "ltxb_qknorm_rope: bad shape B=%d T=%d H=%d", B, T, H);
  LTXB_CHECK_SUPPORTED(dh == 64 || dh == 128, "ltxb_qknorm_rope: head dim %d not in {64,128}", dh);
  LTXB_CHECK_SUPPORTED(H * (dh / 16) <= 2048, "ltxb_qknorm_rope: H*dh=%d too wide", H * dh);
  LTXB_CHECK_ARG(aligned16(x) && ldx % 8 == 0 && ldo % 8 == 0, "ltxb_qknorm_rope: misaligned x/out");
  LTXB_CHECK_ARG(heads_per_group >= 1 && H % heads_per_group == 0, "ltxb_qknorm_rope: %d heads do not split into groups of %d", H, heads_per_group);
  LTXB_CHECK_ARG(weight == nullptr || aligned16(weight), "ltxb_qknorm_rope: misaligned weight");
  LTXB_CHECK_ARG((cos_tab == nullptr) == (sin_tab == nullptr), "ltxb_qknorm_rope: cos/sin come in pairs");
  LTXB_CHECK_ARG(cos_tab == nullptr || weight != nullptr, "ltxb_qknorm_rope: RoPE without the norm is not a path of the reference");
  if (cos_tab) LTXB_CHECK_ARG(aligned16(cos_tab) && aligned16(sin_tab) && (B_pe == 1 || B_pe == B), "ltxb_qknorm_rope: bad rope table");
  // 128 threads per row for the LTX widths (16 rows resident per SM: 1280 tokens are one wave)
  const int items = H * (dh / 16);
  const int per = items <= 128 ? 1 : (items <= 256 ? 2 : (items <= 512 ? 4 : (items <= 1024 ? 8 : 16)));
  const int nthr = std::min(128, (((items + per - 1) / per + 31) / 32) * 32);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xi = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* oi = reinterpret_cast<__nv_bfloat16*>(out);
#define LTXB_LAUNCH_QK(P)                                                                                             \
  LTXB_CUDA(launch_kernel(qknorm_rope_kernel<P>, dim3(B * T, n_seg), dim3(nthr), 0, st, 1, xi, ldx, oi, ldo, heads_per_group, \
                          group_stride, peers, T, H, dh, weight, eps, cos_tab, sin_tab, B_pe, seg_x_stride, seg_w_stride, weight2, n_norm_seg,  \
                          seg_o_stride))
  if (per == 1) LTXB_LAUNCH_QK(1);
  else if (per == 2) LTXB_LAUNCH_QK(2);
  else if (per == 4) LTXB_LAUNCH_QK(4);
  else if (per == 8) LTXB_LAUNCH_QK(8);
  else LTXB_LAUNCH_QK(16);
#undef LTXB_LAUNCH_QK
  return LTXB_OK;
}

extern "C" int ltxb_qknorm_rope(void* x, int64_t ldx, int32_t B, int32_t T, int32_t H, int32_t dh,
                                const float* weight, float eps, const float* cos_tab, const float* sin_tab,
                                int32_t B_pe, void* stream) {
  LTXB_CHECK_ARG(weight, "ltxb_qknorm_rope: null weight");
  return ltxb_qknorm_rope_scatter(x, ldx, x, ldx, H, 0, B, T, H, dh, weight, eps, cos_tab, sin_tab, B_pe, stream);
}

extern "C" int ltxb_qknorm_rope_segments(void* x, int64_t ldx, int32_t n_seg, int64_t seg_stride, int32_t B, int32_t T,
                                         int32_t H, int32_t dh, const float* weight, int64_t w_seg_stride,
                                         const float* weight2, float eps, const float* cos_tab, const float* sin_tab,
                                         int32_t B_pe, void* stream) {
  LTXB_CHECK_ARG(weight, "ltxb_qknorm_rope_segments: null weight");
  return qknorm_rope_launch(x, ldx, x, PeerBases{}, ldx, H, 0, B, T, H, dh, weight, eps, cos_tab, sin_tab, B_pe, stream, n_seg,
                            seg_stride, w_seg_stride, weight2);
}

extern "C" int ltxb_timestep_embed(const float* t, int32_t n, float scale, int32_t dim, void* out, int64_t ldo,
                                   void* stream) {
  LTXB_CHECK_ARG(t && out, "ltxb_timestep_embed: null pointer");
  if (n == 0) return LTXB_OK;
  LTXB_CHECK_ARG(n > 0 && dim > 0 && dim % 2 == 0 && ldo >= dim, "ltxb_timestep_embed: bad shape n=%d dim=%d", n, dim);
  const long long work = static_cast<long long>(n) * (dim / 2);
  LTXB_CUDA(launch_kernel(timestep_embed_kernel, dim3(grid_for(work, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, 
      t, n, scale, dim, reinterpret_cast<__nv_bfloat16*>(out), ldo));
  return LTXB_OK;
}

extern "C" int ltxb_rope_table(const float* positions, int32_t B, int32_t n_axes, int32_t T,
                               const float* max_pos_host, const float* freq, int32_t nfreq, int32_t dim, int32_t H,
                               int32_t use_middle, float* cos_out, float* sin_out, void* stream) {
  LTXB_CHECK_ARG(positions && max_pos_host && freq && cos_out && sin_out, "ltxb_rope_table: null pointer");
  if (B == 0 || T == 0) return LTXB_OK;
  LTXB_CHECK_ARG(B > 0 && T > 0, "ltxb_rope_table: bad shape B=%d T=%d", B, T);
  LTXB_CHECK_SUPPORTED(n_axes >= 1 && n_axes <= 4, "ltxb_rope_table: n_axes=%d not in [1,4]", n_axes);
  LTXB_CHECK_ARG(dim % 2 == 0 && H > 0 && (dim / 2) % H == 0, "ltxb_rope_table: dim=%d not divisible into %d heads", dim, H);
  LTXB_CHECK_ARG(nfreq >= 1 && nfreq * n_axes <= dim / 2, "ltxb_rope_table: nfreq=%d x %d axes exceeds dim/2=%d", nfreq, n_axes, dim / 2);
  RopeAxes ax{};
  for (int i = 0; i < n_axes; ++i) ax.max_pos[i] = max_pos_host[i];
  LTXB_CUDA(launch_kernel(rope_table_kernel, dim3(B * T), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, positions, n_axes, T, ax, freq, nfreq, dim,
                                                                                 H, use_middle, cos_out, sin_out));
  return LTXB_OK;
}

extern "C" int ltxb_silu_bf16(const void* x, void* out, int64_t n, void* stream) {
  LTXB_CHECK_ARG(x && out && n >= 0, "ltxb_silu_bf16: bad argument");
  if (n == 0) return LTXB_OK;
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out), "ltxb_silu_bf16: misaligned");
  LTXB_CUDA(launch_kernel(silu_bf16_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, 
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), n / 8, n));
  return LTXB_OK;
}

extern "C" int ltxb_cast_f32_to_bf16(const float* x, void* out, int64_t n, void* stream) {
  LTXB_CHECK_ARG(x && out && n >= 0, "ltxb_cast_f32_to_bf16: bad argument");
  if (n == 0) return LTXB_OK;
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out), "ltxb_cast_f32_to_bf16: misaligned");
  LTXB_CUDA(launch_kernel(cast_f32_bf16_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, 
      x, reinterpret_cast<__nv_bfloat16*>(out), n / 8, n));
  return LTXB_OK;
}

extern "C" int ltxb_cast_bf16_to_f32(const void* x, float* out, int64_t n, void* stream) {
  LTXB_CHECK_ARG(x && out && n >= 0, "ltxb_cast_bf16_to_f32: bad argument");
  if (n == 0) return LTXB_OK;
  LTXB_CHECK_ARG(aligned16(x) && aligned16(out), "ltxb_cast_bf16_to_f32: misaligned");
  LTXB_CUDA(launch_kernel(cast_bf16_f32_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, 
      reinterpret_cast<const __nv_bfloat16*>(x), out, n / 8, n));
  return LTXB_OK;
}

extern "C" int ltxb_lora_merge_bf16(void* w, int64_t ldw, const float* delta, int64_t ldd, int64_t R, int32_t C,
                                    float strength, void* stream) {
  LTXB_CHECK_ARG(w && delta, "ltxb_lora_merge_bf16: null pointer");
  if (R == 0) return LTXB_OK;
  LTXB_CHECK_ARG(R > 0 && C > 0 && C % 8 == 0, "ltxb_lora_merge_bf16: bad shape R=%lld C=%d (C must be a multiple of 8)", static_cast<long long>(R), C);
  LTXB_CHECK_ARG(aligned16(w) && aligned16(delta) && ldw % 8 == 0 && ldd % 4 == 0 && ldw >= C && ldd >= C,
                 "ltxb_lora_merge_bf16: misaligned operands / leading dimensions");
  LTXB_CUDA(launch_kernel(lora_merge_kernel, dim3(grid_for(R * (C / 8), 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          reinterpret_cast<__nv_bfloat16*>(w), static_cast<long long>(ldw), delta, static_cast<long long>(ldd),
                          static_cast<long long>(R), C, strength));
  return LTXB_OK;
}

extern "C" int ltxb_dequant_affine_bf16(const uint32_t* wq, int64_t ldq, const void* scales, const void* biases, int64_t lds,
                                        int32_t aux_f32, void* out, int64_t ldo, int64_t R, int32_t C, int32_t group_size,
                                        int32_t bits, void* stream) {
  LTXB_CHECK_ARG(wq && scales && biases && out, "ltxb_dequant_affine_bf16: null pointer");
  if (R == 0) return LTXB_OK;
  LTXB_CHECK_SUPPORTED(bits == 2 || bits == 4 || bits == 8, "ltxb_dequant_affine_bf16: bits=%d (2, 4 and 8 are built)", bits);
  LTXB_CHECK_SUPPORTED(group_size == 32 || group_size == 64 || group_size == 128,
                       "ltxb_dequant_affine_bf16: group_size=%d (32, 64 and 128 are built)", group_size);
  LTXB_CHECK_ARG(R > 0 && C > 0 && C % group_size == 0, "ltxb_dequant_affine_bf16: bad shape R=%lld C=%d group=%d",
                 static_cast<long long>(R), C, group_size);
  LTXB_CHECK_ARG(ldq >= static_cast<int64_t>(C) * bits / 32 && lds >= C / group_size && ldo >= C && ldo % 8 == 0 &&
                     aligned16(out) && (reinterpret_cast<uintptr_t>(wq) & 7u) == 0 && (bits != 8 || ldq % 2 == 0),
                 "ltxb_dequant_affine_bf16: misaligned operands / leading dimensions");
  auto launch = [&](auto kernel) {
    return launch_kernel(kernel, dim3(grid_for(R * (C / 8), 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, wq,
                         static_cast<long long>(ldq), scales, biases, static_cast<long long>(lds),
                         reinterpret_cast<__nv_bfloat16*>(out), static_cast<long long>(ldo), static_cast<long long>(R), C, group_size);
  };
  if (aux_f32) {
    if (bits == 8) LTXB_CUDA(launch(dequant_affine_kernel<8, true>));
    else if (bits == 4) LTXB_CUDA(launch(dequant_affine_kernel<4, true>));
    else LTXB_CUDA(launch(dequant_affine_kernel<2, true>));
  } else {
    if (bits == 8) LTXB_CUDA(launch(dequant_affine_kernel<8, false>));
    else if (bits == 4) LTXB_CUDA(launch(dequant_affine_kernel<4, false>));
    else LTXB_CUDA(launch(dequant_affine_kernel<2, false>));
  }
  return LTXB_OK;
}

extern "C" int ltxb_euler_step(float* x, const float* v_pos, const float* v_neg, float cfg_scale,
                               const float* sigma_tok, float sigma, float sigma_next, const float* mask,
                               const float* clean, int64_t n_tok, int32_t C, float* x0_out, void* stream) {
  LTXB_CHECK_ARG(x && v_pos && n_tok >= 0 && C > 0, "ltxb_euler_step: bad argument");
  LTXB_CHECK_ARG(sigma != 0.0f, "ltxb_euler_step: sigma must be non-zero (generate.py:1293-1301 divides by it)");
  LTXB_CHECK_ARG((mask == nullptr) || (clean != nullptr), "ltxb_euler_step: mask needs clean latents");
  if (n_tok == 0) return LTXB_OK;
  LTXB_CUDA(launch_kernel(euler_step_kernel, dim3(grid_for(n_tok * C, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, 
      x, v_pos, v_neg, cfg_scale, sigma_tok, sigma, sigma_next, mask, clean, n_tok, C, x0_out));
  return LTXB_OK;
}

extern "C" int ltxb_timestep_groups(const float* t, int32_t n, int32_t cap, float* values, int32_t* index,
                                    int32_t* count, void* stream) {
  LTXB_CHECK_ARG(t && values && index && count, "ltxb_timestep_groups: null pointer");
  LTXB_CHECK_ARG(n >= 0 && cap >= 1 && cap <= 4096, "ltxb_timestep_groups: bad n=%d / cap=%d (1..4096)", n, cap);
  LTXB_CUDA(launch_kernel(timestep_groups_kernel, dim3(1), dim3(1024), cap * sizeof(unsigned), reinterpret_cast<cudaStream_t>(stream), 1, t, n, cap, values,
                                                                                                    index, count));
  return LTXB_OK;
}

namespace ltxb {
int launch_peer_barrier(const PeerSync& s, cudaStream_t stream) {
  PeerFlags f{};
  for (int i = 0; i < s.n_peers; ++i) f.p[i] = s.flags[i];
  LTXB_CUDA(launch_kernel(peer_barrier_kernel, dim3(1), dim3(32), 0, stream, 1, f, s.n_peers, s.my_rank, s.epoch_counter));
  return LTXB_OK;
}
}  // namespace ltxb

extern "C" int ltxb_peer_barrier(int32_t* const* flag_ptrs, int32_t n_peers, int32_t my_rank, int32_t* epoch_counter,
                                 void* stream) {
  LTXB_CHECK_ARG(flag_ptrs && epoch_counter && n_peers >= 1 && n_peers <= 8 && my_rank >= 0 && my_rank < n_peers,
                 "ltxb_peer_barrier: bad group (n_peers=%d, my_rank=%d)", n_peers, my_rank);
  PeerFlags f{};
  for (int i = 0; i < n_peers; ++i) {
    LTXB_CHECK_ARG(flag_ptrs[i] != nullptr, "ltxb_peer_barrier: null flag pointer %d", i);
    f.p[i] = flag_ptrs[i];
  }
  LTXB_CUDA(launch_kernel(peer_barrier_kernel, dim3(1), dim3(32), 0, reinterpret_cast<cudaStream_t>(stream), 1, f, n_peers,
                          my_rank, epoch_counter));
  return LTXB_OK;
}

extern "C" int ltxb_peer_broadcast(const void* src, int64_t bytes, void* const* dst_ptrs, int32_t n_peers, void* stream) {
  LTXB_CHECK_ARG(src && dst_ptrs && bytes >= 0 && bytes % 16 == 0 && aligned16(src) && n_peers >= 1 && n_peers <= 8,
                 "ltxb_peer_broadcast: need a 16-byte aligned source, a multiple of 16 bytes and 1..8 destinations");
  if (bytes == 0) return LTXB_OK;
  PeerDsts d{};
  for (int i = 0; i < n_peers; ++i) {
    LTXB_CHECK_ARG(dst_ptrs[i] && aligned16(dst_ptrs[i]), "ltxb_peer_broadcast: null / misaligned destination %d", i);
    d.p[i] = reinterpret_cast<uint4*>(dst_ptrs[i]);
  }
  LTXB_CUDA(launch_kernel(peer_broadcast_kernel, dim3(grid_for(bytes / 16, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          1, reinterpret_cast<const uint4*>(src), static_cast<long long>(bytes / 16), d, n_peers));
  return LTXB_OK;
}
