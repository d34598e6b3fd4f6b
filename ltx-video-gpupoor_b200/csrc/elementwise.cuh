// Memory-bound kernels of the denoising path: 16-byte vectorised, one warp per row (row reductions by
// shuffles, values kept in registers so every row is read exactly once), bf16 rounding points mirrored
// from the reference where it is free to do so.
#pragma once
#include "common.cuh"
#include "comm.cuh"

namespace b200 {

DEVI float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

DEVI void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
DEVI void store8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// packed bf16x2 helpers: the reference's in-place bf16 tensor ops (x *= a; x += b) each round once to bf16, which is
// exactly what HMUL2.BF16 / HADD2.BF16 do on two lanes at a time (no F2F round trips through the XU pipe)
DEVI uint32_t bf2_mul(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
DEVI uint32_t bf2_add(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
DEVI void unpack8(const uint4& q, float (&f)[8]) {
  float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
DEVI uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ------------------------------------------------------------------------------------------
// norm + AdaLN modulate:  y = norm(x) * (1 + scale[g]) + shift[g]          (g = row / rows_per_group)
//   kLayerNorm = false: RMSNorm without affine (attention.py:233-251, 314-320; diffusers RMSNorm)
//   kLayerNorm = true : LayerNorm without affine (transformer3d.py:490-502 ; Wan norm1/norm2)
//   optional affine weight/bias (Wan norm3, VAE res_x_y norm3) applied before the modulation.
// x, y: [M, D] bf16 (in-place allowed), scale/shift: row g at (ptr + g*mod_ld), null => no modulation.
// One warp per row, NV uint4 per lane (D = 256*NV); the row stays packed in registers (read once), every load of
// the row and of its modulation vectors is issued before the first use.
// ------------------------------------------------------------------------------------------
template <int NV, bool kLayerNorm>
__global__ void __launch_bounds__(128)
norm_mod_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int M, long long ldx, long long ldy,
                const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift, long long mod_ld,
                int rows_per_group, const __nv_bfloat16* __restrict__ weight, const __nv_bfloat16* __restrict__ bias,
                float eps) {
  constexpr int D = NV * 256;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * 4;
  int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= M) return;
  // rows are walked grid-stride with the NEXT row's loads issued before the current row is reduced, so every warp
  // keeps a full row (NV x 512 B) in flight at all times
  constexpr bool kPrefetch = NV <= 12;       // wider rows would not fit two copies in registers
  uint4 xv[NV], xn[kPrefetch ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(x + row * ldx + (i * 32 + lane) * 8);
  for (; row < M; row += stride) {
    const int nrow = row + stride;
    if (kPrefetch && nrow < M) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xn[kPrefetch ? i : 0] = *reinterpret_cast<const uint4*>(x + nrow * ldx + (i * 32 + lane) * 8);
    }
    const long long g = row / rows_per_group;
    const __nv_bfloat16* sc = scale ? scale + g * mod_ld : nullptr;
    const __nv_bfloat16* sh = shift ? shift + g * mod_ld : nullptr;
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float v[8];
      unpack8(xv[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sum += v[j]; sq += v[j] * v[j]; }
    }
    float mean = 0.f, rs;
    if (kLayerNorm) {
      mean = warp_sum(sum) * (1.0f / D);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (NV > 12) asm volatile("" : "+r"(xv[i].x), "+r"(xv[i].y), "+r"(xv[i].z), "+r"(xv[i].w));
        float v[8];
        unpack8(xv[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; var += d * d; }
      }
      rs = rsqrtf(warp_sum(var) * (1.0f / D) + eps);
    } else {
      rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    }
    // the row is kept PACKED between the passes: hide it from CSE, or the compiler keeps all D unpacked floats live
#pragma unroll
    for (int i = 0; i < NV; ++i) asm volatile("" : "+r"(xv[i].x), "+r"(xv[i].y), "+r"(xv[i].z), "+r"(xv[i].w));
    __nv_bfloat16* yr = y + row * ldy;
    const uint32_t one2 = 0x3f803f80u;           // bf16x2 (1.0, 1.0)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      float v[8], o[8];
      unpack8(xv[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rs;
      if (weight) {
        float w[8];
        unpack8(ldg16(weight + c), w);
        if (bias) {
          float bb[8];
          unpack8(ldg16(bias + c), bb);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = o[j] * w[j] + bb[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = bf16r(o[j]) * w[j];
        }
      }
      uint4 r = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      if (sc) {      // bf16( bf16( bf16(o) * bf16(1 + s) ) + t ), the reference's rounding points
        const uint4 s4 = ldg16(sc + c), t4 = ldg16(sh + c);
        r.x = bf2_add(bf2_mul(r.x, bf2_add(one2, s4.x)), t4.x);
        r.y = bf2_add(bf2_mul(r.y, bf2_add(one2, s4.y)), t4.y);
        r.z = bf2_add(bf2_mul(r.z, bf2_add(one2, s4.z)), t4.z);
        r.w = bf2_add(bf2_mul(r.w, bf2_add(one2, s4.w)), t4.w);
      }
      *reinterpret_cast<uint4*>(yr + c) = r;
      if (NV > 12) asm volatile("" ::: "memory");   // wide rows: keep the modulation loads of later chunks from being hoisted (registers)
    }
    if (kPrefetch) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xv[i] = xn[kPrefetch ? i : 0];
    } else if (nrow < M) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(x + nrow * ldx + (i * 32 + lane) * 8);
    }
  }
}

// ------------------------------------------------------------------------------------------
// `mixed` precision (transformer3d.py:439-442, pipeline_ltx_video.py:1152-1177): the residual stream and the AdaLN
// modulation stay fp32, only the Linear inputs are bf16 (torch.autocast).  Same norm + modulate as norm_mod_kernel with an fp32
// row in, fp32 scale / shift, ONE rounding at the end:  y = bf16( norm(x) * (1 + scale[g]) + shift[g] ).
// ------------------------------------------------------------------------------------------
template <int NV, bool kLayerNorm>
__global__ void __launch_bounds__(128)
norm_mod_f32in_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int M, long long ldx, long long ldy,
                      const float* __restrict__ scale, const float* __restrict__ shift, long long mod_ld, int rows_per_group,
                      float eps) {
  constexpr int D = NV * 256;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * 4;
  for (int row = blockIdx.x * 4 + (threadIdx.x >> 5); row < M; row += stride) {
    const float* xr = x + row * ldx;
    float v[NV][8];
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 8), b = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 8 + 4);
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w; v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) { sum += v[i][j]; sq += v[i][j] * v[i][j]; }
    }
    float mean = 0.f, rs;
    if (kLayerNorm) {
      mean = warp_sum(sum) * (1.0f / D);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; var += d * d; }
      rs = rsqrtf(warp_sum(var) * (1.0f / D) + eps);
    } else {
      rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    }
    const long long g = row / rows_per_group;
    const float* sc = scale ? scale + g * mod_ld : nullptr;
    const float* sh = shift ? shift + g * mod_ld : nullptr;
    __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rs;
      if (sc) {
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + c)), s1 = __ldg(reinterpret_cast<const float4*>(sc + c + 4));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(sh + c)), t1 = __ldg(reinterpret_cast<const float4*>(sh + c + 4));
        const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], 1.0f + ss[j], tt[j]);
      }
      store8(yr + c, o);
    }
  }
}

// ada32[l, g, j, :] = fp32(table[l, j, :]) + fp32(temb[g, j*D:(j+1)*D])  -- the fp32 sum of the bf16 parameters / embeddings that
// `scale_shift_table[None, None] + timestep.float()` forms in mixed mode
__global__ void ada_add_f32_kernel(const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ temb,
                                   float* __restrict__ out, int L, int G, int JD) {
  const long long n8 = static_cast<long long>(L) * G * JD / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 8;
    const int c = static_cast<int>(e % JD);
    const int g = static_cast<int>((e / JD) % G);
    const int l = static_cast<int>(e / JD / G);
    float a[8], b[8];
    load8(table + static_cast<long long>(l) * JD + c, a);
    load8(temb + static_cast<long long>(g) * JD + c, b);
    *reinterpret_cast<float4*>(out + e) = make_float4(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]);
    *reinterpret_cast<float4*>(out + e + 4) = make_float4(a[4] + b[4], a[5] + b[5], a[6] + b[6], a[7] + b[7]);
  }
}

// ------------------------------------------------------------------------------------------
// LTX q/k RMSNorm (over the full inner dim, affine, eps 1e-5) + interleaved-pair RoPE, in place on
// the q and k column slices of a fused QKV buffer (attention.py:1040-1055, 960-975, 477-479).
//   y = bf16(x * rsqrt(mean(x^2)+eps)) * w ;  out = y*cos + rot(y)*sin,  rot: (2i,2i+1) -> (-y[2i+1], y[2i])
// cos/sin: [tokens_per_batch, D] bf16 (row = token index within the batch), null => no RoPE.
// grid.y selects the tensor: 0 = q, 1 = k.  All products / sums are bf16x2 ops with the reference's rounding points.
// ------------------------------------------------------------------------------------------
template <int NV, bool kRope>
__global__ void __launch_bounds__(128)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, int Mq, int Mk, long long ldq, long long ldk,
                    const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                    const __nv_bfloat16* __restrict__ cosT, const __nv_bfloat16* __restrict__ sinT,
                    int tokens_per_batch, float eps) {
  constexpr int D = NV * 256;
  const bool is_k = blockIdx.y == 1;
  __nv_bfloat16* base = is_k ? k : q;
  if (base == nullptr) return;
  const int M = is_k ? Mk : Mq;
  const long long ld = is_k ? ldk : ldq;
  const __nv_bfloat16* w = is_k ? wk : wq;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  __nv_bfloat16* xr = base + row * ld;
  uint4 xv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(xr + (i * 32 + lane) * 8);
  const long long trow = kRope ? static_cast<long long>(row % tokens_per_batch) * D : 0;
  // D <= 2048: the row's cos / sin slices are requested together with the row itself, before the reduction, so one warp keeps
  // 12 KB in flight instead of 4 KB then 8 x 1 KB (the kernel is latency-bound at 19 resident warps / SM: ncu warps_active 30 %)
  constexpr bool kPrefetch = kRope && NV <= 8;
  uint4 cv[kPrefetch ? NV : 1], sv[kPrefetch ? NV : 1];
  if (kPrefetch) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cv[i] = ldg16(cosT + trow + (i * 32 + lane) * 8);
      sv[i] = ldg16(sinT + trow + (i * 32 + lane) * 8);
    }
  }
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float v[8];
    unpack8(xv[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
  }
  const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    float v[8];
    unpack8(xv[i], v);
    const uint4 w4 = ldg16(w + c);
    uint32_t o[4] = {bf2_mul(pack_bf16(v[0] * rs, v[1] * rs), w4.x), bf2_mul(pack_bf16(v[2] * rs, v[3] * rs), w4.y),
                     bf2_mul(pack_bf16(v[4] * rs, v[5] * rs), w4.z), bf2_mul(pack_bf16(v[6] * rs, v[7] * rs), w4.w)};
    if (kRope) {
      const uint4 c4 = kPrefetch ? cv[i] : ldg16(cosT + trow + c), s4 = kPrefetch ? sv[i] : ldg16(sinT + trow + c);
      const uint32_t cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t rot = __byte_perm(o[j], 0, 0x1032) ^ 0x00008000u;     // (-y[2i+1], y[2i])
        o[j] = bf2_add(bf2_mul(o[j], cs[j]), bf2_mul(rot, sn[j]));
      }
    }
    *reinterpret_cast<uint4*>(xr + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// The same arithmetic with the rows of one TOKEN handled by one warp: its q row and k row for every sequence of the batch (the
// guidance conds share the token grid, so they share the token's cos / sin row).  The table row is loaded once and kept in registers
// for the 2*B rows — table traffic L2 -> SM drops from 2*B reads per token to one — and the next row is prefetched while the current
// one is reduced.  q / k: [B*tokens, D] views, row = b * tokens + tok.  Bit-identical to qk_norm_rope_kernel.
template <int NV>
__global__ void __launch_bounds__(128)
qk_norm_rope_tok_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, int B, int tokens, long long ldq, long long ldk,
                        const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                        const __nv_bfloat16* __restrict__ cosT, const __nv_bfloat16* __restrict__ sinT, float eps) {
  constexpr int D = NV * 256;
  const int tok = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= tokens) return;
  uint4 cv[NV], sv[NV], xv[NV], xn[NV];
  auto row_ptr = [&](int r) -> __nv_bfloat16* {
    const int b = r % B;
    return (r < B) ? q + (static_cast<long long>(b) * tokens + tok) * ldq : k + (static_cast<long long>(b) * tokens + tok) * ldk;
  };
  {
    const __nv_bfloat16* x0 = row_ptr(0);
#pragma unroll
    for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(x0 + (i * 32 + lane) * 8);
    const long long trow = static_cast<long long>(tok) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      cv[i] = ldg16(cosT + trow + (i * 32 + lane) * 8);
      sv[i] = ldg16(sinT + trow + (i * 32 + lane) * 8);
    }
  }
#pragma unroll 1
  for (int r = 0; r < 2 * B; ++r) {
    __nv_bfloat16* xr = row_ptr(r);
    if (r + 1 < 2 * B) {
      const __nv_bfloat16* xnext = row_ptr(r + 1);
#pragma unroll
      for (int i = 0; i < NV; ++i) xn[i] = *reinterpret_cast<const uint4*>(xnext + (i * 32 + lane) * 8);
    }
    const __nv_bfloat16* w = (r < B) ? wq : wk;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float v[8];
      unpack8(xv[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
    }
    const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      float v[8];
      unpack8(xv[i], v);
      const uint4 w4 = ldg16(w + c);
      uint32_t o[4] = {bf2_mul(pack_bf16(v[0] * rs, v[1] * rs), w4.x), bf2_mul(pack_bf16(v[2] * rs, v[3] * rs), w4.y),
                       bf2_mul(pack_bf16(v[4] * rs, v[5] * rs), w4.z), bf2_mul(pack_bf16(v[6] * rs, v[7] * rs), w4.w)};
      const uint32_t cs[4] = {cv[i].x, cv[i].y, cv[i].z, cv[i].w}, sn[4] = {sv[i].x, sv[i].y, sv[i].z, sv[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t rot = __byte_perm(o[j], 0, 0x1032) ^ 0x00008000u;     // (-y[2i+1], y[2i])
        o[j] = bf2_add(bf2_mul(o[j], cs[j]), bf2_mul(rot, sn[j]));
      }
      *reinterpret_cast<uint4*>(xr + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) xv[i] = xn[i];
  }
}

// Wide rows (D = 4096: LTX-Video 13B): the token's q / k rows do not fit one warp's registers next to the table row and the prefetched
// next row, so a token is handled by a warp PAIR (each warp NVH chunks = half of the columns, its half of the cos / sin row resident);
// the two partial sums of squares meet in shared memory (slot parity = row parity: a slot is rewritten only after the barrier of the
// following row).  Same arithmetic per element; the row statistic is the sum of two warp sums instead of one (fp32 rounding only).
template <int NVH>
__global__ void __launch_bounds__(128)
qk_norm_rope_tok2_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, int B, int tokens, long long ldq, long long ldk,
                         const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                         const __nv_bfloat16* __restrict__ cosT, const __nv_bfloat16* __restrict__ sinT, float eps) {
  constexpr int D = NVH * 512;
  __shared__ float red[2][2][2];                      // [token of the CTA][row parity][half]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tp = warp >> 1, half = warp & 1;
  const int tok_raw = blockIdx.x * 2 + tp;
  const bool active = tok_raw < tokens;
  const int tok = active ? tok_raw : tokens - 1;      // idle pair: same loads, no stores (keeps the barriers uniform)
  const int col0 = half * NVH * 256;
  uint4 cv[NVH], sv[NVH], xv[NVH], xn[NVH];
  auto row_ptr = [&](int r) -> __nv_bfloat16* {
    const int b = r % B;
    return ((r < B) ? q + (static_cast<long long>(b) * tokens + tok) * ldq : k + (static_cast<long long>(b) * tokens + tok) * ldk) + col0;
  };
  {
    const __nv_bfloat16* x0 = row_ptr(0);
#pragma unroll
    for (int i = 0; i < NVH; ++i) xv[i] = *reinterpret_cast<const uint4*>(x0 + (i * 32 + lane) * 8);
    const long long trow = static_cast<long long>(tok) * D + col0;
#pragma unroll
    for (int i = 0; i < NVH; ++i) {
      cv[i] = ldg16(cosT + trow + (i * 32 + lane) * 8);
      sv[i] = ldg16(sinT + trow + (i * 32 + lane) * 8);
    }
  }
#pragma unroll 1
  for (int r = 0; r < 2 * B; ++r) {
    __nv_bfloat16* xr = row_ptr(r);
    if (r + 1 < 2 * B) {
      const __nv_bfloat16* xnext = row_ptr(r + 1);
#pragma unroll
      for (int i = 0; i < NVH; ++i) xn[i] = *reinterpret_cast<const uint4*>(xnext + (i * 32 + lane) * 8);
    }
    const __nv_bfloat16* w = ((r < B) ? wq : wk) + col0;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NVH; ++i) {
      float v[8];
      unpack8(xv[i], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
    }
    sq = warp_sum(sq);
    if (lane == 0) red[tp][r & 1][half] = sq;
    __syncthreads();
    const float rs = rsqrtf((red[tp][r & 1][0] + red[tp][r & 1][1]) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NVH; ++i) {
      const int c = (i * 32 + lane) * 8;
      float v[8];
      unpack8(xv[i], v);
      const uint4 w4 = ldg16(w + c);
      uint32_t o[4] = {bf2_mul(pack_bf16(v[0] * rs, v[1] * rs), w4.x), bf2_mul(pack_bf16(v[2] * rs, v[3] * rs), w4.y),
                       bf2_mul(pack_bf16(v[4] * rs, v[5] * rs), w4.z), bf2_mul(pack_bf16(v[6] * rs, v[7] * rs), w4.w)};
      const uint32_t cs[4] = {cv[i].x, cv[i].y, cv[i].z, cv[i].w}, sn[4] = {sv[i].x, sv[i].y, sv[i].z, sv[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t rot = __byte_perm(o[j], 0, 0x1032) ^ 0x00008000u;     // (-y[2i+1], y[2i])
        o[j] = bf2_add(bf2_mul(o[j], cs[j]), bf2_mul(rot, sn[j]));
      }
      if (active) *reinterpret_cast<uint4*>(xr + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < NVH; ++i) xv[i] = xn[i];
  }
}

DEVI void load_cos_sin(const float* __restrict__ cosT, const float* __restrict__ sinT, long long off, float (&cs)[8], float (&sn)[8]) {
  const float4 c0 = __ldg(reinterpret_cast<const float4*>(cosT + off)), c1 = __ldg(reinterpret_cast<const float4*>(cosT + off + 4));
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(sinT + off)), s1 = __ldg(reinterpret_cast<const float4*>(sinT + off + 4));
  cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
  sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
}

// ------------------------------------------------------------------------------------------
// Wan q/k RMSNorm (full inner dim, affine, eps 1e-6, two bf16 roundings: model.py:104-111) + 3-axis RoPE
// applied per head with fp32 [tokens, HD] cos/sin tables, fp32 math, one rounding
// (posemb_layers.py:222-276).  token = token_offset + row % tokens_per_batch  (token_offset = the rank's
// first global token under Ulysses sequence parallelism, xdit_context_parallel.py:52-57).
// ------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(128, (NV > 12 ? 3 : 1))      // wide rows: cap the allocation at 168 registers (3 CTAs / SM), not 255
qk_norm_rope_wan_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k, int Mq, int Mk, long long ldq,
                        long long ldk, const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                        const float* __restrict__ cosT, const float* __restrict__ sinT, int head_dim,
                        int tokens_per_batch, int token_offset, float eps) {
  constexpr int D = NV * 256;
  const bool is_k = blockIdx.y == 1;
  __nv_bfloat16* base = is_k ? k : q;
  if (base == nullptr) return;
  const int M = is_k ? Mk : Mq;
  const long long ld = is_k ? ldk : ldq;
  const __nv_bfloat16* w = is_k ? wk : wq;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  __nv_bfloat16* xr = base + row * ld;
  uint4 xv[NV];                                       // the row stays PACKED between the passes (160 unpacked floats at D = 5120 would spill)
#pragma unroll
  for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(xr + (i * 32 + lane) * 8);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float t[8];
    unpack8(xv[i], t);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += t[j] * t[j];
  }
  const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) asm volatile("" : "+r"(xv[i].x), "+r"(xv[i].y), "+r"(xv[i].z), "+r"(xv[i].w));   // keep it packed (no CSE with the pass above)
  const long long trow = cosT ? static_cast<long long>(token_offset + row % tokens_per_batch) * head_dim : 0;
  // every head of a token shares the token's cos / sin row, and a lane's column offset inside its head is the same in every
  // 256-column chunk when head_dim divides 256 (it is 128): the 16 table values are loaded ONCE per row instead of once per chunk
  // (the per-chunk form moved 4x more table bytes than q/k bytes from L2 to the SM)
  const bool hoist = cosT && (256 % head_dim) == 0;
  float cs[8], sn[8];
  if (hoist) load_cos_sin(cosT, sinT, trow + (lane * 8) % head_dim, cs, sn);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (i * 32 + lane) * 8;
    float o[8];
    {
      const uint4 w4 = ldg16(w + c);
      float vv[8];
      unpack8(xv[i], vv);
      unpack8(make_uint4(bf2_mul(pack_bf16(vv[0] * rs, vv[1] * rs), w4.x), bf2_mul(pack_bf16(vv[2] * rs, vv[3] * rs), w4.y),
                         bf2_mul(pack_bf16(vv[4] * rs, vv[5] * rs), w4.z), bf2_mul(pack_bf16(vv[6] * rs, vv[7] * rs), w4.w)), o);
    }
    if (cosT) {
      if (!hoist) load_cos_sin(cosT, sinT, trow + c % head_dim, cs, sn);
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        r[j] = o[j] * cs[j] - o[j + 1] * sn[j];
        r[j + 1] = o[j + 1] * cs[j + 1] + o[j] * sn[j + 1];
      }
      store8(xr + c, r);
    } else {
      store8(xr + c, o);
    }
    if (NV > 12) asm volatile("" ::: "memory");       // wide rows: keep later chunks' weight loads from being hoisted (registers)
  }
}

// Same arithmetic, fused with the Ulysses head-scatter (xdit_context_parallel.py:149-184): the rows of the local
// fused QKV projection [B*n_loc, 3*D] are normalised / rotated (q, k) or passed through (v) and each head group's
// slice is stored straight into the receive buffer of the rank that owns those heads,
//   recv_g[T, b, sel, h_local, :]   (T = global token, sel = q|k|v, g = head / Hp),
// i.e. [N, B, 3, Hp, d] in global token order -- exactly the layout the attention kernel's TMA maps read.
// grid.y = 3 (q, k, v), or 2 when the V third already left from the QKV GEMM's epilogue (GemmParams::vs_peers).  The launch covers rows [row0, row_end) of the M local rows; `signal_ctas` CTAs (of this and the other
// launches of the same exchange) arrive before the last one publishes the epoch flag on every peer (comm.cuh).
// Rows are walked grid-stride by a BOUNDED grid (2 CTAs per SM and selector) with the next row's loads in flight: (a) every CTA
// ends with a system-scope fence that waits for its peer stores to be acknowledged over NVLink, which the round-1 form paid once
// per 4 rows; (b) a grid of thousands of 128-thread CTAs fills every SM's thread slots, so a GEMM launched next to it on another
// stream (the next token chunk's QKV projection) could not become resident and the overlap never happened.
template <int NV>
__global__ void __launch_bounds__(128)
qk_norm_rope_wan_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int row0, int row_end, const __nv_bfloat16* __restrict__ wq,
                                const __nv_bfloat16* __restrict__ wk, const float* __restrict__ cosT,
                                const float* __restrict__ sinT, int head_dim, int tokens_per_batch, int token_offset,
                                float eps, int B, int Hp, const PeerPtrs pp, unsigned int signal_ctas) {
  constexpr int D = NV * 256;
  constexpr bool kPrefetch = NV <= 8;               // wider rows would not fit two copies in registers
  const int sel = blockIdx.y;                       // 0 q, 1 k, 2 v
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * 4;
  int row = row0 + blockIdx.x * 4 + (threadIdx.x >> 5);
  const __nv_bfloat16* w = sel == 1 ? wk : wq;
  const int group_cols = Hp * head_dim;
  uint4 xv[NV], xn[kPrefetch ? NV : 1];
  if (row < row_end) {
#pragma unroll
    for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(qkv + row * ld + sel * D + (i * 32 + lane) * 8);
  }
  for (; row < row_end; row += stride) {
    const int nrow = row + stride;
    if (kPrefetch && nrow < row_end) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xn[kPrefetch ? i : 0] = *reinterpret_cast<const uint4*>(qkv + nrow * ld + sel * D + (i * 32 + lane) * 8);
    }
    const int b = row / tokens_per_batch, n = row - b * tokens_per_batch;
    const long long T = token_offset + n;
    float rs = 1.f;
    if (sel < 2) {
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[8];
        unpack8(xv[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) sq += v[j] * v[j];
      }
      rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    }
    const long long trow = T * head_dim;
    const bool hoist = (256 % head_dim) == 0;       // see qk_norm_rope_wan_kernel: one table read per row, not per chunk
    float cs[8], sn[8];
    if (sel < 2 && hoist) load_cos_sin(cosT, sinT, trow + (lane * 8) % head_dim, cs, sn);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * 32 + lane) * 8;
      uint4 r = xv[i];
      if (sel < 2) {
        float v[8], o[8];
        unpack8(xv[i], v);
        const uint4 w4 = ldg16(w + c);
        unpack8(make_uint4(bf2_mul(pack_bf16(v[0] * rs, v[1] * rs), w4.x), bf2_mul(pack_bf16(v[2] * rs, v[3] * rs), w4.y),
                           bf2_mul(pack_bf16(v[4] * rs, v[5] * rs), w4.z), bf2_mul(pack_bf16(v[6] * rs, v[7] * rs), w4.w)), o);
        if (!hoist) load_cos_sin(cosT, sinT, trow + c % head_dim, cs, sn);
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          q[j] = o[j] * cs[j] - o[j + 1] * sn[j];
          q[j + 1] = o[j + 1] * cs[j + 1] + o[j] * sn[j + 1];
        }
        r = make_uint4(pack_bf16(q[0], q[1]), pack_bf16(q[2], q[3]), pack_bf16(q[4], q[5]), pack_bf16(q[6], q[7]));
      }
      const int g = c / group_cols, cc = c - g * group_cols;
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(pp.data[g]) + ((T * B + b) * 3 + sel) * group_cols + cc;
      *reinterpret_cast<uint4*>(dst) = r;
    }
    if (kPrefetch) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xv[i] = xn[kPrefetch ? i : 0];
    } else if (nrow < row_end) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xv[i] = *reinterpret_cast<const uint4*>(qkv + nrow * ld + sel * D + (i * 32 + lane) * 8);
    }
  }
  peer_signal_done(pp, signal_ctas);
}

// out = sum_j coef[j] * x[j]   (fp32, up to 6 terms; the UniPC predictor/corrector and CFG combine are
// linear combinations with host-computed scalars: fm_solvers_unipc.py:321,458-484,590-626; text2video.py:562)
struct LinCombParams {
  const float* x[6];
  float c[6];
  int terms;
};
__global__ void lincomb_f32_kernel(float* __restrict__ out, long long n, const LinCombParams p) {
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      if (j < p.terms) {
        const float4 v = *reinterpret_cast<const float4*>(p.x[j] + i);
        acc.x += p.c[j] * v.x; acc.y += p.c[j] * v.y; acc.z += p.c[j] * v.z; acc.w += p.c[j] * v.w;
      }
    }
    *reinterpret_cast<float4*>(out + i) = acc;
  }
}

// ------------------------------------------------------------------------------------------
// RectifiedFlowScheduler.step with PER-TOKEN timesteps (rf.py:361-375): every token finds the next schedule entry strictly below
// its own timestep - 1e-6 (0 if none) and takes   prev = x - (t - lower) * v,   or with `noise` (rf.py:370-373)
// prev = (1 - next) * (x - t * v) + next * noise, next = t - (t - lower).   fp32 with the reference's rounding points (separate multiply and add, no
// contraction), so the result equals the PyTorch expression bit for bit.  x, v, noise, out: [tokens, channels] fp32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rf_step_tokens_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ noise,
                      const float* __restrict__ tok_t, long long tokens, int channels, const float* __restrict__ schedule, int num_steps) {
  const int c4 = channels / 4;
  const long long n4 = tokens * c4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    const float t = __ldg(tok_t + i / c4);
    float lower = 0.f;
    for (int s = 0; s < num_steps; ++s) {
      const float ts = __ldg(schedule + s);
      if (ts < __fsub_rn(t, 1e-6f)) { lower = ts; break; }       // the schedule descends: the first hit is the largest
    }
    const float dt = __fsub_rn(t, lower);
    const float4 xv = reinterpret_cast<const float4*>(x)[i], vv = reinterpret_cast<const float4*>(v)[i];
    float4 o;
    if (noise) {
      const float4 z = reinterpret_cast<const float4*>(noise)[i];
      const float nxt = __fsub_rn(t, dt);                        // next_timestep = timestep - dt, as the reference rounds it (:372)
      const float a = __fsub_rn(1.0f, nxt);
      o.x = __fadd_rn(__fmul_rn(a, __fsub_rn(xv.x, __fmul_rn(t, vv.x))), __fmul_rn(nxt, z.x));
      o.y = __fadd_rn(__fmul_rn(a, __fsub_rn(xv.y, __fmul_rn(t, vv.y))), __fmul_rn(nxt, z.y));
      o.z = __fadd_rn(__fmul_rn(a, __fsub_rn(xv.z, __fmul_rn(t, vv.z))), __fmul_rn(nxt, z.z));
      o.w = __fadd_rn(__fmul_rn(a, __fsub_rn(xv.w, __fmul_rn(t, vv.w))), __fmul_rn(nxt, z.w));
    } else {
      o.x = __fsub_rn(xv.x, __fmul_rn(dt, vv.x));
      o.y = __fsub_rn(xv.y, __fmul_rn(dt, vv.y));
      o.z = __fsub_rn(xv.z, __fmul_rn(dt, vv.z));
      o.w = __fsub_rn(xv.w, __fmul_rn(dt, vv.w));
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------
// ada[l, g, j, :] = table[l, j, :] + temb[g, j*D:(j+1)*D]   (bf16 add; attention.py:239-241)
// ------------------------------------------------------------------------------------------
__global__ void ada_add_kernel(const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ temb,
                               __nv_bfloat16* __restrict__ out, int L, int G, int JD /* = 6*D */) {
  const long long n8 = static_cast<long long>(L) * G * JD / 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 8;
    const int c = static_cast<int>(e % JD);
    const int g = static_cast<int>((e / JD) % G);
    const int l = static_cast<int>(e / JD / G);
    float a[8], b[8];
    load8(table + static_cast<long long>(l) * JD + c, a);
    load8(temb + static_cast<long long>(g) * JD + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    store8(out + e, a);
  }
}

// ------------------------------------------------------------------------------------------
// elementwise activation / blend helpers
// ------------------------------------------------------------------------------------------
// mode 0: copy, 1: gelu-tanh, 2: silu, 3: exact gelu   (LTXB200_ACT_*)
__global__ void act_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n, int mode) {
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 8) {
    float v[8];
    load8(x + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v[j] = mode == 2 ? __fdividef(v[j], 1.0f + __expf(-v[j])) : (mode == 1 ? gelu_tanh(v[j]) : (mode == 3 ? gelu_erf(v[j]) : v[j]));
    store8(y + i, v);
  }
}

// STG "AttentionValues" blend (attention.py:1134-1139): a[b] = a[b]*m[b] + v[b]*(1-m[b]),
// a: [B, rows, D] contiguous, v: column slice with row stride ldv.
__global__ void stg_blend_kernel(__nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ v, long long ldv,
                                 const float* __restrict__ mask, int B, long long rows, int D) {
  const long long per_b = rows * D / 8;
  const long long n8 = per_b * B;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_b);
    const float m = mask[b];
    if (m == 1.0f) continue;
    const long long e = i * 8;
    const long long r = e / D;
    const int c = static_cast<int>(e - r * D);
    float x[8], y[8];
    load8(a + e, x);
    load8(v + r * ldv + c, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = bf16r(bf16r(x[j] * m) + bf16r(y[j] * (1.0f - m)));
    store8(a + e, x);
  }
}

// sinusoidal timestep embedding (diffusers get_timestep_embedding, flip_sin_to_cos=True, shift 0):
// out[i, 0:half] = cos(t_i * f_k), out[i, half:] = sin(t_i * f_k), f_k = exp(-ln(10000) k / half)
__global__ void timestep_embed_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int n, int dim,
                                      int cos_first) {
  const int half = dim / 2;
  const int i = blockIdx.x;
  if (i >= n) return;
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float f = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
    const float a = t[i] * f;
    const float c = cosf(a), s = sinf(a);
    out[static_cast<long long>(i) * dim + k] = __float2bfloat16_rn(cos_first ? c : s);
    out[static_cast<long long>(i) * dim + half + k] = __float2bfloat16_rn(cos_first ? s : c);
  }
}

// fp32 <-> bf16 casts (vectorised)
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 8) {
    const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
    *reinterpret_cast<uint4*>(y + i) = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}

// ------------------------------------------------------------------------------------------
// Guidance (CFG with cfg-star projection, STG, std-rescale) + rectified-flow Euler step
// (pipeline_ltx_video.py:1183-1222,1309-1342 ; rf.py:350-375).  Three launches:
//   1) guidance_reduce_kernel<0>: per-block partials of  sum(text*uncond), sum(uncond^2)
//   2) guidance_reduce_kernel<1>: partials of sum/sumsq of text and of the combined prediction
//   3) guidance_step_kernel: combine, rescale, x <- x - dt*v (per-token dt, conditioning mask)
// noise_pred: [conds, n] bf16 with conds ordered (uncond?, text, perturbed?).  Partials are reduced in
// a fixed order (deterministic).
// ------------------------------------------------------------------------------------------
struct GuidanceParams {
  const __nv_bfloat16* pred;   // cond c at pred + c*cond_stride
  long long cond_stride;
  long long n;                 // elements per cond (tokens * channels)
  int channels;
  int has_cfg, has_stg, do_rescale;
  float guidance_scale, stg_scale, rescale;
  float* partials;             // [2 phases][4][kGuidanceBlocks]
  // step
  float* latents;              // [n] fp32, updated in place
  const float* timesteps;      // [num_steps] descending
  int num_steps;
  float t;                     // current global timestep
  const float* cond_mask;      // [tokens] or null
  __nv_bfloat16* latents_bf16; // optional bf16 copy of the updated latents (next model input)
  const float* noise;          // stochastic sampling (rf.py:370-373): fresh N(0,1) per element, or null (Euler step)
};
constexpr int kGuidanceBlocks = 148;

DEVI float block_sum_256(float v, float* sm) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 8) r = sm[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;   // valid in thread 0
}

DEVI float sum_partials(const float* p) {
  float s = 0.f;
  for (int i = 0; i < kGuidanceBlocks; ++i) s += p[i];
  return s;
}

DEVI float cfg_alpha(const GuidanceParams& g) {
  const float dot = bf16r(sum_partials(g.partials + 0 * kGuidanceBlocks));
  const float nrm = bf16r(bf16r(sum_partials(g.partials + 1 * kGuidanceBlocks)) + 1e-8f);
  return bf16r(dot / nrm);
}

// combined prediction for element i (before std-rescale)
DEVI float guidance_combine(const GuidanceParams& g, long long i, float alpha, float* text_out) {
  const int conds = 1 + g.has_cfg + g.has_stg;
  float out, text;
  if (g.has_cfg) {
    float un = __bfloat162float(g.pred[i]);
    text = __bfloat162float(g.pred[g.cond_stride + i]);
    if (g.guidance_scale != 0.f && g.guidance_scale != 1.f) {
      un = bf16r(alpha * un);
      out = bf16r(un + bf16r(g.guidance_scale * bf16r(text - un)));
    } else {
      out = text;
    }
  } else {
    text = __bfloat162float(g.pred[i]);
    out = text;
  }
  if (g.has_stg) {
    const float pert = __bfloat162float(g.pred[static_cast<long long>(conds - 1) * g.cond_stride + i]);
    out = bf16r(out + bf16r(g.stg_scale * bf16r(text - pert)));
  }
  *text_out = text;
  return out;
}

template <int kPhase>
__global__ void __launch_bounds__(256) guidance_reduce_kernel(const GuidanceParams g) {
  __shared__ float sm[8];
  float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
  float alpha = 0.f;
  if (kPhase == 1 && g.has_cfg) alpha = cfg_alpha(g);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < g.n; i += static_cast<long long>(gridDim.x) * 256) {
    if (kPhase == 0) {
      const float un = __bfloat162float(g.pred[i]), tx = __bfloat162float(g.pred[g.cond_stride + i]);
      a += bf16r(tx * un);
      b += bf16r(un * un);
    } else {
      float tx;
      const float o = guidance_combine(g, i, alpha, &tx);
      a += tx; b += tx * tx; c += o; d += o * o;
    }
  }
  float* P = g.partials + kPhase * 4 * kGuidanceBlocks;
  float r;
  r = block_sum_256(a, sm); if (threadIdx.x == 0) P[0 * kGuidanceBlocks + blockIdx.x] = r;
  r = block_sum_256(b, sm); if (threadIdx.x == 0) P[1 * kGuidanceBlocks + blockIdx.x] = r;
  if (kPhase == 1) {
    r = block_sum_256(c, sm); if (threadIdx.x == 0) P[2 * kGuidanceBlocks + blockIdx.x] = r;
    r = block_sum_256(d, sm); if (threadIdx.x == 0) P[3 * kGuidanceBlocks + blockIdx.x] = r;
  }
}

__global__ void __launch_bounds__(256) guidance_step_kernel(const GuidanceParams g) {
  float alpha = 0.f, factor = 1.f;
  if (g.has_cfg) alpha = cfg_alpha(g);
  if (g.has_stg && g.do_rescale && g.stg_scale > 0.f) {
    const float* P = g.partials + 4 * kGuidanceBlocks;
    const double n = static_cast<double>(g.n);
    const double st = sum_partials(P), st2 = sum_partials(P + kGuidanceBlocks);
    const double so = sum_partials(P + 2 * kGuidanceBlocks), so2 = sum_partials(P + 3 * kGuidanceBlocks);
    // unbiased std, as torch.std
    const float std_t = bf16r(static_cast<float>(sqrt(fmax((st2 - st * st / n) / (n - 1.0), 0.0))));
    const float std_o = bf16r(static_cast<float>(sqrt(fmax((so2 - so * so / n) / (n - 1.0), 0.0))));
    factor = bf16r(std_t / std_o);
    factor = bf16r(bf16r(g.rescale * factor) + (1.0f - g.rescale));
  }
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < g.n; i += static_cast<long long>(gridDim.x) * 256) {
    float tx;
    float v = guidance_combine(g, i, alpha, &tx);
    if (factor != 1.f) v = bf16r(v * factor);
    // per-token dt (rf.py:350-367): next lower timestep strictly below t_tok - 1e-6 (0 if none)
    float t_tok = g.t;
    bool update = true;
    if (g.cond_mask) {
      const float cm = g.cond_mask[i / g.channels];
      t_tok = fminf(g.t, 1.0f - cm);
      update = (g.t - 1e-6f) < (1.0f - cm);
    }
    float lower = 0.f;
    for (int s = 0; s < g.num_steps; ++s) {
      const float ts = g.timesteps[s];
      if (ts < t_tok - 1e-6f) { lower = ts; break; }     // timesteps descend: first hit is the max
    }
    const float dt = t_tok - lower;
    float x = g.latents[i];
    if (update) {
      if (g.noise) x = (1.0f - lower) * (x - t_tok * v) + lower * g.noise[i];   // add_noise(x0, noise, t - dt) (rf.py:382-392)
      else x = x - dt * v;
    }
    g.latents[i] = x;
    if (g.latents_bf16) g.latents_bf16[i] = __float2bfloat16_rn(x);
  }
}

// ------------------------------------------------------------------------------------------
// Wan CFG / CFG-Zero* combine in fp32 (text2video.py:31-42,551-562):
//   alpha = <cond, uncond> / (||uncond||^2 + 1e-8)   (only when use_alpha)
//   out   = alpha*uncond + g * (cond - alpha*uncond)
// phase 0 writes per-block partial sums (deterministic order), phase 1 combines.
// ------------------------------------------------------------------------------------------
template <int kPhase>
__global__ void __launch_bounds__(256)
cfg_combine_f32_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, float* __restrict__ out,
                       long long n, float g, int use_alpha, float* __restrict__ partials) {
  __shared__ float sm[8];
  if (kPhase == 0) {
    float a = 0.f, b = 0.f;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
      const float c = cond[i], u = uncond[i];
      a += c * u; b += u * u;
    }
    float r = block_sum_256(a, sm); if (threadIdx.x == 0) partials[blockIdx.x] = r;
    r = block_sum_256(b, sm); if (threadIdx.x == 0) partials[kGuidanceBlocks + blockIdx.x] = r;
  } else {
    float alpha = 1.f;
    if (use_alpha) alpha = sum_partials(partials) / (sum_partials(partials + kGuidanceBlocks) + 1e-8f);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
      const float u = alpha * uncond[i];
      out[i] = u + g * (cond[i] - u);
    }
  }
}

// ------------------------------------------------------------------------------------------
// VAE helpers (NDHWC bf16 activations)
// ------------------------------------------------------------------------------------------
// PixelNorm over channels (eps 1e-8) followed by SiLU (pixel_norm.py:12, causal_video_autoencoder.py:1212-1240).
// A group of G = min(32, C/8) lanes owns one voxel; NV = C / (8*G) uint4 per lane.
template <int C>
__global__ void __launch_bounds__(256)
pixelnorm_silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long voxels, float eps,
                      int apply_silu, const __nv_bfloat16* __restrict__ scale = nullptr,
                      const __nv_bfloat16* __restrict__ shift = nullptr) {
  constexpr int G = (C / 8 < 32) ? C / 8 : 32;
  constexpr int NV = C / (8 * G);
  constexpr int VPW = 32 / G;                                   // voxels per warp
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long vox = warp_global * VPW + lane / G;
  const int gl = lane % G;
  const bool ok = vox < voxels;
  float v[NV][8];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (ok) load8(x + vox * C + (i * G + gl) * 8, v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += v[i][j] * v[i][j];
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float inv = 1.0f / sqrtf(sq * (1.0f / C) + eps);
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float o[8], sc[8], sh[8];
    if (scale) {        // timestep-conditioned decoder: x * (1 + scale) + shift per channel (causal_video_autoencoder.py:1224-1237)
      load8(scale + (i * G + gl) * 8, sc);
      load8(shift + (i * G + gl) * 8, sh);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = bf16r(v[i][j] * inv);
      if (scale) t = bf16r(bf16r(t * bf16r(1.0f + sc[j])) + sh[j]);
      o[j] = apply_silu ? __fdividef(t, 1.0f + __expf(-t)) : t;
    }
    store8(y + vox * C + (i * G + gl) * 8, o);
  }
}

// LayerNorm with affine over a NARROW channel vector (C = 64 / 128: the res_x_y shortcut of the LTX VAE encoder,
// causal_video_autoencoder.py:1244-1250), G = C/8 lanes per row; rows are contiguous [rows, C].
template <int C>
__global__ void __launch_bounds__(256)
layernorm_narrow_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long rows,
                        const __nv_bfloat16* __restrict__ weight, const __nv_bfloat16* __restrict__ bias, float eps) {
  constexpr int G = C / 8;
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long row = warp_global * RPW + lane / G;
  const int gl = lane % G;
  const bool ok = row < rows;
  float v[8];
  if (ok) load8(x + row * C + gl * 8, v);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += v[j];
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / C);
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; var += d * d; }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rs = rsqrtf(var * (1.0f / C) + eps);
  if (!ok) return;
  float w[8], b[8], o[8];
  load8(weight + gl * 8, w);
  load8(bias + gl * 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rs * w[j] + b[j];
  store8(y + row * C + gl * 8, o);
}

// Wan VAE RMS_norm (wan/modules/vae.py:41-58) + optional SiLU on NDHWC voxels:
//   y = x / max(||x||_2, 1e-12) * sqrt(c_real) * gamma[c]      (F.normalize over the channel dim)
// C is the stored (64-padded) channel count, c_real the model's; pad channels hold zeros and gamma = 0 there.
__host__ __device__ constexpr int l2_group(int c8) { int g = 32; while (c8 % g) g >>= 1; return g; }   // largest power of two <= 32 dividing C/8
template <int C>
__global__ void __launch_bounds__(256)
l2norm_silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long voxels,
                   const __nv_bfloat16* __restrict__ gamma, float scale, int apply_silu) {
  constexpr int G = l2_group(C / 8);
  constexpr int NV = C / (8 * G);
  constexpr int VPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long vox = warp_global * VPW + lane / G;
  const int gl = lane % G;
  const bool ok = vox < voxels;
  float v[NV][8];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (ok) load8(x + vox * C + (i * G + gl) * 8, v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += v[i][j] * v[i][j];
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float inv = scale / fmaxf(sqrtf(sq), 1e-12f);
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float g[8], o[8];
    load8(gamma + (i * G + gl) * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = bf16r(v[i][j] * inv * g[j]);
      o[j] = apply_silu ? __fdividef(t, 1.0f + __expf(-t)) : t;
    }
    store8(y + vox * C + (i * G + gl) * 8, o);
  }
}

// nearest x2 spatial upsample of NDHWC frames (Upsample(scale_factor=(2,2), mode='nearest-exact'), vae.py:61-88):
// out[f, 2h+a, 2w+b, :] = in[f, h, w, :].  One thread per 16 B of OUTPUT.
__global__ void upsample2x_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long frames,
                                       int H, int W, int C) {
  const int c8 = C / 8;
  const long long n = frames * (2LL * H) * (2LL * W) * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c8);
    long long r = i / c8;
    const int ow = static_cast<int>(r % (2 * W)); r /= 2 * W;
    const int oh = static_cast<int>(r % (2 * H));
    const long long f = r / (2 * H);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((f * H + (oh >> 1)) * W + (ow >> 1)) * C) + c);
    reinterpret_cast<uint4*>(y)[i] = v;
  }
}

// row softmax: P[r, :] = softmax(scale * S[r, :]) ; S fp32 [rows, ld_s], P bf16 [rows, ld_p]; one warp per row.
// (the single 384-wide head of the Wan VAE AttentionBlock, vae.py:249-272, does not fit the flash kernel's head dims)
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, int rows, int cols, long long ld_s,
                    long long ld_p, float scale_log2) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = S + row * ld_s;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, s[c]);
  m = warp_max(m) * scale_log2;
  float l = 0.f;
  for (int c = lane; c < cols; c += 32) l += fast_exp2(fmaf(s[c], scale_log2, -m));
  l = warp_sum(l);
  const float inv = 1.0f / l;
  __nv_bfloat16* pr = P + row * ld_p;
  for (int c = lane; c < cols; c += 32) pr[c] = __float2bfloat16_rn(fast_exp2(fmaf(s[c], scale_log2, -m)) * inv);
}

// latents [B, C, F, H, W] fp32/bf16 (NCDHW) -> per-channel de-normalise (x*std+mean) -> NDHWC bf16
// (vae_encode.py:239-247).  Small tensor: one thread per output element.
template <typename T>
__global__ void latent_to_ndhwc_kernel(const T* __restrict__ z, __nv_bfloat16* __restrict__ out, int B, int C,
                                       long long FHW, const float* __restrict__ stdv, const float* __restrict__ meanv) {
  const long long n = static_cast<long long>(B) * C * FHW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long s = (i / C) % FHW;
    const long long b = i / C / FHW;
    float v = static_cast<float>(z[(b * C + c) * FHW + s]);
    if (stdv) v = bf16r(bf16r(bf16r(v) * bf16r(stdv[c])) + bf16r(meanv[c]));
    out[i] = __float2bfloat16_rn(v);
  }
}

// out = bf16(a*x + b*y): TeaCache residual bookkeeping (wan/modules/model.py:1051-1054 `x += previous_residual`,
// :1090-1099 `torch.sub(x, ori)`); fp32 arithmetic, one rounding, like ATen's bf16 add/sub.  out may alias x or y.
__global__ void axpby_kernel(const __nv_bfloat16* x, const __nv_bfloat16* y, __nv_bfloat16* out, long long n, float a, float b) {
  for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 8; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 8) {
    float u[8], v[8];
    load8(x + i, u);
    load8(y + i, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = fmaf(b, v[j], a * u[j]);
    store8(out + i, u);
  }
}

// TeaCache distance statistics (model.py:1039): out[0] = sum |bf16(a - b)|, out[1] = sum |b| over n elements
// (one 256-thread block: the operands are the [1, dim] time embeddings of two consecutive steps).
__global__ void rel_l1_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n,
                              float* __restrict__ out) {
  __shared__ float sm[8];
  float d = 0.f, m = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = __bfloat162float(a[i]), y = __bfloat162float(b[i]);
    d += fabsf(bf16r(x - y));
    m += fabsf(y);
  }
  d = block_sum_256(d, sm);
  m = block_sum_256(m, sm);
  if (threadIdx.x == 0) { out[0] = d; out[1] = m; }
}

// ------------------------------------------------------------------------------------------
// LTX multi-scale flow: GroupNorm(32)+SiLU of the LatentUpsampler (latent_upsampler.py:15-39,73-75), AdaIN latent
// filter (pipeline_ltx_video.py:1709-1737), latent re-normalisation, bilinear frame resize (:1890-1901).  The tensors are a few
// MB (latent resolution), L2-resident: these kernels are latency-bound, two launches per GroupNorm, deterministic partial sums.
// ------------------------------------------------------------------------------------------
constexpr int kGnGroups = 32;

// x [B, voxels, C] bf16 (NDHWC), C % 256 == 0.  grid (chunks, B), 256 threads; partial [B, chunks, 32, 2] = (sum, sum of squares)
__global__ void groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, long long voxels, int C) {
  __shared__ float sh[2][256];
  const int tpv = C / 8, vpb = 256 / tpv;
  const int slot = threadIdx.x % tpv, vrow = threadIdx.x / tpv;
  const long long per = (voxels + gridDim.x - 1) / gridDim.x;
  const long long v0 = blockIdx.x * per, v1 = (v0 + per < voxels) ? v0 + per : voxels;
  const __nv_bfloat16* xb = x + static_cast<long long>(blockIdx.y) * voxels * C;
  float s = 0.f, q = 0.f;
  for (long long v = v0 + vrow; v < v1; v += vpb) {
    float f[8];
    load8(xb + v * C + slot * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += f[j]; q = fmaf(f[j], f[j], q); }
  }
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < kGnGroups) {
    const int tpg = tpv / kGnGroups;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < vpb; ++r)
      for (int k = 0; k < tpg; ++k) { const int i = r * tpv + threadIdx.x * tpg + k; a += sh[0][i]; b += sh[1][i]; }
    float* o = partial + ((static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * kGnGroups + threadIdx.x) * 2;
    o[0] = a; o[1] = b;
  }
}

// y = [silu]( bf16(groupnorm(x) * gamma + beta) [+ residual] ) with ATen's rounding points (norm output, sum, activation each bf16)
__global__ void groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                       const float* __restrict__ partial, int chunks, long long voxels, int C,
                                       const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                       const __nv_bfloat16* __restrict__ residual, float eps, int apply_silu) {
  __shared__ float mean[kGnGroups], rstd[kGnGroups];
  if (threadIdx.x < kGnGroups) {
    float a = 0.f, b = 0.f;
    const float* pp = partial + (static_cast<long long>(blockIdx.y) * chunks * kGnGroups + threadIdx.x) * 2;
    for (int c = 0; c < chunks; ++c) { a += pp[c * kGnGroups * 2]; b += pp[c * kGnGroups * 2 + 1]; }
    const float n = static_cast<float>(voxels) * (C / kGnGroups);
    const float m = a / n;
    mean[threadIdx.x] = m;
    rstd[threadIdx.x] = rsqrtf(fmaxf(b / n - m * m, 0.f) + eps);
  }
  __syncthreads();
  const int tpv = C / 8, vpb = 256 / tpv;
  const int slot = threadIdx.x % tpv, vrow = threadIdx.x / tpv;
  const int g = slot / (tpv / kGnGroups);
  const float m = mean[g], r = rstd[g];
  float ga[8], be[8];
  load8(gamma + slot * 8, ga);
  load8(beta + slot * 8, be);
  const long long base = static_cast<long long>(blockIdx.y) * voxels * C;
  for (long long v = static_cast<long long>(blockIdx.x) * vpb + vrow; v < voxels; v += static_cast<long long>(gridDim.x) * vpb) {
    const long long e = base + v * C + slot * 8;
    float f[8];
    load8(x + e, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = bf16r(fmaf((f[j] - m) * r, ga[j], be[j]));
    if (residual) {
      float h[8];
      load8(residual + e, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = bf16r(f[j] + h[j]);
    }
    if (apply_silu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
    }
    store8(y + e, f);
  }
}

// AdaIN: one block per (batch, channel) row.  x [rows, n], ref [rows, m] fp32; out = lerp(x, (x - mean_x)/std_x * std_ref + mean_ref, factor)
// with torch.std_mean's unbiased std (two passes over the L2-resident row).
DEVI float block_sum_256_all(float v, float* sm, float* bcast) {      // the sum, in every thread
  const float r = block_sum_256(v, sm);
  if (threadIdx.x == 0) *bcast = r;
  __syncthreads();
  const float out = *bcast;
  __syncthreads();
  return out;
}

__global__ void adain_kernel(const float* __restrict__ x, const float* __restrict__ ref, float* __restrict__ out, long long n,
                             long long m, float factor) {
  __shared__ float sm[8];
  __shared__ float bc;
  const float* xr = x + blockIdx.x * n;
  const float* rr = ref + blockIdx.x * m;
  float a = 0.f, b = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) a += xr[i];
  for (long long i = threadIdx.x; i < m; i += 256) b += rr[i];
  const float mx = block_sum_256_all(a, sm, &bc) / n, mr = block_sum_256_all(b, sm, &bc) / m;
  a = 0.f; b = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) { const float d = xr[i] - mx; a = fmaf(d, d, a); }
  for (long long i = threadIdx.x; i < m; i += 256) { const float d = rr[i] - mr; b = fmaf(d, d, b); }
  const float sx = sqrtf(block_sum_256_all(a, sm, &bc) / (n - 1)), sr = sqrtf(block_sum_256_all(b, sm, &bc) / (m - 1));
  float* o = out + blockIdx.x * n;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float v = xr[i], t = (v - mx) / sx * sr + mr;
    o[i] = factor < 0.5f ? v + factor * (t - v) : t - (t - v) * (1.0f - factor);        // torch.lerp
  }
}

// NDHWC bf16 -> NCDHW fp32 with the inverse of latent_to_ndhwc's affine: (x - mean[c]) / std[c]  (normalize_latents, vae_encode.py:228-237)
__global__ void latent_from_ndhwc_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int B, int C, long long FHW,
                                         const float* __restrict__ stdv, const float* __restrict__ meanv) {
  const long long n = static_cast<long long>(B) * C * FHW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long s = i % FHW;
    const int c = static_cast<int>((i / FHW) % C);
    const long long b = i / FHW / C;
    float v = __bfloat162float(x[(b * FHW + s) * C + c]);
    if (stdv) v = (v - meanv[c]) / stdv[c];
    out[i] = v;
  }
}

// F.interpolate(mode="bilinear", align_corners=False) on [planes, h, w] -> [planes, H, W] fp32
__global__ void bilinear_resize_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int h, int w, int H, int W) {
  const float sh = static_cast<float>(h) / H, sw = static_cast<float>(w) / W;
  const long long n = planes * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % W), oy = static_cast<int>((i / W) % H);
    const long long p = i / W / H;
    const float fy = fmaxf((oy + 0.5f) * sh - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * sw - 0.5f, 0.f);
    const int y0 = min(static_cast<int>(fy), h - 1), x0 = min(static_cast<int>(fx), w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const float* xp = x + p * h * w;
    const float top = xp[y0 * w + x0] * (1.f - lx) + xp[y0 * w + x1] * lx;
    const float bot = xp[y1 * w + x0] * (1.f - lx) + xp[y1 * w + x1] * lx;
    y[i] = top * (1.f - ly) + bot * ly;
  }
}

}  // namespace b200
