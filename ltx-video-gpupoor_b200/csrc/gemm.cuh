// Persistent warp-specialised bf16 GEMM / implicit-GEMM conv3d for sm_100a.
//
//   D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue),  fp32 accumulation in TMEM.
//
// One CTA per SM, 192 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> global).  Operands are staged by TMA into 128B-swizzled
// shared memory (K-major for both), kStages-deep mbarrier ring; the accumulator is double-buffered in
// TMEM (2 x BN columns) so tile i's epilogue overlaps tile i+1's main loop.
//
// kConv = true turns the A producer into an implicit-GEMM gather for a 3x3x3 convolution over an
// NDHWC activation: the M tile is a BH x BW spatial patch of one frame, each k-block is one
// (tap, 64-channel slice); spatial zero padding comes from TMA out-of-bounds fill, temporal
// replicate padding from clamping the frame coordinate (causal_conv3d.py:44-59).
#pragma once
#include "common.cuh"
#include "comm.cuh"

namespace b200 {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;

enum GemmAct : int { kActNone = 0, kActGeluTanh = 1, kActSilu = 2, kActGeluErf = 3 };
enum GemmStore : int {
  kStoreRowMajor = 0,   // out[m*ldc + n]
  kStoreConvD2S = 1,    // conv: depth-to-space 2x2x2 (+ drop first frame), channel order (p1,p2,p3,c)
  kStoreConvUnpatch = 2 // conv: final unpatchify to NCFHW, channel order (c,q,r), patch 4x4
};

struct GemmParams {
  int M, N, K;
  // epilogue
  void* out;
  long long ldc;
  const __nv_bfloat16* bias;       // [N] or null
  int act;
  const __nv_bfloat16* residual;   // [M, ldr] or null:   out = residual + gate * val
  long long ldr;
  const __nv_bfloat16* gate;       // [ceil(M/rows_per_gate), gate_ld] or null
  int rows_per_gate;
  long long gate_ld;
  // Ulysses sequence parallelism: the columns >= vs_col0 of a row-major bf16 output (the V third of the fused QKV projection, which needs
  // no normalisation) are stored straight into the receive buffer of the rank that owns their heads, laid out [N, B, 3, H/P, d] in global
  // token order (elementwise.cuh, qk_norm_rope_wan_scatter_kernel): one third of the exchange leaves from this epilogue, spread over the
  // GEMM's run time, instead of being re-read and stored by the scatter kernel.  vs_peers.P == 0: off.
  PeerPtrs vs_peers;
  int vs_col0, vs_group_cols, vs_B, vs_tokens_per_batch, vs_token_offset, vs_row0;
  // `mixed` precision: fp32 residual stream and fp32 gate (same strides, in elements); used instead of the bf16 pointers when set
  const float* residual32;
  const float* gate32;
  int out_f32;
  int store_mode;
  // fused PixelNorm (+ SiLU) of the OUTPUT row (row-major store, the whole channel vector in one N tile: N <= BN):
  //   norm_mode 1: `out` as usual and out2 = silu(pixelnorm(bf16(out row)));   2: only out2 (the raw row is not stored)
  // (ResnetBlock3D.norm1 / norm2 + non_linearity of the LTX VAE, causal_video_autoencoder.py:1212-1240, in the producing conv)
  int norm_mode;
  __nv_bfloat16* out2;             // [M, N] bf16, row stride N
  float norm_eps;
  // conv geometry (kConv only): activation [B, T, H, W, Cin], tile = BH x BW patch
  int cB, cT, cH, cW, cCin;
  int cBH, cBW;
  int c_tiles_h, c_tiles_w;
  int c_causal;                    // 1: taps (t-2,t-1,t) ; 0: (t-1,t,t+1)
  int c_taps_t;                    // temporal taps: 3 or 1
  int c_taps_hw;                   // spatial taps per axis: 3 (zero padding) or 1
  int c_tpad_zero;                 // temporal padding: 0 = replicate (clamped frame index, LTX), 1 = zeros (TMA OOB fill, Wan)
  int c_st, c_shw;                 // output strides (1 or 2): cT/cH/cW are OUTPUT dims, the TMA map strides over the input
  int cTin;                        // input frames (temporal clamp)
  int c_off_hw;                    // added to the spatial tap coordinate: 0 = centred taps (h-1,h,h+1); 1 = taps (h,h+1,h+2), i.e.
                                   // nn.ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) of the Wan encoder (wan/modules/vae.py:90-93)
  // tile rasterisation of the persistent schedule: 0 = M fastest (one weight panel per wave, all of A re-read per
  // N tile), 1 = N fastest (a wave = a few row panels x every N tile: A streams through once, W stays in L2)
  int n_fastest;
};

template <int BN, int kCtas = 1>
struct GemmSmem {
  static constexpr int kStageBytesA = kGemmBM * kGemmBK * 2;          // 16 KB
  static constexpr int kStageBytesB = (BN / kCtas) * kGemmBK * 2;     // CTA pair: each CTA holds half of the B tile
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = ((BN == 256 && kCtas == 1) || BN == 512) ? 4 : 6;      // BN = 512 (CTA pair): 4 x 48 KB
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024 alignment slack
};

// Epilogue of one accumulator row (one thread = one TMEM lane = one output row / voxel): BN fp32 columns at t_addr -> bias /
// activation / gate / residual -> store in the requested layout.  Shared by the GEMM / implicit-GEMM kernel below and the
// halo-tiled convolution (conv_halo.cuh).
template <int BN>
DEVI void epilogue_row(const GemmParams& p, uint32_t t_addr, int tn, bool row_ok, long long m_lin, int ob, int ot, int oh, int ow) {
  const __nv_bfloat16* gate_row = nullptr;
  if (p.gate) gate_row = p.gate + (m_lin / p.rows_per_gate) * p.gate_ld;
  const __nv_bfloat16* res_row = p.residual ? p.residual + m_lin * p.ldr : nullptr;
  const float* gate32_row = p.gate32 ? p.gate32 + (m_lin / p.rows_per_gate) * p.gate_ld : nullptr;
  const float* res32_row = p.residual32 ? p.residual32 + m_lin * p.ldr : nullptr;
  float sumsq = 0.f;

  // One 32-column chunk of the accumulator row: bias / activation / gate / residual / store.  The chunk's TMEM load and its
  // bias slice were requested one chunk earlier (see the driver loop below), so neither latency is exposed here.
  auto process = [&](const uint32_t (&v)[32], const uint4 (&bias4)[4], int c0) {
    const int n0 = tn * BN + c0;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    const bool full = (n0 + 32 <= p.N);
    if (row_ok) {
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (full || n0 + j < p.N) {
            const uint4 bq = bias4[j >> 3];
            const float2 b0 = unpack_bf16(bq.x), b1 = unpack_bf16(bq.y), b2 = unpack_bf16(bq.z), b3 = unpack_bf16(bq.w);
            // packed adds (FADD2): the epilogue runs one warp per scheduler, every instruction it does not issue is latency it does not pay
            unpack_f32x2(add_f32x2(pack_f32x2(f[j], f[j + 1]), pack_f32x2(b0.x, b0.y)), f[j], f[j + 1]);
            unpack_f32x2(add_f32x2(pack_f32x2(f[j + 2], f[j + 3]), pack_f32x2(b1.x, b1.y)), f[j + 2], f[j + 3]);
            unpack_f32x2(add_f32x2(pack_f32x2(f[j + 4], f[j + 5]), pack_f32x2(b2.x, b2.y)), f[j + 4], f[j + 5]);
            unpack_f32x2(add_f32x2(pack_f32x2(f[j + 6], f[j + 7]), pack_f32x2(b3.x, b3.y)), f[j + 6], f[j + 7]);
          }
        }
      }
      if (p.act == kActGeluTanh) {
#pragma unroll
#ifdef LTXB200_GELU_SCALAR
        for (int j = 0; j < 32; ++j) f[j] = gelu_tanh(f[j]);
#else
        for (int j = 0; j < 32; j += 2) unpack_f32x2(gelu_tanh_f32x2(pack_f32x2(f[j], f[j + 1])), f[j], f[j + 1]);
#endif
      } else if (p.act == kActSilu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __fdividef(f[j], 1.0f + __expf(-f[j]));
      } else if (p.act == kActGeluErf) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
      }
      if (gate_row) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (full || n0 + j < p.N) {
            const uint4 gq = *reinterpret_cast<const uint4*>(gate_row + n0 + j);
            const float2 g0 = unpack_bf16(gq.x), g1 = unpack_bf16(gq.y), g2 = unpack_bf16(gq.z), g3 = unpack_bf16(gq.w);
            f[j] *= g0.x; f[j + 1] *= g0.y; f[j + 2] *= g1.x; f[j + 3] *= g1.y;
            f[j + 4] *= g2.x; f[j + 5] *= g2.y; f[j + 6] *= g3.x; f[j + 7] *= g3.y;
          }
        }
      }
      if (gate32_row) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (full || n0 + j < p.N) {
            const float4 g4 = *reinterpret_cast<const float4*>(gate32_row + n0 + j);
            f[j] *= g4.x; f[j + 1] *= g4.y; f[j + 2] *= g4.z; f[j + 3] *= g4.w;
          }
        }
      }
      if (res32_row) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (full || n0 + j < p.N) {
            const float4 r4 = *reinterpret_cast<const float4*>(res32_row + n0 + j);
            f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
          }
        }
      }
      if (res_row) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (full || n0 + j < p.N) {
            const uint4 rq = *reinterpret_cast<const uint4*>(res_row + n0 + j);
            const float2 r0 = unpack_bf16(rq.x), r1 = unpack_bf16(rq.y), r2 = unpack_bf16(rq.z), r3 = unpack_bf16(rq.w);
            f[j] += r0.x; f[j + 1] += r0.y; f[j + 2] += r1.x; f[j + 3] += r1.y;
            f[j + 4] += r2.x; f[j + 5] += r2.y; f[j + 6] += r3.x; f[j + 7] += r3.y;
          }
        }
      }
    }
    if (p.norm_mode) {                         // warp-uniform; the TMEM store is warp-collective, so rows outside the tensor take part too
      // the statistic is taken over the values as they are stored (bf16), like the standalone kernel that read them back
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        f[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
        if (full || n0 + j < p.N) sumsq = fmaf(f[j], f[j], sumsq);
      }
      uint32_t wb[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) wb[j] = __float_as_uint(f[j]);
      tmem_st32(t_addr + c0, wb);              // kept in the accumulator's own columns for the second pass
    }
    if (row_ok && p.norm_mode != 2) {
      // ---- store ----
      if (p.store_mode == kStoreRowMajor) {
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + m_lin * p.ldc + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (full || n0 + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + m_lin * p.ldc + n0;
          if (p.vs_peers.P > 0 && n0 >= p.vs_col0) {         // warp-uniform: a 32-column chunk never straddles a head group
            const long long row = p.vs_row0 + m_lin;
            const int b = static_cast<int>(row / p.vs_tokens_per_batch);
            const long long T = p.vs_token_offset + (row - static_cast<long long>(b) * p.vs_tokens_per_batch);
            const int c = n0 - p.vs_col0, g = c / p.vs_group_cols;
            o = static_cast<__nv_bfloat16*>(p.vs_peers.data[g]) + ((T * p.vs_B + b) * 3 + 2) * p.vs_group_cols + (c - g * p.vs_group_cols);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (full || n0 + j < p.N)
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                            pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
        }
      } else if (p.store_mode == kStoreConvD2S) {
        // N = 8*C, channel n = ((p1*2+p2)*2+p3)*C + c  ->  out[b, 2t+p1-1, 2h+p2, 2w+p3, c]  (NDHWC, T' = 2T-1)
        const int C = p.N >> 3;
        const int q = n0 / C, c = n0 - q * C;             // 32-col chunk never straddles (C % 32 == 0)
        const int p1 = q >> 2, p2 = (q >> 1) & 1, p3 = q & 1;
        const int t2 = 2 * ot + p1 - 1;
        if (t2 >= 0) {
          const long long vox = ((static_cast<long long>(ob) * (2 * p.cT - 1) + t2) * (2 * p.cH) + (2 * oh + p2)) * (2 * p.cW) + (2 * ow + p3);
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + vox * C + c;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                          pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
        }
      } else {
        // kStoreConvUnpatch: N = Cimg*16, n = (c*4+q)*4 + r -> out[b, c, t, 4h+q, 4w+r]  (NCFHW, fp32 or bf16)
        const long long HW4 = static_cast<long long>(4 * p.cH) * (4 * p.cW);
        const int Cimg = p.N >> 4;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + j;
          if (n < p.N) {
            const int c = n >> 4, q = (n >> 2) & 3;
            const long long off = ((static_cast<long long>(ob) * Cimg + c) * p.cT + ot) * HW4 +
                                  static_cast<long long>(4 * oh + q) * (4 * p.cW) + 4 * ow;
            if (p.out_f32)
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            else
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) =
                  make_uint2(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]));
          }
        }
      }
    }
  };
  auto load_bias = [&](uint4 (&bias4)[4], int c0) {
    const int n0 = tn * BN + c0;
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        if (n0 + j < p.N) bias4[j >> 3] = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + j));
    }
  };
  // driver: two chunks per iteration on alternating register sets; while chunk c is processed the TMEM load and the bias
  // slice of chunk c+1 are in flight (tcgen05.wait::ld covers every outstanding load, so the next one is issued after it)
  const int ncols = (p.N - tn * BN) < BN ? (p.N - tn * BN) : BN;          // warp-uniform
  uint32_t va[32], vb[32];
  uint4 ba[4], bb[4];
  tmem_ld32(t_addr, va);
  load_bias(ba, 0);
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 64) {
    tmem_wait_ld();
    if (c0 + 32 < ncols) { tmem_ld32(t_addr + c0 + 32, vb); load_bias(bb, c0 + 32); }
    process(va, ba, c0);
    if (c0 + 32 < ncols) {
      tmem_wait_ld();
      if (c0 + 64 < ncols) { tmem_ld32(t_addr + c0 + 64, va); load_bias(ba, c0 + 64); }
      process(vb, bb, c0 + 32);
    }
  }
  if (p.norm_mode) {
    // second pass: x * rsqrt(mean(x^2) + eps) -> bf16 -> SiLU -> out2, from the rounded values parked in TMEM
    tmem_wait_st();
    const float inv = 1.0f / sqrtf(sumsq / static_cast<float>(p.N) + p.norm_eps);
    __nv_bfloat16* o2 = p.out2 + m_lin * p.N;
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      tmem_ld32(t_addr + c0, va);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (c0 + j < p.N) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = __bfloat162float(__float2bfloat16_rn(__uint_as_float(va[j + e]) * inv));
              o[e] = __fdividef(t, 1.0f + __expf(-t));
            }
            *reinterpret_cast<uint4*>(o2 + c0 + j) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
          }
        }
      }
    }
  }
}

// kCtas = 2: the tile is 256 x BN on a CTA PAIR (for the convolution: two consecutive 128-voxel patches) (cluster of 2, tcgen05 cta_group::2).  Each CTA loads its 128 rows
// of A and half of the B tile (so the shared-memory fill and operand traffic per SM drop by a third), the leader issues
// M = 256 MMAs whose accumulator halves land in each CTA's own TMEM, and each CTA runs the epilogue of its 128 rows.
template <int BN, bool kConv, int kCtas = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  static_assert(kCtas == 1 || kCtas == 2, "one CTA or a CTA pair");
  static_assert(BN <= 256 || (BN == 512 && kCtas == 2 && !kConv), "the 256 x 512 tile exists for the plain CTA-pair GEMM only");
  // BN = 512: a 256 x 512 pair tile — the accumulator takes all 512 TMEM columns of each CTA, so it is SINGLE-buffered (the epilogue of
  // a tile is not overlapped with the next mainloop) in exchange for a third less operand traffic per flop: A 16 KB + B 32 KB per CTA and
  // k-block for 8.4 MFLOP = 175 flop/B against 131 for 256 x 256.  The GEMMs of the step run at the L2 -> SM limit (11-12 TB/s,
  // profiles/r02_ncu_summary.md §8), so this pays where the mainloop is long against the epilogue: K >= 4096 (FFN-down).
  constexpr int kAccBufs = (BN == 512) ? 1 : 2;
  constexpr int kMmaN = (BN > 256) ? 256 : BN;          // columns per tcgen05.mma (N <= 256)
  constexpr int kNSub = BN / kMmaN;
  using S = GemmSmem<BN, kCtas>;
  const uint32_t cta_rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const int cta_first = blockIdx.x / kCtas, cta_stride = gridDim.x / kCtas;     // persistent walk in units of CTA groups
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = kConv ? (p.cB * p.cT * p.c_tiles_h * p.c_tiles_w + kCtas - 1) / kCtas : (p.M + kGemmBM * kCtas - 1) / (kGemmBM * kCtas);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kb_per_tap = kConv ? (p.cCin + kGemmBK - 1) / kGemmBK : 1;
  const int taps = kConv ? p.c_taps_t * p.c_taps_hw * p.c_taps_hw : 1;
  const int num_kb = kConv ? taps * kb_per_tap : (p.K + kGemmBK - 1) / kGemmBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * kCtas);        // the leader's MMA thread waits for the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCtas == 2) tmem_alloc_2cta<kAccBufs * BN>(tmem_slot);
    else tmem_alloc<kAccBufs * BN>(tmem_slot);
  }
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all();              // the peer must not signal barriers that are not initialised yet
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {     // elect.sync (not lane == 0): ptxas then knows a single lane runs this and issues the
                           // uniform-datapath TMA / MMA instructions directly instead of through per-lane ELECT loops
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta_first; tile < num_tiles; tile += cta_stride) {
        const int tm = (p.n_fastest ? tile / tiles_n : tile % tiles_m) * kCtas + static_cast<int>(cta_rank);   // this CTA's 128-row block
        const int tn = p.n_fastest ? tile % tiles_n : tile / tiles_m;
        int cb = 0, ct = 0, ch0 = 0, cw0 = 0;
        if (kConv) {
          int r = tm;                            // an odd patch count leaves the last pair's second CTA with cb == cB: TMA zero-fills
          cw0 = (r % p.c_tiles_w) * p.cBW; r /= p.c_tiles_w;
          ch0 = (r % p.c_tiles_h) * p.cBH; r /= p.c_tiles_h;
          ct = r % p.cT; cb = r / p.cT;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kStageBytesA;
          // CTA pair: both CTAs' bytes complete on the LEADER's barrier (the MMAs are issued there); only the leader arms it
          if (kCtas == 1 || cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], kCtas * S::kStageBytes);
          const int brow = tn * BN + static_cast<int>(cta_rank) * (BN / kCtas);      // this CTA's part of the B tile
          if (kConv) {
            const int tap = kb / kb_per_tap, cblk = kb - tap * kb_per_tap;
            const int thw = p.c_taps_hw * p.c_taps_hw;
            const int kt = tap / thw, kh = (tap % thw) / p.c_taps_hw, kw = tap % p.c_taps_hw;     // tap-major K: (kt, kh, kw, ci)
            int tt = ct, hh = ch0, ww = cw0;
            tt = ct * p.c_st; hh = ch0 * p.c_shw; ww = cw0 * p.c_shw;
            if (p.c_taps_t == 3) {
              tt += p.c_causal ? kt - 2 : kt - 1;
              if (!p.c_tpad_zero) tt = tt < 0 ? 0 : (tt > p.cTin - 1 ? p.cTin - 1 : tt);        // replicate; else TMA zero-fills t < 0
            }
            if (p.c_taps_hw == 3) {
              hh += kh - 1 + p.c_off_hw;
              ww += kw - 1 + p.c_off_hw;
            }
            if (kCtas == 2) {
              tma_load_5d_2sm(sa, &tmA, &full_bar[stage], cblk * kGemmBK, ww, hh, tt, cb);
              tma_load_2d_2sm(sb, &tmB, &full_bar[stage], tap * p.cCin + cblk * kGemmBK, brow);
            } else {
              tma_load_5d(sa, &tmA, &full_bar[stage], cblk * kGemmBK, ww, hh, tt, cb);
              tma_load_2d(sb, &tmB, &full_bar[stage], tap * p.cCin + cblk * kGemmBK, brow);
            }
          } else if (kCtas == 2) {
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * kGemmBK, tm * kGemmBM);
            if (kNSub == 1) {
              tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * kGemmBK, brow);
            } else {
              // each MMA reads columns [0, 128) of its 256 from the leader's B rows and [128, 256) from the peer's: CTA r keeps the B
              // rows {sub * 256 + r * 128 ...} of the tile, so that accumulator column c is output column c
#pragma unroll
              for (int sub = 0; sub < kNSub; ++sub)
                tma_load_2d_2sm(sb + sub * (kMmaN / 2) * kGemmBK * 2, &tmB, &full_bar[stage], kb * kGemmBK,
                                tn * BN + sub * kMmaN + static_cast<int>(cta_rank) * (kMmaN / 2));
            }
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kGemmBK, tm * kGemmBM);
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * kGemmBK, brow);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (CTA pair: the leader only) =================
    const int issue_tiles = (kCtas == 2 && cta_rank != 0) ? 0 : num_tiles;      // the peer's MMA warp only allocates / frees TMEM
    constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM * kCtas, kMmaN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cta_first; tile < issue_tiles; tile += cta_stride) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t sb = sa + S::kStageBytesA;
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            const uint64_t ad = umma_smem_desc_sw128(sa + k * 32, 16, 1024);
#pragma unroll
            for (int sub = 0; sub < kNSub; ++sub) {
              const uint64_t bd = umma_smem_desc_sw128(sb + sub * (kMmaN / kCtas) * kGemmBK * 2 + k * 32, 16, 1024);
              if (kCtas == 2) umma_ss_2cta(d_tmem + sub * kMmaN, ad, bd, idesc, (kb | k) ? 1u : 0u);
              else umma_ss(d_tmem + sub * kMmaN, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
          }
          if (kCtas == 2) {
            umma_commit_2cta(&empty_bar[stage], 3u);                       // frees the slot in BOTH CTAs
            if (kb == num_kb - 1) umma_commit_2cta(&tfull_bar[acc], 3u);   // both epilogues
          } else {
            umma_commit(&empty_bar[stage]);                       // frees the smem slot when MMAs retire
            if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator ready for the epilogue
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue =================
    const int sub = warp & 3;                 // TMEM sub-partition this warp may read
    const int row_in_tile = sub * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cta_first; tile < num_tiles; tile += cta_stride) {
      const int tm = (p.n_fastest ? tile / tiles_n : tile % tiles_m) * kCtas + static_cast<int>(cta_rank);   // this CTA's 128-row block
      const int tn = p.n_fastest ? tile % tiles_n : tile / tiles_m;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(sub * 32) << 16);

      // ---- where does this thread's row go? ----
      bool row_ok;
      long long m_lin = 0;                    // linear row index (for bias-free row-major / gate / residual)
      int ob = 0, ot = 0, oh = 0, ow = 0;
      if (kConv) {
        int r = tm;
        const int w0 = (r % p.c_tiles_w) * p.cBW; r /= p.c_tiles_w;
        const int h0 = (r % p.c_tiles_h) * p.cBH; r /= p.c_tiles_h;
        ot = r % p.cT; ob = r / p.cT;
        oh = h0 + row_in_tile / p.cBW;
        ow = w0 + row_in_tile % p.cBW;
        row_ok = (oh < p.cH) && (ow < p.cW) && (ob < p.cB);
        m_lin = ((static_cast<long long>(ob) * p.cT + ot) * p.cH + oh) * p.cW + ow;
      } else {
        m_lin = static_cast<long long>(tm) * kGemmBM + row_in_tile;
        row_ok = m_lin < p.M;
      }
      epilogue_row<BN>(p, t_addr, tn, row_ok, m_lin, ob, ot, oh, ow);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 2 && cta_rank != 0) mbar_arrive_remote(&tempty_bar[acc], 0);
        else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all();              // the peer may still multicast into this CTA's barriers / read its shared memory
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kCtas == 2) tmem_dealloc_2cta<kAccBufs * BN>(tmem_base);
    else tmem_dealloc<kAccBufs * BN>(tmem_base);
  }
}

}  // namespace b200
