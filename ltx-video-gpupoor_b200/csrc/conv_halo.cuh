// Halo-tiled implicit-GEMM 3x3x3 (or 1x3x3) convolution for narrow outputs (Cout <= 128), sm_100a.
//
// Why: ncu on the plain implicit GEMM (gemm.cuh, kConv) at the last stage of the LTX VAE decoder (Cin = Cout = 128,
// 121x128x192 voxels: 8 launches = 56 % of the decode) shows the tensor pipe 30 % active and 41.9 GB per launch crossing
// L2 -> SM at 10.2 TB/s = 5700 B/clk, the chip-wide L2 cap (profiles/r02_ncu_summary.md): every 128-voxel tile re-reads
// its activation window once per tap (27 x 16 KB) and streams the whole weight tensor (54 x 16 KB) for 113 MFLOP, 65 flop/B.
// This kernel raises that to ~150 flop/B so the same convolution is bound by the tensor pipe instead:
//   * the output tile is a 16 x 16 spatial patch of one frame = TWO 128-row accumulators (left / right 16 x 8 half) that share
//     every weight tile in shared memory (weights cross L2 -> SM once per 256 voxels, not once per 128);
//   * for each (temporal tap, 64-channel slice, kw) ONE activation box of 18 rows x 16 columns (the patch plus its vertical halo,
//     shifted by kw - 1 horizontally) is loaded by TMA and serves the three vertical taps kh = 0,1,2 of both halves: rows are
//     stored [h][w] with 16 w per h, so the half `sub`, tap `kh` operand is the 1024-byte-aligned window starting at
//     (kh * 16 + sub * 8) * 128 B with 8-row groups 2048 B apart (the UMMA descriptor's stride-byte-offset) -- no copies,
//     no unaligned swizzle atoms.  Activation traffic per tile: 18 boxes x 36 KB instead of 2 x 54 x 16 KB.
// Spatial zero padding = TMA out-of-bounds fill (the box may start at -1), temporal padding as in gemm.cuh.
// Warp roles, barriers, TMEM double buffering and the epilogue (epilogue_row) are those of gemm.cuh.
#pragma once
#include "gemm.cuh"

namespace b200 {

constexpr int kHaloTile = 16;                      // output patch: 16 x 16 voxels of one frame

template <int BN>
struct ConvHaloSmem {
  static constexpr int kBoxW = kHaloTile, kBoxH = kHaloTile + 2;
  static constexpr int kStageBytesA = kBoxH * kBoxW * 128;        // 36 KB: 288 rows of 64 channels
  static constexpr int kStageBytesB = BN * 128;                   // one tap's weights for a 64-channel slice
  static constexpr int kStagesA = 3;
  static constexpr int kStagesB = (BN == 128) ? 6 : 8;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStagesA * kStageBytesA + kStagesB * kStageBytesB + kBarBytes + 1024;
  static_assert(kStageBytesA % 1024 == 0 && kStageBytesB % 1024 == 0, "stages keep the 1024 B swizzle-atom alignment");
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv3d_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = ConvHaloSmem<BN>;
  constexpr int SA = S::kStagesA, SB = S::kStagesB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * S::kStageBytesA;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + SB * S::kStageBytesB);
  uint64_t* a_empty = a_full + SA;
  uint64_t* b_full = a_empty + SA;
  uint64_t* b_empty = b_full + SB;
  uint64_t* tfull_bar = b_empty + SB;       // [2]
  uint64_t* tempty_bar = tfull_bar + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = p.c_tiles_w, tiles_h = p.c_tiles_h;
  const int num_tiles = p.cB * p.cT * tiles_h * tiles_w;
  const int kb_per_tap = p.cCin / kGemmBK;
  const int taps_t = p.c_taps_t;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<4 * BN>(tmem_slot);     // 2 buffers x 2 halves x BN columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int r = tile;
        const int w0 = (r % tiles_w) * kHaloTile; r /= tiles_w;
        const int h0 = (r % tiles_h) * kHaloTile; r /= tiles_h;
        const int ct = r % p.cT, cb = r / p.cT;
        for (int kt = 0; kt < taps_t; ++kt) {
          int tt = ct;
          if (taps_t == 3) {
            tt += p.c_causal ? kt - 2 : kt - 1;
            if (!p.c_tpad_zero) tt = tt < 0 ? 0 : (tt > p.cTin - 1 ? p.cTin - 1 : tt);        // replicate; else TMA zero-fills t < 0
          }
          for (int cblk = 0; cblk < kb_per_tap; ++cblk) {
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
              mbar_wait(&a_empty[sa], pa ^ 1);
              mbar_arrive_expect_tx(&a_full[sa], S::kStageBytesA);
              tma_load_5d(sA + sa * S::kStageBytesA, &tmA, &a_full[sa], cblk * kGemmBK, w0 - 1 + kw, h0 - 1, tt, cb);
              if (++sa == SA) { sa = 0; pa ^= 1; }
#pragma unroll 1
              for (int kh = 0; kh < 3; ++kh) {
                const int tap = (kt * 3 + kh) * 3 + kw;                                        // tap-major K: (kt, kh, kw, ci)
                mbar_wait(&b_empty[sb], pb ^ 1);
                mbar_arrive_expect_tx(&b_full[sb], S::kStageBytesB);
                tma_load_2d(sB + sb * S::kStageBytesB, &tmB, &b_full[sb], tap * p.cCin + cblk * kGemmBK, 0);
                if (++sb == SB) { sb = 0; pb ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(kGemmBM, BN, 0, 0);
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    const int a_steps = taps_t * kb_per_tap * 3;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 2 * BN;
      for (int as = 0; as < a_steps; ++as) {
        mbar_wait(&a_full[sa], pa);
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a0 = smem_u32(sA + sa * S::kStageBytesA) + kh * (S::kBoxW * 128);
            const uint32_t b0 = smem_u32(sB + sb * S::kStageBytesB);
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
              for (int k = 0; k < kGemmBK / 16; ++k) {
                const uint64_t ad = umma_smem_desc_sw128(a0 + sub * 1024 + k * 32, 16, S::kBoxW * 128);   // 8-row groups one h-row apart
                const uint64_t bd = umma_smem_desc_sw128(b0 + k * 32, 16, 1024);
                umma_ss(d_tmem + sub * BN, ad, bd, idesc, (as | kh | k) ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[sb]);
            if (kh == 2) umma_commit(&a_empty[sa]);
            if (kh == 2 && as == a_steps - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;                   // TMEM sub-partition this warp may read
    const int row = q * 32 + lane;            // row of a half: (h, w) = (row / 8, row % 8)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int r = tile;
      const int w0 = (r % tiles_w) * kHaloTile; r /= tiles_w;
      const int h0 = (r % tiles_h) * kHaloTile; r /= tiles_h;
      const int ot = r % p.cT, ob = r / p.cT;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int oh = h0 + (row >> 3), ow = w0 + sub * 8 + (row & 7);
        const bool row_ok = (oh < p.cH) && (ow < p.cW);
        const long long m_lin = ((static_cast<long long>(ob) * p.cT + ot) * p.cH + oh) * p.cW + ow;
        const uint32_t t_addr = tmem_base + acc * 2 * BN + sub * BN + (static_cast<uint32_t>(q * 32) << 16);
        epilogue_row<BN>(p, t_addr, 0, row_ok, m_lin, ob, ot, oh, ow);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<4 * BN>(tmem_base);
  }
}

}  // namespace b200
