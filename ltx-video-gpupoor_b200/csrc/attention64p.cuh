// d = 64 self-attention with the softmax software-pipelined ACROSS key blocks (LTX-Video: 6144 keys, 32 heads x 64).
//
// Same work decomposition and math as attention_fwd_kernel<64, false> (attention.cuh: work item = (b, h, 256 query rows) = two
// 128-row tiles sharing every 128-key K/V block; one query row per softmax thread; lazy O rescale; packed FFMA2 / FADD2
// arithmetic with a share of the exponentials on the FMA pipe), restricted to the case that dominates the LTX step: no key bias,
// no key lengths, Lk a multiple of 128, plain output.  Everything else stays on attention_fwd_kernel.
//
// What differs is the schedule of a softmax thread.  In attention_fwd_kernel a key block is three phases in sequence — wait for
// S, TMEM -> registers with the row max, then the exponentials — and ncu shows where that loses (profiles/r01b_ncu_attn_d64_*):
// the XU pipe that bounds d = 64 (16 ex2 / clk / SM against 512 clk of MMA per 128x128 block) is 65 % busy, the softmax warps
// sit in the S wait 13-18 % of the time and in fixed-latency dependency stalls for most of the rest: with two warps per
// scheduler there is nothing to issue while one warp runs its MUFU-free max pass and the other waits.  Here the three phases of
// CONSECUTIVE blocks overlap inside one thread: while block G's exponentials are computed 32 scores at a time, the registers
// each chunk frees are refilled at once with block G+1's scores (tcgen05.ld), and their row max is folded in between the next
// chunks' exponentials.  No register is added (the 128-score array is reused in place), no TMEM read is added, and the
// instruction stream of a block becomes one homogeneous mix of MUFU, FMA-pipe and ALU work with the TMEM latency under it.
//
// That needs S(G+1) to exist while P(G) is still being produced, so each tile owns its S, P and O columns outright
// (2 x 128 S + 2 x 64 P + 2 x 64 O = 512 TMEM columns) instead of rotating three aliased S/P buffers; each tile has its own
// MMA-issuing thread, so one tile's barrier waits never delay the other tile's MMAs:
//     S_t(G+2) is issued as soon as S_t(G+1) is in registers (s_read), right behind O_t += P_t(G).V(G) (p_full);
//     P_t(G) is stored once P_t(G-1).V has retired (pv_done) — the first store sits behind the first chunk's exponentials.
#pragma once
#include "attention.cuh"

namespace b200 {

struct Attn64PCfg {
  static constexpr int D = 64, BN = 128;
  static constexpr int kQBytes = kAttnBM * D * 2;            // one Q tile
  static constexpr int kKBytes = BN * D * 2;                 // one K (or V) block
  static constexpr int kStages = (128 * 1024) / (2 * kKBytes);
  static constexpr int kBarBytes = 512;
  static constexpr int kTotal = 2 * kQBytes + 2 * kStages * kKBytes + kBarBytes + 1024;
  static constexpr int kThreads = 12 * 32;                   // 8 softmax warps + TMA warp + one MMA warp per tile + 1 idle
  static constexpr uint32_t kColS = 0, kColP = 256, kColO = 384;
  static constexpr int kSoftmaxRegs = 208, kOtherRegs = 64;
};

// barrier waits / arrivals on a precomputed 32-bit shared address (one register per softmax thread; the generic-pointer helpers
// make ptxas rebuild the address — ~20 instructions — at every use once the 128-score array has taken the registers)
DEVI void mbar_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
DEVI void mbar_arrive_a(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(addr) : "memory"); }

// row max of 32 scores already in registers, into two independent accumulators
DEVI void max_chunk(const uint32_t* v, float& mx0, float& mx1) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    mx0 = fmax3(mx0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
    mx1 = fmax3(mx1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
  }
}

__global__ void __launch_bounds__(Attn64PCfg::kThreads, 1)
attention64p_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Attn64PCfg;
  constexpr int D = C::D, BN = C::BN, kStages = C::kStages;
  constexpr int kTmaWarp = 8, kMmaWarp0 = 9;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                 // [2][128][64]
  uint8_t* sK = sQ + 2 * C::kQBytes;                  // [kStages][128][64]
  uint8_t* sV = sK + kStages * C::kKBytes;            // [kStages][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * C::kKBytes);
  uint64_t* q_full = bars;                // [2]       TMA -> MMA
  uint64_t* q_empty = q_full + 2;         // [2]       last S of the item issued: the Q tile may be overwritten
  uint64_t* k_full = q_empty + 2;         // [kStages]
  uint64_t* k_empty = k_full + kStages;   // [kStages] both tiles' S of the block retired (2 commits)
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;   // [kStages] both tiles' P.V of the block retired (2 commits)
  uint64_t* s_full = v_empty + kStages;   // [2]       S_t(G) complete
  uint64_t* s_read = s_full + 2;          // [2]       S_t(G) is in the softmax threads' registers (4 warps arrive)
  uint64_t* p_full = s_read + 2;          // [2]       P_t(G) written (4 warps arrive)
  uint64_t* pv_done = p_full + 2;         // [2]       P_t(G).V retired: P_t may be rewritten, O_t may be rescaled
  uint64_t* o_done = pv_done + 2;         // [2]       last P.V of the item retired
  uint64_t* o_free = o_done + 2;          // [2]       epilogue has read O_t (4 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.Lk / BN;
  const int n_items = (static_cast<int>(blockIdx.x) < p.total) ? (p.total - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  const uint32_t totalG = static_cast<uint32_t>(n_items) * static_cast<uint32_t>(nblk);      // key blocks this CTA walks (per tile)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_empty[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&s_read[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_done[t], 1);
      mbar_init(&o_free[t], 4);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 2);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C::kOtherRegs));
    if (warp == kTmaWarp && elect_one()) {
      // ================= TMA producer =================
      uint32_t kc = 0;
      int it = 0;
      for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
        const int qp = w % p.pairs, bh = w / p.pairs, h = bh % p.H, b = bh / p.H;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait_backoff(&q_empty[t], (it & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[t], C::kQBytes);
          tma_load_4d(sQ + t * C::kQBytes, &tmQ, &q_full[t], 0, h, qp * 256 + t * kAttnBM, b);
        }
        for (int j = 0; j < nblk; ++j, ++kc) {
          const int st = kc % kStages;
          const uint32_t ph = (kc / kStages) & 1;
          mbar_wait_backoff(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], C::kKBytes);
          tma_load_4d(sK + st * C::kKBytes, &tmK, &k_full[st], 0, h, j * BN, b);
          mbar_wait_backoff(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], C::kKBytes);
          tma_load_4d(sV + st * C::kKBytes, &tmV, &v_full[st], 0, h, j * BN, b);
        }
      }
    } else if ((warp == kMmaWarp0 || warp == kMmaWarp0 + 1) && elect_one()) {
      // ================= MMA issuer of tile t =================
      const int t = warp - kMmaWarp0;
      constexpr uint32_t idesc_s = umma_idesc_bf16(kAttnBM, BN, 0, 0);   // S = Q K^T  (both K-major)
      constexpr uint32_t idesc_o = umma_idesc_bf16(kAttnBM, D, 0, 1);    // O += P V   (V is MN-major)
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ + t * C::kQBytes), 16, 1024);
      const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV), BN * 128, 1024);
      const uint32_t tS = tmem_base + C::kColS + t * BN, tP = tmem_base + C::kColP + t * (BN / 2), tO = tmem_base + C::kColO + t * D;
      // all items have nblk key blocks: block G of this CTA is key block G % nblk of item G / nblk and sits in ring stage G % kStages
      auto issue_s = [&](uint32_t G) {
        const uint32_t it = G / nblk, j = G - it * nblk;
        const int st = G % kStages;
        mbar_wait_parked(&k_full[st], (G / kStages) & 1);
        if (j == 0) mbar_wait_parked(&q_full[t], it & 1);
        if (G > 0) mbar_wait_parked(&s_read[t], (G - 1) & 1);            // S_t(G-1) has left TMEM
        tc_fence_after();
        const uint64_t ka = kdesc + static_cast<uint32_t>(st * (C::kKBytes >> 4));
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) umma_ss(tS, qdesc + static_cast<uint32_t>((ks * 32) >> 4), ka + static_cast<uint32_t>((ks * 32) >> 4), idesc_s, ks ? 1u : 0u);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[st]);
        if (j + 1 == static_cast<uint32_t>(nblk)) umma_commit(&q_empty[t]);
      };
      if (totalG > 0) issue_s(0);
      if (totalG > 1) issue_s(1);
#pragma unroll 1
      for (uint32_t G = 0; G < totalG; ++G) {
        const uint32_t it = G / nblk, j = G - it * nblk;
        const int sv = G % kStages;
        mbar_wait_parked(&p_full[t], G & 1);
        if (j == 0) mbar_wait_parked(&o_free[t], (it & 1) ^ 1);
        mbar_wait_parked(&v_full[sv], (G / kStages) & 1);
        tc_fence_after();
        // B = V[16 keys (K), 64 (N)], N contiguous: 8-key groups 1024 B apart (SBO)
        const uint64_t va = vdesc + static_cast<uint32_t>(sv * (C::kKBytes >> 4));
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) umma_ts(tO, tP + ks * 8, va + static_cast<uint32_t>(ks * (2048 >> 4)), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[sv]);
        if (j + 1 == static_cast<uint32_t>(nblk)) umma_commit(&o_done[t]);
        if (G + 2 < totalG) issue_s(G + 2);
      }
    }
  } else {
    // ================= softmax / correction / epilogue: warpgroup t owns query tile t =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::kSoftmaxRegs));
    const int t = warp >> 2, sub = warp & 3, row = sub * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tS = tmem_base + C::kColS + t * BN + lane_addr;
    const uint32_t tP = tmem_base + C::kColP + t * (BN / 2) + lane_addr;
    const uint32_t tO = tmem_base + C::kColO + t * D + lane_addr;
    const float sc = p.scale_log2;
    // one opaque base register (a volatile mov: not rematerialisable) + compile-time offsets
    uint32_t bar_base = smem_u32(bars) + 8u * static_cast<uint32_t>(t);
    asm volatile("mov.u32 %0, %0;" : "+r"(bar_base));
    constexpr uint32_t kOffSFull = (4 + 4 * kStages) * 8;
    const uint32_t a_s_full = bar_base + kOffSFull, a_s_read = a_s_full + 16, a_p_full = a_s_full + 32;
    const uint32_t a_pv_done = a_s_full + 48, a_o_done = a_s_full + 64, a_o_free = a_s_full + 80;
    uint32_t v[BN];                    // the scores of the block whose exponentials are due (raw, unscaled)
    float mx0 = -INFINITY, mx1 = -INFINITY;
    if (totalG > 0) {                  // the very first block: nothing to overlap it with
      mbar_wait_a(a_s_full, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < BN; c += 32) tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < BN; c += 32) max_chunk(&v[c], mx0, mx1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_s_read);
    }
    uint32_t G = 0;
    int it = 0;
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      const int qp = w % p.pairs, bh = w / p.pairs, h = bh % p.H, b = bh / p.H;
      float m_ref = 0.f, m_run = 0.f, l = 0.f;
#pragma unroll 1
      for (int j = 0; j < nblk; ++j, ++G) {
        const bool has_next = G + 1 < totalG;
        const float m_blk = fmaxf(mx0, mx1) * sc;              // scale > 0: max commutes with the scaling
        bool p_free = true;                                    // P_t(G-1).V known retired (first block of an item: the epilogue saw o_done)
        if (j == 0) {
          m_run = m_blk;
          m_ref = (m_blk == -INFINITY) ? 0.f : m_blk;
        } else {
          m_run = fmaxf(m_run, m_blk);
          const bool need = m_run > m_ref + 8.0f;
          p_free = __any_sync(0xffffffffu, need);              // rescale_o waits for pv_done itself when any row needs it
          rescale_o<D>(tO, need, m_run, m_ref, l, &pv_done[t], (G - 1) & 1);
        }
        mx0 = -INFINITY; mx1 = -INFINITY;                      // from here on: the running max of block G+1
        uint64_t ls[2] = {0ull, 0ull};
        uint64_t X[16];
        uint32_t pk[16];
        // The loads / max of block G+1 are issued unconditionally (for the CTA's last block they read stale scores nobody
        // uses): the stretch between the two barrier waits stays ONE basic block, so ptxas interleaves the MUFU-free max work
        // and the TMEM latency with the exponentials.
        // ---- scores 0..31 ----
        scale_chunk(&v[0], sc, -m_ref, X);
        exp_pairs<AttnPoly<64>::kPer8, AttnPoly<64>::kDeg, 0, 16>(X, pk, ls);
        if (!p_free) {
          mbar_wait_a(a_pv_done, (G - 1) & 1);
          tc_fence_after();
        }
        tmem_st16(tP, pk);
        // ---- scores 32..63; once they are scaled, the first two chunks of block G+1 move into the freed registers ----
        scale_chunk(&v[32], sc, -m_ref, X);
        if (has_next) {
          mbar_wait_a(a_s_full, (G + 1) & 1);
          tc_fence_after();
        }
        tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        exp_pairs<AttnPoly<64>::kPer8, AttnPoly<64>::kDeg, 0, 16>(X, pk, ls);
        tmem_st16(tP + 16, pk);
        // ---- scores 64..95 ----
        scale_chunk(&v[64], sc, -m_ref, X);
        tmem_wait_ld();
        tmem_ld32(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
        max_chunk(&v[0], mx0, mx1);
        max_chunk(&v[32], mx0, mx1);
        exp_pairs<AttnPoly<64>::kPer8, AttnPoly<64>::kDeg, 0, 16>(X, pk, ls);
        tmem_st16(tP + 32, pk);
        // ---- scores 96..127 ----
        scale_chunk(&v[96], sc, -m_ref, X);
        tmem_wait_ld();
        tmem_ld32(tS + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
        max_chunk(&v[64], mx0, mx1);
        exp_pairs<AttnPoly<64>::kPer8, AttnPoly<64>::kDeg, 0, 8>(X, pk, ls);
        tmem_wait_ld();
        max_chunk(&v[96], mx0, mx1);
        exp_pairs<AttnPoly<64>::kPer8, AttnPoly<64>::kDeg, 8, 8>(X, pk, ls);
        tmem_st16(tP + 48, pk);
        float l0, l1;
        unpack_f32x2(add_f32x2(ls[0], ls[1]), l0, l1);
        l += l0 + l1;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_a(a_p_full);
          if (has_next) mbar_arrive_a(a_s_read);
        }
      }
      // ---- epilogue: O / l -> bf16 -> global (the next item's first scores wait in v meanwhile) ----
      mbar_wait_a(a_o_done, it & 1);
      tc_fence_after();
      const float inv = (l > 0.f) ? __fdividef(1.0f, l) : 0.f;
      const int q = qp * 256 + t * kAttnBM + row;
      __nv_bfloat16* orow = p.out + static_cast<long long>(b) * p.out_bs + static_cast<long long>(q) * p.out_ld + h * D;
#pragma unroll 1
      for (int c = 0; c < D; c += 32) {
        uint32_t o[32];
        tmem_ld32(tO + c, o);
        tmem_wait_ld();
        if (q < p.Lq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[i + e]) * inv;
            *reinterpret_cast<uint4*>(orow + c + i) =
                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_o_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
