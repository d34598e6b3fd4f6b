// d = 128 attention on a CTA PAIR (cluster of 2, tcgen05 cta_group::2) — the same split-phase two-tile pipeline as attention_fwd_kernel<128>
// (attention.cuh), with every MMA 256 rows tall: a pair walks work items of 512 query rows = two 256-row tiles; CTA r keeps rows
// r*128..r*128+127 of each tile (its own Q rows, S / P / O in its own TMEM lanes, its own softmax warpgroups and epilogue) and only HALF
// of every K / V block:
//   S = Q K^T   (M 256, N 64 per half block, K-major B): the 64 keys of a half block are split 32 / 32 between the CTAs
//   O += P V    (M 256, N 128 = d, MN-major B):          the 128 value columns are split 64 / 64 between the CTAs
// so the shared-memory operand reads and the TMA fill per SM drop by a third (S) and a half (P.V), and one thread issues for two SMs.
// Why: the free-running-MMA ablation (profiles/r02_attn64_ab.md §10) caps the single-CTA kernel's own MMA stream at ~1530 TF/s — below
// what cuDNN's kernel reaches per clock — i.e. the 128 x 128 single-CTA instruction mix is the limit, not the softmax around it.
// The leader (cluster rank 0) issues all MMAs; commits are multicast to both CTAs; the peer's softmax warps arrive on the leader's
// handshake barriers through the cluster address space; TMA loads of both CTAs complete on the leader's barriers.
// Measured (profiles/r02_attn64_ab.md §12): level with the single-CTA kernel in burst, -1 % sustained at a 3.7 % higher SM clock.
// OPT-IN (LTXB200_ATTN128_2CTA=1).  Reference semantics as attention.cuh (utils/attention.py:99-116).
#pragma once
#include "attention.cuh"

namespace b200 {

struct Attn2CtaCfg {
  static constexpr int D = 128, BN = 128;
  static constexpr int kQBytes = kAttnBM * D * 2;          // one Q tile: this CTA's 128 rows
  static constexpr int kKBytes = (BN / 2) * D * 2;         // this CTA's 64 keys of a block: [half s][chunk c][32 keys][128 B]
  static constexpr int kVBytes = BN * (D / 2) * 2;         // this CTA's 64 value columns of a block: [128 keys][128 B]
  static constexpr int kStagesK = 5, kStagesV = 4;
  static constexpr int kBarBytes = 512;
  static constexpr int kTotal = 2 * kQBytes + kStagesK * kKBytes + kStagesV * kVBytes + kBarBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
  static constexpr int kThreads = 12 * 32;                 // 8 softmax warps + TMA warp + MMA warp + 2 idle
  static constexpr int kSoftmaxRegs = 208, kOtherRegs = 64;
  static constexpr int kRowsPerItem = 4 * kAttnBM;         // 512 query rows per pair item
};

template <bool kMasked>
__global__ void __launch_bounds__(Attn2CtaCfg::kThreads, 1)
attention128p2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Attn2CtaCfg;
  constexpr int D = C::D, BN = C::BN;
  constexpr int kStagesK = C::kStagesK, kStagesV = C::kStagesV;
  constexpr int kSBufs = 2;
  constexpr int kSmWarps = 8, kTmaWarp = 8, kMmaWarp = 9;
  constexpr uint32_t kColS0 = 0, kColO0 = kSBufs * BN;
  const uint32_t rank = cluster_ctarank();
  const int cl_first = blockIdx.x >> 1, cl_stride = gridDim.x >> 1;
  const int groups = (p.Lq + C::kRowsPerItem - 1) / C::kRowsPerItem;
  const int total = p.B * p.H * groups;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                 // [2][128][D]
  uint8_t* sK = sQ + 2 * C::kQBytes;                  // [kStagesK][2 halves][2 chunks][32][128 B]
  uint8_t* sV = sK + kStagesK * C::kKBytes;           // [kStagesV][128][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesV * C::kVBytes);
  uint64_t* q_full = bars;                // [2]   leader: both CTAs' Q rows of tile t have landed
  uint64_t* q_empty = q_full + 2;         // [2]   multicast commit
  uint64_t* k_full = q_empty + 2;         // [kStagesK] leader
  uint64_t* k_empty = k_full + kStagesK;  //       multicast commit
  uint64_t* v_full = k_empty + kStagesK;  // [kStagesV] leader
  uint64_t* v_empty = v_full + kStagesV;  //       multicast commit
  uint64_t* s_full = v_empty + kStagesV;  // [2]   multicast commit: S(n) complete in both CTAs' TMEM
  uint64_t* p_full = s_full + kSBufs;     // [2]   leader: 8 warps (4 per CTA)
  uint64_t* pv_done = p_full + kSBufs;    // [2]   multicast commit
  uint64_t* o_done = pv_done + 2;         // [2]   multicast commit
  uint64_t* o_free = o_done + 2;          // [2]   leader: 8 warps
  uint64_t* s_read = o_free + 2;          // [2]   leader: 8 warps
  uint64_t* p_half = s_read + kSBufs;     // [2]   leader: 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + kSBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto item_keys = [&](int b) -> int { return p.key_lens ? __ldg(p.key_lens + b) : p.Lk; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_empty[t], 1);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_done[t], 1);
      mbar_init(&o_free[t], 8);
    }
    for (int i = 0; i < kSBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 8);
      mbar_init(&s_read[i], 8);
      mbar_init(&p_half[i], 8);
    }
    for (int i = 0; i < kStagesK; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kStagesV; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                                  // the peer must not signal barriers that are not initialised yet
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kSmWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C::kOtherRegs));
    if (warp == kTmaWarp && elect_one()) {
      // ================= TMA producer (both CTAs; the leader arms the barriers with both CTAs' bytes) =================
      uint32_t kc = 0;
      int it = 0;
      for (int w = cl_first; w < total; w += cl_stride, ++it) {
        const int qg = w % groups, bh = w / groups, h = bh % p.H, b = bh / p.H;
        const int nblk = (item_keys(b) + BN - 1) / BN;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait_backoff(&q_empty[t], (it & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&q_full[t], 2 * C::kQBytes);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_4d_2sm(sQ + t * C::kQBytes + c * (kAttnBM * 128), &tmQ, &q_full[t], c * 64, h,
                            qg * C::kRowsPerItem + t * 256 + static_cast<int>(rank) * kAttnBM, b);
        }
        auto load_k = [&](int j) {
          const uint32_t kk = kc + j;
          const int st = kk % kStagesK;
          mbar_wait_backoff(&k_empty[st], ((kk / kStagesK) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&k_full[st], 2 * C::kKBytes);
          // half s of the block (keys 64 s ..): this CTA's 32 keys of it, in two 64-column chunks
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int c = 0; c < 2; ++c)
              tma_load_4d_2sm(sK + st * C::kKBytes + s * 8192 + c * 4096, &tmK, &k_full[st], c * 64, h,
                              j * BN + s * 64 + static_cast<int>(rank) * 32, b);
        };
        load_k(0);
        for (int j = 0; j < nblk; ++j) {
          if (j + 1 < nblk) load_k(j + 1);
          const uint32_t vv = kc + j;
          const int sv = vv % kStagesV;
          mbar_wait_backoff(&v_empty[sv], ((vv / kStagesV) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&v_full[sv], 2 * C::kVBytes);
          tma_load_4d_2sm(sV + sv * C::kVBytes, &tmV, &v_full[sv], static_cast<int>(rank) * 64, h, j * BN, b);
        }
        kc += nblk;
      }
    } else if (warp == kMmaWarp && rank == 0 && elect_one()) {
      // ================= MMA issuer (leader CTA only) =================
      constexpr uint32_t idesc_sh = umma_idesc_bf16(2 * kAttnBM, BN / 2, 0, 0);   // S, one half of the keys
      constexpr uint32_t idesc_o = umma_idesc_bf16(2 * kAttnBM, D, 0, 1);         // O += P V (V MN-major)
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV), BN * 128, 1024);
      struct SCur { int w, it, nblk, m; uint32_t kc0, N; } sc{cl_first, 0, 0, 0, 0u, 0u};
      auto s_enter = [&]() { if (sc.w < total) sc.nblk = (item_keys((sc.w / groups) / p.H) + BN - 1) / BN; };
      auto s_advance = [&]() {
        ++sc.m; ++sc.N;
        if (sc.m >= 2 * sc.nblk) { sc.kc0 += sc.nblk; sc.m = 0; sc.w += cl_stride; ++sc.it; s_enter(); }
      };
      // S at the cursor, half `part` of the keys (1 = keys 64.., issued first; 0 = keys 0..63, completes the step)
      auto issue_s = [&](int part) {
        const int t = sc.m & 1, j = sc.m >> 1;
        const uint32_t kpos = sc.kc0 + j;
        const int st = kpos % kStagesK;
        if (t == 0 && part == 1) mbar_wait_parked(&k_full[st], (kpos / kStagesK) & 1);
        if (j == 0) mbar_wait_parked(&q_full[t], sc.it & 1);
        tc_fence_after();
        const uint32_t buf = sc.N % kSBufs;
        const uint64_t qa = qdesc + static_cast<uint32_t>(t * (C::kQBytes >> 4));
        const uint64_t ka = kdesc + static_cast<uint32_t>(st * (C::kKBytes >> 4)) + static_cast<uint32_t>(part * (8192 >> 4));
        const uint32_t ts = tmem_base + kColS0 + buf * BN + part * (BN / 2);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t offa = ((ks >> 2) * (kAttnBM * 128) + (ks & 3) * 32) >> 4;
          const uint32_t offb = ((ks >> 2) * (32 * 128) + (ks & 3) * 32) >> 4;
          umma_ss_2cta(ts, qa + offa, ka + offb, idesc_sh, ks ? 1u : 0u);
        }
        if (part == 0) {
          umma_commit_2cta(&s_full[buf], 3u);
          if (j + 1 == sc.nblk) umma_commit_2cta(&q_empty[t], 3u);
          if (t == 1) umma_commit_2cta(&k_empty[st], 3u);
        }
      };
      auto issue_pv = [&](int t, uint32_t buf, int sv, int ks0, int ks1, bool acc) {
        const uint64_t va = vdesc + static_cast<uint32_t>(sv * (C::kVBytes >> 4));
        const uint32_t to = tmem_base + kColO0 + t * D, tp = tmem_base + kColS0 + buf * BN;
#pragma unroll
        for (int ks = ks0; ks < ks1; ++ks)
          umma_ts_2cta(to, tp + ks * 8, va + static_cast<uint32_t>(ks * (2048 >> 4)), idesc_o, (acc || ks > ks0) ? 1u : 0u);
      };
      s_enter();
#pragma unroll 1
      for (int i = 0; i < kSBufs && sc.w < total; ++i) { issue_s(1); issue_s(0); s_advance(); }
      uint32_t vc = 0, N = 0;
      int it = 0;
      for (int w = cl_first; w < total; w += cl_stride, ++it) {
        const int nblk = (item_keys((w / groups) / p.H) + BN - 1) / BN;
        const int nsteps = 2 * nblk;
#pragma unroll 1
        for (int n = 0; n < nsteps; ++n, ++N) {
          const int t = n & 1, j = n >> 1;
          const int sv = vc % kStagesV;
          const uint32_t buf = N % kSBufs, par = (N / kSBufs) & 1;
          const bool more = sc.w < total;
          if (more) { mbar_wait_parked(&s_read[buf], par); issue_s(1); }
          mbar_wait_parked(&p_half[buf], par);
          if (j == 0) mbar_wait_parked(&o_free[t], (it & 1) ^ 1);
          if (t == 0) mbar_wait_parked(&v_full[sv], (vc / kStagesV) & 1);
          tc_fence_after();
          issue_pv(t, buf, sv, 0, BN / 32, j > 0);
          mbar_wait_parked(&p_full[buf], par);
          tc_fence_after();
          issue_pv(t, buf, sv, BN / 32, BN / 16, true);
          umma_commit_2cta(&pv_done[t], 3u);
          if (t == 1) { umma_commit_2cta(&v_empty[sv], 3u); ++vc; }
          if (j + 1 == nblk) umma_commit_2cta(&o_done[t], 3u);
          if (more) { issue_s(0); s_advance(); }
        }
      }
    }
  } else {
    // ================= softmax / correction / epilogue: warpgroup t owns this CTA's 128 rows of tile t =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::kSoftmaxRegs));
    const int t = warp >> 2;
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tO = tmem_base + kColO0 + t * D + lane_addr;
    uint32_t G = 0;
    int it = 0;
    for (int w = cl_first; w < total; w += cl_stride, ++it) {
      const int qg = w % groups, bh = w / groups, h = bh % p.H, b = bh / p.H;
      const float* bias = p.key_bias ? p.key_bias + static_cast<long long>(b) * p.Lk : nullptr;
      const int Lk_b = item_keys(b);
      const int nblk = (Lk_b + BN - 1) / BN;
      float m_ref = 0.f, m_run = 0.f, l = 0.f;
      for (int j = 0; j < nblk; ++j, ++G) {
        const uint32_t N = 2 * G + t, buf = N % kSBufs;
        mbar_wait(&s_full[buf], (N / kSBufs) & 1);
        tc_fence_after();
        const int kbase = j * BN;
        const uint32_t tS = tmem_base + kColS0 + buf * BN + lane_addr;
        if (kMasked && (bias != nullptr || kbase + BN > Lk_b))
          softmax_block<D, BN, true, D, true, BN / 2, true>(tS, tO, j == 0, kbase, Lk_b, bias, p.scale_log2, m_ref, m_run, l, &pv_done[t], (G - 1) & 1,
                                                           &s_read[buf], &p_half[buf], lane);
        else
          softmax_block<D, BN, false, D, true, BN / 2, true>(tS, tO, j == 0, kbase, Lk_b, bias, p.scale_log2, m_ref, m_run, l, &pv_done[t], (G - 1) & 1,
                                                            &s_read[buf], &p_half[buf], lane);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) softmax_arrive<true>(&p_full[buf]);
      }
      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(&o_done[t], it & 1);
      tc_fence_after();
      const float inv = (l > 0.f) ? __fdividef(1.0f, l) : 0.f;
      const int q = qg * C::kRowsPerItem + t * 256 + static_cast<int>(rank) * kAttnBM + row;
      __nv_bfloat16* orow;
      if (p.peers.P > 0) {
        const int dst = q / p.tokens_per_peer, nl = q - dst * p.tokens_per_peer;
        orow = static_cast<__nv_bfloat16*>(p.peers.data[dst < p.peers.P ? dst : 0]) +
               (static_cast<long long>(b) * p.tokens_per_peer + nl) * p.out_ld + (p.head_offset + h) * D;
      } else {
        orow = p.out + static_cast<long long>(b) * p.out_bs + static_cast<long long>(q) * p.out_ld + h * D;
      }
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
            if (p.accumulate) {
              const uint4 prev = *reinterpret_cast<const uint4*>(orow + c + i);
              const float2 a = unpack_bf16(prev.x), b2 = unpack_bf16(prev.y), c2 = unpack_bf16(prev.z), d2 = unpack_bf16(prev.w);
              f[0] += a.x; f[1] += a.y; f[2] += b2.x; f[3] += b2.y; f[4] += c2.x; f[5] += c2.y; f[6] += d2.x; f[7] += d2.y;
            }
            *reinterpret_cast<uint4*>(orow + c + i) =
                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) softmax_arrive<true>(&o_free[t]);
    }
  }

  tc_fence_before();
  cluster_sync_all();                                  // the peer may still multicast into this CTA's barriers / read its shared memory
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
  if (p.peers.P > 0) peer_signal_done(p.peers, gridDim.x);
}

}  // namespace b200
