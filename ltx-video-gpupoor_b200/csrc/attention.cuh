// Flash-style non-causal attention forward for sm_100a: tcgen05 MMAs, S/P/O in TMEM, online softmax.
//
//   O[b, q, h, :] = softmax_k( scale * Q[b,q,h,:].K[b,k,h,:] + key_bias[b,k] ) V[b,k,h,:]
//
// Layout [B, L, H, d] for Q/K/V with arbitrary (16B-aligned) token and batch strides, so q/k/v may be
// column slices of one fused QKV projection.  One CTA = one 128-row Q tile of one (b, h); two CTAs are
// co-resident per SM so one CTA's softmax overlaps the other's MMAs.  192 threads: warp 0 = TMA
// producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = softmax / correction / epilogue with one
// query row per thread (TMEM lane == row).
//
// TMEM columns: [0,128) S fp32, overwritten in place by P (bf16 pairs, 64 columns) ; [128,128+d) O fp32.
// P is consumed straight from TMEM as the A operand of the P.V MMA (tcgen05.mma "TS" form); V is read
// from shared memory as an MN-major B operand, so no transposes are materialised.
// Rescaling of O is lazy (only when the running max grows by more than 2^8).
// Reference semantics: utils/attention.py:99-116 (sdpa_wrapper) incl. the additive key mask.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kAttnThreads = 192;
constexpr int kAttnBM = 128;   // query rows per CTA
constexpr int kAttnBN = 128;   // keys per block

struct AttnParams {
  int B, H, Lq, Lk;
  float scale_log2;              // softmax scale * log2(e)
  const float* key_bias;         // [B, Lk] additive (natural-log domain) or null
  __nv_bfloat16* out;            // [B, Lq, H*d] contiguous rows, row stride out_ld
  long long out_ld, out_bs;
};

template <int D>
struct AttnSmem {
  static constexpr int kStages = (D == 64) ? 2 : 1;
  static constexpr int kQBytes = kAttnBM * D * 2;
  static constexpr int kKBytes = kAttnBN * D * 2;
  static constexpr int kVBytes = kAttnBN * D * 2;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kQBytes + kStages * (kKBytes + kVBytes) + kBarBytes + 1024;
};

// One 128-key block of the online softmax for one query row (one thread): S (TMEM fp32) -> P (TMEM, bf16 pairs),
// running max m_ref (log2 domain) and row sum l updated, O rescaled lazily.  With kEarlyS, P has its own TMEM
// columns: `s_free` is signalled as soon as the whole S row sits in registers (so the MMA warp can already issue
// the next block's Q.K^T), and `pv_done` (previous block's P.V retired) is awaited before O or P are touched.
template <int D, bool kPredicated, bool kEarlyS>
DEVI void softmax_block(uint32_t tS, uint32_t tP, uint32_t tO, int j, int kbase, int Lk, const float* bias, float sc,
                        float& m_ref, float& l, uint64_t* s_free, uint64_t* pv_done, int lane) {
  const float kLog2e = 1.4426950408889634f;
  // ---- pass 1: block row-max (TMEM loads software-pipelined: chunk c+1 is in flight while c is reduced) ----
  float mx0 = -INFINITY, mx1 = -INFINITY;
  {
    uint32_t va[32], vb[32];
    tmem_ld32(tS, va);
#pragma unroll
    for (int c = 0; c < kAttnBN; c += 64) {
      tmem_wait_ld();
      tmem_ld32(tS + c + 32, vb);
      if (kPredicated) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = kbase + c + i;
          float s = __uint_as_float(va[i]) * sc;
          if (bias && k < Lk) s += __ldg(bias + k) * kLog2e;
          if (k >= Lk) s = -INFINITY;
          mx0 = fmaxf(mx0, s);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(va[i]), __uint_as_float(va[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(va[i + 2]), __uint_as_float(va[i + 3])));
        }
      }
      tmem_wait_ld();
      if (c + 64 < kAttnBN) tmem_ld32(tS + c + 64, va);
      if (kPredicated) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = kbase + c + 32 + i;
          float s = __uint_as_float(vb[i]) * sc;
          if (bias && k < Lk) s += __ldg(bias + k) * kLog2e;
          if (k >= Lk) s = -INFINITY;
          mx1 = fmaxf(mx1, s);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(vb[i]), __uint_as_float(vb[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3])));
        }
      }
    }
  }
  float m_blk = fmaxf(mx0, mx1);
  if (!kPredicated) m_blk *= sc;                 // scale > 0: max commutes with the scaling
  // ---- lazy rescale of O and l ----
  bool need;
  if (j == 0) {
    m_ref = (m_blk == -INFINITY) ? 0.f : m_blk;
    need = false;
  } else {
    need = m_blk > m_ref + 8.0f;
  }
  if (kEarlyS && j > 0) {          // P(j-1).V must have retired before O is rescaled or P is overwritten
    mbar_wait(pv_done, (j - 1) & 1);
    tc_fence_after();
  }
  if (__any_sync(0xffffffffu, need)) {
    const float m_new = need ? m_blk : m_ref;
    const float alpha = fast_exp2(m_ref - m_new);
    m_ref = m_new;
    l *= alpha;
#pragma unroll 1
    for (int c = 0; c < D; c += 32) {
      uint32_t o[32];
      tmem_ld32(tO + c, o);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st32(tO + c, o);
    }
  }
  // ---- pass 2: P = exp2(s*scale - m_ref) -> bf16 pairs into TMEM, l += rowsum ----
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  uint32_t va[32], vb[32];
  tmem_ld32(tS, va);
#pragma unroll
  for (int c = 0; c < kAttnBN; c += 64) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t (&v)[32] = half ? vb : va;
      uint32_t (&nx)[32] = half ? va : vb;
      const int cc = c + half * 32;
      tmem_wait_ld();
      if (cc + 32 < kAttnBN) tmem_ld32(tS + cc + 32, nx);
      if (kEarlyS && cc + 32 == kAttnBN) {     // every S column of this row is now in registers: S may be overwritten
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
      }
      float e[32];
      if (kPredicated) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = kbase + cc + i;
          float s = __uint_as_float(v[i]) * sc;
          if (bias && k < Lk) s += __ldg(bias + k) * kLog2e;
          e[i] = (k < Lk) ? fast_exp2(s - m_ref) : 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = fmaf(__uint_as_float(v[i]), sc, -m_ref);
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = fast_exp2(e[i]);
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        l0 += e[i]; l1 += e[i + 1]; l2 += e[i + 2]; l3 += e[i + 3];
        pk[i >> 1] = pack_bf16(e[i], e[i + 1]);
        pk[(i >> 1) + 1] = pack_bf16(e[i + 2], e[i + 3]);
      }
      tmem_st16(tP + (cc >> 1), pk);
    }
  }
  l += (l0 + l1) + (l2 + l3);
}

template <int D, bool kMasked>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using S = AttnSmem<D>;
  constexpr int kStages = S::kStages;
  constexpr int kChunks = D / 64;                    // 64-wide (128 B) column chunks per row
  constexpr uint32_t kTmemCols = 256;
  // d=64: P has its own columns so S(j+1) can be issued while softmax(j) is still exponentiating (kEarlyS);
  // d=128: 128 (S) + 128 (O) fill the 256-column budget of a 2-CTA/SM kernel, so P aliases S.
  constexpr bool kEarlyS = (D == 64);
  constexpr uint32_t kColS = 0, kColP = kEarlyS ? 128 : 0, kColO = kEarlyS ? 192 : 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + S::kQBytes;
  uint8_t* sV = sK + kStages * S::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStages * S::kVBytes);
  uint64_t* q_full = bars;               // 1
  uint64_t* k_full = bars + 1;           // kStages
  uint64_t* k_empty = k_full + 2;
  uint64_t* v_full = k_empty + 2;
  uint64_t* v_empty = v_full + 2;
  uint64_t* s_full = v_empty + 2;        // 1
  uint64_t* p_full = s_full + 1;         // 1
  uint64_t* o_done = p_full + 1;         // 1
  uint64_t* s_free = o_done + 1;         // 1  softmax finished READING S(j)          (kEarlyS)
  uint64_t* pv_done = s_free + 1;        // 1  P(j).V MMA retired: P and O may be touched (kEarlyS)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kAttnBM, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.Lk + kAttnBN - 1) / kAttnBN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    mbar_init(s_free, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, S::kQBytes);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) tma_load_4d(sQ + c * (kAttnBM * 128), &tmQ, q_full, c * 64, h, q0, b);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait_backoff(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], S::kKBytes);
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          tma_load_4d(sK + st * S::kKBytes + c * (kAttnBN * 128), &tmK, &k_full[st], c * 64, h, j * kAttnBN, b);
        mbar_wait_backoff(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], S::kVBytes);
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
          tma_load_4d(sV + st * S::kVBytes + c * (kAttnBN * 128), &tmV, &v_full[st], c * 64, h, j * kAttnBN, b);
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_s = umma_idesc_bf16(kAttnBM, kAttnBN, 0, 0);   // S = Q K^T  (both K-major)
    constexpr uint32_t idesc_o = umma_idesc_bf16(kAttnBM, D, 0, 1);         // O += P V   (V is MN-major)
    mbar_wait(q_full, 0);
    auto issue_s = [&](int stage) {
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + stage * S::kKBytes);
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t off = (ks >> 2) * (kAttnBM * 128) + (ks & 3) * 32;
        umma_ss(tmem_base + kColS, umma_smem_desc_sw128(qa + off, 16, 1024), umma_smem_desc_sw128(ka + off, 16, 1024),
                idesc_s, ks ? 1u : 0u);
      }
      umma_commit(&k_empty[stage]);
      umma_commit(s_full);
    };
    auto issue_pv = [&](int stage, int j) {
      const uint32_t va = smem_u32(sV + stage * S::kVBytes);
#pragma unroll
      for (int ks = 0; ks < kAttnBN / 16; ++ks) {
        // B = V[16 keys (K), D (N)], N contiguous: 8-key groups 1024 B apart (SBO), 64-col groups one chunk apart (LBO)
        const uint64_t vd = umma_smem_desc_sw128(va + ks * 2048, kAttnBN * 128, 1024);
        umma_ts(tmem_base + kColO, tmem_base + kColP + ks * 8, vd, idesc_o, (j | ks) ? 1u : 0u);
      }
      umma_commit(&v_empty[stage]);
      if (kEarlyS) umma_commit(pv_done);
      if (j == nblk - 1) umma_commit(o_done);
    };
    if (kEarlyS) {
      // S(j+1) is issued as soon as softmax(j) has read S(j) out of TMEM, ahead of P(j).V, so the next block's
      // scores are ready when the softmax warps come back for them.
      int st_k = 0, st_v = 0;
      uint32_t ph_k = 0, ph_v = 0;
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (lane == 0) issue_s(0);
      __syncwarp();
      if (++st_k == kStages) { st_k = 0; ph_k ^= 1; }
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          mbar_wait(&k_full[st_k], ph_k);
          mbar_wait(s_free, j & 1);
          tc_fence_after();
          if (lane == 0) issue_s(st_k);
          __syncwarp();
          if (++st_k == kStages) { st_k = 0; ph_k ^= 1; }
        }
        mbar_wait(&v_full[st_v], ph_v);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        if (lane == 0) issue_pv(st_v, j);
        __syncwarp();
        if (++st_v == kStages) { st_v = 0; ph_v ^= 1; }
      }
    } else {
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        if (lane == 0) issue_s(st);
        __syncwarp();
        mbar_wait(&v_full[st], ph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        if (lane == 0) issue_pv(st, j);
        __syncwarp();
        if (++st == kStages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ================= softmax / correction / epilogue =================
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tS = tmem_base + kColS + lane_addr;
    const uint32_t tP = tmem_base + kColP + lane_addr;
    const uint32_t tO = tmem_base + kColO + lane_addr;
    const float* bias = p.key_bias ? p.key_bias + static_cast<long long>(b) * p.Lk : nullptr;
    float m_ref = 0.f, l = 0.f;

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * kAttnBN;
      // full blocks without a bias take the lean path; the tail block / biased blocks take the predicated one
      if (kMasked && (bias != nullptr || kbase + kAttnBN > p.Lk))
        softmax_block<D, true, kEarlyS>(tS, tP, tO, j, kbase, p.Lk, bias, p.scale_log2, m_ref, l, s_free, pv_done, lane);
      else
        softmax_block<D, false, kEarlyS>(tS, tP, tO, j, kbase, p.Lk, bias, p.scale_log2, m_ref, l, s_free, pv_done, lane);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(o_done, 0);
    tc_fence_after();
    const float inv = (l > 0.f) ? __fdividef(1.0f, l) : 0.f;
    const int q = q0 + row;
    __nv_bfloat16* orow = p.out + static_cast<long long>(b) * p.out_bs + static_cast<long long>(q) * p.out_ld + h * D;
#pragma unroll 1
    for (int c = 0; c < D; c += 32) {
      uint32_t o[32];
      tmem_ld32(tO + c, o);
      tmem_wait_ld();
      if (q < p.Lq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          *reinterpret_cast<uint4*>(orow + c + i) = make_uint4(
              pack_bf16(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv),
              pack_bf16(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv),
              pack_bf16(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv),
              pack_bf16(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace b200
