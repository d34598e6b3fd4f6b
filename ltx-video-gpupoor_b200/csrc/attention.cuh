// Flash-style non-causal attention forward for sm_100a: tcgen05 MMAs, S/P/O in TMEM, online softmax.
//
//   O[b, q, h, :] = softmax_k( scale * Q[b,q,h,:].K[b,k,h,:] + key_bias[b,k] ) V[b,k,h,:]
//
// Layout [B, L, H, d] for Q/K/V with arbitrary (16B-aligned) token and batch strides, so q/k/v may be
// column slices of one fused QKV projection.  Reference semantics: utils/attention.py:99-116
// (sdpa_wrapper) incl. the additive key mask.
//
// Structure (persistent, warp-specialised, two query tiles per CTA, one CTA per SM):
//   * a CTA walks work items (b, h, 256 query rows) round-robin; each item is two 128-row Q tiles that share
//     every 128-key K/V block brought in by TMA (one producer warp, kStages-deep K and V rings);
//   * the unit of work is a step n = 2*j + t (key block j, tile t).  S(n) = Q_t.K(j)^T lives in TMEM buffer
//     n % kSBufs.  One MMA warp issues, in order:  O_t += P(n).V(j)  then  S(n + kSBufs)  into the buffer P(n)
//     just left.  With kSBufs = 3 (d = 64: 3*128 + 2*64 = 512 columns) the scores of a tile's next block are
//     computed while its current block is still in the softmax, so the softmax warpgroups never wait for the
//     tensor pipe; with kSBufs = 2 (d = 128: 2*128 + 2*128 columns) the two tiles ping-pong;
//   * two softmax warpgroups (one per tile, one query row per thread = TMEM lane): the S row is read from TMEM
//     ONCE into registers, row max -> lazy rescale of O (only when the max grows by > 2^8) -> exp2 -> P (bf16
//     pairs) written over the first half of that S buffer and consumed straight from TMEM as the A operand of
//     the P.V MMA ("TS" form); V is an MN-major B operand, so nothing is transposed;
//   * the same warpgroups normalise and store O when their tile is finished, while the MMA warp already runs
//     the next work item's first S blocks.
// TMEM columns: S buffers (kSBufs x 128) | O_0 | O_1.
#pragma once
#include "common.cuh"
#include "comm.cuh"

// LTXB200_ATTN_L2PF = n: the TMA producer requests the K / V blocks n key blocks ahead into L2 (0 = off)
#ifndef LTXB200_ATTN_L2PF
#define LTXB200_ATTN_L2PF 0
#endif
// the producer's waits for a free ring stage: sleep-and-probe (64 ns naps) or parked on the barrier (-DLTXB200_ATTN_PRODUCER_PARK)
#ifdef LTXB200_ATTN_PRODUCER_PARK
#define LTXB200_PRODUCER_WAIT(bar, par) mbar_wait_parked(bar, par)
#else
#define LTXB200_PRODUCER_WAIT(bar, par) mbar_wait_backoff(bar, par)
#endif

namespace b200 {

constexpr int kAttnBM = 128;   // query rows per tile (= TMEM lanes)

struct AttnParams {
  int B, H, Lq, Lk;
  float scale_log2;              // softmax scale * log2(e)
  const float* key_bias;         // [B, Lk] additive (natural-log domain) or null
  const int* key_lens;           // [B] valid keys per batch element (1 <= key_lens[b] <= Lk) or null: keys beyond are ignored, i.e. the
                                 // right-padded prompt mask of cross-attention WITHOUT a bias pass and without the padded key blocks
  __nv_bfloat16* out;            // [B, Lq, H*d] contiguous rows, row stride out_ld
  long long out_ld, out_bs;
  int pairs;                     // ceil(Lq / 256)
  int total;                     // B * H * pairs work items
  // Ulysses return path (xdit_context_parallel.py:186-190) fused into the epilogue: when peers.P > 0 the row of
  // query token q goes to rank q / tokens_per_peer, at row b*tokens_per_peer + q % tokens_per_peer and head
  // head_offset + h of that rank's [B*n_loc, H_total*d] matrix (out_ld = its row stride), over NVLink.
  PeerPtrs peers;
  int tokens_per_peer, head_offset;
  int accumulate;                // out += result (bf16 read-modify-write): the second attention of WanI2VCrossAttention (x += img_x)
};

constexpr int kAttnBN = 128;   // keys per block

template <int D, bool kMasked = true>
struct AttnCfg {
  // kHalf (d = 64, -DLTXB200_ATTN64_HALFROW): every query row is shared by TWO threads (64 score columns each), i.e. 16 softmax warps =
  // 4 per scheduler instead of 2, which is what the exp loop needs to hide its latencies (mufu_bench2: 18.0 -> 22.3 exp/clk/SM);
  // the two halves exchange their block maxima through shared memory and a 64-thread named barrier.
#if defined(LTXB200_ATTN64_HALFROW) && defined(LTXB200_ATTN128_HALFROW)
  static constexpr bool kHalf = true;
#elif defined(LTXB200_ATTN64_HALFROW)
  static constexpr bool kHalf = (D == 64);
#elif defined(LTXB200_ATTN128_HALFROW)
  // d = 128 is bound by the latency of ONE tile's chain (S -> softmax -> P.V -> next S: TMEM holds a single S buffer per tile); two threads
  // per row were meant to halve the softmax link of that chain.  Parity-green, measured 1196 vs 1222 TF/s on one box: the link is bound by
  // the tile's 16 384 exponentials on the shared XU pipe, not by the instructions of one thread.  Off by default.
  static constexpr bool kHalf = (D == 128);
#else
  static constexpr bool kHalf = false;
#endif
  // kOnes (d = 64, unmasked kernel, -DLTXB200_ATTN64_ONES): the softmax ROW SUM comes out of the tensor pipe.  Every V stage is followed in
  // shared memory by a constant tile of bf16 ones, which the P.V MMA reads as 16 more value columns (N = 80): accumulator columns 64..79
  // of a tile hold sum_k P[row, k] — fp32, over exactly the bf16-rounded P the MMA multiplies — so the softmax threads drop the 64 packed
  // adds per 128 scores (10 % of their instruction stream) and the lazy rescale of O rescales the sum with it.  TMEM then only fits three
  // score buffers of 96 columns (3 * 96 + 2 * 80 = 448 <= 512).  Parity-green and measured NEUTRAL (968.7 vs 965.0 TF/s on one box,
  // profiles/r02_attn64_ab.md §4): what the adds save, the per-block fixed costs of 64 instead of 48 key blocks take back.  Off by default.
#ifdef LTXB200_ATTN64_ONES
  static constexpr bool kOnes = (D == 64) && !kMasked && !kHalf;
#else
  static constexpr bool kOnes = false;
#endif
#ifdef LTXB200_ATTN128_BN64
  // d = 128 alternative: 64-key blocks with FOUR S/P buffers (4*64 + 2*128 = 512 columns) = two per tile, so a tile's next scores are
  // computed while its current block is in the softmax (no split-phase tricks)
  static constexpr int BN = (D == 128) ? 64 : (kOnes ? 96 : kAttnBN);
#else
  static constexpr int BN = kOnes ? 96 : kAttnBN;
#endif
  static constexpr int kOW = D + (kOnes ? 16 : 0);                            // accumulator columns per tile: O (| 16 copies of the row sum)
  static constexpr int kSBufs = (D == 64) ? 3 : (BN == 64 ? 4 : 2);           // S/P buffers in TMEM, used in rotation by the steps
  // With only two S buffers (d = 128 fills TMEM) a tile's next scores cannot be computed ahead in a spare buffer, so
  // the block is pipelined in halves instead: keys 64..127 of S(n+2) are issued as soon as softmax(n) has READ S(n)
  // (P(n) only overwrites columns 0..63), P.V of keys 0..63 starts when the first half of P(n) is written, and only
  // P.V of keys 64..127 plus the low half of S(n+2) remain between "P complete" and "next S ready".
  static constexpr bool kSplit = (kSBufs == 2);
  // k2Mma (-DLTXB200_ATTN128_2MMA): one MMA-issuing thread PER TILE (the two idle warps of the warpgroup) instead of one for both.  A single
  // in-order thread that waits for tile 0's P cannot issue tile 1's S_hi although its scores have been read, and vice versa; with two
  // threads each tile's chain (S -> softmax -> P.V -> S) advances on its own and the tensor pipe interleaves whatever has been issued.
#ifdef LTXB200_ATTN128_2MMA
  static constexpr bool k2Mma = kSplit;
#else
  static constexpr bool k2Mma = false;
#endif
  // kSplit34: the early P signal comes after 3/4 of the block instead of 1/2, and the low half of the next scores is issued as two
  // 32-key MMAs — keys 0..31 right behind P.V of keys 0..95, keys 32..63 behind P.V of keys 96..127 — so that only a quarter of P.V and
  // a quarter of S remain between "P complete" and "next S ready" (256 instead of 512 tensor-pipe cycles on the tile's critical chain)
#ifdef LTXB200_ATTN128_SPLIT34
  static constexpr bool kSplit34 = kSplit && BN == 128 && !kHalf;
#else
  static constexpr bool kSplit34 = false;
#endif
  static_assert(kSBufs * BN + 2 * kOW <= 512, "TMEM budget");
  static constexpr int kQBytes = kAttnBM * D * 2;            // one Q tile
  static constexpr int kKBytes = BN * D * 2;                 // one K (or V) block
  static constexpr int kVBytes = kKBytes + (kOnes ? BN * 128 : 0);            // V stage: the block (+ the ones tile, [BN][64] bf16 1.0)
  static constexpr int kStages = kOnes ? 5 : (128 * 1024) / (2 * kKBytes);    // 128 KB of K/V in flight (kOnes: 5 x 36 KB)
  static constexpr int kBarBytes = 512 + (kHalf ? 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 4 : 0);   // + max / sum exchange slots of the half rows
  // K ring one stage deeper than the V ring where shared memory allows it (d = 128: 3 x 32 KB of K, 2 x 32 KB of V): the score MMAs of
  // block j + 1 are issued while block j is still in the softmax, i.e. K is needed a block earlier than V (-DLTXB200_ATTN_KSTAGES_EQ: equal rings)
#ifdef LTXB200_ATTN_KSTAGES_EQ
  static constexpr int kStagesK = kStages;
#else
  static constexpr int kStagesK = kStages + ((2 * kQBytes + (kStages + 1) * kKBytes + kStages * kVBytes + kBarBytes + 1024 <= 227 * 1024 && kStages < 3) ? 1 : 0);
#endif
  static constexpr int kStagesV = kStages;
  static constexpr int kTotal = 2 * kQBytes + kStagesK * kKBytes + kStagesV * kVBytes + kBarBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
  static constexpr int kSoftmaxWarps = kHalf ? 16 : 8;
  static constexpr int kThreads = (kSoftmaxWarps + 4) * 32;  // softmax warps + TMA warp + MMA warp + 2 idle (warpgroup alignment)
  static constexpr uint32_t kTmemCols = 512;
  static constexpr int kSoftmaxRegs = kHalf ? 104 : 208, kOtherRegs = 64;  // setmaxnreg: the softmax warpgroups take the registers
};

DEVI float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Lazy rescale of O_t and l: the reference max moves to m_new only for rows whose running max outgrew it by more
// than 2^8 (warp-uniform branch: the TMEM ops are warp-collective), after the previous block's P.V MMA has retired.
template <int D>
DEVI void rescale_o(uint32_t tO, bool need, float m_new, float& m_ref, float& l, uint64_t* pv_done, uint32_t pv_parity) {   // D = columns of O this thread owns
  if (__any_sync(0xffffffffu, need)) {
    mbar_wait(pv_done, pv_parity);
    tc_fence_after();
    const float alpha = need ? fast_exp2(m_ref - m_new) : 1.0f;
    if (need) m_ref = m_new;
    l *= alpha;
#pragma unroll 1
    for (int c = 0; c < D; c += 16) {
      uint32_t o[16];
      tmem_ld16(tO + c, o);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
      tmem_st16(tO + c, o);
    }
  }
}

// exp2 of 32 scores (already in registers) -> 16 packed bf16 pairs at tP, partial row sums in ls[2] (packed fp32 pairs).
// The scale/shift and the row sum run as FFMA2 / FADD2 (half the issue slots of the scalar form), and kPolyPer8 of every
// 8 element pairs take their exponentials from the FMA-pipe polynomial instead of MUFU.EX2, which is what bounds the
// d = 64 kernel (profiles/scripts/mufu_bench2.cu: 14.3 -> 18.0 exp/clk/SM at 2 softmax warps per scheduler).
// Share (of every 8 pairs) and degree of the polynomial exponentials, per head dim.  Measured on one box (profiles/r02_ncu_summary.md,
// attention section): d = 64 — degree 3 x 2/8: 916 TF/s, degree 2 x 2/8: 935, degree 3 x 3/8: 917, degree 2 x 3/8: 967, x 4/8: 919, x 5/8: 882.
#ifndef LTXB200_ATTN_POLY_D64
#define LTXB200_ATTN_POLY_D64 3
#endif
#ifndef LTXB200_ATTN_POLY_DEG_D64
#define LTXB200_ATTN_POLY_DEG_D64 2
#endif
#ifndef LTXB200_ATTN_POLY_D128
#define LTXB200_ATTN_POLY_D128 2
#endif
#ifndef LTXB200_ATTN_POLY_DEG_D128
#define LTXB200_ATTN_POLY_DEG_D128 3
#endif
template <int D> struct AttnPoly {
  static constexpr int kPer8 = (D == 64) ? LTXB200_ATTN_POLY_D64 : LTXB200_ATTN_POLY_D128;
  static constexpr int kDeg = (D == 64) ? LTXB200_ATTN_POLY_DEG_D64 : LTXB200_ATTN_POLY_DEG_D128;
};
template <bool kScaled, int kPolyPer8, int kDeg, bool kSum = true>
DEVI void exp_chunk_pk(const uint32_t* v, float sc, float neg_m, uint32_t (&pk)[16], uint64_t (&ls)[2]) {
  const uint64_t NM = pack_f32x2(neg_m, neg_m);
  const uint64_t SC = pack_f32x2(sc, sc);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const uint64_t V = pack_u32x2(v[2 * p], v[2 * p + 1]);
    const uint64_t X = kScaled ? add_f32x2(V, NM)        // v already holds scale*s + bias
                               : fma_f32x2(V, SC, NM);
    float e0, e1;
    const bool poly = !kScaled && (((p & 7) + 1) * kPolyPer8 / 8 != (p & 7) * kPolyPer8 / 8);   // spread over the 8 pairs
#ifdef LTXB200_ABL_NOEXP                         // ablation build (wrong results): what the block costs without its exponentials
    unpack_f32x2(X, e0, e1);
#else
    if (poly) {
      exp2_poly_f32x2<kDeg>(X, e0, e1);
    } else {
      float x0, x1;
      unpack_f32x2(X, x0, x1);
      e0 = fast_exp2(x0);
      e1 = fast_exp2(x1);
    }
#endif
#ifndef LTXB200_ABL_NOSUM
    if (kSum) ls[p & 1] = add_f32x2(ls[p & 1], pack_f32x2(e0, e1));
#endif
    pk[p] = pack_bf16(e0, e1);
  }
}
// the two halves of exp_chunk_pk as separate steps (attention64p.cuh issues the next block's TMEM loads between them):
// scale_chunk: 32 raw scores -> 16 packed pairs x = s*scale - m;   exp_pairs: kN pairs [p0, p0 + kN) -> bf16 pairs + row sums
DEVI void scale_chunk(const uint32_t* v, float sc, float neg_m, uint64_t (&X)[16]) {
  const uint64_t NM = pack_f32x2(neg_m, neg_m);
  const uint64_t SC = pack_f32x2(sc, sc);
#pragma unroll
  for (int p = 0; p < 16; ++p) X[p] = fma_f32x2(pack_u32x2(v[2 * p], v[2 * p + 1]), SC, NM);
}
template <int kPolyPer8, int kDeg, int kP0, int kN>
DEVI void exp_pairs(const uint64_t (&X)[16], uint32_t (&pk)[16], uint64_t (&ls)[2]) {
#pragma unroll
  for (int p = kP0; p < kP0 + kN; ++p) {
    float e0, e1;
    const bool poly = (((p & 7) + 1) * kPolyPer8 / 8 != (p & 7) * kPolyPer8 / 8);   // spread over the 8 pairs
    if (poly) {
      exp2_poly_f32x2<kDeg>(X[p], e0, e1);
    } else {
      float x0, x1;
      unpack_f32x2(X[p], x0, x1);
      e0 = fast_exp2(x0);
      e1 = fast_exp2(x1);
    }
    ls[p & 1] = add_f32x2(ls[p & 1], pack_f32x2(e0, e1));
    pk[p] = pack_bf16(e0, e1);
  }
}

template <bool kScaled, int kPolyPer8, int kDeg, bool kSum = true>
DEVI void exp_chunk(const uint32_t* v, float sc, float neg_m, uint32_t tP, uint64_t (&ls)[2]) {
  uint32_t pk[16];
  exp_chunk_pk<kScaled, kPolyPer8, kDeg, kSum>(v, sc, neg_m, pk, ls);
  tmem_st16(tP, pk);
}

// One key block of the online softmax for one query row (one thread): S (fp32, BN columns at tS) -> registers in
// ONE TMEM pass -> block max -> (lazy) rescale -> exp2 -> P (bf16 pairs) over the first BN/2 columns of tS.
// m_ref: reference max of the row (log2 domain), m_run: largest score seen so far, l: row sum relative to m_ref.
// kToLeader (CTA-pair kernel, attention128p2.cuh): the handshake barriers live in the cluster's rank-0 CTA, whose thread issues the MMAs of both
template <bool kToLeader>
DEVI void softmax_arrive(uint64_t* bar) {
#ifdef LTXB200_ATTN128_2CTA_RELEASE               // A/B: the releasing remote arrive (a cluster-scope fence per hop)
  if (kToLeader) mbar_arrive_remote(bar, 0);
#else
  if (kToLeader) mbar_arrive_remote_relaxed(bar, 0);
#endif
  else mbar_arrive(bar);
}
template <int D, int BN, bool kPredicated, int kOW = D, bool kSum = true, int kEarly = BN / 2, bool kToLeader = false>   // kOW: accumulator columns the lazy rescale covers; kSum: row sum kept here; kEarly: keys written when p_half is signalled
DEVI void softmax_block(uint32_t tS, uint32_t tO, bool first, int kbase, int Lk, const float* bias, float sc,
                        float& m_ref, float& m_run, float& l, uint64_t* pv_done, uint32_t pv_parity,
                        uint64_t* s_read, uint64_t* p_half, int lane) {
#ifdef LTXB200_ABL_NOTMEM                        // diagnostic build (wrong results): the block's barrier handshakes without its TMEM traffic and arithmetic
  if (s_read) { tc_fence_before(); __syncwarp(); if (lane == 0) softmax_arrive<kToLeader>(s_read); }
  if (p_half) { tc_fence_before(); __syncwarp(); if (lane == 0) softmax_arrive<kToLeader>(p_half); }
  m_ref = 0.f; m_run = 0.f; l = 1.f;
  return;
#endif
  const float kLog2e = 1.4426950408889634f;
  const bool bias_vec = bias != nullptr && ((reinterpret_cast<uintptr_t>(bias) | (static_cast<uintptr_t>(kbase) << 2)) & 15) == 0;
  // A partial key block WITHOUT a bias (the tail of a sequence; the last block of a right-padded prompt given as a key length) takes the
  // lean path too: the raw scores of the keys beyond Lk are replaced by -inf (exp2 gives 0, the polynomial 2^-126) and everything else —
  // max on the raw scores, scale folded into the exp, packed arithmetic, polynomial share — is the code of a full block.
  const bool lean_tail = kPredicated && bias == nullptr;        // warp-uniform
  uint32_t v[BN];
  float mx0 = -INFINITY, mx1 = -INFINITY;
  // kLoadAll: every chunk of the row is requested before the first wait (one exposed TMEM latency per block instead of one per chunk —
  // the chunk-by-chunk form hides the latency behind the OTHER warp of the scheduler, which a tile that runs alone does not have)
#ifdef LTXB200_ATTN_LOADALL
  constexpr bool kLoadAll = (LTXB200_ATTN_LOADALL == 2) || (LTXB200_ATTN_LOADALL == 1 && D == 128);
#else
  constexpr bool kLoadAll = false;
#endif
  if (kLoadAll) {
#pragma unroll
    for (int c = 0; c < BN; c += 32) tmem_ld32(tS + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
  } else {
    tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
  }
#pragma unroll
  for (int c = 0; c < BN; c += 32) {
    if (!kLoadAll || c == 0) tmem_wait_ld();
    if (!kLoadAll && c + 32 < BN) tmem_ld32(tS + c + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c + 32]));
    if (lean_tail) {
      const int nv = Lk - kbase - c;                            // valid keys in this 32-column chunk
      if (nv < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= nv) v[c + i] = 0xff800000u;                  // -inf
      }
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(v[c + i]), __uint_as_float(v[c + i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(v[c + i + 2]), __uint_as_float(v[c + i + 3]));
      }
    } else if (kPredicated) {
      if (kbase + BN <= Lk && bias_vec) {
        // full block with a 16B-aligned bias row: vector loads (every lane reads the same addresses: one broadcast
        // transaction each), no tail predication
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + kbase + c + i));
          const float s0 = fmaf(__uint_as_float(v[c + i]), sc, b4.x * kLog2e), s1 = fmaf(__uint_as_float(v[c + i + 1]), sc, b4.y * kLog2e);
          const float s2 = fmaf(__uint_as_float(v[c + i + 2]), sc, b4.z * kLog2e), s3 = fmaf(__uint_as_float(v[c + i + 3]), sc, b4.w * kLog2e);
          v[c + i] = __float_as_uint(s0); v[c + i + 1] = __float_as_uint(s1);
          v[c + i + 2] = __float_as_uint(s2); v[c + i + 3] = __float_as_uint(s3);
          mx0 = fmax3(mx0, s0, s1);
          mx1 = fmax3(mx1, s2, s3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = kbase + c + i;
          float s = __uint_as_float(v[c + i]) * sc;
          if (bias && k < Lk) s = fmaf(__ldg(bias + k), kLog2e, s);
          if (k >= Lk) s = -INFINITY;
          v[c + i] = __float_as_uint(s);
          mx0 = fmaxf(mx0, s);
        }
      }
    } else {
      // four independent max chains per 32-score chunk (a softmax warp that has its scheduler to itself — d = 128, where the two tiles
      // alternate — is bound by dependency latency, not by issue slots; measured +0-3 % at d = 128, neutral at d = 64)
      float mx2 = -INFINITY, mx3 = -INFINITY;
#ifdef LTXB200_ABL_NOMAX                         // ablation build (wrong results): the block without its max pass
      mx0 = 0.f;
#else
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(v[c + i]), __uint_as_float(v[c + i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(v[c + i + 2]), __uint_as_float(v[c + i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(v[c + i + 4]), __uint_as_float(v[c + i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(v[c + i + 6]), __uint_as_float(v[c + i + 7]));
      }
#endif
      mx0 = fmaxf(mx0, mx2);
      mx1 = fmaxf(mx1, mx3);
    }
  }
  if (s_read) {                                  // the whole S row is in registers: columns 64.. may be overwritten
    tc_fence_before();
    __syncwarp();
    if (lane == 0) softmax_arrive<kToLeader>(s_read);
  }
  float m_blk = fmaxf(mx0, mx1);
  if (!kPredicated || lean_tail) m_blk *= sc;    // scale > 0: max commutes with the scaling
  if (first) {
    m_run = m_blk;
    m_ref = (m_blk == -INFINITY) ? 0.f : m_blk;
  } else {
    m_run = fmaxf(m_run, m_blk);
    rescale_o<kOW>(tO, m_run > m_ref + 8.0f, m_run, m_ref, l, pv_done, pv_parity);
  }
  uint64_t ls[2] = {0ull, 0ull};
#pragma unroll
  for (int c = 0; c < BN; c += 32) {
    if (kPredicated && !lean_tail) exp_chunk<true, AttnPoly<D>::kPer8, AttnPoly<D>::kDeg, kSum>(&v[c], sc, -m_ref, tS + (c >> 1), ls);
    else exp_chunk<false, AttnPoly<D>::kPer8, AttnPoly<D>::kDeg, kSum>(&v[c], sc, -m_ref, tS + (c >> 1), ls);
    if (p_half && c + 32 == kEarly) {            // P of keys 0..kEarly-1 is in TMEM: their P.V may start
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) softmax_arrive<kToLeader>(p_half);
    }
  }
  if (kSum) {
    float l0, l1;
    unpack_f32x2(add_f32x2(ls[0], ls[1]), l0, l1);
    l += l0 + l1;
  }
}

// Half-row form of softmax_block (kHalf): this thread owns columns [half*BN/2, +BN/2) of the row's scores, its partner thread (same
// TMEM lane, the warp four further on) the other half.  Block maxima meet in shared memory (slot parity = block parity, so a slot is
// rewritten only after the partner has passed the following barrier); P goes to this thread's half of the P columns, the row sum
// stays partial (the halves are added in the epilogue), the lazy rescale touches this thread's half of the O columns.
template <int D, int BN, bool kPredicated>
DEVI void softmax_block_half(uint32_t tS, uint32_t tO_half, int half, bool first, int kbase, int Lk, const float* bias, float sc,
                             float& m_ref, float& m_run, float& l, uint64_t* pv_done, uint32_t pv_parity, float* x_mine,
                             const float* x_partner, int bar_id, uint64_t* s_read = nullptr, uint64_t* p_half = nullptr, int lane = 0) {
  constexpr int NC = BN / 2;
  const float kLog2e = 1.4426950408889634f;
  const int c_off = half * NC;
  uint32_t v[NC];
  float mx0 = -INFINITY, mx1 = -INFINITY;
  tmem_ld32(tS + c_off, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
#pragma unroll
  for (int c = 0; c < NC; c += 32) {
    tmem_wait_ld();
    if (c + 32 < NC) tmem_ld32(tS + c_off + c + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c + 32]));
    if (kPredicated) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = kbase + c_off + c + i;
        float s = __uint_as_float(v[c + i]) * sc;
        if (bias && k < Lk) s = fmaf(__ldg(bias + k), kLog2e, s);
        if (k >= Lk) s = -INFINITY;
        v[c + i] = __float_as_uint(s);
        mx0 = fmaxf(mx0, s);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(v[c + i]), __uint_as_float(v[c + i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(v[c + i + 2]), __uint_as_float(v[c + i + 3]));
      }
    }
  }
  float m_blk = fmaxf(mx0, mx1);
  if (!kPredicated) m_blk *= sc;
  *x_mine = m_blk;
  // both halves have their scores in registers and their maximum published once this barrier is passed (so the partner may also
  // overwrite "my" score columns with its half of P)
  tc_fence_before();
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  if (s_read && half == 0) {                     // split mode: the whole S row is in registers (both halves): columns BN/2.. may be overwritten
    if (lane == 0) mbar_arrive(s_read);
  }
  tc_fence_after();
  m_blk = fmaxf(m_blk, *x_partner);
  if (first) {
    m_run = m_blk;
    m_ref = (m_blk == -INFINITY) ? 0.f : m_blk;
  } else {
    m_run = fmaxf(m_run, m_blk);
    rescale_o<D / 2>(tO_half, m_run > m_ref + 8.0f, m_run, m_ref, l, pv_done, pv_parity);
  }
  uint64_t ls[2] = {0ull, 0ull};
#pragma unroll
  for (int c = 0; c < NC; c += 32) exp_chunk<kPredicated, AttnPoly<D>::kPer8, AttnPoly<D>::kDeg>(&v[c], sc, -m_ref, tS + ((c_off + c) >> 1), ls);
  if (p_half && half == 0) {                     // split mode: P of keys 0..BN/2-1 (this half's) is in TMEM: their P.V may start
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_half);
  }
  float l0, l1;
  unpack_f32x2(add_f32x2(ls[0], ls[1]), l0, l1);
  l += l0 + l1;
}

template <int D, bool kMasked>
__global__ void __launch_bounds__(AttnCfg<D, kMasked>::kThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = AttnCfg<D, kMasked>;
  constexpr int BN = C::BN;
  constexpr int kOW = C::kOW;
  constexpr bool kOnes = C::kOnes;
  constexpr int kStagesK = C::kStagesK, kStagesV = C::kStagesV;
  constexpr int kSBufs = C::kSBufs;
  constexpr bool kSplit = C::kSplit;
  constexpr int kChunks = D / 64;                    // 64-wide (128 B) column chunks per row
  constexpr bool kHalf = C::kHalf;
  constexpr int kSmWarps = C::kSoftmaxWarps;
  constexpr int kTmaWarp = kSmWarps, kMmaWarp = kSmWarps + 1;
  constexpr uint32_t kColS0 = 0, kColO0 = kSBufs * BN;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                 // [2][128][D]
  uint8_t* sK = sQ + 2 * C::kQBytes;                  // [kStagesK][BN][D]
  uint8_t* sV = sK + kStagesK * C::kKBytes;           // [kStagesV][BN][D] (kOnes: each stage followed by its [BN][64] tile of ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesV * C::kVBytes);
  uint64_t* q_full = bars;                // [2]       TMA -> MMA
  uint64_t* q_empty = q_full + 2;         // [2]       last S of the item issued: Q tile may be overwritten
  uint64_t* k_full = q_empty + 2;         // [kStagesK]
  uint64_t* k_empty = k_full + kStagesK;
  uint64_t* v_full = k_empty + kStagesK;  // [kStagesV]
  uint64_t* v_empty = v_full + kStagesV;
  uint64_t* s_full = v_empty + kStagesV;  // [kSBufs]  S(n) complete in buffer n % kSBufs
  uint64_t* p_full = s_full + kSBufs;     // [kSBufs]  P(n) written over it (4 warps arrive)
  uint64_t* pv_done = p_full + kSBufs;    // [2]       P.V of tile t's block retired (lazy-rescale guard)
  uint64_t* o_done = pv_done + 2;         // [2]       last P.V of the item retired
  uint64_t* o_free = o_done + 2;          // [2]       epilogue has read O_t (4 warps arrive)
  uint64_t* s_read = o_free + 2;          // [kSBufs]  softmax holds S(n) in registers (kSplit; 4 warps arrive)
  uint64_t* p_half = s_read + kSBufs;     // [kSBufs]  first half of P(n) written (kSplit; 4 warps arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_half + kSBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // key blocks of a work item: all of Lk, or the batch element's own valid prefix (key_lens)
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
      mbar_init(&o_free[t], kSmWarps / 2);
    }
    for (int i = 0; i < kSBufs; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kSmWarps / 2);
      mbar_init(&s_read[i], 4);
      mbar_init(&p_half[i], 4);
    }
    for (int i = 0; i < kStagesK; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], C::k2Mma ? 2 : 1);     // k2Mma: both tiles' issuing threads release a stage
    }
    for (int i = 0; i < kStagesV; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], C::k2Mma ? 2 : 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<C::kTmemCols>(tmem_slot);
  if constexpr (kOnes) {
    // the constant value columns behind every V stage: all 64 columns of the tile are 1.0, so the 128-byte swizzle is immaterial
    for (int st = 0; st < kStagesV; ++st) {
      uint4* ones = reinterpret_cast<uint4*>(sV + st * C::kVBytes + C::kKBytes);
      for (int i = threadIdx.x; i < BN * 128 / 16; i += C::kThreads) ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    }
    fence_proxy_async();                               // generic-proxy stores -> visible to the MMA's shared-memory reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kSmWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C::kOtherRegs));
    if (warp == kTmaWarp && elect_one()) {
      // ================= TMA producer =================
      uint32_t kc = 0;
      int it = 0;
      for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
        const int qp = w % p.pairs, bh = w / p.pairs, h = bh % p.H, b = bh / p.H;
        const int nblk = (item_keys(b) + BN - 1) / BN;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait_backoff(&q_empty[t], (it & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[t], C::kQBytes);
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            tma_load_4d(sQ + t * C::kQBytes + c * (kAttnBM * 128), &tmQ, &q_full[t], c * 64, h, qp * 256 + t * kAttnBM, b);
        }
        // K(j + 1) is requested before V(j): the K ring runs one block ahead of the V ring
        auto load_k = [&](int j) {
          const uint32_t kk = kc + j;
          const int st = kk % kStagesK;
#if LTXB200_ATTN_L2PF > 0
          // K / V blocks kL2Pf ahead of the ring are requested into L2: the CTAs of a head walk its keys in near lock-step, so every block's
          // first touch is a DRAM miss that all of them wait for; the rings (2-4 stages) are too shallow to hide that latency themselves
          if (j + LTXB200_ATTN_L2PF < nblk) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              tma_prefetch_l2_4d(&tmK, c * 64, h, (j + LTXB200_ATTN_L2PF) * BN, b);
              tma_prefetch_l2_4d(&tmV, c * 64, h, (j + LTXB200_ATTN_L2PF) * BN, b);
            }
          }
#endif
          LTXB200_PRODUCER_WAIT(&k_empty[st], ((kk / kStagesK) & 1) ^ 1);
#ifdef LTXB200_ABL_NOLOAD                        // diagnostic build (wrong results): only the first fill of every ring stage is a real load
          if (kk >= static_cast<uint32_t>(kStagesK)) { mbar_arrive(&k_full[st]); return; }
#endif
          mbar_arrive_expect_tx(&k_full[st], C::kKBytes);
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            tma_load_4d(sK + st * C::kKBytes + c * (BN * 128), &tmK, &k_full[st], c * 64, h, j * BN, b);
        };
        load_k(0);
        for (int j = 0; j < nblk; ++j) {
          if (j + 1 < nblk) load_k(j + 1);
          const uint32_t vv = kc + j;
          const int sv = vv % kStagesV;
          LTXB200_PRODUCER_WAIT(&v_empty[sv], ((vv / kStagesV) & 1) ^ 1);
#ifdef LTXB200_ABL_NOLOAD
          if (vv >= static_cast<uint32_t>(kStagesV)) { mbar_arrive(&v_full[sv]); continue; }
#endif
          mbar_arrive_expect_tx(&v_full[sv], C::kKBytes);
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            tma_load_4d(sV + sv * C::kVBytes + c * (BN * 128), &tmV, &v_full[sv], c * 64, h, j * BN, b);
        }
        kc += nblk;
      }
    } else if ((warp == kMmaWarp || (C::k2Mma && warp == kMmaWarp + 1)) && elect_one()) {
      // ================= MMA issuer (k2Mma: this thread issues for tile `mt` only; otherwise for both) =================
#ifdef LTXB200_ABL_FREEMMA                       // diagnostic build (wrong results): the MMA thread never waits for the softmax warps (which do
#define LTXB200_SM_WAIT(bar, par) ((void)0)      // nothing): the rate of the tensor pipe + TMA rings alone on this kernel's instruction mix
#else
#define LTXB200_SM_WAIT(bar, par) mbar_wait_parked(bar, par)
#endif
      constexpr int kStep = C::k2Mma ? 2 : 1;           // stride of this thread through the global steps N = 2 * block + tile
      const int mt = C::k2Mma ? warp - kMmaWarp : 0;
      constexpr uint32_t idesc_s = umma_idesc_bf16(kAttnBM, BN, 0, 0);   // S = Q K^T  (both K-major)
      constexpr uint32_t idesc_sh = umma_idesc_bf16(kAttnBM, BN / 2, 0, 0);   // one half of the keys
      constexpr uint32_t idesc_sq = umma_idesc_bf16(kAttnBM, BN / 4, 0, 0);   // one quarter of the keys (kSplit34)
      constexpr uint32_t idesc_o = umma_idesc_bf16(kAttnBM, kOW, 0, 1);  // O += P V   (V is MN-major; kOnes: 16 more columns of ones behind it)
      // descriptors = (constant high bits | start address >> 4); tile / stage / k-step offsets are added to the
      // low word at issue time (the 14-bit address field cannot carry: shared memory is < 256 KB)
      const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ), 16, 1024);
      const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK), 16, 1024);
      const uint64_t vdesc = umma_smem_desc_sw128(smem_u32(sV), BN * 128, 1024);
      // Two cursors walk the CTA's work items.  The P.V cursor follows the softmax (step N: O_t += P(N).V); the S cursor runs kSBufs
      // steps ahead of it and keeps running ACROSS item boundaries: the first score blocks of the next item are issued while the last
      // P.V steps of the current one are still being produced, so an item with few key blocks (cross-attention: 2-4) does not pay a
      // tensor-pipe round trip per item.  Steps are numbered globally (N = 2 * block + tile), S(N) lives in buffer N % kSBufs.
      struct SCur { int w, it, nblk, m; uint32_t kc0, N; } sc{static_cast<int>(blockIdx.x), 0, 0, mt, 0u, static_cast<uint32_t>(mt)};
      auto s_enter = [&]() { if (sc.w < p.total) sc.nblk = (item_keys((sc.w / p.pairs) / p.H) + BN - 1) / BN; };
      auto s_advance = [&]() {
        sc.m += kStep; sc.N += kStep;
        if (sc.m >= 2 * sc.nblk) { sc.kc0 += sc.nblk; sc.m = mt; sc.w += gridDim.x; ++sc.it; s_enter(); }
      };
      // S at the cursor: tile m&1, key block m>>1 of item sc.w.
      // part: -1 = all BN keys; 1 = keys BN/2.. (issued first in split mode); 0 = keys 0..BN/2-1 (completes the step);
      //       kSplit34: 2 = keys 0..BN/4-1, 3 = keys BN/4..BN/2-1 (completes the step)
      auto issue_s = [&](int part) {
        const int t = sc.m & 1, j = sc.m >> 1;
        const uint32_t kpos = sc.kc0 + j;
        const int st = kpos % kStagesK;
        if ((t == 0 || C::k2Mma) && (part == 1 || part < 0)) mbar_wait_parked(&k_full[st], (kpos / kStagesK) & 1);
        if (j == 0) mbar_wait_parked(&q_full[t], sc.it & 1);
        tc_fence_after();
        const uint32_t buf = sc.N % kSBufs;
        const int key0 = part == 1 ? BN / 2 : (part == 3 ? BN / 4 : 0);              // first key (= first score column) of this part
        const uint64_t qa = qdesc + static_cast<uint32_t>(t * (C::kQBytes >> 4));
        const uint64_t ka = kdesc + static_cast<uint32_t>(st * (C::kKBytes >> 4)) + static_cast<uint32_t>((key0 * 128) >> 4);
        const uint32_t ts = tmem_base + kColS0 + buf * BN + key0;
        const uint32_t idesc = part < 0 ? idesc_s : (part >= 2 ? idesc_sq : idesc_sh);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t offa = ((ks >> 2) * (kAttnBM * 128) + (ks & 3) * 32) >> 4;
          const uint32_t offb = ((ks >> 2) * (BN * 128) + (ks & 3) * 32) >> 4;
          umma_ss(ts, qa + offa, ka + offb, idesc, ks ? 1u : 0u);
        }
        if (part != 1 && part != 2) {
          umma_commit(&s_full[buf]);
          if (j + 1 == sc.nblk) umma_commit(&q_empty[t]);
          if (t == 1 || C::k2Mma) umma_commit(&k_empty[st]);
        }
      };
      // O_t += P(n)[:, keys] . V(j)[keys, :]   (k-steps [ks0, ks1) of 16 keys)
      auto issue_pv = [&](int t, uint32_t buf, int sv, int ks0, int ks1, bool acc) {
        // B = V[16 keys (K), D (N)], N contiguous: 8-key groups 1024 B apart (SBO), 64-col groups one chunk apart (LBO)
        const uint64_t va = vdesc + static_cast<uint32_t>(sv * (C::kVBytes >> 4));
        const uint32_t to = tmem_base + kColO0 + t * kOW, tp = tmem_base + kColS0 + buf * BN;
#pragma unroll
        for (int ks = ks0; ks < ks1; ++ks)
          umma_ts(to, tp + ks * 8, va + static_cast<uint32_t>(ks * (2048 >> 4)), idesc_o, (acc || ks > ks0) ? 1u : 0u);
      };
      s_enter();
#pragma unroll 1
      for (int i = 0; i < kSBufs / kStep && sc.w < p.total; ++i) { issue_s(-1); s_advance(); }
      uint32_t vc = 0;               // V ring counter
      uint32_t N = mt;               // global step of the P.V cursor
      int it = 0;
      for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
        const int nblk = (item_keys((w / p.pairs) / p.H) + BN - 1) / BN;
        const int nsteps = 2 * nblk;
        // step n: O_t += P(n).V(j), then the S cursor's block into the buffer P(n) leaves
#pragma unroll 1
        for (int n = mt; n < nsteps; n += kStep, N += kStep) {
          const int t = n & 1, j = n >> 1;
          const int sv = vc % kStagesV;
          const uint32_t buf = N % kSBufs, par = (N / kSBufs) & 1;
          const bool more = sc.w < p.total;        // the cursor then stands at step N + kSBufs, i.e. on this step's buffer
          if (kSplit) {
            if (more) { LTXB200_SM_WAIT(&s_read[buf], par); issue_s(1); }
            LTXB200_SM_WAIT(&p_half[buf], par);
            if (j == 0) LTXB200_SM_WAIT(&o_free[t], (it & 1) ^ 1);
            if (t == 0 || C::k2Mma) mbar_wait_parked(&v_full[sv], (vc / kStagesV) & 1);
            tc_fence_after();
            constexpr int kEarlySteps = C::kSplit34 ? 3 * BN / 64 : BN / 32;      // k-steps (16 keys) covered by the early P signal
            issue_pv(t, buf, sv, 0, kEarlySteps, j > 0);
            if (C::kSplit34 && more) issue_s(2);                                   // score columns 0..BN/4-1 = P of keys 0..BN/2-1: consumed
            LTXB200_SM_WAIT(&p_full[buf], par);
            tc_fence_after();
            issue_pv(t, buf, sv, kEarlySteps, BN / 16, true);
          } else {
            LTXB200_SM_WAIT(&p_full[buf], par);
            if (j == 0) LTXB200_SM_WAIT(&o_free[t], (it & 1) ^ 1);
            if (t == 0) mbar_wait_parked(&v_full[sv], (vc / kStagesV) & 1);
            tc_fence_after();
            issue_pv(t, buf, sv, 0, BN / 16, j > 0);
          }
          umma_commit(&pv_done[t]);
          if (t == 1 || C::k2Mma) { umma_commit(&v_empty[sv]); ++vc; }
          if (j + 1 == nblk) umma_commit(&o_done[t]);
          if (more) { issue_s(kSplit ? (C::kSplit34 ? 3 : 0) : -1); s_advance(); }
        }
      }
#ifdef LTXB200_ABL_FREEMMA
      if (it > 0) mbar_wait(&o_done[1], (it - 1) & 1);       // nobody else waits for the pipe to drain
#endif
    }
  } else {
    // ================= softmax / correction / epilogue: warpgroup t owns query tile t =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::kSoftmaxRegs));
    const int t = kHalf ? (warp >> 3) : (warp >> 2);          // query tile of this warp
    const int half = kHalf ? ((warp >> 2) & 1) : 0;           // which half of the score columns (kHalf)
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(sub * 32) << 16;
    const uint32_t tO = tmem_base + kColO0 + t * kOW + lane_addr;
    // kHalf exchange slots behind the barriers: xmax[tile][half][parity][row], xsum[tile][half][row]
    float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);
    float* xsum = xmax + 2 * 2 * 2 * 128;
    const int pair_bar = 1 + t * 4 + sub;      // named barrier of the two warps that share these 32 rows
    uint32_t G = 0;                  // global key-block counter of this tile; its step is N = 2*G + t
    int it = 0;
#ifdef LTXB200_ABL_FREEMMA
    if (false)
#endif
    for (int w = blockIdx.x; w < p.total; w += gridDim.x, ++it) {
      const int qp = w % p.pairs, bh = w / p.pairs, h = bh % p.H, b = bh / p.H;
      const float* bias = p.key_bias ? p.key_bias + static_cast<long long>(b) * p.Lk : nullptr;
      const int Lk_b = item_keys(b);
      const int nblk = (Lk_b + BN - 1) / BN;
      float m_ref = 0.f, m_run = 0.f, l = 0.f;
      for (int j = 0; j < nblk; ++j, ++G) {
        const uint32_t N = 2 * G + t, buf = N % kSBufs;
        mbar_wait(&s_full[buf], (N / kSBufs) & 1);
        tc_fence_after();
#ifdef LTXB200_ATTN128_PINGPONG
        // strict ping-pong of the two softmax warpgroups (named barriers 9 / 10, 256 threads each: 128 waiting + 128 arriving): a tile's
        // softmax section runs alone on the SM while the other tile's MMAs are in flight
        if (kSplit && !kHalf) {
          if (G == 0 && t == 1) asm volatile("bar.arrive 9, 256;" ::: "memory");      // tile 0 goes first
          asm volatile("bar.sync %0, 256;" ::"r"(9 + t) : "memory");
        }
#endif
        const int kbase = j * BN;
        const uint32_t tS = tmem_base + kColS0 + buf * BN + lane_addr;
        // full blocks without a bias take the lean path; the tail block / biased blocks take the predicated one
        if constexpr (kHalf) {
          float* xm = xmax + ((t * 2 + half) * 2 + (G & 1)) * 128 + row;
          const float* xp = xmax + ((t * 2 + (half ^ 1)) * 2 + (G & 1)) * 128 + row;
          if (kMasked && (bias != nullptr || kbase + BN > Lk_b))
            softmax_block_half<D, BN, true>(tS, tO + half * (D / 2), half, j == 0, kbase, Lk_b, bias, p.scale_log2, m_ref, m_run, l, &pv_done[t],
                                            (G - 1) & 1, xm, xp, pair_bar, kSplit ? &s_read[buf] : nullptr, kSplit ? &p_half[buf] : nullptr, lane);
          else
            softmax_block_half<D, BN, false>(tS, tO + half * (D / 2), half, j == 0, kbase, Lk_b, bias, p.scale_log2, m_ref, m_run, l, &pv_done[t],
                                             (G - 1) & 1, xm, xp, pair_bar, kSplit ? &s_read[buf] : nullptr, kSplit ? &p_half[buf] : nullptr, lane);
        } else
        if (kMasked && (bias != nullptr || kbase + BN > Lk_b))
          softmax_block<D, BN, true, kOW, !kOnes, C::kSplit34 ? 3 * BN / 4 : BN / 2>(tS, tO, j == 0, kbase, Lk_b, bias, p.scale_log2, m_ref, m_run, l, &pv_done[t], (G - 1) & 1,
                                                  kSplit ? &s_read[buf] : nullptr, kSplit ? &p_half[buf] : nullptr, lane);
        else
          softmax_block<D, BN, false, kOW, !kOnes, C::kSplit34 ? 3 * BN / 4 : BN / 2>(tS, tO, j == 0, kbase, Lk_b, bias, p.scale_log2, m_ref, m_run, l, &pv_done[t], (G - 1) & 1,
                                                   kSplit ? &s_read[buf] : nullptr, kSplit ? &p_half[buf] : nullptr, lane);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
#ifdef LTXB200_ATTN128_PINGPONG
        if (kSplit && !kHalf) asm volatile("bar.arrive %0, 256;" ::"r"(9 + (t ^ 1)) : "memory");
#endif
      }
      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(&o_done[t], it & 1);
      tc_fence_after();
      if constexpr (kHalf) {                      // the row sum is the sum of the two halves' partial sums
        xsum[(t * 2 + half) * 128 + row] = l;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        l += xsum[(t * 2 + (half ^ 1)) * 128 + row];
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");      // the slot is free for the next item
      }
      if constexpr (kOnes) {                      // the row sum is the tile's accumulator column D (copies in D+1 .. D+15)
        uint32_t lr[16];
        tmem_ld16(tO + D, lr);
        tmem_wait_ld();
        l = __uint_as_float(lr[0]);
      }
      const float inv = (l > 0.f) ? __fdividef(1.0f, l) : 0.f;
      const int q = qp * 256 + t * kAttnBM + row;
      __nv_bfloat16* orow;
      if (p.peers.P > 0) {
        const int dst = q / p.tokens_per_peer, nl = q - dst * p.tokens_per_peer;
        orow = static_cast<__nv_bfloat16*>(p.peers.data[dst < p.peers.P ? dst : 0]) +
               (static_cast<long long>(b) * p.tokens_per_peer + nl) * p.out_ld + (p.head_offset + h) * D;
      } else {
        orow = p.out + static_cast<long long>(b) * p.out_bs + static_cast<long long>(q) * p.out_ld + h * D;
      }
#pragma unroll 1
      for (int c = kHalf ? half * (D / 2) : 0; c < (kHalf ? (half + 1) * (D / 2) : D); c += 32) {
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
      if (lane == 0) mbar_arrive(&o_free[t]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
  if (p.peers.P > 0) peer_signal_done(p.peers, gridDim.x);
}

}  // namespace b200
