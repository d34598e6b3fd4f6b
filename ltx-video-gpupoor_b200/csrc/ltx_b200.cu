// C-ABI entry points of libltx_b200.so (see include/ltx_b200.h).  Host-side launch logic only.
#include "../../include/ltx_b200.h"

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "attention.cuh"
#include "attention64p.cuh"
#include "attention128p2.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "conv_halo.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "tensormap.h"

using namespace b200;

static std::atomic<long long> g_launches{0};
static int g_num_sms = 0;

static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int launch_status() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaPeekAtLastError() == cudaSuccess ? kOk : kErrCuda;
}
static inline int ew_blocks(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int ltxb200_abi_version(void) { return 4; }
extern "C" long long ltxb200_launch_count(void) { return g_launches.load(); }
extern "C" const char* ltxb200_error_string(int code) {
  switch (code) {
    case kOk: return "ok";
    case kErrBadShape: return "bad shape";
    case kErrBadAlign: return "pointer or stride not 16-byte aligned";
    case kErrCuda: return "CUDA launch error";
    case kErrTensorMap: return "cuTensorMapEncodeTiled failed";
    case kErrUnsupported: return "unsupported configuration";
    default: return "unknown error";
  }
}

// ------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------
template <int BN, bool kConv>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_tiles,
                       cudaStream_t st) {
  using S = GemmSmem<BN>;
  static bool configured = false;
  auto kern = gemm_bf16_kernel<BN, kConv>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal) != cudaSuccess) return kErrCuda;
    configured = true;
  }
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kern<<<grid, kGemmThreads, S::kTotal, st>>>(ta, tb, p);
  return launch_status();
}

// CTA-pair variant: cluster of 2 along M, grid = 2 x min(super tiles, SMs / 2)
template <int BN, bool kConv>
static int launch_gemm_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_tiles, cudaStream_t st) {
  using S = GemmSmem<BN, 2>;
  static bool configured = false;
  auto kern = gemm_bf16_kernel<BN, kConv, 2>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal) != cudaSuccess) return kErrCuda;
    configured = true;
  }
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (num_tiles < pairs ? num_tiles : pairs));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, tb, p) != cudaSuccess) return kErrCuda;
  return launch_status();
}

// halo-tiled convolution (conv_halo.cuh): one CTA per SM, persistent over 16x16 patches
template <int BN>
static int launch_conv_halo(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_tiles, cudaStream_t st) {
  using S = ConvHaloSmem<BN>;
  static bool configured = false;
  auto kern = conv3d_halo_kernel<BN>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal) != cudaSuccess) return kErrCuda;
    configured = true;
  }
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kern<<<grid, kGemmThreads, S::kTotal, st>>>(ta, tb, p);
  return launch_status();
}

static int fill_peers(PeerPtrs* pp, int P, int rank, void* const* data_ptrs, void* const* flag_ptrs, unsigned int epoch, void* counter);

struct VScatter {          // see GemmParams::vs_peers
  PeerPtrs peers;
  int col0, group_cols, B, tokens_per_batch, token_offset, row0;
};

static int gemm_impl(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                     void* out, int64_t ldc, int out_f32, const void* bias, int act,
                     const void* residual, int64_t ldr, const void* gate, int64_t gate_ld,
                     int rows_per_gate, int mod_f32, void* stream, const VScatter* vs = nullptr) {
  if (M <= 0 || N <= 0 || K <= 0 || (K & 7) || (N & 7)) return kErrBadShape;
  if (!aligned16(A) || !aligned16(W) || !aligned16(out) || (lda & 7) || (ldw & 7) || (ldc & (out_f32 ? 3 : 7)))
    return kErrBadAlign;
  if ((bias && !aligned16(bias)) || (residual && (!aligned16(residual) || (ldr & (mod_f32 ? 3 : 7)))) ||
      (gate && (!aligned16(gate) || (gate_ld & (mod_f32 ? 3 : 7)) || rows_per_gate <= 0)))
    return kErrBadAlign;
  int BN = (N <= 128) ? 128 : 256;
  static const int env_2cta = getenv("LTXB200_GEMM_2CTA") ? atoi(getenv("LTXB200_GEMM_2CTA")) : 1;
  const bool two_cta = env_2cta && BN == 256 && M > 128;
  // 256 x 512 pair tiles (single-buffered accumulator, a third less operand traffic per flop), OPT-IN: parity-green, and measured
  // slower on FFN-down (18432 x 2048 x 8192): 1485 vs 1498 TF/s burst; sustained at the power cap 1247 TF/s @ 1477 MHz vs 1290 @ 1275 MHz —
  // the chip clocks 16 % higher on the lighter operand traffic but the un-overlapped epilogue and the two half-width MMAs per k-step cost more.
  // LTXB200_GEMM_BN512: 0 = never (default), 1 = K >= 4096, 2 = whenever the shape allows
  static const int env_bn512 = getenv("LTXB200_GEMM_BN512") ? atoi(getenv("LTXB200_GEMM_BN512")) : 0;
  const bool wide = two_cta && !vs && (N % 512 == 0) && env_bn512 && (env_bn512 == 2 || K >= 4096);
  if (wide) BN = 512;
  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t str[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {kGemmBK, kGemmBM};
    if (make_tmap_bf16(&ta, A, 2, dims, str, box)) return kErrTensorMap;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {kGemmBK, static_cast<uint32_t>(wide ? 128 : (two_cta ? BN / 2 : BN))};     // CTA pair: each CTA loads half of the B tile (256 x 512: in two boxes)
    if (make_tmap_bf16(&tb, W, 2, dims, str, box)) return kErrTensorMap;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.out = out; p.ldc = ldc; p.out_f32 = out_f32;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.act = act;
  if (mod_f32) { p.residual32 = static_cast<const float*>(residual); p.gate32 = static_cast<const float*>(gate); }
  else { p.residual = static_cast<const __nv_bfloat16*>(residual); p.gate = static_cast<const __nv_bfloat16*>(gate); }
  p.ldr = ldr; p.gate_ld = gate_ld; p.rows_per_gate = rows_per_gate > 0 ? rows_per_gate : 1;
  p.store_mode = kStoreRowMajor;
  if (vs) {
    p.vs_peers = vs->peers; p.vs_col0 = vs->col0; p.vs_group_cols = vs->group_cols; p.vs_B = vs->B;
    p.vs_tokens_per_batch = vs->tokens_per_batch; p.vs_token_offset = vs->token_offset; p.vs_row0 = vs->row0;
  }
  // the weights of this path always fit the 126 MB L2 (<= 34 MB); A often does not (FFN-down: 302 MB)
  p.n_fastest = (static_cast<long long>(N) * K * 2 <= (48ll << 20)) ? 1 : 0;
  if (const char* e = getenv("LTXB200_GEMM_RASTER")) p.n_fastest = atoi(e);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (wide) return launch_gemm_2cta<512, false>(ta, tb, p, ((M + 2 * kGemmBM - 1) / (2 * kGemmBM)) * (N / 512), st);
  if (two_cta) return launch_gemm_2cta<256, false>(ta, tb, p, ((M + 2 * kGemmBM - 1) / (2 * kGemmBM)) * ((N + BN - 1) / BN), st);
  const int tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + BN - 1) / BN);
  return BN == 128 ? launch_gemm<128, false>(ta, tb, p, tiles, st) : launch_gemm<256, false>(ta, tb, p, tiles, st);
}

extern "C" int ltxb200_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                                 void* out, int64_t ldc, int out_f32, const void* bias, int act,
                                 const void* residual, int64_t ldr, const void* gate, int64_t gate_ld,
                                 int rows_per_gate, void* stream) {
  return gemm_impl(A, lda, W, ldw, M, N, K, out, ldc, out_f32, bias, act, residual, ldr, gate, gate_ld, rows_per_gate, 0, stream);
}

extern "C" int ltxb200_gemm_qkv_vscatter_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int K, int D, void* out,
                                             int64_t ldc, const void* bias, int head_dim, int tokens_per_batch, int token_offset,
                                             int row0, int B, int P, int rank, void* const* recv_ptrs, void* stream) {
  if (D <= 0 || head_dim <= 0 || (head_dim % 32) || (D % head_dim) || ((D / head_dim) % P) || B <= 0 || tokens_per_batch <= 0 || row0 < 0)
    return kErrBadShape;
  VScatter vs{};
  // no flag is touched here: the q/k scatter kernel that follows in the stream publishes the exchange's epoch
  if (!recv_ptrs || P < 1 || P > kMaxPeers || rank < 0 || rank >= P) return kErrBadShape;
  for (int r = 0; r < P; ++r) {
    if (!recv_ptrs[r] || !aligned16(recv_ptrs[r])) return kErrBadAlign;
    vs.peers.data[r] = recv_ptrs[r];
  }
  vs.peers.P = P; vs.peers.rank = rank;
  vs.col0 = 2 * D; vs.group_cols = D / P; vs.B = B; vs.tokens_per_batch = tokens_per_batch; vs.token_offset = token_offset; vs.row0 = row0;
  return gemm_impl(A, lda, W, ldw, M, 3 * D, K, out, ldc, 0, bias, kActNone, nullptr, 0, nullptr, 0, 1, 0, stream, &vs);
}

extern "C" int ltxb200_gemm_bf16_f32res(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                                        float* out, int64_t ldc, const void* bias, int act, const float* residual, int64_t ldr,
                                        const float* gate, int64_t gate_ld, int rows_per_gate, void* stream) {
  return gemm_impl(A, lda, W, ldw, M, N, K, out, ldc, 1, bias, act, residual, ldr, gate, gate_ld, rows_per_gate, 1, stream);
}

// ------------------------------------------------------------------------------------------
// conv3d (implicit GEMM)
// ------------------------------------------------------------------------------------------
// x: [B, Tin, Hin, Win, Cin]; out: [B, T, H, W, Cout] with T = (Tin - 1) / st + 1 etc. (st, shw = output strides)
static int conv_impl(const void* x, const void* w, const void* bias, void* out, int B, int Tin, int Hin, int Win, int Cin, int Cout,
                     int taps_t, int taps_hw, int causal, int tpad_zero, int store_mode, int out_f32, const void* residual,
                     void* stream, int st_t = 1, int st_hw = 1, int off_hw = 0, int norm_mode = 0, void* out2 = nullptr,
                     float norm_eps = 0.f) {
  if (norm_mode) {      // fused PixelNorm + SiLU second output: one N tile must hold the whole channel vector
    if ((norm_mode != 1 && norm_mode != 2) || !out2 || !aligned16(out2) || Cout > 256 || store_mode != LTXB200_CONV_STORE_NDHWC || out_f32)
      return kErrUnsupported;
    if (norm_mode == 1 && !out) return kErrBadAlign;
    if (norm_mode == 2 && !out) out = out2;           // never written; keeps the pointer checks below uniform
  }
  if (B <= 0 || Tin <= 0 || Hin <= 0 || Win <= 0 || (Cin % 64) || (Cout & 7)) return kErrBadShape;
  if ((st_t != 1 && st_t != 2) || (st_hw != 1 && st_hw != 2)) return kErrUnsupported;
  if ((st_t != 1 || st_hw != 1) && (!causal || store_mode != LTXB200_CONV_STORE_NDHWC || residual)) return kErrUnsupported;
  if ((st_t == 2 && taps_t != 3) || (st_hw == 2 && taps_hw != 3) || (off_hw != 0 && (off_hw != 1 || taps_hw != 3))) return kErrUnsupported;
  if (Hin - 1 - off_hw < 0 || Win - 1 - off_hw < 0) return kErrBadShape;
  const int T = (Tin - 1) / st_t + 1, H = (Hin - 1 - off_hw) / st_hw + 1, W = (Win - 1 - off_hw) / st_hw + 1;
  if ((taps_t != 1 && taps_t != 3) || (taps_hw != 1 && taps_hw != 3)) return kErrUnsupported;
  if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias)) || (residual && !aligned16(residual)))
    return kErrBadAlign;
  if (store_mode == LTXB200_CONV_STORE_D2S && ((Cout % 8) || ((Cout / 8) % 32))) return kErrBadShape;
  if (store_mode == LTXB200_CONV_STORE_UNPATCH && (Cout % 16)) return kErrBadShape;
  if (store_mode != LTXB200_CONV_STORE_NDHWC && residual) return kErrUnsupported;
  // Narrow outputs (Cout <= 128) with 3x3 spatial taps and unit stride: the halo-tiled kernel (conv_halo.cuh), which re-uses every
  // activation box for three taps and every weight tile for two 128-voxel halves (the plain implicit GEMM is bound by L2 -> SM
  // traffic at these shapes: profiles/r02_ncu_summary.md).  LTXB200_CONV_HALO=0 keeps the plain kernel (A/B, tests).
  const char* eh = getenv("LTXB200_CONV_HALO");
  const int env_halo = eh ? atoi(eh) : 1;
  if (env_halo && taps_hw == 3 && st_t == 1 && st_hw == 1 && off_hw == 0 && Cout <= 128) {
    const int taps = taps_t * 9;
    const int BNh = Cout <= 64 ? 64 : 128;
    CUtensorMap ta, tb;
    {
      uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                          static_cast<uint64_t>(Tin), static_cast<uint64_t>(B)};
      uint64_t str[4] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                         static_cast<uint64_t>(Hin) * Win * Cin * 2, static_cast<uint64_t>(Tin) * Hin * Win * Cin * 2};
      uint32_t box[5] = {kGemmBK, kHaloTile, kHaloTile + 2, 1, 1};
      if (make_tmap_bf16(&ta, x, 5, dims, str, box)) return kErrTensorMap;
    }
    {
      uint64_t dims[2] = {static_cast<uint64_t>(taps) * Cin, static_cast<uint64_t>(Cout)};
      uint64_t str[1] = {static_cast<uint64_t>(taps) * Cin * 2};
      uint32_t box[2] = {kGemmBK, static_cast<uint32_t>(BNh)};
      if (make_tmap_bf16(&tb, w, 2, dims, str, box)) return kErrTensorMap;
    }
    GemmParams p{};
    p.M = B * T * H * W; p.N = Cout; p.K = taps * Cin;
    p.out = out; p.ldc = Cout; p.out_f32 = out_f32;
    p.bias = static_cast<const __nv_bfloat16*>(bias);
    p.act = kActNone;
    p.residual = static_cast<const __nv_bfloat16*>(residual); p.ldr = Cout;
    p.gate = nullptr; p.rows_per_gate = 1;
    p.store_mode = store_mode == LTXB200_CONV_STORE_NDHWC ? kStoreRowMajor
                   : (store_mode == LTXB200_CONV_STORE_D2S ? kStoreConvD2S : kStoreConvUnpatch);
    p.cB = B; p.cT = T; p.cH = H; p.cW = W; p.cCin = Cin; p.cBH = kHaloTile; p.cBW = kHaloTile;
    p.c_tiles_h = (H + kHaloTile - 1) / kHaloTile; p.c_tiles_w = (W + kHaloTile - 1) / kHaloTile;
    p.c_causal = causal ? 1 : 0; p.c_taps_t = taps_t; p.c_taps_hw = 3; p.c_tpad_zero = tpad_zero ? 1 : 0;
    p.c_st = 1; p.c_shw = 1; p.cTin = Tin; p.c_off_hw = 0;
    p.norm_mode = norm_mode; p.out2 = static_cast<__nv_bfloat16*>(out2); p.norm_eps = norm_eps;
    const long long tiles = static_cast<long long>(B) * T * p.c_tiles_h * p.c_tiles_w;
    if (tiles > 0x7fffffffLL) return kErrBadShape;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return BNh == 64 ? launch_conv_halo<64>(ta, tb, p, static_cast<int>(tiles), st) : launch_conv_halo<128>(ta, tb, p, static_cast<int>(tiles), st);
  }
  // pick the 128-voxel patch shape with the least padding
  static const int shapes[8][2] = {{8, 16}, {16, 8}, {4, 32}, {32, 4}, {2, 64}, {64, 2}, {1, 128}, {128, 1}};
  int best = 0;
  long long best_area = -1;
  for (int i = 0; i < 8; ++i) {
    const long long th = (H + shapes[i][0] - 1) / shapes[i][0], tw = (W + shapes[i][1] - 1) / shapes[i][1];
    const long long area = th * tw;
    if (best_area < 0 || area < best_area) { best_area = area; best = i; }
  }
  const int BH = shapes[best][0], BW = shapes[best][1];
  const int BN = (Cout <= 128) ? 128 : 256;
  // CTA pairs for the convolution are implemented and parity-tested but OFF by default: the full 768x512x121 VAE decode measured
  // 58.7 ms without and 58.3 ms with them (the decode is bound by its activation traffic, not by operand fill).
  // LTXB200_CONV_2CTA: unset / 0 = never, 1 = when there are enough patches to fill the machine twice, 2 = always (tests).
  const char* e2 = getenv("LTXB200_CONV_2CTA");
  const int env_2cta = e2 ? atoi(e2) : 0;
  const bool two_cta = env_2cta == 2 || (env_2cta == 1 && static_cast<long long>(B) * T * ((H + BH - 1) / BH) * ((W + BW - 1) / BW) >= 2 * num_sms());
  const int taps = taps_t * taps_hw * taps_hw;
  CUtensorMap ta, tb;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Win), static_cast<uint64_t>(Hin),
                        static_cast<uint64_t>(Tin), static_cast<uint64_t>(B)};
    uint64_t str[4] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Win) * Cin * 2,
                       static_cast<uint64_t>(Hin) * Win * Cin * 2, static_cast<uint64_t>(Tin) * Hin * Win * Cin * 2};
    // strided convolution: the box spans st_hw * BW source pixels and the traversal stride picks every st_hw-th one
    uint32_t box[5] = {kGemmBK, static_cast<uint32_t>(BW * st_hw), static_cast<uint32_t>(BH * st_hw), 1, 1};
    uint32_t es[5] = {1, static_cast<uint32_t>(st_hw), static_cast<uint32_t>(st_hw), 1, 1};
    if (make_tmap_bf16(&ta, x, 5, dims, str, box, es)) return kErrTensorMap;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(taps) * Cin, static_cast<uint64_t>(Cout)};
    uint64_t str[1] = {static_cast<uint64_t>(taps) * Cin * 2};
    uint32_t box[2] = {kGemmBK, static_cast<uint32_t>(two_cta ? BN / 2 : BN)};      // CTA pair: each CTA loads half of the weight tile
    if (make_tmap_bf16(&tb, w, 2, dims, str, box)) return kErrTensorMap;
  }
  GemmParams p{};
  p.M = B * T * H * W; p.N = Cout; p.K = taps * Cin;
  p.out = out; p.ldc = Cout; p.out_f32 = out_f32;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.act = kActNone;
  p.residual = static_cast<const __nv_bfloat16*>(residual); p.ldr = Cout;
  p.gate = nullptr; p.rows_per_gate = 1;
  p.store_mode = store_mode == LTXB200_CONV_STORE_NDHWC ? kStoreRowMajor
                 : (store_mode == LTXB200_CONV_STORE_D2S ? kStoreConvD2S : kStoreConvUnpatch);
  p.cB = B; p.cT = T; p.cH = H; p.cW = W; p.cCin = Cin; p.cBH = BH; p.cBW = BW;
  p.c_tiles_h = (H + BH - 1) / BH; p.c_tiles_w = (W + BW - 1) / BW;
  p.c_causal = causal ? 1 : 0; p.c_taps_t = taps_t; p.c_taps_hw = taps_hw; p.c_tpad_zero = tpad_zero ? 1 : 0;
  p.c_st = st_t; p.c_shw = st_hw; p.cTin = Tin; p.c_off_hw = off_hw;
  p.norm_mode = norm_mode; p.out2 = static_cast<__nv_bfloat16*>(out2); p.norm_eps = norm_eps;
  p.n_fastest = (static_cast<long long>(Cout) * taps * Cin * 2 <= (48ll << 20)) ? 1 : 0;
  if (const char* e = getenv("LTXB200_GEMM_RASTER")) p.n_fastest = atoi(e);
  const int patches = B * T * p.c_tiles_h * p.c_tiles_w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (two_cta) {
    const int tiles2 = ((patches + 1) / 2) * ((Cout + BN - 1) / BN);
    return BN == 128 ? launch_gemm_2cta<128, true>(ta, tb, p, tiles2, st) : launch_gemm_2cta<256, true>(ta, tb, p, tiles2, st);
  }
  const int tiles = patches * ((Cout + BN - 1) / BN);
  return BN == 128 ? launch_gemm<128, true>(ta, tb, p, tiles, st) : launch_gemm<256, true>(ta, tb, p, tiles, st);
}

extern "C" int ltxb200_conv3d_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H,
                                   int W, int Cin, int Cout, int causal, int store_mode, int out_f32,
                                   const void* residual, void* stream) {
  return conv_impl(x, w, bias, out, B, T, H, W, Cin, Cout, 3, 3, causal, 0, store_mode, out_f32, residual, stream);
}

extern "C" int ltxb200_conv3d_norm_bf16(const void* x, const void* w, const void* bias, void* out, void* out2, int B, int T, int H,
                                        int W, int Cin, int Cout, int causal, const void* residual, int norm_mode, float eps,
                                        void* stream) {
  if (norm_mode != 1 && norm_mode != 2) return kErrBadShape;
  return conv_impl(x, w, bias, out, B, T, H, W, Cin, Cout, 3, 3, causal, 0, LTXB200_CONV_STORE_NDHWC, 0, residual, stream, 1, 1, 0,
                   norm_mode, out2, eps);
}

extern "C" int ltxb200_conv3d_strided_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W,
                                           int Cin, int Cout, int stride_t, int stride_hw, void* stream) {
  return conv_impl(x, w, bias, out, B, T, H, W, Cin, Cout, 3, 3, 1, 0, LTXB200_CONV_STORE_NDHWC, 0, nullptr, stream, stride_t,
                   stride_hw);
}

extern "C" int ltxb200_conv_taps_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W,
                                      int Cin, int Cout, int taps_t, int taps_hw, int causal_zero_pad, const void* residual,
                                      void* stream) {
  // causal_zero_pad: 0 = causal taps (t-2,t-1,t), replicate padding; 1 = causal, zero padding; 2 = centred taps (t-1,t,t+1), zero padding
  return conv_impl(x, w, bias, out, B, T, H, W, Cin, Cout, taps_t, taps_hw, causal_zero_pad == 2 ? 0 : 1, causal_zero_pad != 0,
                   LTXB200_CONV_STORE_NDHWC, 0, residual, stream);
}

extern "C" int ltxb200_conv_taps_strided_bf16(const void* x, const void* w, const void* bias, void* out, int B, int T, int H, int W,
                                              int Cin, int Cout, int taps_t, int taps_hw, int stride_t, int stride_hw,
                                              int off_hw, void* stream) {
  return conv_impl(x, w, bias, out, B, T, H, W, Cin, Cout, taps_t, taps_hw, 1, 1, LTXB200_CONV_STORE_NDHWC, 0, nullptr, stream,
                   stride_t, stride_hw, off_hw);
}

// ------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------
static int make_qkv_tmap(CUtensorMap* m, const void* base, int B, int H, int L, int d, int64_t ld, int64_t bs, int box_rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(H), static_cast<uint64_t>(L), static_cast<uint64_t>(B)};
  uint64_t str[3] = {static_cast<uint64_t>(d) * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(bs) * 2};
  uint32_t box[4] = {64, 1, static_cast<uint32_t>(box_rows), 1};
  return make_tmap_bf16(m, base, 4, dims, str, box);
}

template <int D, bool kMasked>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                       cudaStream_t st) {
  using C = AttnCfg<D, kMasked>;
  static bool configured = false;
  auto kern = attention_fwd_kernel<D, kMasked>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal) != cudaSuccess) return kErrCuda;
    configured = true;
  }
  const int grid = p.total < num_sms() ? p.total : num_sms();   // persistent: each CTA walks work items round-robin
  kern<<<grid, C::kThreads, C::kTotal, st>>>(tq, tk, tv, p);
  return launch_status();
}

// d = 64, no bias / key lengths, Lk % 128 == 0, plain output: the cross-block pipelined kernel (attention64p.cuh), OPT-IN with
// LTXB200_ATTN64P=1.  Parity-green, and measured 2-3 % behind attention_fwd_kernel<64, false> on the LTX shape (same box, every
// polynomial setting but one; profiles/r02_attn64_ab.md), so the three-buffer kernel stays the default.
static int launch_attn64p(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, cudaStream_t st) {
  using C = Attn64PCfg;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention64p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal) != cudaSuccess) return kErrCuda;
    configured = true;
  }
  const int grid = p.total < num_sms() ? p.total : num_sms();
  attention64p_kernel<<<grid, C::kThreads, C::kTotal, st>>>(tq, tk, tv, p);
  return launch_status();
}
// d = 128 on a CTA pair (attention128p2.cuh), OPT-IN with LTXB200_ATTN128_2CTA=1: cluster of 2, grid = 2 x min(items of 512 query rows, SMs / 2)
template <bool kMasked>
static int launch_attn128p2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, cudaStream_t st) {
  using C = Attn2CtaCfg;
  static bool configured = false;
  auto kern = attention128p2_kernel<kMasked>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal) != cudaSuccess) return kErrCuda;
    configured = true;
  }
  const long long items = static_cast<long long>(p.B) * p.H * ((p.Lq + C::kRowsPerItem - 1) / C::kRowsPerItem);
  const int pairs = num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * static_cast<unsigned>(items < pairs ? items : pairs));
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, p) != cudaSuccess) return kErrCuda;
  return launch_status();
}
static bool attn128p2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LTXB200_ATTN128_2CTA");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}
static bool attn64p_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LTXB200_ATTN64P");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

static int attention_impl(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                          const void* v, int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso, int B, int H,
                          int Lq, int Lk, int d, float scale, const float* key_bias, const PeerPtrs* peers,
                          int tokens_per_peer, int head_offset, int accumulate, void* stream, const int* key_lens = nullptr) {
  if (B <= 0 || H <= 0 || Lq <= 0 || Lk <= 0 || B > 65535 || H > 65535) return kErrBadShape;
  if (d != 64 && d != 128) return kErrUnsupported;
  if (!peers && !out) return kErrBadAlign;
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || (out && !aligned16(out)) || (ldq & 7) || (ldk & 7) || (ldv & 7) ||
      (ldo & 7) || (bsq & 7) || (bsk & 7) || (bsv & 7) || (bso & 7))
    return kErrBadAlign;
  // keys per block (K / V box rows) depend on the kernel: the unmasked d = 64 kernel walks 96-key blocks (AttnCfg::kOnes)
  const int BNu = d == 64 ? AttnCfg<64, false>::BN : AttnCfg<128, false>::BN, BNm = d == 64 ? AttnCfg<64, true>::BN : AttnCfg<128, true>::BN;
  const bool use64p = d == 64 && !key_bias && !key_lens && Lk % Attn64PCfg::BN == 0 && Lk >= 2 * Attn64PCfg::BN && !peers && !accumulate && attn64p_enabled();
  const bool masked = (key_bias != nullptr) || (Lk % BNu != 0) || (key_lens != nullptr);
  const int BN = use64p ? Attn64PCfg::BN : (masked ? BNm : BNu);
  const bool use128p2 = d == 128 && BN == Attn2CtaCfg::BN && attn128p2_enabled();
  CUtensorMap tq, tk, tv;
  if (make_qkv_tmap(&tq, q, B, H, Lq, d, ldq, bsq, kAttnBM) || make_qkv_tmap(&tk, k, B, H, Lk, d, ldk, bsk, use128p2 ? 32 : BN) ||
      make_qkv_tmap(&tv, v, B, H, Lk, d, ldv, bsv, BN))
    return kErrTensorMap;
  AttnParams p{};
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk;
  const float sc = scale > 0.f ? scale : 1.0f / sqrtf(static_cast<float>(d));
  p.scale_log2 = sc * 1.4426950408889634f;
  p.key_bias = key_bias;
  p.key_lens = key_lens;
  p.out = static_cast<__nv_bfloat16*>(out); p.out_ld = ldo; p.out_bs = bso;
  p.pairs = (Lq + 2 * kAttnBM - 1) / (2 * kAttnBM);
  const long long total = static_cast<long long>(B) * H * p.pairs;
  if (total > 0x7fffffffLL) return kErrBadShape;
  p.total = static_cast<int>(total);
  if (peers) { p.peers = *peers; p.tokens_per_peer = tokens_per_peer; p.head_offset = head_offset; }
  p.accumulate = accumulate;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use64p) return launch_attn64p(tq, tk, tv, p, st);
  if (use128p2) return masked ? launch_attn128p2<true>(tq, tk, tv, p, st) : launch_attn128p2<false>(tq, tk, tv, p, st);
  if (d == 64) return masked ? launch_attn<64, true>(tq, tk, tv, p, st) : launch_attn<64, false>(tq, tk, tv, p, st);
  return masked ? launch_attn<128, true>(tq, tk, tv, p, st) : launch_attn<128, false>(tq, tk, tv, p, st);
}

extern "C" int ltxb200_attention_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                                      const void* v, int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso,
                                      int B, int H, int Lq, int Lk, int d, float scale, const float* key_bias,
                                      void* stream) {
  return attention_impl(q, ldq, bsq, k, ldk, bsk, v, ldv, bsv, out, ldo, bso, B, H, Lq, Lk, d, scale, key_bias, nullptr, 0, 0, 0, stream);
}

extern "C" int ltxb200_attention_klens_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                                            const void* v, int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso,
                                            int B, int H, int Lq, int Lk, int d, float scale, const int* key_lens, void* stream) {
  if (!key_lens) return kErrBadAlign;
  return attention_impl(q, ldq, bsq, k, ldk, bsk, v, ldv, bsv, out, ldo, bso, B, H, Lq, Lk, d, scale, nullptr, nullptr, 0, 0, 0, stream, key_lens);
}

extern "C" int ltxb200_attention_acc_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                                          const void* v, int64_t ldv, int64_t bsv, void* out, int64_t ldo, int64_t bso,
                                          int B, int H, int Lq, int Lk, int d, float scale, const float* key_bias,
                                          void* stream) {
  return attention_impl(q, ldq, bsq, k, ldk, bsk, v, ldv, bsv, out, ldo, bso, B, H, Lq, Lk, d, scale, key_bias, nullptr, 0, 0, 1, stream);
}

// ------------------------------------------------------------------------------------------
// peer-memory exchange (Ulysses sequence parallelism over NVLink P2P)
// ------------------------------------------------------------------------------------------
static int fill_peers(PeerPtrs* pp, int P, int rank, void* const* data_ptrs, void* const* flag_ptrs, unsigned int epoch,
                      void* counter) {
  if (P < 1 || P > kMaxPeers || rank < 0 || rank >= P || !data_ptrs || !flag_ptrs || !counter) return kErrBadShape;
  *pp = PeerPtrs{};
  for (int r = 0; r < P; ++r) {
    if (!data_ptrs[r] || !flag_ptrs[r] || !aligned16(data_ptrs[r])) return kErrBadAlign;
    pp->data[r] = data_ptrs[r];
    pp->flags[r] = static_cast<unsigned int*>(flag_ptrs[r]);
  }
  pp->P = P; pp->rank = rank; pp->epoch = epoch; pp->counter = static_cast<unsigned int*>(counter);
  return kOk;
}

extern "C" int ltxb200_comm_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_64B) {
  if (!dev_ptr || !ipc_handle_64B || bytes == 0) return kErrBadShape;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return kErrCuda;
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return kErrCuda; }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return kErrCuda; }
  memcpy(ipc_handle_64B, &h, sizeof(h));
  *dev_ptr = p;
  return kOk;
}
extern "C" int ltxb200_comm_open(const void* ipc_handle_64B, void** dev_ptr) {
  if (!dev_ptr || !ipc_handle_64B) return kErrBadShape;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64B, sizeof(h));
  return cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? kOk : kErrCuda;
}
extern "C" int ltxb200_comm_close(void* dev_ptr) { return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? kOk : kErrCuda; }
extern "C" int ltxb200_comm_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? kOk : kErrCuda; }

extern "C" int ltxb200_comm_wait_status(const void* flags, int P, unsigned int epoch, void* status_dev, void* status_host,
                                        unsigned int timeout_ms, void* stream) {
  if (!flags || P < 1 || P > kMaxPeers) return kErrBadShape;
  static const unsigned int env_ms = getenv("LTXB200_COMM_TIMEOUT_MS") ? static_cast<unsigned int>(atoi(getenv("LTXB200_COMM_TIMEOUT_MS"))) : 20000u;
  const unsigned long long ns = static_cast<unsigned long long>(timeout_ms ? timeout_ms : env_ms) * 1000000ull;
  comm_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned int*>(flags), P, epoch,
                                                                    static_cast<unsigned int*>(status_dev),
                                                                    static_cast<unsigned int*>(status_host), ns);
  return launch_status();
}

extern "C" int ltxb200_comm_wait(const void* flags, int P, unsigned int epoch, void* stream) {
  return ltxb200_comm_wait_status(flags, P, epoch, nullptr, nullptr, 0, stream);
}

extern "C" int ltxb200_peer_allgather(const void* src, int64_t seg_bytes, int nseg, int P, int rank, void* const* dst_ptrs,
                                      void* const* flag_ptrs, unsigned int epoch, void* counter, void* stream) {
  if (!src || seg_bytes <= 0 || (seg_bytes & 15) || nseg <= 0) return kErrBadShape;
  if (!aligned16(src)) return kErrBadAlign;
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, P, rank, dst_ptrs, flag_ptrs, epoch, counter)) return rc;
  const long long vec = seg_bytes / 16;
  peer_allgather_kernel<<<ew_blocks(vec * nseg, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(src), vec, nseg, pp);
  return launch_status();
}

// grid.x of one scatter launch over `rows` rows (grid-stride: at most 2 CTAs of 4 warps per SM and selector), and the CTAs it
// contributes to the arrival counter (x 3 selectors; x 2 when V leaves from the GEMM epilogue)
static inline unsigned int scatter_blocks(int rows) {
  const unsigned int want = static_cast<unsigned int>((rows + 3) / 4), cap = static_cast<unsigned int>(num_sms()) * 2u;
  return want < cap ? want : cap;
}
static inline unsigned int scatter_ctas(int rows) { return scatter_blocks(rows) * 3u; }

extern "C" int ltxb200_qk_norm_rope_wan_scatter_rows_bf16(const void* qkv, int64_t ld, int M, int row0, int rows, int D, const void* wq,
                                                          const void* wk, const float* cos_table, const float* sin_table,
                                                          int head_dim, int tokens_per_batch, int token_offset, float eps, int B,
                                                          int P, int rank, void* const* recv_ptrs, void* const* flag_ptrs,
                                                          unsigned int epoch, void* counter, unsigned int signal_ctas, int nsel, void* stream) {
  if (M <= 0 || D <= 0 || (D % 256) || B <= 0 || tokens_per_batch <= 0 || M != B * tokens_per_batch) return kErrBadShape;
  if ((nsel != 2 && nsel != 3) || row0 < 0 || rows <= 0 || row0 + rows > M || signal_ctas < scatter_blocks(rows) * nsel) return kErrBadShape;
  if (!qkv || !aligned16(qkv) || (ld & 7) || !wq || !wk || !cos_table || !sin_table || !aligned16(cos_table) || !aligned16(sin_table))
    return kErrBadAlign;
  if (head_dim <= 0 || (head_dim & 7) || (D % head_dim) || ((D / head_dim) % P)) return kErrBadShape;
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, P, rank, recv_ptrs, flag_ptrs, epoch, counter)) return rc;
  const int Hp = D / head_dim / P;
  dim3 grid(scatter_blocks(rows), nsel);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const __nv_bfloat16*>(qkv);
  auto WQ = static_cast<const __nv_bfloat16*>(wq);
  auto WK = static_cast<const __nv_bfloat16*>(wk);
#define QKS_CASE(n) \
  case n: qk_norm_rope_wan_scatter_kernel<n><<<grid, 128, 0, st>>>(X, ld, row0, row0 + rows, WQ, WK, cos_table, sin_table, head_dim, tokens_per_batch, token_offset, eps, B, Hp, pp, signal_ctas); break;
  switch (D / 256) {
    QKS_CASE(1) QKS_CASE(2) QKS_CASE(4) QKS_CASE(6) QKS_CASE(8) QKS_CASE(12) QKS_CASE(16) QKS_CASE(20)
    default: return kErrUnsupported;
  }
#undef QKS_CASE
  return launch_status();
}

extern "C" unsigned int ltxb200_scatter_signal_ctas(int rows) { return rows > 0 ? scatter_ctas(rows) : 0u; }   // per 3 selectors

extern "C" int ltxb200_qk_norm_rope_wan_scatter_bf16(const void* qkv, int64_t ld, int M, int D, const void* wq, const void* wk,
                                                     const float* cos_table, const float* sin_table, int head_dim,
                                                     int tokens_per_batch, int token_offset, float eps, int B, int P,
                                                     int rank, void* const* recv_ptrs, void* const* flag_ptrs,
                                                     unsigned int epoch, void* counter, void* stream) {
  return ltxb200_qk_norm_rope_wan_scatter_rows_bf16(qkv, ld, M, 0, M, D, wq, wk, cos_table, sin_table, head_dim, tokens_per_batch,
                                                    token_offset, eps, B, P, rank, recv_ptrs, flag_ptrs, epoch, counter,
                                                    M > 0 ? scatter_ctas(M) : 0u, 3, stream);
}

extern "C" int ltxb200_attention_scatter_bf16(const void* q, int64_t ldq, int64_t bsq, const void* k, int64_t ldk, int64_t bsk,
                                              const void* v, int64_t ldv, int64_t bsv, int64_t ldo, int B, int H, int Lq,
                                              int Lk, int d, float scale, const float* key_bias, int P, int rank,
                                              void* const* out_ptrs, void* const* flag_ptrs, unsigned int epoch,
                                              void* counter, int tokens_per_peer, int head_offset, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_peers(&pp, P, rank, out_ptrs, flag_ptrs, epoch, counter)) return rc;
  if (tokens_per_peer <= 0 || static_cast<long long>(tokens_per_peer) * P < Lq || head_offset < 0) return kErrBadShape;
  return attention_impl(q, ldq, bsq, k, ldk, bsk, v, ldv, bsv, nullptr, ldo, 0, B, H, Lq, Lk, d, scale, key_bias, &pp,
                        tokens_per_peer, head_offset, 0, stream);
}

// ------------------------------------------------------------------------------------------
// memory-bound kernels
// ------------------------------------------------------------------------------------------
template <bool LN>
static int launch_norm(int NV, const __nv_bfloat16* x, __nv_bfloat16* y, int M, int64_t ldx, int64_t ldy,
                       const __nv_bfloat16* sc, const __nv_bfloat16* sh, int64_t mod_ld, int rpg,
                       const __nv_bfloat16* w, const __nv_bfloat16* b, float eps, cudaStream_t st) {
  int grid = (M + 3) / 4;
  const int cap = num_sms() * 4;            // rows are walked grid-stride (4 resident CTAs per SM at ~120 registers)
  if (grid > cap) grid = cap;
#define NORM_CASE(n) \
  case n: norm_mod_kernel<n, LN><<<grid, 128, 0, st>>>(x, y, M, ldx, ldy, sc, sh, mod_ld, rpg, w, b, eps); break;
  switch (NV) {
    NORM_CASE(1) NORM_CASE(2) NORM_CASE(4) NORM_CASE(5) NORM_CASE(6) NORM_CASE(8) NORM_CASE(12) NORM_CASE(16) NORM_CASE(20)
    default: return kErrUnsupported;
  }
#undef NORM_CASE
  return launch_status();
}

extern "C" int ltxb200_norm_mod_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int M, int D, const void* scale,
                                     const void* shift, int64_t mod_ld, int rows_per_group, const void* weight,
                                     const void* bias, float eps, int layer_norm, void* stream) {
  if (M <= 0 || D <= 0) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y) || (ldx & 7) || (ldy & 7) || (scale && (!aligned16(scale) || (mod_ld & 7))) ||
      (shift && !aligned16(shift)) || (weight && !aligned16(weight)) || (bias && !aligned16(bias)))
    return kErrBadAlign;
  if ((scale == nullptr) != (shift == nullptr)) return kErrBadShape;
  if (D == 64 || D == 128) {          // narrow affine LayerNorm (VAE encoder shortcut): contiguous rows, no modulation
    if (!layer_norm || !weight || !bias || scale || ldx != D || ldy != D) return kErrUnsupported;
    const int rpw = 32 / (D / 8);
    const long long warps = (static_cast<long long>(M) + rpw - 1) / rpw, blocks = (warps + 7) / 8;
    cudaStream_t s2 = static_cast<cudaStream_t>(stream);
    if (D == 64)
      layernorm_narrow_kernel<64><<<static_cast<unsigned>(blocks), 256, 0, s2>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), M, static_cast<const __nv_bfloat16*>(weight), static_cast<const __nv_bfloat16*>(bias), eps);
    else
      layernorm_narrow_kernel<128><<<static_cast<unsigned>(blocks), 256, 0, s2>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), M, static_cast<const __nv_bfloat16*>(weight), static_cast<const __nv_bfloat16*>(bias), eps);
    return launch_status();
  }
  if (D % 256) return kErrBadShape;
  const int rpg = rows_per_group > 0 ? rows_per_group : M;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto Y = static_cast<__nv_bfloat16*>(y);
  auto SC = static_cast<const __nv_bfloat16*>(scale);
  auto SH = static_cast<const __nv_bfloat16*>(shift);
  auto Wt = static_cast<const __nv_bfloat16*>(weight);
  auto Bs = static_cast<const __nv_bfloat16*>(bias);
  return layer_norm ? launch_norm<true>(D / 256, X, Y, M, ldx, ldy, SC, SH, mod_ld, rpg, Wt, Bs, eps, st)
                    : launch_norm<false>(D / 256, X, Y, M, ldx, ldy, SC, SH, mod_ld, rpg, Wt, Bs, eps, st);
}

extern "C" int ltxb200_norm_mod_f32in(const float* x, int64_t ldx, void* y, int64_t ldy, int M, int D, const float* scale,
                                      const float* shift, int64_t mod_ld, int rows_per_group, float eps, int layer_norm,
                                      void* stream) {
  if (M <= 0 || D <= 0 || (D % 256)) return kErrBadShape;
  if (!x || !y || !aligned16(x) || !aligned16(y) || (ldx & 3) || (ldy & 7) || (scale && (!aligned16(scale) || (mod_ld & 3))) ||
      (shift && !aligned16(shift)))
    return kErrBadAlign;
  if ((scale == nullptr) != (shift == nullptr)) return kErrBadShape;
  const int rpg = rows_per_group > 0 ? rows_per_group : M;
  int grid = (M + 3) / 4;
  const int cap = num_sms() * 4;
  if (grid > cap) grid = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto Y = static_cast<__nv_bfloat16*>(y);
#define NMF_CASE(n) \
  case n: if (layer_norm) norm_mod_f32in_kernel<n, true><<<grid, 128, 0, st>>>(x, Y, M, ldx, ldy, scale, shift, mod_ld, rpg, eps); \
          else norm_mod_f32in_kernel<n, false><<<grid, 128, 0, st>>>(x, Y, M, ldx, ldy, scale, shift, mod_ld, rpg, eps); break;
  switch (D / 256) {
    NMF_CASE(2) NMF_CASE(4) NMF_CASE(8) NMF_CASE(16)
    default: return kErrUnsupported;
  }
#undef NMF_CASE
  return launch_status();
}

extern "C" int ltxb200_ada_add_f32(const void* table, const void* temb, float* out, int L, int G, int JD, void* stream) {
  if (L <= 0 || G <= 0 || JD <= 0 || (JD & 7)) return kErrBadShape;
  if (!aligned16(table) || !aligned16(temb) || !aligned16(out)) return kErrBadAlign;
  const long long n8 = static_cast<long long>(L) * G * JD / 8;
  ada_add_f32_kernel<<<ew_blocks(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(temb), out, L, G, JD);
  return launch_status();
}

extern "C" int ltxb200_qk_norm_rope_bf16(void* q, int64_t ldq, int Mq, void* k, int64_t ldk, int Mk, int D,
                                         const void* wq, const void* wk, const void* cos_table,
                                         const void* sin_table, int tokens_per_batch, float eps, void* stream) {
  if (D <= 0 || (D % 256) || (q == nullptr && k == nullptr)) return kErrBadShape;
  if ((q && (!aligned16(q) || (ldq & 7) || !wq)) || (k && (!aligned16(k) || (ldk & 7) || !wk))) return kErrBadAlign;
  if ((cos_table == nullptr) != (sin_table == nullptr)) return kErrBadShape;
  if (cos_table && tokens_per_batch <= 0) return kErrBadShape;
  const int M = Mq > Mk ? Mq : Mk;
  dim3 grid((M + 3) / 4, k ? 2 : 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto Q = static_cast<__nv_bfloat16*>(q);
  auto Kp = static_cast<__nv_bfloat16*>(k);
  auto WQ = static_cast<const __nv_bfloat16*>(wq);
  auto WK = static_cast<const __nv_bfloat16*>(wk);
  auto C = static_cast<const __nv_bfloat16*>(cos_table);
  auto Sn = static_cast<const __nv_bfloat16*>(sin_table);
  // self-attention shape (q and k rows of the same B*tokens grid, RoPE): one warp per TOKEN, the table row read once for all 2*B rows
  static const int env_tok = getenv("LTXB200_ROPE_TOKEN_MAJOR") ? atoi(getenv("LTXB200_ROPE_TOKEN_MAJOR")) : 1;
  if (env_tok && C && q && k && Mq == Mk && (Mq % tokens_per_batch) == 0 && D == 4096) {          // LTX-Video 13B: a warp pair per token
    const int Bq = Mq / tokens_per_batch;
    qk_norm_rope_tok2_kernel<8><<<dim3((tokens_per_batch + 1) / 2), 128, 0, st>>>(Q, Kp, Bq, tokens_per_batch, ldq, ldk, WQ, WK, C, Sn, eps);
    return launch_status();
  }
  if (env_tok && C && q && k && Mq == Mk && (Mq % tokens_per_batch) == 0 && D <= 2048) {
    const int Bq = Mq / tokens_per_batch;
    dim3 gt((tokens_per_batch + 3) / 4);
#define QKT_CASE(n) \
  case n: qk_norm_rope_tok_kernel<n><<<gt, 128, 0, st>>>(Q, Kp, Bq, tokens_per_batch, ldq, ldk, WQ, WK, C, Sn, eps); break;
    switch (D / 256) {
      QKT_CASE(2) QKT_CASE(4) QKT_CASE(6) QKT_CASE(8)
      default: return kErrUnsupported;
    }
#undef QKT_CASE
    return launch_status();
  }
#define QK_CASE(n) \
  case n: if (C) qk_norm_rope_kernel<n, true><<<grid, 128, 0, st>>>(Q, Kp, q ? Mq : 0, Mk, ldq, ldk, WQ, WK, C, Sn, tokens_per_batch, eps); \
          else qk_norm_rope_kernel<n, false><<<grid, 128, 0, st>>>(Q, Kp, q ? Mq : 0, Mk, ldq, ldk, WQ, WK, C, Sn, tokens_per_batch, eps); break;
  switch (D / 256) {
    QK_CASE(2) QK_CASE(4) QK_CASE(6) QK_CASE(8) QK_CASE(12) QK_CASE(16) QK_CASE(20)
    default: return kErrUnsupported;
  }
#undef QK_CASE
  return launch_status();
}

extern "C" int ltxb200_qk_norm_rope_wan_bf16(void* q, int64_t ldq, int Mq, void* k, int64_t ldk, int Mk, int D,
                                             const void* wq, const void* wk, const float* cos_table,
                                             const float* sin_table, int head_dim, int tokens_per_batch,
                                             int token_offset, float eps, void* stream) {
  if (D <= 0 || (D % 256) || (q == nullptr && k == nullptr)) return kErrBadShape;
  if ((q && (!aligned16(q) || (ldq & 7) || !wq)) || (k && (!aligned16(k) || (ldk & 7) || !wk))) return kErrBadAlign;
  if ((cos_table == nullptr) != (sin_table == nullptr)) return kErrBadShape;
  if (cos_table && (tokens_per_batch <= 0 || head_dim <= 0 || (head_dim & 7) || (D % head_dim) || !aligned16(cos_table) || !aligned16(sin_table)))
    return kErrBadShape;
  const int M = Mq > Mk ? Mq : Mk;
  dim3 grid((M + 3) / 4, k ? 2 : 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto Q = static_cast<__nv_bfloat16*>(q);
  auto Kp = static_cast<__nv_bfloat16*>(k);
  auto WQ = static_cast<const __nv_bfloat16*>(wq);
  auto WK = static_cast<const __nv_bfloat16*>(wk);
#define QKW_CASE(n) \
  case n: qk_norm_rope_wan_kernel<n><<<grid, 128, 0, st>>>(Q, Kp, q ? Mq : 0, Mk, ldq, ldk, WQ, WK, cos_table, sin_table, head_dim, tokens_per_batch, token_offset, eps); break;
  switch (D / 256) {
    QKW_CASE(1) QKW_CASE(2) QKW_CASE(4) QKW_CASE(6) QKW_CASE(8) QKW_CASE(12) QKW_CASE(16) QKW_CASE(20)
    default: return kErrUnsupported;
  }
#undef QKW_CASE
  return launch_status();
}

extern "C" int ltxb200_lincomb_f32(float* out, int64_t n, int terms, const float* const* xs, const float* coefs,
                                   void* stream) {
  if (n <= 0 || (n & 3) || terms <= 0 || terms > 6 || !xs || !coefs) return kErrBadShape;
  if (!aligned16(out)) return kErrBadAlign;
  LinCombParams p{};
  p.terms = terms;
  for (int j = 0; j < terms; ++j) {
    if (!aligned16(xs[j])) return kErrBadAlign;
    p.x[j] = xs[j];
    p.c[j] = coefs[j];
  }
  lincomb_f32_kernel<<<ew_blocks(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, p);
  return launch_status();
}

extern "C" int ltxb200_rf_step_tokens_f32(float* out, const float* x, const float* v, const float* noise, const float* tok_timesteps,
                                          int64_t tokens, int channels, const float* schedule, int num_steps, void* stream) {
  if (tokens <= 0 || channels <= 0 || (channels & 3) || num_steps < 0 || !out || !x || !v || !tok_timesteps || (num_steps > 0 && !schedule))
    return kErrBadShape;
  if (!aligned16(out) || !aligned16(x) || !aligned16(v) || (noise && !aligned16(noise))) return kErrBadAlign;
  rf_step_tokens_kernel<<<ew_blocks(tokens * (channels / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, x, v, noise, tok_timesteps, tokens, channels, schedule, num_steps);
  return launch_status();
}

extern "C" int ltxb200_ada_add_bf16(const void* table, const void* temb, void* out, int L, int G, int JD, void* stream) {
  if (L <= 0 || G <= 0 || JD <= 0 || (JD & 7)) return kErrBadShape;
  if (!aligned16(table) || !aligned16(temb) || !aligned16(out)) return kErrBadAlign;
  const long long n8 = static_cast<long long>(L) * G * JD / 8;
  ada_add_kernel<<<ew_blocks(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(temb), static_cast<__nv_bfloat16*>(out), L, G, JD);
  return launch_status();
}

extern "C" int ltxb200_act_bf16(const void* x, void* y, int64_t n, int mode, void* stream) {
  if (n <= 0 || (n & 7)) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y)) return kErrBadAlign;
  act_kernel<<<ew_blocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n, mode);
  return launch_status();
}

extern "C" int ltxb200_stg_blend_bf16(void* a, const void* v, int64_t ldv, const float* mask, int B, int64_t rows,
                                      int D, void* stream) {
  if (B <= 0 || rows <= 0 || D <= 0 || (D & 7)) return kErrBadShape;
  if (!aligned16(a) || !aligned16(v) || (ldv & 7)) return kErrBadAlign;
  // v rows are addressed as (b*rows + r) * ldv: the value slice of the fused QKV buffer
  stg_blend_kernel<<<ew_blocks(B * rows * D / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(v), ldv, mask, B, rows, D);
  return launch_status();
}

extern "C" int ltxb200_axpby_bf16(const void* x, const void* y, void* out, int64_t n, float a, float b, void* stream) {
  if (n <= 0 || (n & 7)) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y) || !aligned16(out)) return kErrBadAlign;
  axpby_kernel<<<ew_blocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(out), n, a, b);
  return launch_status();
}

extern "C" int ltxb200_rel_l1_bf16(const void* a, const void* b, int64_t n, float* out2, void* stream) {
  if (n <= 0) return kErrBadShape;
  if (!a || !b || !out2) return kErrBadAlign;
  rel_l1_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(a),
                                                                 static_cast<const __nv_bfloat16*>(b), n, out2);
  return launch_status();
}

extern "C" int ltxb200_groupnorm_silu_bf16(const void* x, void* y, int B, int64_t voxels, int C, const void* gamma, const void* beta,
                                           const void* residual, float eps, int apply_silu, float* scratch, void* stream) {
  if (B <= 0 || voxels <= 0 || C <= 0 || (C % 256) || C > 2048) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || (residual && !aligned16(residual)) || !scratch)
    return kErrBadAlign;
  const int vpb = 256 / (C / 8);
  long long want = (voxels + vpb - 1) / vpb;
  const int chunks = static_cast<int>(want < LTXB200_GROUPNORM_CHUNKS ? want : LTXB200_GROUPNORM_CHUNKS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  groupnorm_stats_kernel<<<dim3(chunks, B), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), scratch, voxels, C);
  const long long cap = num_sms() * 8ll / B;
  const int blocks = static_cast<int>(want < 1 ? 1 : (want > cap ? (cap < 1 ? 1 : cap) : want));
  groupnorm_apply_kernel<<<dim3(blocks, B), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), scratch, chunks, voxels, C,
      static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(beta),
      static_cast<const __nv_bfloat16*>(residual), eps, apply_silu);
  return launch_status();
}

extern "C" int ltxb200_adain_f32(const float* x, const float* ref, float* out, int rows, int64_t n, int64_t m, float factor,
                                 void* stream) {
  if (rows <= 0 || n < 2 || m < 2) return kErrBadShape;
  if (!x || !ref || !out) return kErrBadAlign;
  adain_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ref, out, n, m, factor);
  return launch_status();
}

extern "C" int ltxb200_latent_from_ndhwc(const void* x, float* out, int B, int C, int64_t FHW, const float* stdv, const float* meanv,
                                         void* stream) {
  if (B <= 0 || C <= 0 || FHW <= 0) return kErrBadShape;
  if (!x || !out || ((stdv == nullptr) != (meanv == nullptr))) return kErrBadAlign;
  const long long n = static_cast<long long>(B) * C * FHW;
  latent_from_ndhwc_kernel<<<ew_blocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), out, B, C, FHW, stdv, meanv);
  return launch_status();
}

extern "C" int ltxb200_bilinear_resize_f32(const float* x, float* y, int64_t planes, int h, int w, int H, int W, void* stream) {
  if (planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return kErrBadShape;
  if (!x || !y) return kErrBadAlign;
  bilinear_resize_kernel<<<ew_blocks(planes * H * W, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, planes, h, w, H, W);
  return launch_status();
}

extern "C" int ltxb200_timestep_embed(const float* t, void* out, int n, int dim, void* stream) {
  if (n <= 0 || dim <= 0 || (dim & 1)) return kErrBadShape;
  timestep_embed_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, static_cast<__nv_bfloat16*>(out), n, dim, 1);
  return launch_status();
}

extern "C" int ltxb200_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
  if (n <= 0 || (n & 7)) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y)) return kErrBadAlign;
  cast_f32_to_bf16_kernel<<<ew_blocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y), n);
  return launch_status();
}

static int guidance_step_impl(const void* pred, int64_t cond_stride, int64_t n, int channels, int has_cfg,
                              int has_stg, int do_rescale, float guidance_scale, float stg_scale,
                              float rescale, float* latents, void* latents_bf16, const float* timesteps,
                              int num_steps, float t, const float* cond_mask, float* scratch, const float* noise, void* stream) {
  if (n <= 0 || channels <= 0 || num_steps <= 0 || !latents || !pred || !timesteps) return kErrBadShape;
  if ((has_cfg || has_stg) && !scratch) return kErrBadShape;
  GuidanceParams g{};
  g.pred = static_cast<const __nv_bfloat16*>(pred);
  g.cond_stride = cond_stride; g.n = n; g.channels = channels;
  g.has_cfg = has_cfg ? 1 : 0; g.has_stg = has_stg ? 1 : 0; g.do_rescale = do_rescale ? 1 : 0;
  g.guidance_scale = guidance_scale; g.stg_scale = stg_scale; g.rescale = rescale;
  g.partials = scratch;
  g.latents = latents; g.latents_bf16 = static_cast<__nv_bfloat16*>(latents_bf16);
  g.timesteps = timesteps; g.num_steps = num_steps; g.t = t; g.cond_mask = cond_mask; g.noise = noise;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = kOk;
  if (g.has_cfg && guidance_scale != 0.f && guidance_scale != 1.f) {
    guidance_reduce_kernel<0><<<kGuidanceBlocks, 256, 0, st>>>(g);
    rc = launch_status();
    if (rc) return rc;
  } else {
    g.has_cfg = g.has_cfg;   // alpha unused when scale is 0/1 (combine falls back to text)
  }
  if (g.has_stg && g.do_rescale && stg_scale > 0.f) {
    guidance_reduce_kernel<1><<<kGuidanceBlocks, 256, 0, st>>>(g);
    rc = launch_status();
    if (rc) return rc;
  }
  guidance_step_kernel<<<kGuidanceBlocks, 256, 0, st>>>(g);
  return launch_status();
}

extern "C" int ltxb200_guidance_step(const void* pred, int64_t cond_stride, int64_t n, int channels, int has_cfg,
                                     int has_stg, int do_rescale, float guidance_scale, float stg_scale,
                                     float rescale, float* latents, void* latents_bf16, const float* timesteps,
                                     int num_steps, float t, const float* cond_mask, float* scratch, void* stream) {
  return guidance_step_impl(pred, cond_stride, n, channels, has_cfg, has_stg, do_rescale, guidance_scale, stg_scale, rescale,
                            latents, latents_bf16, timesteps, num_steps, t, cond_mask, scratch, nullptr, stream);
}

extern "C" int ltxb200_guidance_step_stochastic(const void* pred, int64_t cond_stride, int64_t n, int channels, int has_cfg,
                                                int has_stg, int do_rescale, float guidance_scale, float stg_scale,
                                                float rescale, float* latents, void* latents_bf16, const float* timesteps,
                                                int num_steps, float t, const float* cond_mask, float* scratch,
                                                const float* noise, void* stream) {
  if (!noise) return kErrBadAlign;
  return guidance_step_impl(pred, cond_stride, n, channels, has_cfg, has_stg, do_rescale, guidance_scale, stg_scale, rescale,
                            latents, latents_bf16, timesteps, num_steps, t, cond_mask, scratch, noise, stream);
}

extern "C" int ltxb200_cfg_combine_f32(const float* cond, const float* uncond, float* out, int64_t n, float guide_scale,
                                       int use_alpha, float* scratch, void* stream) {
  if (n <= 0 || !cond || !uncond || !out) return kErrBadShape;
  if (use_alpha && !scratch) return kErrBadShape;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_alpha) {
    cfg_combine_f32_kernel<0><<<kGuidanceBlocks, 256, 0, st>>>(cond, uncond, out, n, guide_scale, 1, scratch);
    const int rc = launch_status();
    if (rc) return rc;
  }
  cfg_combine_f32_kernel<1><<<kGuidanceBlocks, 256, 0, st>>>(cond, uncond, out, n, guide_scale, use_alpha ? 1 : 0, scratch);
  return launch_status();
}

static int pixelnorm_impl(const void* x, void* y, int64_t voxels, int C, float eps, int apply_silu, const void* scale,
                          const void* shift, void* stream) {
  if (voxels <= 0) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y) || ((scale == nullptr) != (shift == nullptr)) || (scale && (!aligned16(scale) || !aligned16(shift))))
    return kErrBadAlign;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto Y = static_cast<__nv_bfloat16*>(y);
  auto SC = static_cast<const __nv_bfloat16*>(scale);
  auto SH = static_cast<const __nv_bfloat16*>(shift);
#define PN_CASE(c)                                                                             \
  case c: {                                                                                    \
    constexpr int G = (c / 8 < 32) ? c / 8 : 32;                                               \
    const long long warps = (voxels + (32 / G) - 1) / (32 / G);                                \
    const long long blocks = (warps + 7) / 8;                                                  \
    pixelnorm_silu_kernel<c><<<static_cast<unsigned>(blocks), 256, 0, st>>>(X, Y, voxels, eps, apply_silu, SC, SH); \
  } break;
  switch (C) {
    PN_CASE(64) PN_CASE(128) PN_CASE(256) PN_CASE(512) PN_CASE(1024)
    default: return kErrUnsupported;
  }
#undef PN_CASE
  return launch_status();
}

extern "C" int ltxb200_pixelnorm_silu_bf16(const void* x, void* y, int64_t voxels, int C, float eps, int apply_silu,
                                           void* stream) {
  return pixelnorm_impl(x, y, voxels, C, eps, apply_silu, nullptr, nullptr, stream);
}

extern "C" int ltxb200_pixelnorm_mod_silu_bf16(const void* x, void* y, int64_t voxels, int C, float eps, const void* scale,
                                               const void* shift, int apply_silu, void* stream) {
  if (!scale || !shift) return kErrBadAlign;
  return pixelnorm_impl(x, y, voxels, C, eps, apply_silu, scale, shift, stream);
}

extern "C" int ltxb200_l2norm_silu_bf16(const void* x, void* y, int64_t voxels, int C, int c_real, const void* gamma,
                                        int apply_silu, void* stream) {
  if (voxels <= 0 || c_real <= 0 || c_real > C || !gamma) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma)) return kErrBadAlign;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto Y = static_cast<__nv_bfloat16*>(y);
  auto G_ = static_cast<const __nv_bfloat16*>(gamma);
  const float scale = sqrtf(static_cast<float>(c_real));
#define L2_CASE(c)                                                                             \
  case c: {                                                                                    \
    constexpr int G = l2_group(c / 8);                                                         \
    const long long warps = (voxels + (32 / G) - 1) / (32 / G);                                \
    const long long blocks = (warps + 7) / 8;                                                  \
    l2norm_silu_kernel<c><<<static_cast<unsigned>(blocks), 256, 0, st>>>(X, Y, voxels, G_, scale, apply_silu); \
  } break;
  switch (C) {
    L2_CASE(64) L2_CASE(128) L2_CASE(192) L2_CASE(256) L2_CASE(384) L2_CASE(512)
    default: return kErrUnsupported;
  }
#undef L2_CASE
  return launch_status();
}

extern "C" int ltxb200_upsample2x_nhwc_bf16(const void* x, void* y, int64_t frames, int H, int W, int C, void* stream) {
  if (frames <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return kErrBadShape;
  if (!aligned16(x) || !aligned16(y)) return kErrBadAlign;
  const long long n = frames * 4LL * H * W * (C / 8);
  upsample2x_nhwc_kernel<<<ew_blocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), frames, H, W, C);
  return launch_status();
}

extern "C" int ltxb200_softmax_rows_f32_bf16(const float* s, int64_t ld_s, void* p, int64_t ld_p, int rows, int cols,
                                             float scale, void* stream) {
  if (rows <= 0 || cols <= 0 || !s || !p) return kErrBadShape;
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, static_cast<__nv_bfloat16*>(p), rows, cols, ld_s, ld_p, scale * 1.4426950408889634f);
  return launch_status();
}

extern "C" int ltxb200_latent_to_ndhwc(const void* z, int is_f32, void* out, int B, int C, int64_t FHW,
                                       const float* stdv, const float* meanv, void* stream) {
  if (B <= 0 || C <= 0 || FHW <= 0) return kErrBadShape;
  if ((stdv == nullptr) != (meanv == nullptr)) return kErrBadShape;
  const long long n = static_cast<long long>(B) * C * FHW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    latent_to_ndhwc_kernel<float><<<ew_blocks(n, 256), 256, 0, st>>>(static_cast<const float*>(z), static_cast<__nv_bfloat16*>(out), B, C, FHW, stdv, meanv);
  else
    latent_to_ndhwc_kernel<__nv_bfloat16><<<ew_blocks(n, 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(z), static_cast<__nv_bfloat16*>(out), B, C, FHW, stdv, meanv);
  return launch_status();
}
