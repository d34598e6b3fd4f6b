// Microbenchmark: exp2 throughput per SM for the softmax inner loop (B200).  Standalone: nvcc -arch=sm_100a.
//   f32 MUFU.EX2, f16x2 / bf16x2 MUFU.EX2, FMA-pipe polynomial exp2, and mixes of MUFU + polynomial.
#include "../../ltx-video-gpupoor_b200/csrc/common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// Cody-Waite + degree-3 polynomial on the FMA pipe; valid for x in [-126, 126]
__device__ __forceinline__ float ex2poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;           // 1.5 * 2^23: round-to-nearest integer lands in the low mantissa bits
  const float r = t - 12582912.f;
  const float f = x - r;                     // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041087f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed, long long* cyc) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = 0xb800b800u + i;   // small negative halves
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {            // 16 x f32 MUFU
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2f(a[i]) - 1.5f;
    } else if (MODE == 1) {     // 16 x f16x2 MUFU = 32 exps
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = ex2h2(h[i]) ^ 0x80008000u;
    } else if (MODE == 2) {     // 16 x bf16x2 MUFU = 32 exps
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = ex2b2(h[i]) ^ 0x80008000u;
    } else if (MODE == 3) {     // 16 x polynomial
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2poly(a[i]) - 1.5f;
    } else if (MODE == 4) {     // 12 MUFU + 4 poly
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ((i & 3) == 3 ? ex2poly(a[i]) : ex2f(a[i])) - 1.5f;
    } else if (MODE == 5) {     // 8 MUFU + 8 poly
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ((i & 1) ? ex2poly(a[i]) : ex2f(a[i])) - 1.5f;
    } else if (MODE == 6) {     // 10 MUFU + 6 poly
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ((i % 8) >= 5 ? ex2poly(a[i]) : ex2f(a[i])) - 1.5f;
    } else if (MODE == 7) {     // softmax-like: fma + MUFU + add + pack, all f32 MUFU
      float s = 0.f; uint32_t pk = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float e0 = ex2f(fmaf(a[i], 0.18f, -seed)), e1 = ex2f(fmaf(a[i + 1], 0.18f, -seed));
        s += e0 + e1;
        __nv_bfloat162 v = __floats2bfloat162_rn(e0, e1);
        pk ^= *reinterpret_cast<uint32_t*>(&v);
        a[i] = e0 - 1.5f; a[i + 1] = e1 - 1.5f;
      }
      h[0] ^= pk; a[0] += s;
    } else if (MODE == 8) {     // softmax-like with 25% polynomial
      float s = 0.f; uint32_t pk = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float x0 = fmaf(a[i], 0.18f, -seed), x1 = fmaf(a[i + 1], 0.18f, -seed);
        float e0 = ex2f(x0), e1 = ((i & 2) ? ex2poly(x1) : ex2f(x1));
        s += e0 + e1;
        __nv_bfloat162 v = __floats2bfloat162_rn(e0, e1);
        pk ^= *reinterpret_cast<uint32_t*>(&v);
        a[i] = e0 - 1.5f; a[i + 1] = e1 - 1.5f;
      }
      h[0] ^= pk; a[0] += s;
    } else if (MODE == 9) {     // softmax-like via f16x2: fma f32, cvt pack f16x2, MUFU f16x2 (output is packed P)
      uint32_t pk = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float x0 = fmaf(a[i], 0.18f, -seed), x1 = fmaf(a[i + 1], 0.18f, -seed);
        __half2 hx = __floats2half2_rn(x0, x1);
        uint32_t e = ex2h2(*reinterpret_cast<uint32_t*>(&hx));
        pk ^= e;
        a[i] = x0 * 0.5f; a[i + 1] = x1 * 0.5f;
      }
      h[0] ^= pk;
    }
  }
  const long long t1 = clock64();
  float s = 0.f; uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += a[i]; x ^= h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(x & 0xff);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int exps_per_iter, int warps_per_smsp) {
  const int blocks = 148 * warps_per_smsp, iters = 4096;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, blocks * 128 * 4)); CK(cudaMalloc(&cyc, blocks * 8));
  k<MODE><<<blocks, 128>>>(out, 16, 0.37f, cyc);
  CK(cudaDeviceSynchronize());
  k<MODE><<<blocks, 128>>>(out, iters, 0.37f, cyc);
  CK(cudaDeviceSynchronize());
  long long h[148 * 8]; CK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  // per SM: warps_per_smsp blocks x 128 threads, each doing iters*exps_per_iter exps in `avg` cycles
  const double per_clk = (double)warps_per_smsp * 128 * iters * exps_per_iter / avg;
  printf("%-46s warps/SMSP=%d  %.1f exp/clk/SM  (%.2f cyc per warp-level 32 exps per SMSP)\n", name, warps_per_smsp, per_clk, 32.0 * 4 / per_clk);
  cudaFree(out); cudaFree(cyc);
}


// ---------------- TMEM read/write bandwidth: `warps` warps each issue LDTM.x32 (4 KB) / STTM.x32 back to back ----------------
template <int MODE>   // 0 = ld only, 1 = st only, 2 = ld + 32 FMNMX-ish consumers
__global__ void __launch_bounds__(512) tmem_bw(long long* cyc, float* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) b200::tmem_alloc<512>(&slot);
  b200::tc_fence_before();
  __syncthreads();
  b200::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  b200::tmem_st32(base, v); b200::tmem_st32(base + 32, v); b200::tmem_st32(base + 64, v); b200::tmem_st32(base + 96, v);
  b200::tmem_wait_st();
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 32) {
      if (MODE == 0 || MODE == 2) {
        uint32_t r[32];
        b200::tmem_ld32(base + c, r);
        if (MODE == 2) { b200::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc = fmaxf(acc, __uint_as_float(r[i])); }
        else { asm volatile("" :: "r"(r[0]), "r"(r[31])); }
      } else {
        b200::tmem_st32(base + c, v);
      }
    }
  }
  if (MODE == 0) b200::tmem_wait_ld();
  if (MODE == 1) b200::tmem_wait_st();
  const long long t1 = clock64();
  if (threadIdx.x % 32 == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  b200::tc_fence_before();
  __syncthreads();
  if (warp == 0) { b200::tc_fence_after(); b200::tmem_dealloc<512>(slot); }
}
template <int MODE>
void run_tmem(const char* name, int warps) {
  long long* cyc; float* out; const int iters = 2048;
  CK(cudaMalloc(&cyc, 148 * 16 * 8)); CK(cudaMalloc(&out, 148 * 512 * 4));
  tmem_bw<MODE><<<148, warps * 32>>>(cyc, out, 8); CK(cudaDeviceSynchronize());
  tmem_bw<MODE><<<148, warps * 32>>>(cyc, out, iters); CK(cudaDeviceSynchronize());
  long long h[148 * 16]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  printf("TMEM %-28s warps=%2d  %.1f B/clk/SM  (%.1f cyc per x32 op per warp)\n", name, warps, (double)warps * iters * 4 * 4096 / mx, mx / (iters * 4.0));
  cudaFree(cyc); cudaFree(out);
}

int main() {
  for (int w : {1, 4, 8, 16}) { run_tmem<0>("LDTM.x32 only", w); run_tmem<2>("LDTM.x32 + wait + 32 FMNMX", w); run_tmem<1>("STTM.x32 only", w); }

  for (int w = 1; w <= 4; w *= 2) {
    run<0>("f32 MUFU.EX2", 16, w);
    run<1>("f16x2 MUFU.EX2", 32, w);
    run<2>("bf16x2 MUFU.EX2", 32, w);
    run<3>("poly3 on FMA pipe", 16, w);
    run<4>("12 MUFU + 4 poly", 16, w);
    run<6>("10 MUFU + 6 poly", 16, w);
    run<5>("8 MUFU + 8 poly", 16, w);
    run<7>("softmax-like (fma+ex2+sum+pack) all MUFU", 16, w);
    run<8>("softmax-like 25% poly", 16, w);
    run<9>("softmax-like f16x2 (fma+cvt+ex2.f16x2)", 16, w);
  }
  return 0;
}
