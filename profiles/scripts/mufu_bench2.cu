// Microbenchmark 2 (round 1, session 4): does the softmax inner loop get faster with PACKED f32x2 arithmetic
// (FFMA2 / FADD2: one issue slot for two elements) and part of the exponentials evaluated by a degree-3 polynomial on the
// FMA pipe instead of MUFU.EX2?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench2 mufu_bench2.cu
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
#define DEVI __device__ __forceinline__

DEVI float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
DEVI uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
DEVI void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
DEVI uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DEVI uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEVI uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEVI uint32_t pack_bf16(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }

// two exp2 on the FMA pipe: Cody-Waite split with the 1.5*2^23 magic constant + degree-3 polynomial on [-0.5, 0.5]
DEVI void ex2poly2(uint64_t X, float& e0, float& e1) {
  float x0, x1;
  unpack2(X, x0, x1);
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  const uint64_t Xc = pack2(x0, x1);
  const uint64_t MAG = pack2(12582912.f, 12582912.f);
  const uint64_t T = add2(Xc, MAG);
  const uint64_t R = sub2(T, MAG);
  const uint64_t Fr = sub2(Xc, R);
  uint64_t P = fma2(Fr, pack2(0.0555041087f, 0.0555041087f), pack2(0.2402265070f, 0.2402265070f));
  P = fma2(P, Fr, pack2(0.6931471806f, 0.6931471806f));
  P = fma2(P, Fr, pack2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack2(P, p0, p1); unpack2(T, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
DEVI float ex2poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;
  const float r = t - 12582912.f;
  const float f = x - r;
  float p = fmaf(f, 0.0555041087f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// MODE: 0 scalar all-MUFU (as the shipped kernel: FFMA + MUFU + FADD + F2FP/2)
//       1 packed scale / sum, all MUFU
//       2..5 packed, POLY pairs out of every 8 pairs = 1, 2, 3, 4  (12.5 %, 25 %, 37.5 %, 50 %)
//       6 scalar, 25 % scalar polynomial (round-1 mode 8)
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed, long long* cyc) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f;
  uint32_t acc = 0;
  float l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 6) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float x0 = fmaf(a[i], 0.18f, -seed), x1 = fmaf(a[i + 1], 0.18f, -seed), x2 = fmaf(a[i + 2], 0.18f, -seed), x3 = fmaf(a[i + 3], 0.18f, -seed);
        float e0 = ex2f(x0), e1 = ex2f(x1), e2 = ex2f(x2), e3 = (MODE == 6) ? ex2poly(x3) : ex2f(x3);
        s[0] += e0; s[1] += e1; s[2] += e2; s[3] += e3;
        acc ^= pack_bf16(e0, e1) + pack_bf16(e2, e3);
        a[i] = e0 - 1.5f; a[i + 1] = e1 - 1.5f; a[i + 2] = e2 - 1.5f; a[i + 3] = e3 - 1.5f;
      }
      l += (s[0] + s[1]) + (s[2] + s[3]);
    } else {
      constexpr int POLY = MODE - 1;      // poly pairs per 8 pairs
      const uint64_t SC = pack2(0.18f, 0.18f), NM = pack2(-seed, -seed), M15 = pack2(-1.5f, -1.5f);
      uint64_t S0 = pack2(0.f, 0.f), S1 = S0;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint64_t X = fma2(pack2(a[i], a[i + 1]), SC, NM);
        float e0, e1;
        if (((i >> 1) & 7) < POLY) ex2poly2(X, e0, e1);
        else { float x0, x1; unpack2(X, x0, x1); e0 = ex2f(x0); e1 = ex2f(x1); }
        const uint64_t E = pack2(e0, e1);
        if (i & 2) S1 = add2(S1, E); else S0 = add2(S0, E);
        acc ^= pack_bf16(e0, e1);
        const uint64_t A = add2(E, M15);
        unpack2(A, a[i], a[i + 1]);
      }
      float s0, s1; unpack2(add2(S0, S1), s0, s1);
      l += s0 + s1;
    }
  }
  const long long t1 = clock64();
  float s = l;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 0xff);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  const int blocks = 148 * warps_per_smsp, iters = 4096;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, blocks * 128 * 4)); CK(cudaMalloc(&cyc, blocks * 8));
  k<MODE><<<blocks, 128>>>(out, 16, 0.37f, cyc);
  CK(cudaDeviceSynchronize());
  k<MODE><<<blocks, 128>>>(out, iters, 0.37f, cyc);
  CK(cudaDeviceSynchronize());
  long long h[148 * 8]; CK(cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  const double per_clk = (double)warps_per_smsp * 128 * iters * 32 / avg;
  printf("%-52s warps/SMSP=%d  %.1f exp/clk/SM\n", name, warps_per_smsp, per_clk);
  cudaFree(out); cudaFree(cyc);
}

__global__ void acc_check(float* err) {
  float worst = 0.f;
  for (int i = threadIdx.x; i < 200000; i += blockDim.x) {
    const float x = -20.f + i * 1e-4f;
    float e0, e1;
    ex2poly2(pack2(x, x - 0.37f), e0, e1);
    worst = fmaxf(worst, fabsf(e0 - exp2f(x)) / exp2f(x));
    worst = fmaxf(worst, fabsf(e1 - exp2f(x - 0.37f)) / exp2f(x - 0.37f));
  }
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(worst));
}

int main() {
  float* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  acc_check<<<1, 256>>>(err); float h; CK(cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost));
  printf("poly3 exp2 max relative error on [-20, 0]: %.3e\n", h);
  for (int w = 1; w <= 4; w *= 2) {
    run<0>("scalar fma+MUFU+add+pack (shipped loop)", w);
    run<6>("scalar, 25% scalar poly", w);
    run<1>("packed FFMA2/FADD2, all MUFU", w);
    run<2>("packed, 12.5% poly", w);
    run<3>("packed, 25% poly", w);
    run<4>("packed, 37.5% poly", w);
    run<5>("packed, 50% poly", w);
  }
  return 0;
}
