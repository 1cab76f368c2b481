// Counter-based sampler shared by the generator kernels (replaces unseeded tf.multinomial, reference
// imodel.py:179).  Canonical evaluation order: oracle/wavenet_oracle.py sample_from_logits -- reproduced bit
// for bit (separately rounded multiplies / adds only, fixed scan order).
#pragma once
#include "common.cuh"
#include "tables.inc"

namespace wn {

__device__ __forceinline__ uint32_t philox4x32_10_w0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}

__device__ __forceinline__ float sampler_uniform(uint64_t seed, uint64_t step, uint32_t stream) {
  const uint32_t w0 = philox4x32_10_w0((uint32_t)step, (uint32_t)(step >> 32), stream, 0u, (uint32_t)seed,
                                       (uint32_t)(seed >> 32));
  return (float)(w0 >> 8) * 5.9604644775390625e-8f;  // 24 bits -> [0,1)
}

// exp(x), x <= 0, with separately rounded multiplies and adds only (never contracted to fma)
__device__ __forceinline__ float det_exp(float x) {
  const float t = __fmul_rn(x, __uint_as_float(kLog2eBits));
  if (t < -120.f) return 0.f;
  const float n = rintf(t);
  const float f = __fsub_rn(t, n);
  constexpr uint32_t kCoef[7] = WN_EXP2_COEF_BITS;
  float p = __uint_as_float(kCoef[6]);
#pragma unroll
  for (int i = 5; i >= 0; --i) p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(kCoef[i]));
  const float scale = __int_as_float(((int)n + 127) << 23);
  return __fmul_rn(p, scale);
}

// one warp samples one row of 256 logits; lane owns entries 8*lane .. 8*lane+7
__device__ __forceinline__ int warp_sample(const float* lg, float u) {
  const int lane = threadIdx.x & 31;
  float v[8];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = lg[lane * 8 + j];
    mx = fmaxf(mx, v[j]);
  }
  mx = warp_max(mx);
  float s[8];
  float run = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    run = __fadd_rn(run, det_exp(__fsub_rn(v[j], mx)));
    s[j] = run;
  }
  float incl = run;  // Kogge-Stone inclusive scan over lane totals
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl = __fadd_rn(incl, o);
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0.f;
  const float total = __shfl_sync(0xffffffffu, incl, 31);
  const float thr = __fmul_rn(u, total);
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) cnt += (__fadd_rn(excl, s[j]) <= thr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  return min(cnt, 255);
}

// two rows in lock step (the generator's warps own two streams each): the same arithmetic per row, so the result is
// bit-identical to two warp_sample calls; the two dependency chains (running sums, scans, shuffles) overlap
__device__ __forceinline__ void warp_sample2(const float* lg0, const float* lg1, float u0, float u1, int& r0, int& r1) {
  const int lane = threadIdx.x & 31;
  const float* lg[2] = {lg0, lg1};
  const float u[2] = {u0, u1};
  float v[2][8], mx[2], s[2][8], incl[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(lg[k] + lane * 8);
    const float4 b = *reinterpret_cast<const float4*>(lg[k] + lane * 8 + 4);
    v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w;
    v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
    mx[k] = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[k] = fmaxf(mx[k], v[k][j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xffffffffu, mx[0], o));
    mx[1] = fmaxf(mx[1], __shfl_xor_sync(0xffffffffu, mx[1], o));
  }
  float e[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    e[0][j] = det_exp(__fsub_rn(v[0][j], mx[0]));
    e[1][j] = det_exp(__fsub_rn(v[1][j], mx[1]));
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float run = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      run = __fadd_rn(run, e[k][j]);
      s[k][j] = run;
    }
    incl[k] = run;
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float o0 = __shfl_up_sync(0xffffffffu, incl[0], d);
    const float o1 = __shfl_up_sync(0xffffffffu, incl[1], d);
    if (lane >= d) {
      incl[0] = __fadd_rn(incl[0], o0);
      incl[1] = __fadd_rn(incl[1], o1);
    }
  }
  int cnt[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float excl = __shfl_up_sync(0xffffffffu, incl[k], 1);
    if (lane == 0) excl = 0.f;
    const float total = __shfl_sync(0xffffffffu, incl[k], 31);
    const float thr = __fmul_rn(u[k], total);
    cnt[k] = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) cnt[k] += (__fadd_rn(excl, s[k][j]) <= thr) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt[0] += __shfl_xor_sync(0xffffffffu, cnt[0], o);
    cnt[1] += __shfl_xor_sync(0xffffffffu, cnt[1], o);
  }
  r0 = min(cnt[0], 255);
  r1 = min(cnt[1], 255);
}

}  // namespace wn
