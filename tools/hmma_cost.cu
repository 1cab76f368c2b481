// Microbenchmark for the generator redesign: what does ONE warp pay for a layer's worth of mma.sync.m16n8k16 (32 MMAs as
// 8 independent accumulators x 4 k-steps, B fragments from shared memory) and for its gate (32 tanh.approx per lane)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/hmma_cost tools/hmma_cost.cu && gpurun_out/hmma_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(long long* out, float* sink, int iters, int nwarps_active) {
  __shared__ uint2 bsm[32 * 32];  // 32 (n-tile, k-step) blocks x 32 lanes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) bsm[i] = make_uint2(0x3f803f80u + i, 0x3f803f80u);
  __syncthreads();
  if (warp >= nwarps_active) return;
  uint32_t a[4][4];
  for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) a[k][j] = 0x3f003f00u + lane + k + j;
  float acc[8][4] = {};
  float g = 0.01f * lane;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE & 1) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint2 b = bsm[(nt * 4 + ks) * 32 + lane];
          mma16816(acc[nt], a[ks], b.x, b.y);
        }
    }
    if (MODE & 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = (MODE & 1) ? acc[j >> 2][j & 3] : g + j;
        const float q = (MODE & 1) ? acc[4 + (j >> 2)][j & 3] : g - j;
        g += tanh_fast(s) * (0.5f * tanh_fast(0.5f * q) + 0.5f);
      }
      if (MODE & 1) { acc[0][0] = g; }
    }
    if (MODE & 4) {  // residual: 8 more MMAs that depend on the gate
      uint32_t z[4] = {__float_as_uint(g), __float_as_uint(g) + 1, __float_as_uint(g) + 2, __float_as_uint(g) + 3};
      float r[4][4] = {};
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint2 b = bsm[(nt * 2 + ks) * 32 + lane];
          mma16816(r[nt], z, b.x, b.y);
        }
      a[0][0] = __float_as_uint(r[0][0] + r[1][1] + r[2][2] + r[3][3]);
    }
  }
  long long t1 = clock64();
  float s = g;
  for (int nt = 0; nt < 8; ++nt) for (int j = 0; j < 4; ++j) s += acc[nt][j];
  sink[threadIdx.x] = s;
  if (lane == 0) out[warp] = t1 - t0;
}

template <int MODE>
void run(const char* what, int nw) {
  long long* d; float* s; cudaMalloc(&d, 64 * 8); cudaMalloc(&s, 1024 * 4);
  const int iters = 2000;
  k<MODE><<<1, 128, 0>>>(d, s, iters, nw);
  cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%-60s warps %d: %.1f cycles per iteration (warp 0)\n", what, nw, (double)h[0] / iters);
  cudaFree(d); cudaFree(s);
}
int main() {
  run<1>("conv: 32 mma.sync (8 accumulators x 4 k-steps), B from smem", 1);
  run<2>("gate: 16 x (tanh + sigmoid) per lane", 1);
  run<3>("conv + gate (dependent)", 1);
  run<7>("conv + gate + residual (8 dependent MMAs): one layer", 1);
  run<7>("one layer, 4 warps on 4 schedulers", 4);
  return 0;
}
