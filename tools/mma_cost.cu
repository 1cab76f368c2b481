// Micro-benchmark (developer aid, GPU box): tensor-pipe cost of one tcgen05.mma (M = 128, K = 16) as a function of N
// and of the operand majorness, issued back to back by one thread (what the layer kernels' MMA warp does).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lb_wavenet_b200/csrc -o tools/bin/mma_cost tools/mma_cost.cu
#include <cstdio>
#include <cstdlib>

#include "umma.cuh"

using namespace wn::umma;

__global__ void __launch_bounds__(128) k_cost(int N, int a_mn, int b_mn, int sw, int reps, long long* out, int alt) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tm_s;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tm_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn != 0, b_mn != 0);
    const uint64_t da = a_mn ? make_mnmajor_desc(a, sw, 8192) : make_kmajor_desc(a, sw, 0);
    const uint64_t db = b_mn ? make_mnmajor_desc(b, sw, 8192) : make_kmajor_desc(b, sw, 0);
    // warm-up
    for (int i = 0; i < 8; ++i) mma_bf16_ss(tm_s, da, db, idesc, i != 0);
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) mma_bf16_ss(tm_s + (i & alt) * 256, da, db, idesc, true);
    const long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm_s, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 2000;
  for (int sw : {64}) {
    for (int mn = 0; mn < 4; mn += 3) {
     for (int alt = 0; alt < 2; ++alt)
      for (int N : {32, 64, 96, 128}) {
        printf("alt %d ", alt);
        k_cost<<<1, 128, 100 * 1024>>>(N, mn & 1, mn >> 1, sw, reps, d, alt);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
          printf("error %s\n", cudaGetErrorString(e));
          return 1;
        }
        printf("sw %3d  A %s  B %s  N %3d : issue %6.1f cyc/mma, complete %6.1f cyc/mma\n", sw, (mn & 1) ? "MN" : "K ",
               (mn >> 1) ? "MN" : "K ", N, (double)h[0] / reps, (double)h[1] / reps);
      }
    }
  }
  return 0;
}
