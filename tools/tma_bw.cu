// Micro-benchmark (developer aid, GPU box): how fast can one SM stream 8 KB activation tiles on chip?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_bw tools/tma_bw.cu -lcuda && /tmp/tma_bw
// mode 0: TMA 3-D box {32 x 128 rows} of a [rows][32] bf16 tensor, SWIZZLE_64B   (what the layer kernels use)
// mode 1: TMA 2-D box {64 x 64 rows} of the same bytes viewed as [rows/2][64], SWIZZLE_128B
// mode 2: cp.async.bulk (1-D, 8 KB contiguous, no swizzle)
// mode 3: TMA 3-D box {32 x 256 rows}, SWIZZLE_64B (16 KB per request)
// mode 4: mode 0 loads plus two 8 KB TMA stores per loaded tile (the forward layer kernel's 1 : 2 read : write mix)
// Each CTA: one producer thread keeps `nst` tiles in flight, one consumer thread recycles them immediately.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64) k_stream(const __grid_constant__ CUtensorMap map, const unsigned char* base, int n_tiles, int nst, int tile_bytes) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[32], empty[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_my; ++i) {
      const int tile = blockIdx.x + i * gridDim.x, s = i % nst;
      mbar_wait(&empty[s], ((i / nst) & 1) ^ 1);
      mbar_expect(&full[s], tile_bytes);
      unsigned char* dst = smem + (size_t)s * tile_bytes;
      if (MODE == 0 || MODE == 4) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"((uint64_t)&map), "r"(s32(&full[s])), "r"(0), "r"(tile * 128), "r"(0) : "memory");
      } else if (MODE == 3) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"((uint64_t)&map), "r"(s32(&full[s])), "r"(0), "r"(tile * 256), "r"(0) : "memory");
      } else if (MODE == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)), "l"((uint64_t)&map), "r"(s32(&full[s])), "r"(0), "r"(tile * 64) : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"((uint64_t)(base + (size_t)tile * tile_bytes)), "r"(tile_bytes), "r"(s32(&full[s])) : "memory");
      }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < n_my; ++i) {
      const int s = i % nst;
      mbar_wait(&full[s], (i / nst) & 1);
      if (MODE == 4) {  // write the tile back twice (second half of the buffer), then recycle the stage
        const int tile = blockIdx.x + i * gridDim.x;
        unsigned char* src = smem + (size_t)s * tile_bytes;
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&map), "r"(s32(src)), "r"(0), "r"((n_tiles + 2 * tile) * 128), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&map), "r"(s32(src)), "r"(0), "r"((n_tiles + 2 * tile + 1) * 128), "r"(0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      mbar_arrive(&empty[s]);
    }
  }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = (size_t)2 << 30;  // 2 GiB streamed per run (>> L2)
  unsigned char* d;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemset(d, 1, bytes));
  cudaDriverEntryPointQueryResult q;
  void* p = nullptr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_enc enc = (PFN_enc)p;
  const uint64_t rows = bytes / 64;
  CUtensorMap m0, m1, m3;
  {
    cuuint64_t dims[3] = {32, rows, 1}, str[2] = {64, rows * 64};
    cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
    if (enc(&m0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return printf("enc0 failed\n");
    cuuint32_t box3[3] = {32, 256, 1};
    if (enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return printf("enc3 failed\n");
  }
  {
    cuuint64_t dims[2] = {64, rows / 2}, str[1] = {128};
    cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
    if (enc(&m1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return printf("enc1 failed\n");
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(k_stream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int mode = 0; mode < 5; ++mode) {
    if (mode == 1 || mode == 2 || mode == 3) continue;  // (box shape / bulk-copy variants: no difference, see gpurun_out/tma_bw.txt)
    const int tile_bytes = mode == 3 ? 16384 : 8192;
    const int n_tiles = (int)(bytes / tile_bytes) / (mode == 4 ? 3 : 1);  // mode 4: first third read, rest written
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
      for (int kb_in_flight : {32, 64, 96, 160}) {
        const int nst = kb_in_flight * 1024 / tile_bytes / ctas_per_sm;
        if (nst < 1 || nst > 32) continue;
        const size_t smem = (size_t)nst * tile_bytes + 1024;
        const int grid = 148 * ctas_per_sm;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          if (mode == 0) k_stream<0><<<grid, 64, smem>>>(m0, d, n_tiles, nst, tile_bytes);
          if (mode == 1) k_stream<1><<<grid, 64, smem>>>(m1, d, n_tiles, nst, tile_bytes);
          if (mode == 2) k_stream<2><<<grid, 64, smem>>>(m0, d, n_tiles, nst, tile_bytes);
          if (mode == 3) k_stream<3><<<grid, 64, smem>>>(m3, d, n_tiles, nst, tile_bytes);
          if (mode == 4) k_stream<4><<<grid, 64, smem>>>(m0, d, n_tiles, nst, tile_bytes);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        printf("mode %d  ctas/sm %d  in-flight %3d KB/SM (nst %2d)  %7.1f GB/s (read + written)\n", mode, ctas_per_sm, kb_in_flight, nst,
               (mode == 4 ? (double)n_tiles * tile_bytes * 3 : (double)bytes) / best * 1e-6);
      }
    }
  }
  return 0;
}
