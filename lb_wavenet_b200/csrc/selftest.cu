// Device self-test of the TMA + tcgen05 + TMEM building blocks (umma.cuh) and the host-side
// tensor-map encoder.  C[M,N] = A[M,K] * B[N,K]^T with A, B bf16 row-major (both K-major).
//   mode 0: operand tiles land in shared memory through TMA with hardware swizzle;
//   mode 1: operand tiles are written by the threads with umma::swizzled_offset (the layout the
//           fused kernels' epilogues use when an activation tile becomes the next MMA's A operand).
#include "common.cuh"
#include "umma.cuh"

namespace wn {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    WN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (p == nullptr || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return WN_ERR_CUDA;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                        bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, swizzle %d)", (int)r, rank, swizzle_bytes);
    return WN_ERR_CUDA;
  }
  return WN_OK;
}

using namespace umma;

// one CTA per 128-row tile of C; 128 threads; K consumed in blocks of one swizzle span
template <int SW>
__global__ void __launch_bounds__(128) k_selftest_gemm(const __grid_constant__ CUtensorMap map_a,
                                                       const __grid_constant__ CUtensorMap map_b,
                                                       const bf16* __restrict__ A, const bf16* __restrict__ B,
                                                       float* __restrict__ C, int M, int N, int K, int mode) {
  constexpr int KB = SW / 2;  // bf16 elements per swizzle span
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_tile = smem;                  // 128 x SW bytes
  unsigned char* b_tile = smem + 128 * SW;       // N x SW bytes (N <= 256)
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * 128;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;

  if (tid == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, N);
  const int nkb = K / KB;
  for (int kb = 0; kb < nkb; ++kb) {
    const uint32_t par = kb & 1;
    if (mode == 0) {
      if (tid == 0) {
        mbar_expect_tx(&full_bar, (uint32_t)((128 + N) * SW));
        tma_load_2d(a_tile, &map_a, &full_bar, kb * KB, m0);
        tma_load_2d(b_tile, &map_b, &full_bar, kb * KB, 0);
      }
    } else {
      // 16-byte chunks written by the threads at their swizzled positions
      constexpr int CPR = SW / 16;
      for (int idx = tid; idx < (128 + N) * CPR; idx += 128) {
        const int r = idx / CPR, c = idx % CPR;
        uint4 v = make_uint4(0, 0, 0, 0);
        unsigned char* dst;
        if (r < 128) {
          if (m0 + r < M) v = *reinterpret_cast<const uint4*>(A + (size_t)(m0 + r) * K + kb * KB + c * 8);
          dst = a_tile + swizzled_offset(r, c * 16, SW);
        } else {
          v = *reinterpret_cast<const uint4*>(B + (size_t)(r - 128) * K + kb * KB + c * 8);
          dst = b_tile + swizzled_offset(r - 128, c * 16, SW);
        }
        *reinterpret_cast<uint4*>(dst) = v;
      }
      fence_proxy_async_smem();
      __syncthreads();
    }
    if (tid == 0) {
      if (mode == 0) mbar_wait(&full_bar, par);
      tc_fence_after_sync();
#pragma unroll
      for (int k = 0; k < KB / 16; ++k) {
        const uint64_t ad = make_kmajor_desc(smem_u32(a_tile), SW, k * 32);
        const uint64_t bd = make_kmajor_desc(smem_u32(b_tile), SW, k * 32);
        mma_bf16_ss(tmem_base, ad, bd, idesc, (kb | k) != 0);
      }
      mma_commit(&mma_bar);
    }
    // everyone waits until the MMAs have consumed the tiles before they are overwritten
    mbar_wait(&mma_bar, par);
    tc_fence_after_sync();
    __syncthreads();
  }
  // epilogue: warp w owns TMEM lanes [32w, 32w+32) == rows m0 + 32w + lane
  const int row = m0 + warp * 32 + (tid & 31);
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, r);
    tmem_ld_wait();
    if (row < M) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < N) C[(size_t)row * N + n0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// CTA-pair variant (cta_group::2): a cluster of two CTAs computes 256 rows of C against ONE copy of each B block --
// each CTA loads its own 128 rows of A and half of the N rows of B; the leader issues M = 256 MMAs that read both
// CTAs' shared memory and write both CTAs' tensor memory.  SW128, K blocks of 64, one stage (a recipe check, not a
// fast kernel).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k_selftest_gemm_pair(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     float* __restrict__ C, int M, int N, int K) {
  constexpr int SW = 128, KB = 64;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* a_tile = smem;                  // 128 rows x 128 bytes
  unsigned char* b_tile = smem + 128 * SW;       // N / 2 rows x 128 bytes
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int m0 = (int)(blockIdx.x >> 1) * 256 + (int)rank * 128;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (tid == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 0) tmem_alloc_pair(&tmem_base_s, ncols);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any remote completion / multicast arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(256, N);
  const int nkb = K / KB;
  for (int kb = 0; kb < nkb; ++kb) {
    const uint32_t par = kb & 1;
    if (tid == 0) {
      if (rank == 0) mbar_expect_tx(&full_bar, (uint32_t)(2 * (128 + N / 2) * SW));  // both CTAs' loads
      tma_load_2d_pair(a_tile, &map_a, &full_bar, kb * KB, m0);
      tma_load_2d_pair(b_tile, &map_b, &full_bar, kb * KB, (int)rank * (N / 2));
      if (rank == 0) {
        mbar_wait(&full_bar, par);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          const uint64_t ad = make_kmajor_desc(smem_u32(a_tile), SW, k * 32);
          const uint64_t bd = make_kmajor_desc(smem_u32(b_tile), SW, k * 32);
          mma_bf16_ss_pair(tmem_base, ad, bd, idesc, (kb | k) != 0);
        }
        mma_commit_pair(&mma_bar, 3);  // both CTAs: the tiles may be overwritten, the accumulators read
      }
    }
    mbar_wait(&mma_bar, par);
    tc_fence_after_sync();
    __syncthreads();
  }
  const int row = m0 + warp * 32 + (tid & 31);
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, r);
    tmem_ld_wait();
    if (row < M) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n0 + j < N) C[(size_t)row * N + n0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the other may still be served by it
  if (warp == 0) tmem_dealloc_pair(tmem_base, ncols);
}

}  // namespace wn

using namespace wn;

extern "C" int wn_selftest_umma_gemm(const void* d_a, const void* d_b, float* d_c, int32_t M, int32_t N, int32_t K,
                                     int32_t swizzle, void* stream_) {
  const int mode = swizzle < 0 ? 1 : 0;
  const int sw = swizzle < 0 ? -swizzle : swizzle;
  if (!d_a || !d_b || !d_c || M < 1 || N < 16 || N > 256 || N % 16 || (sw != 32 && sw != 64 && sw != 128) ||
      K % (sw / 2) != 0 || K < sw / 2) {
    set_error("wn_selftest_umma_gemm: invalid argument (N multiple of 16 <= 256, K multiple of swizzle/2)");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  CUtensorMap ma, mb;
  const uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)M}, dims_b[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t str[1] = {(uint64_t)K * 2};
  const uint32_t box_a[2] = {(uint32_t)(sw / 2), 128}, box_b[2] = {(uint32_t)(sw / 2), (uint32_t)N};
  int rc;
  if ((rc = make_tensor_map_bf16(&ma, d_a, 2, dims_a, str, box_a, sw))) return rc;
  if ((rc = make_tensor_map_bf16(&mb, d_b, 2, dims_b, str, box_b, sw))) return rc;
  const size_t smem = (size_t)(128 + 256) * sw + 1024;
  const dim3 grid((M + 127) / 128);
  const bf16* A = reinterpret_cast<const bf16*>(d_a);
  const bf16* B = reinterpret_cast<const bf16*>(d_b);
  if (sw == 128) {
    WN_CUDA_CHECK(cudaFuncSetAttribute(k_selftest_gemm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_selftest_gemm<128><<<grid, 128, smem, st>>>(ma, mb, A, B, d_c, M, N, K, mode);
  } else if (sw == 64) {
    k_selftest_gemm<64><<<grid, 128, smem, st>>>(ma, mb, A, B, d_c, M, N, K, mode);
  } else {
    k_selftest_gemm<32><<<grid, 128, smem, st>>>(ma, mb, A, B, d_c, M, N, K, mode);
  }
  WN_LAUNCH_CHECK();
  return WN_OK;
}

extern "C" int wn_selftest_umma_gemm_pair(const void* d_a, const void* d_b, float* d_c, int32_t M, int32_t N, int32_t K,
                                          void* stream_) {
  if (!d_a || !d_b || !d_c || M < 1 || N < 32 || N > 256 || N % 32 || K % 64 != 0 || K < 64) {
    set_error("wn_selftest_umma_gemm_pair: invalid argument (N multiple of 32 <= 256, K multiple of 64)");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  CUtensorMap ma, mb;
  const uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)M}, dims_b[2] = {(uint64_t)K, (uint64_t)N};
  const uint64_t str[1] = {(uint64_t)K * 2};
  const uint32_t box_a[2] = {64, 128}, box_b[2] = {64, (uint32_t)(N / 2)};
  int rc;
  if ((rc = make_tensor_map_bf16(&ma, d_a, 2, dims_a, str, box_a, 128))) return rc;
  if ((rc = make_tensor_map_bf16(&mb, d_b, 2, dims_b, str, box_b, 128))) return rc;
  const size_t smem = (size_t)(128 + 128) * 128 + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_selftest_gemm_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int pairs = (M + 255) / 256;
  k_selftest_gemm_pair<<<2 * pairs, 128, smem, st>>>(ma, mb, d_c, M, N, K);
  WN_LAUNCH_CHECK();
  return WN_OK;
}
