// tc_gemm_test.cu -- developer self-test of the tcgen05 building blocks (tc_common.cuh): C[128 x N] = A[128 x K] * W[N x K]^T
// with split-precision TF32 (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM).  Not part of the public ABI; used by
// tests/test_gpu_tc_gemm.py to pin descriptor encodings / swizzle / TMEM addressing before the fused kernel relies on them.
#include "../common.cuh"
#include "../tc_common.cuh"

namespace b200 {

// 160 threads: warps 0-3 stage operands and run the epilogue, warp 4 issues the MMAs.
template <int N>
__global__ void __launch_bounds__(160, 1)
tc_gemm_test_kernel(int K, const float *__restrict__ A, const float *__restrict__ W, float *__restrict__ C, int passes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand k-blocks: A_hi, A_lo (128 rows), W_hi, W_lo (N rows); 128 B per row
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a_hi = base, *a_lo = a_hi + 128 * 128, *w_hi = a_lo + 128 * 128, *w_lo = w_hi + N * 128;
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 4) tc::tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) {
    tc::mbar_init(&bar_mma, 1);
    tc::mbar_fence_init();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  constexpr uint32_t idesc = tc::make_idesc_tf32(128, N);

  const int nkb = K / 32;
  for (int kb = 0; kb < nkb; ++kb) {
    if (warp < 4) {
      // thread t stages row t of A and rows t, t+128 (N=256) of W for this k-block
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(A + (size_t)tid * K + kb * 32 + c * 4);
        float4 h, l;
        tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
        tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
        const uint32_t off = tc::sw128_offset(tid, c);
        *reinterpret_cast<float4 *>(a_hi + off) = h;
        *reinterpret_cast<float4 *>(a_lo + off) = l;
      }
      for (int r = tid; r < N; r += 128)
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4 *>(W + (size_t)r * K + kb * 32 + c * 4);
          float4 h, l;
          tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
          tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
          const uint32_t off = tc::sw128_offset(r, c);
          *reinterpret_cast<float4 *>(w_hi + off) = h;
          *reinterpret_cast<float4 *>(w_lo + off) = l;
        }
      tc::fence_proxy_async_smem();
    }
    __syncthreads();
    if (tid == 128) {
      tc::tc_fence_after_sync();
      const uint64_t da_hi = tc::make_desc_sw128(tc::smem_addr(a_hi)), da_lo = tc::make_desc_sw128(tc::smem_addr(a_lo));
      const uint64_t dw_hi = tc::make_desc_sw128(tc::smem_addr(w_hi)), dw_lo = tc::make_desc_sw128(tc::smem_addr(w_lo));
      for (int ks = 0; ks < 4; ++ks) {       // UMMA_K = 8 floats = 32 bytes = +2 in the descriptor's 16-byte units
        const uint64_t adv = (uint64_t)(ks * 2);
        tc::mma_tf32(tmem_d, da_hi + adv, dw_hi + adv, idesc, (kb | ks) != 0);
        if (passes == 4) {  // small terms in their own accumulator (columns N..2N), summed in the epilogue
          tc::mma_tf32(tmem_d + N, da_lo + adv, dw_hi + adv, idesc, (kb | ks) != 0);
          tc::mma_tf32(tmem_d + N, da_hi + adv, dw_lo + adv, idesc, 1);
        } else {
          if (passes >= 2) tc::mma_tf32(tmem_d, da_lo + adv, dw_hi + adv, idesc, 1);
          if (passes >= 3) tc::mma_tf32(tmem_d, da_hi + adv, dw_lo + adv, idesc, 1);
        }
      }
      tc::mma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, (uint32_t)(kb & 1));  // operands may be overwritten, accumulator is up to date
  }
  tc::tc_fence_after_sync();
  if (warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_wait();
      if (passes == 4 && N == 128) {
        uint32_t r2[32];
        tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N + c0), r2);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) C[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
      } else {
        for (int j = 0; j < 32; ++j) C[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<256>(tmem_d);
}

}  // namespace b200

using namespace b200;

// developer hook (not declared in include/): A (128,K), W (N,K), C (128,N) device fp32; K % 32 == 0; N in {128, 256}
extern "C" int b200_debug_tc_gemm(int N, int K, const float *A, const float *W, float *C, int passes, void *stream) {
  B200_CHECK_ARG((N == 128 || N == 256) && K > 0 && K % 32 == 0, "tc_gemm: unsupported shape N=%d K=%d", N, K);
  const size_t smem = 1024 + 2 * 128 * 128 + 2 * (size_t)N * 128;
  if (N == 128) {
    B200_CUDA_OK(cudaFuncSetAttribute(tc_gemm_test_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_test_kernel<128><<<1, 160, smem, (cudaStream_t)stream>>>(K, A, W, C, passes);
  } else {
    B200_CUDA_OK(cudaFuncSetAttribute(tc_gemm_test_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_test_kernel<256><<<1, 160, smem, (cudaStream_t)stream>>>(K, A, W, C, passes);
  }
  B200_LAUNCH_OK("tc_gemm_test_kernel");
  return 0;
}
