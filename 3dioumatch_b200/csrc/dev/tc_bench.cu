// tc_bench.cu -- developer micro-benchmarks / self-tests of tcgen05 issue modes on sm_100a (not part of the public ABI).
//
//   b200_debug_tc_rate : cycles per tcgen05.mma (kind::tf32, M = 128, K = 8) issued back to back by one elected lane,
//                        for  mode 0: A and B from shared memory (SS), N = 128
//                             mode 1: SS, N = 256      mode 2: SS, N = 64
//                             mode 3: A from TMEM (TS), N = 128      mode 4: TS, N = 256      mode 5: TS, N = 64
//                        `stress` > 0 adds that many warps streaming float4 stores to shared memory meanwhile
//                        (the epilogue / gather traffic the fused kernel produces next to its MMAs).
//   b200_debug_tc_gemm_ts : C[128 x N] = A[128 x K] * W[N x K]^T with the A operand staged in TENSOR MEMORY
//                        (tcgen05.st) and split-precision TF32, pinning the TMEM A-operand layout before the fused
//                        kernel relies on it.
#include "../common.cuh"
#include "../tc_common.cuh"

namespace b200 {

__device__ __forceinline__ bool tcb_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- issue-rate benchmark -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
tc_rate_kernel(int mode, int iters, int stress, unsigned long long *out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a_s = base;                 // 4 k-blocks of A: 4 x 16 KB
  uint8_t *b_s = base + 4 * 16384;     // 2 k-blocks of B with up to 256 rows: 2 x 32 KB
  uint8_t *junk = b_s + 2 * 32768;     // 32 KB written by the stress warps
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (4 * 16384 + 2 * 32768) / 4; i += blockDim.x) reinterpret_cast<float *>(base)[i] = 0.f;
  if (warp == 0) tc::tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
    stop_flag = 0;
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  const int n = (mode % 3) == 0 ? 128 : ((mode % 3) == 1 ? 256 : 64);
  const bool ts = mode >= 3;
  const uint32_t idesc = tc::make_idesc_tf32(128, n);
  if (warp == 0) {
    const uint32_t a_addr = tc::smem_addr(a_s), b_addr = tc::smem_addr(b_s);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t da = tc::make_desc_sw128(a_addr + (uint32_t)(it & 3) * 16384u);
      const uint64_t db = tc::make_desc_sw128(b_addr + (uint32_t)(it & 1) * 32768u);
      if (tcb_elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);
          if (ts) {  // accumulators in columns [0, n), A operand read from columns 256.. (K = 8 columns per MMA)
            mma_tf32_ts(tmem_d, tmem_d + 256u + (uint32_t)((it & 3) * 32 + ks * 8), db + adv, idesc, 1u);
            mma_tf32_ts(tmem_d, tmem_d + 384u + (uint32_t)((it & 3) * 32 + ks * 8), db + adv, idesc, 1u);
          } else {
            tc::mma_tf32(tmem_d, da + adv, db + adv, idesc, 1u);
            tc::mma_tf32(tmem_d + (n <= 128 ? 256u : 0u), da + adv, db + adv, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (tcb_elect_one()) tc::mma_commit(&bar);
    __syncwarp();
    tc::mbar_wait(&bar, 0u);
    const long long t1 = clock64();
    if (lane == 0) {
      out[0] = (unsigned long long)(t1 - t0);
      out[1] = (unsigned long long)iters * 8ull;
      stop_flag = 1;
    }
  } else if (warp <= stress) {
    // shared-memory store traffic next to the MMAs: 512 B per warp instruction
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    unsigned long long stores = 0;
    while (!stop_flag) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        *reinterpret_cast<float4 *>(junk + ((warp * 2048 + u * 4096 + lane * 16) & 32767)) = v;
      stores += 8;
    }
    if (lane == 0) atomicAdd(&out[2], stores * 512ull);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem_d);
}

// ---- TS-mode GEMM self-test ---------------------------------------------------------------------------------------------
// 160 threads: warps 0-3 stage A into TMEM (row = lane) and W into shared memory, warp 4 issues.  K <= 96 per call
// (A_hi in columns [256, 256+K), A_lo in [384, 384+K)); accumulators: products [0, N), corrections [128, 128+N), N <= 128.
__global__ void __launch_bounds__(160, 1)
tc_gemm_ts_kernel(int N, int K, const float *__restrict__ A, const float *__restrict__ W, float *__restrict__ C) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nkb = K / 32;
  uint8_t *w_hi = base, *w_lo = base + (size_t)nkb * 16384;  // [kb][128 rows x 128 B]
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 4) tc::tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    tc::mbar_init(&bar_mma, 1);
    tc::mbar_fence_init();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  if (warp < 4) {
    const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int kb = 0; kb < nkb; ++kb) {
      uint32_t hi[32], lo[32];
      for (int j = 0; j < 32; ++j) {
        float h, l;
        tc::split_tf32(A[(size_t)tid * K + kb * 32 + j], h, l);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(l);
      }
      tmem_st_32x32(lane_addr + 256u + (uint32_t)(kb * 32), hi);
      tmem_st_32x32(lane_addr + 384u + (uint32_t)(kb * 32), lo);
      for (int r = tid; r < N; r += 128)
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4 *>(W + (size_t)r * K + kb * 32 + c * 4);
          float4 h, l;
          tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
          tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
          const uint32_t off = tc::sw128_offset(r, c);
          *reinterpret_cast<float4 *>(w_hi + (size_t)kb * 16384 + off) = h;
          *reinterpret_cast<float4 *>(w_lo + (size_t)kb * 16384 + off) = l;
        }
    }
    tmem_st_wait();
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  if (warp == 4) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    const uint32_t wh = tc::smem_addr(w_hi), wl = tc::smem_addr(w_lo);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint64_t dwh = tc::make_desc_sw128(wh + (uint32_t)kb * 16384u), dwl = tc::make_desc_sw128(wl + (uint32_t)kb * 16384u);
      if (tcb_elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);
          const uint32_t col = (uint32_t)(kb * 32 + ks * 8);
          mma_tf32_ts(tmem_d, tmem_d + 256u + col, dwh + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
          mma_tf32_ts(tmem_d + 128u, tmem_d + 384u + col, dwh + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
          mma_tf32_ts(tmem_d + 128u, tmem_d + 256u + col, dwl + adv, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (tcb_elect_one()) tc::mma_commit(&bar_mma);
    __syncwarp();
  }
  tc::mbar_wait(&bar_mma, 0u);
  tc::tc_fence_after_sync();
  if (warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32], r2[32];
      tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + 128u + (uint32_t)c0, r2);
      tc::tmem_ld_wait();
      for (int j = 0; j < 32; ++j) C[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<512>(tmem_d);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_debug_tc_rate(int mode, int iters, int stress, unsigned long long *out3_device, void *stream) {
  B200_CHECK_ARG(mode >= 0 && mode <= 5 && iters > 0 && stress >= 0 && stress <= 15, "tc_rate: bad arguments");
  const size_t smem = 1024 + 4 * 16384 + 2 * 32768 + 32768;
  B200_CUDA_OK(cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_rate_kernel<<<1, 512, smem, (cudaStream_t)stream>>>(mode, iters, stress, out3_device);
  B200_LAUNCH_OK("tc_rate_kernel");
  return 0;
}

// A (128,K), W (N,K), C (128,N) device fp32; K in {32,64,96}; N multiple of 32, <= 128
extern "C" int b200_debug_tc_gemm_ts(int N, int K, const float *A, const float *W, float *C, void *stream) {
  B200_CHECK_ARG(N >= 32 && N <= 128 && N % 32 == 0 && K >= 32 && K <= 96 && K % 32 == 0, "tc_gemm_ts: unsupported N=%d K=%d",
                 N, K);
  const size_t smem = 1024 + 2 * (size_t)(K / 32) * 16384;
  B200_CUDA_OK(cudaFuncSetAttribute(tc_gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_gemm_ts_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(N, K, A, W, C);
  B200_LAUNCH_OK("tc_gemm_ts_kernel");
  return 0;
}
