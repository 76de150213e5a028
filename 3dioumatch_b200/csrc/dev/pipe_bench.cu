// pipe_bench.cu -- developer micro-benchmark (not part of the ABI headers): issue rate of the instruction kinds the FPS
// update is made of, at 1 / 2 / 4 warps per scheduler.  DESIGN.md section 7 item 1: the 16-warp FPS shape runs at IPC ~0.6
// and the question is which pipe (FMA: FFMA / FFMA2 / FADD2, ALU: FMNMX / FSETP+FSEL / SEL) sets that.
// One CTA on one SM; every warp runs `iters` rounds of 64 independent-chain instructions of one kind (8 chains x 8);
// result = SM cycles per warp-instruction per scheduler  ( = elapsed * 4 / (warps * instructions) ).
#include "../common.cuh"

namespace b200 {

template <int KIND>
__global__ void __launch_bounds__(512, 1) pipe_kernel(int iters, float seed, unsigned long long *cycles, float *sink) {
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + (float)(threadIdx.x + i);
    b[i] = seed * 0.5f + (float)i;
  }
  f32x2 p[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i] = pack2(a[i], b[i]);
    q[i] = pack2(b[i], a[i]);
  }
  int sel[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) {
          asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if (KIND == 1) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(q[i]));
        } else if (KIND == 2) {
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
        } else if (KIND == 3) {
          asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if (KIND == 4) {  // one tournament step: compare, select the value, select the index
          asm volatile(
              "{\n.reg .pred g;\nsetp.gt.f32 g, %2, %0;\nselp.f32 %0, %2, %0, g;\nselp.s32 %1, %3, %1, g;\n}\n"
              : "+f"(a[i]), "+r"(sel[i])
              : "f"(b[i]), "r"(i + r));
        } else {
          asm volatile("add.s32 %0, %0, %1;" : "+r"(sel[i]) : "r"(i + r));
        }
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    unpack2(p[i], lo, hi);
    acc += a[i] + lo + hi + (float)sel[i];
  }
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0) cycles[0] = (unsigned long long)(t1 - t0);
}

}  // namespace b200

using namespace b200;

// out[kind * 3 + w] for kind in {FFMA, FFMA2, FADD2, FMNMX, FSETP+2xSEL, IADD}, w in {4, 8, 16 warps}: cycles per
// warp-instruction per scheduler (KIND 4 counts its three instructions)
extern "C" int b200_debug_pipe_rates(int iters, float *out18) {
  unsigned long long *cyc = nullptr;
  float *sink = nullptr;
  B200_CUDA_OK(cudaMalloc(&cyc, sizeof(unsigned long long)));
  B200_CUDA_OK(cudaMalloc(&sink, sizeof(float)));
  const int warps[3] = {4, 8, 16};
  for (int kind = 0; kind < 6; ++kind) {
    for (int w = 0; w < 3; ++w) {
      for (int rep = 0; rep < 2; ++rep) {  // second run is the measurement
        switch (kind) {
          case 0: pipe_kernel<0><<<1, warps[w] * 32>>>(iters, 1.0f, cyc, sink); break;
          case 1: pipe_kernel<1><<<1, warps[w] * 32>>>(iters, 1.0f, cyc, sink); break;
          case 2: pipe_kernel<2><<<1, warps[w] * 32>>>(iters, 1.0f, cyc, sink); break;
          case 3: pipe_kernel<3><<<1, warps[w] * 32>>>(iters, 1.0f, cyc, sink); break;
          case 4: pipe_kernel<4><<<1, warps[w] * 32>>>(iters, 1.0f, cyc, sink); break;
          default: pipe_kernel<5><<<1, warps[w] * 32>>>(iters, 1.0f, cyc, sink); break;
        }
      }
      B200_CUDA_OK(cudaDeviceSynchronize());
      unsigned long long c = 0;
      B200_CUDA_OK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
      const double instr = (double)iters * 64.0 * (kind == 4 ? 3.0 : 1.0);
      out18[kind * 3 + w] = (float)((double)c * 4.0 / (warps[w] * instr));
    }
  }
  cudaFree(cyc);
  cudaFree(sink);
  return 0;
}
