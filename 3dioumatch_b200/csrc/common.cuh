// common.cuh -- shared host/device helpers of libb200pc.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace b200 {

// ---- error reporting (C ABI: int status + b200_last_error()) -------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define B200_CHECK_ARG(cond, ...)  \
  do {                             \
    if (!(cond)) {                 \
      b200::set_error(__VA_ARGS__); \
      return 1;                    \
    }                              \
  } while (0)

#define B200_CUDA_OK(expr)                                                                    \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

#define B200_LAUNCH_OK(name)                                                                   \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      b200::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 3;                                                                                \
    }                                                                                          \
    b200::count_launch();                                                                      \
  } while (0)

// Per-device host caches.  The library is called from one python thread per replica under nn.DataParallel
// (train.py:187-191 of the reference), each bound to its own device: nothing below may be a process-global scalar.
constexpr int B200_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < B200_MAX_DEVICES) ? dev : 0;
}

inline int num_sms() {
  static std::atomic<int> n[B200_MAX_DEVICES] = {};
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is kept by the runtime PER DEVICE: remember the largest opt-in made for a
// kernel on each device (one DynSmemOptIn object per kernel instantiation, usually a function-local static).
struct DynSmemOptIn {
  std::atomic<size_t> bytes[B200_MAX_DEVICES] = {};
  template <typename Kernel>
  cudaError_t ensure(Kernel kern, size_t need) {
    const int dev = current_device();
    if (need <= bytes[dev].load(std::memory_order_acquire)) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
    if (e == cudaSuccess) {  // racing threads of one device may both set it: the attribute only ever grows
      size_t cur = bytes[dev].load(std::memory_order_relaxed);
      while (cur < need && !bytes[dev].compare_exchange_weak(cur, need, std::memory_order_release)) {
      }
    }
    return e;
  }
};

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Stream-ordered scratch memory.  The device's default pool gives freed blocks back to the driver at every
// synchronisation unless a release threshold is set, which turns each eager call after a sync into a driver
// allocation (hundreds of microseconds); keep them pooled instead.
inline cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t stream) {
  static std::atomic<bool> tuned[B200_MAX_DEVICES] = {};
  const int dev = current_device();
  if (!tuned[dev].load(std::memory_order_relaxed)) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    tuned[dev].store(true, std::memory_order_relaxed);
  }
  return cudaMallocAsync(p, bytes, stream);
}

// Stream-ordered scratch that is returned to the pool on every exit path of a launcher (error returns included).
struct ScratchGuard {
  void *ptr = nullptr;
  cudaStream_t stream = nullptr;
  ScratchGuard() = default;
  ScratchGuard(const ScratchGuard &) = delete;
  ScratchGuard &operator=(const ScratchGuard &) = delete;
  cudaError_t alloc(size_t bytes, cudaStream_t s) {
    stream = s;
    return scratch_alloc(&ptr, bytes, s);
  }
  ~ScratchGuard() {
    if (ptr) cudaFreeAsync(ptr, stream);
  }
};

// ---- the reference's squared-distance rounding sequence -------------------------------------
// nvcc contracts (a*a + b*b + c*c) of the reference kernels (ball_query_gpu.cu:36-37,
// sampling_gpu.cu:105,108-109, interpolate_gpu.cu:38) for sm_100a into
//     fma(c,c, fma(a,a, mul(b,b)))
// Index parity is bit-exact, so the sequence is pinned with intrinsics the compiler may
// not re-associate or re-contract.
__device__ __forceinline__ float sq3(float a, float b, float c) {
  float t = __fmul_rn(b, b);
  t = __fmaf_rn(a, a, t);
  return __fmaf_rn(c, c, t);
}
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  return sq3(__fsub_rn(ax, bx), __fsub_rn(ay, by), __fsub_rn(az, bz));
}

// ---- packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2: one issue slot per two IEEE-rounded fp32 operations) ---------
// Same rounding per element as the scalar intrinsics above, so the squared-distance sequence stays bit-exact.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// (a - b)^2 summed over x, y, z for two points at once: fma(dz,dz, fma(dx,dx, dy*dy)) per element
__device__ __forceinline__ f32x2 sqdist3_x2(f32x2 ax, f32x2 ay, f32x2 az, f32x2 bx, f32x2 by, f32x2 bz) {
  f32x2 dx, dy, dz, t;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(ax), "l"(bx));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(ay), "l"(by));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(az), "l"(bz));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dy));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dx), "l"(t));
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dz), "l"(t));
  return t;
}

}  // namespace b200
