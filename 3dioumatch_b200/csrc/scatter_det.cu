// scatter_det.cu -- deterministic gradients of gather_points / group_points / three_interpolate.
//
// The reference accumulates these gradients with atomicAdd in whatever order the blocks happen to run
// (sampling_gpu.cu:39-52, group_points_gpu.cu:48-68, interpolate_gpu.cu:121-148), so two runs of the same step differ in
// the last bits.  Here the scatter is inverted first: a stable radix sort of (scene * N + target index) with the entry
// position as payload lists, for every target point, the entries that feed it in ascending position; one thread per
// (target, channel chunk) then sums its segment in that fixed order with plain fp32 adds.  The order is the one a
// sequential loop over the entries uses, so the result is bit-identical to the CPU oracle (oracle/pointnet2_oracle.c) and
// to itself from run to run.
#include <cub/device/device_radix_sort.cuh>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"

namespace b200 {

constexpr int SD_THREADS = 256;
constexpr int SD_CCHUNK = 8;

__global__ void __launch_bounds__(SD_THREADS)
det_key_kernel(long long total, int E, int Ntgt, const int32_t *__restrict__ idx, unsigned *__restrict__ keys,
               int *__restrict__ vals) {
  const long long i = (long long)blockIdx.x * SD_THREADS + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / E), e = (int)(i - (long long)b * E);
  keys[i] = (unsigned)b * (unsigned)Ntgt + (unsigned)idx[i];
  vals[i] = e;
}

// seg_start / seg_end are zero-filled first: a target nobody points at keeps the empty segment [0, 0)
__global__ void __launch_bounds__(SD_THREADS)
det_bounds_kernel(long long total, unsigned nseg, const unsigned *__restrict__ keys, int *__restrict__ seg_start,
                  int *__restrict__ seg_end) {
  const long long i = (long long)blockIdx.x * SD_THREADS + threadIdx.x;
  if (i >= total) return;
  const unsigned k = keys[i];
  if (k >= nseg) return;  // an index outside [0, N) of the last scene: never written out of bounds (undefined in the reference)
  if (i == 0 || keys[i - 1] != k) seg_start[k] = (int)i;
  if (i == total - 1 || keys[i + 1] != k) seg_end[k] = (int)i + 1;
}

// grad_points[b,c,t] = sum over the entries e of target t, ascending e, of  weight[b,e] * grad_out[b,c,e / per_src]
template <bool WEIGHTED>
__global__ void __launch_bounds__(SD_THREADS)
det_reduce_kernel(int C, int Ntgt, int E, int per_src, const float *__restrict__ grad_out,
                  const float *__restrict__ weight, const int *__restrict__ vals, const int *__restrict__ seg_start,
                  const int *__restrict__ seg_end, float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const int t = blockIdx.x * SD_THREADS + threadIdx.x;
  if (t >= Ntgt) return;
  const size_t seg = (size_t)b * Ntgt + t;
  const int s0 = seg_start[seg], s1 = seg_end[seg];
  const int Esrc = E / per_src;
  const int c_begin = blockIdx.y * SD_CCHUNK, c_end = min(C, c_begin + SD_CCHUNK);
  for (int c = c_begin; c < c_end; ++c) {
    const float *go = grad_out + ((size_t)b * C + c) * Esrc;
    float acc = 0.f;
    for (int i = s0; i < s1; ++i) {
      const int e = vals[i];
      const float g = go[per_src == 1 ? e : e / per_src];
      acc = WEIGHTED ? __fadd_rn(acc, __fmul_rn(g, weight[(size_t)b * E + e])) : __fadd_rn(acc, g);
    }
    grad_points[((size_t)b * C + c) * Ntgt + t] = acc;
  }
}

static int end_bit_for(unsigned long long n) {
  int bits = 1;
  while (bits < 32 && (1ull << bits) < n) ++bits;
  return bits;
}

struct DetLayout {
  size_t keys_in, keys_out, vals_in, vals_out, seg_start, seg_end, cub, total_bytes, cub_bytes;
};

static DetLayout det_layout(int B, int Ntgt, long long total) {
  DetLayout L;
  auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  L.keys_in = off; off += a256(sizeof(unsigned) * (size_t)total);
  L.keys_out = off; off += a256(sizeof(unsigned) * (size_t)total);
  L.vals_in = off; off += a256(sizeof(int) * (size_t)total);
  L.vals_out = off; off += a256(sizeof(int) * (size_t)total);
  L.seg_start = off; off += a256(sizeof(int) * (size_t)B * Ntgt);
  L.seg_end = off; off += a256(sizeof(int) * (size_t)B * Ntgt);
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned *)nullptr, (unsigned *)nullptr, (const int *)nullptr,
                                  (int *)nullptr, (int)total, 0, end_bit_for((unsigned long long)B * Ntgt));
  L.cub = off; L.cub_bytes = temp; off += a256(temp);
  L.total_bytes = off;
  return L;
}

// E entries per scene pointing into [0, Ntgt); per_src entries share one grad_out column (3 for three_interpolate)
static int det_scatter(const char *what, int B, int C, int Ntgt, int E, int per_src, const float *grad_out,
                       const int32_t *idx, const float *weight, float *grad_points, void *workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  B200_CHECK_ARG(B >= 0 && C >= 0 && Ntgt >= 0 && E >= 0, "%s: bad sizes", what);
  if (B == 0 || C == 0 || Ntgt == 0) return 0;
  B200_CHECK_ARG(grad_points, "%s: null pointer", what);
  if (E == 0) {
    B200_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * Ntgt, stream));
    return 0;
  }
  B200_CHECK_ARG(grad_out && idx, "%s: null pointer", what);
  B200_CHECK_ARG(B <= 65535, "%s: B=%d exceeds grid.z", what, B);
  const long long total = (long long)B * E;
  B200_CHECK_ARG(total < (1ll << 31) && (unsigned long long)B * Ntgt < (1ull << 32), "%s: problem too large", what);
  const DetLayout L = det_layout(B, Ntgt, total);
  B200_CHECK_ARG(workspace && workspace_bytes >= L.total_bytes, "%s: workspace too small (%zu < %zu bytes)", what,
                 workspace_bytes, L.total_bytes);
  char *ws = (char *)workspace;
  unsigned *keys_in = (unsigned *)(ws + L.keys_in), *keys_out = (unsigned *)(ws + L.keys_out);
  int *vals_in = (int *)(ws + L.vals_in), *vals_out = (int *)(ws + L.vals_out);
  int *seg_start = (int *)(ws + L.seg_start), *seg_end = (int *)(ws + L.seg_end);
  const unsigned blocks = (unsigned)((total + SD_THREADS - 1) / SD_THREADS);
  det_key_kernel<<<blocks, SD_THREADS, 0, stream>>>(total, E, Ntgt, idx, keys_in, vals_in);
  B200_LAUNCH_OK("det_key_kernel");
  size_t temp = L.cub_bytes;
  B200_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + L.cub, temp, keys_in, keys_out, vals_in, vals_out, (int)total, 0,
                                               end_bit_for((unsigned long long)B * Ntgt), stream));
  count_launch(3);
  B200_CUDA_OK(cudaMemsetAsync(seg_start, 0, (L.cub - L.seg_start), stream));  // seg_start and seg_end are adjacent
  det_bounds_kernel<<<blocks, SD_THREADS, 0, stream>>>(total, (unsigned)B * (unsigned)Ntgt, keys_out, seg_start, seg_end);
  B200_LAUNCH_OK("det_bounds_kernel");
  dim3 grid(ceil_div(Ntgt, SD_THREADS), ceil_div(C, SD_CCHUNK), B);
  if (weight)
    det_reduce_kernel<true><<<grid, SD_THREADS, 0, stream>>>(C, Ntgt, E, per_src, grad_out, weight, vals_out, seg_start,
                                                            seg_end, grad_points);
  else
    det_reduce_kernel<false><<<grid, SD_THREADS, 0, stream>>>(C, Ntgt, E, per_src, grad_out, nullptr, vals_out,
                                                             seg_start, seg_end, grad_points);
  B200_LAUNCH_OK("det_reduce_kernel");
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200pn2_scatter_det_workspace(int B, int n_targets, int entries_per_scene) {
  if (B <= 0 || n_targets <= 0 || entries_per_scene <= 0) return 0;
  return det_layout(B, n_targets, (long long)B * entries_per_scene).total_bytes;
}

extern "C" int b200pn2_gather_points_grad_det(int B, int C, int N, int m, const float *grad_out, const int32_t *idx,
                                              float *grad_points, void *workspace, size_t workspace_bytes,
                                              b200_stream_t s) {
  return det_scatter("gather_points_grad_det", B, C, N, m, 1, grad_out, idx, nullptr, grad_points, workspace,
                     workspace_bytes, (cudaStream_t)s);
}

extern "C" int b200pn2_group_points_grad_det(int B, int C, int N, int M, int ns, const float *grad_out,
                                             const int32_t *idx, float *grad_points, void *workspace,
                                             size_t workspace_bytes, b200_stream_t s) {
  B200_CHECK_ARG(M >= 0 && ns >= 0 && (long long)M * ns < (1ll << 31), "group_points_grad_det: bad sizes");
  return det_scatter("group_points_grad_det", B, C, N, M * ns, 1, grad_out, idx, nullptr, grad_points, workspace,
                     workspace_bytes, (cudaStream_t)s);
}

extern "C" int b200pn2_three_interpolate_grad_det(int B, int C, int n, int m, const float *grad_out,
                                                  const int32_t *idx, const float *weight, float *grad_points,
                                                  void *workspace, size_t workspace_bytes, b200_stream_t s) {
  B200_CHECK_ARG(n >= 0 && (long long)n * 3 < (1ll << 31), "three_interpolate_grad_det: bad sizes");
  B200_CHECK_ARG(weight || n == 0, "three_interpolate_grad_det: null pointer");
  return det_scatter("three_interpolate_grad_det", B, C, m, n * 3, 3, grad_out, idx, weight, grad_points, workspace,
                     workspace_bytes, (cudaStream_t)s);
}
