// ball_query.cu -- radius neighbour search for sm_100a.
//
// Replaces query_ball_point_kernel (pointnet2/_ext_src/src/ball_query_gpu.cu:14-59): B thread blocks, each
// thread serially scanning all N points from L2 for its centres.
//
// B200 design: a warp owns QW centres; the CTA (8 warps = 32 centres) streams the scene's points through a
// shared-memory tile once for all 32 centres (32x reuse of every L2 byte), lanes test 32 consecutive points
// against the warp's centres, hits are ranked with ballot + popc so that the output is exactly the reference's
// "first nsample hits in ascending point index, padded with the first hit, zeros when empty".  A warp stops
// when its centres are full; the CTA stops when all warps are.  Grid = (ceil(M/32), B) -> 512 CTAs for SA1.
#include "../../include/b200_pointnet2.h"
#include "common.cuh"

namespace b200 {

constexpr int BQ_WARPS = 8;
constexpr int BQ_QW = 4;       // centres per warp
constexpr int BQ_TILE = 1024;  // points per shared-memory tile (12 KB)

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int N, int M, float radius2, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int32_t *__restrict__ idx, int *__restrict__ unit_list,
                  int *__restrict__ unit_total) {
  __shared__ float s_pts[BQ_TILE * 3];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = (blockIdx.x * BQ_WARPS + warp) * BQ_QW;
  const float *pts = xyz + (size_t)b * N * 3;
  const unsigned lt_mask = (1u << lane) - 1u;

  float cx[BQ_QW], cy[BQ_QW], cz[BQ_QW];
  int cnt[BQ_QW], first[BQ_QW];
#pragma unroll
  for (int q = 0; q < BQ_QW; ++q) {
    const int c = c0 + q;
    if (c < M) {
      const float *p = new_xyz + ((size_t)b * M + c) * 3;
      cx[q] = p[0]; cy[q] = p[1]; cz[q] = p[2];
      cnt[q] = 0;
    } else {
      cx[q] = cy[q] = cz[q] = 0.f;
      cnt[q] = nsample;  // nothing to do
    }
    first[q] = 0;
  }
  bool wdone = true;
#pragma unroll
  for (int q = 0; q < BQ_QW; ++q) wdone = wdone && (cnt[q] >= nsample);

  for (int t0 = 0; t0 < N; t0 += BQ_TILE) {
    if (__syncthreads_and(wdone)) break;  // also: previous tile fully consumed
    const int tn = min(BQ_TILE, N - t0);
    for (int i = tid; i < tn * 3; i += BQ_WARPS * 32) s_pts[i] = pts[(size_t)t0 * 3 + i];
    __syncthreads();
    if (!wdone) {
      for (int s = 0; s < tn; s += 32) {
        const int k = s + lane;
        const bool inb = k < tn;
        const float x = inb ? s_pts[k * 3 + 0] : 0.f;
        const float y = inb ? s_pts[k * 3 + 1] : 0.f;
        const float z = inb ? s_pts[k * 3 + 2] : 0.f;
#pragma unroll
        for (int q = 0; q < BQ_QW; ++q) {
          if (cnt[q] < nsample) {  // warp-uniform
            const float d2 = sqdist3(cx[q], cy[q], cz[q], x, y, z);  // ball_query_gpu.cu:36-37 (new - x)
            const bool hit = inb && (d2 < radius2);                  // :38 strict
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) {
              if (cnt[q] == 0) first[q] = t0 + s + __ffs(mask) - 1;
              const int slot = cnt[q] + __popc(mask & lt_mask);
              if (hit && slot < nsample) idx[((size_t)b * M + c0 + q) * nsample + slot] = t0 + k;
              cnt[q] += __popc(mask);
            }
          }
        }
        wdone = true;
#pragma unroll
        for (int q = 0; q < BQ_QW; ++q) wdone = wdone && (cnt[q] >= nsample);
        if (wdone) break;
      }
    }
  }
  // :39-43 the first hit pre-fills every slot; an empty ball keeps the zero-initialised row (ball_query.cpp:24-26)
#pragma unroll
  for (int q = 0; q < BQ_QW; ++q) {
    const int c = c0 + q;
    if (c < M) {
      const int have = min(cnt[q], nsample);
      for (int l = have + lane; l < nsample; l += 32) idx[((size_t)b * M + c) * nsample + l] = first[q];
      if (unit_list && lane == 0) {
        // the fused SA kernel only feeds the 16-slot units that hold distinct neighbours through the MLP (sa_tcp.cu):
        // slots [0, have) are distinct ascending indices, the rest copies of slot 0
        const int units = ((max(have, 1) - 1) >> 4) + 1;
        const int base = atomicAdd(unit_total, units);
        for (int u = 0; u < units; ++u) unit_list[base + u] = (b * M + c) * 8 + u;
      }
    }
  }
}

// ball_query_grid.cu
bool ball_query_grid_wanted(int B, int N, int M, float radius);
int ball_query_grid_launch(int B, int N, int M, float radius, int nsample, const float *new_xyz, const float *xyz,
                           int32_t *idx, cudaStream_t stream, int *unit_list, int *unit_total);

// shared with sa_fused.cu.  unit_list / unit_total (optional, device): every centre appends the 16-slot units of its
// neighbour list that hold distinct neighbours (compacted tiles of the fused SA kernel, sa_tcp.cu)
int ball_query_launch(int B, int N, int M, float radius, int nsample, const float *new_xyz, const float *xyz,
                      int32_t *idx, cudaStream_t stream, int *unit_list, int *unit_total) {
  B200_CHECK_ARG(B >= 0 && N >= 0 && M >= 0 && nsample >= 0, "ball_query: bad sizes B=%d N=%d M=%d nsample=%d", B,
                 N, M, nsample);
  if (B == 0 || M == 0 || nsample == 0) return 0;
  B200_CHECK_ARG(new_xyz && xyz && idx, "ball_query: null pointer");
  B200_CHECK_ARG(B <= 65535, "ball_query: B=%d exceeds grid.y", B);
  if (ball_query_grid_wanted(B, N, M, radius))
    return ball_query_grid_launch(B, N, M, radius, nsample, new_xyz, xyz, idx, stream, unit_list, unit_total);
  const float radius2 = radius * radius;  // ball_query_gpu.cu:27, one fp32 multiply
  dim3 grid(ceil_div(M, BQ_WARPS * BQ_QW), B);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, stream>>>(N, M, radius2, nsample, new_xyz, xyz, idx, unit_list, unit_total);
  B200_LAUNCH_OK("ball_query_kernel");
  return 0;
}

}  // namespace b200

extern "C" int b200pn2_ball_query(int B, int N, int M, float radius, int nsample, const float *new_xyz,
                                  const float *xyz, int32_t *idx, b200_stream_t stream) {
  return b200::ball_query_launch(B, N, M, radius, nsample, new_xyz, xyz, idx, (cudaStream_t)stream, nullptr, nullptr);
}
