// fps_pruned.cu -- furthest point sampling for large clouds with exact spatial pruning (sm_100a).
//
// Same contract and the same bit-exact tie order as fps.cu (reference sampling_gpu.cu:74-178), same cluster machinery
// (register-resident points, redux.sync arg-max, st.async + mbarrier exchange).  The difference is WHICH points a
// thread owns and how often it has to touch them:
//   * a pre-pass sorts every scene along a Morton (Z-order) curve; a thread owns a contiguous run of that order, i.e. a
//     spatially compact cluster of PPT points with a bounding sphere (centre, radius);
//   * when a new sample c is selected, a point's running min-distance can only change if |p - c|^2 < mind(p).  If
//     (|c - centre| - radius)^2 already exceeds the largest min-distance in the cluster, NO point of the thread changes:
//     the thread keeps its cached candidate (value, tie key, index, coordinates) and skips the update loop.
// Skipping an update that would not have changed anything leaves every min-distance, hence every selected index,
// identical to the unpruned algorithm.  After the first few hundred samples a new sample only touches the warps around
// it, so the per-iteration cost falls from "all points" to "the fixed reduction latency".
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <stdlib.h>

#include "../../../include/b200_pointnet2.h"
#include "../common.cuh"

namespace cg = cooperative_groups;

namespace b200 {

// ---- pre-pass: bounding box -> 30-bit Morton key -> stable radix sort by (scene, key) ---------------------------
__device__ __forceinline__ unsigned f2ordered(float f) {  // monotone float -> uint
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(256)
fpp_bbox_kernel(int N, int B, const float *__restrict__ xyz,
                unsigned *__restrict__ bbox /* [B*3] ordered minima, then [B*3] ordered maxima */) {
  const int b = blockIdx.y;
  const float *p = xyz + (size_t)b * N * 3;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = blockIdx.x * 256 + threadIdx.x; k < N; k += gridDim.x * 256)
    for (int d = 0; d < 3; ++d) {
      const float v = p[(size_t)k * 3 + d];
      if (v == v && fabsf(v) < 1e30f) {  // ignore NaN / inf for the box
        mn[d] = fminf(mn[d], v);
        mx[d] = fmaxf(mx[d], v);
      }
    }
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&bbox[b * 3 + d], f2ordered(mn[d]));
      atomicMax(&bbox[B * 3 + b * 3 + d], f2ordered(mx[d]));
    }
  }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256)
fpp_key_kernel(int N, int B, int total, const float *__restrict__ xyz, const unsigned *__restrict__ bbox,
               unsigned long long *__restrict__ keys, int *__restrict__ vals) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int b = i / N, k = i - b * N;
  unsigned q[3];
  for (int d = 0; d < 3; ++d) {
    const float lo = ordered2f(bbox[b * 3 + d]), hi = ordered2f(bbox[B * 3 + b * 3 + d]);
    const float v = xyz[(size_t)i * 3 + d];
    float t = (hi > lo) ? (v - lo) / (hi - lo) : 0.f;
    t = (t == t) ? fminf(fmaxf(t, 0.f), 1.f) : 0.f;
    q[d] = (unsigned)(t * 1023.f);
  }
  const unsigned morton = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
  keys[i] = ((unsigned long long)b << 32) | morton;
  vals[i] = k;
}

struct __align__(16) FppRecord {
  int v;
  unsigned key;
  int k;
  int pad;
  float x, y, z, w;
};

__device__ __forceinline__ unsigned fpp_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// One cluster per scene.  Thread g = rank * THREADS + tid owns the points at Morton positions s = p * T + g
// (T = CS * THREADS, p = 0 .. NB-1), so "block (warp, p)" = 32 consecutive positions of the curve: a compact blob
// with a bounding sphere.  All per-point state lives in shared memory ([p][tid] planes: x, y, z, running min-distance,
// original index); lane p of a warp keeps block p's summary in registers (sphere, largest min-distance, the tie key
// and lane of the point that holds it, and the squared reach beyond which a new sample cannot change the block).
// Neighbouring blocks of the curve sit in neighbouring WARPS, so the few blocks a sample does touch are spread over
// the whole cluster: the per-iteration critical path is one or two block updates instead of every point.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_pruned_kernel(int N, int m, int L, int NB, const float *__restrict__ xyz, const int *__restrict__ order,
                  int32_t *__restrict__ idx) {
  constexpr int NWARP = THREADS / 32;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CS = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (int)CS * THREADS;
  const int g = (int)rank * THREADS + tid;

  const float *pts = xyz + (size_t)b * N * 3;
  const int *ord = order + (size_t)b * N;
  int32_t *out = idx + (size_t)b * m;

  extern __shared__ __align__(16) unsigned char fpp_dyn[];
  float *s_x = (float *)fpp_dyn;              // [NB][THREADS]
  float *s_y = s_x + (size_t)NB * THREADS;
  float *s_z = s_y + (size_t)NB * THREADS;
  float *s_t = s_z + (size_t)NB * THREADS;    // running min-distance (-1 pins a skipped point)
  int *s_k = (int *)(s_t + (size_t)NB * THREADS);

  __shared__ int s_v[2][NWARP];
  __shared__ unsigned s_key[2][NWARP];
  __shared__ int s_kk[2][NWARP];
  __shared__ float s_wx[2][NWARP], s_wy[2][NWARP], s_wz[2][NWARP];
  __shared__ FppRecord s_slot[2][16];
  __shared__ __align__(8) unsigned long long s_bar[2];

  if (CS > 1) {
    if (tid == 0) {
      for (int s = 0; s < 2; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fpp_smem(&s_bar[s])), "r"(1u) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();
  }

  const unsigned bsmask = (1u << L) - 1u;
  auto tie_key = [&](int k) -> unsigned {  // reference order among equal distances: bitrev_L(k mod bs), then k
    const unsigned rev = L > 0 ? (__brev((unsigned)k & bsmask) >> (32 - L)) : 0u;
    return (rev << 22) | ((unsigned)k >> L);
  };

  // ---- load the scene slice; lane p builds the summary of block p -------------------------------------------------
  float ccx = 0.f, ccy = 0.f, ccz = 0.f, crad = 0.f;  // sphere of block `lane`
  float thr = -INFINITY;                               // squared reach; -inf: nothing in the block can ever change
  int wmv = (int)0xbf800000;                           // bits of the block's largest min-distance (-1.0f: none)
  unsigned wkey = 0xffffffffu;
  int wsrc = 0;
  for (int p = 0; p < NB; ++p) {
    const int s = p * T + g;
    float x = 0.f, y = 0.f, z = 0.f, t = -1.0f;
    int k = 0;
    if (s < N) {
      k = ord[s];
      x = pts[(size_t)k * 3 + 0];
      y = pts[(size_t)k * 3 + 1];
      z = pts[(size_t)k * 3 + 2];
      const float mag = sq3(x, y, z);                  // sampling_gpu.cu:105
      t = ((double)mag <= 1e-3) ? -1.0f : 1e10f;       // :106 (double compare) ; sampling.cpp:78-80
    }
    const int o = p * THREADS + tid;
    s_x[o] = x; s_y[o] = y; s_z[o] = z; s_t[o] = t; s_k[o] = k;
    const bool live = t > 0.f;
    // bounding box of the live points -> sphere centre ; a NaN coordinate poisons the radius (-> never pruned)
    const unsigned lox = __reduce_min_sync(0xffffffffu, live ? f2ordered(x) : 0xffffffffu);
    const unsigned loy = __reduce_min_sync(0xffffffffu, live ? f2ordered(y) : 0xffffffffu);
    const unsigned loz = __reduce_min_sync(0xffffffffu, live ? f2ordered(z) : 0xffffffffu);
    const unsigned hix = __reduce_max_sync(0xffffffffu, live ? f2ordered(x) : 0u);
    const unsigned hiy = __reduce_max_sync(0xffffffffu, live ? f2ordered(y) : 0u);
    const unsigned hiz = __reduce_max_sync(0xffffffffu, live ? f2ordered(z) : 0u);
    const bool any_live = __any_sync(0xffffffffu, live);
    const float mx = 0.5f * (ordered2f(lox) + ordered2f(hix));
    const float my = 0.5f * (ordered2f(loy) + ordered2f(hiy));
    const float mz = 0.5f * (ordered2f(loz) + ordered2f(hiz));
    float r2 = 0.f;
    bool bad = false;
    if (live) {
      const float dx = x - mx, dy = y - my, dz = z - mz;
      r2 = dx * dx + dy * dy + dz * dz;
      bad = !(r2 == r2) || r2 > 1e30f;
      if (bad) r2 = 0.f;
    }
    const float rmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(r2)));  // r2 >= 0
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (lane == p) {
      ccx = mx; ccy = my; ccz = mz;
      crad = any_bad ? INFINITY : sqrtf(rmax) * 1.0001f + 1e-6f;
      thr = any_live ? INFINITY : -INFINITY;  // +inf: the first sample touches every live block
    }
  }
  __syncthreads();

  const float x0 = pts[0], y0 = pts[1], z0 = pts[2];
  float cx = x0, cy = y0, cz = z0;  // idx[0] = 0
  if (rank == 0 && tid == 0) out[0] = 0;

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    if (CS > 1 && tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fpp_smem(&s_bar[par])),
                   "r"(CS * (unsigned)sizeof(FppRecord))
                   : "memory");
    // ---- which blocks of this warp can the new sample reach? ------------------------------------------------------
    {
      const float ddx = cx - ccx, ddy = cy - ccy, ddz = cz - ccz;
      const float dc2 = ddx * ddx + ddy * ddy + ddz * ddz;
      unsigned touched = __ballot_sync(0xffffffffu, !(dc2 > thr));  // NaN -> touched
      touched &= NB >= 32 ? 0xffffffffu : ((1u << NB) - 1u);
      while (touched) {
        const int p = __ffs(touched) - 1;
        touched &= touched - 1;
        const int o = p * THREADS + tid;
        const float t = s_t[o];
        const float d = sqdist3(s_x[o], s_y[o], s_z[o], cx, cy, cz);  // :108-109 (x2 - x1)
        const float d2 = fminf(d, t);                                 // :111
        if (__any_sync(0xffffffffu, d2 != t)) {
          s_t[o] = d2;
          const int v = __float_as_int(d2);  // d2 >= +0 or == -1.0f: integer order == float order
          const int bv = __reduce_max_sync(0xffffffffu, v);
          const unsigned key = v == bv ? tie_key(s_k[o]) : 0xffffffffu;
          const unsigned bk = __reduce_min_sync(0xffffffffu, key);
          const int src = __ffs(__ballot_sync(0xffffffffu, v == bv && key == bk)) - 1;
          if (lane == p) {
            wmv = bv; wkey = bk; wsrc = src;
            const float reach = sqrtf(__int_as_float(bv)) + crad;
            thr = bv < 0 ? -INFINITY : reach * reach * 1.001f;
          }
        }
      }
    }

    // ---- warp arg-max over the block summaries: value, then tie key ------------------------------------------------
    int bvv = __reduce_max_sync(0xffffffffu, lane < NB ? wmv : (int)0x80000000);
    unsigned wk_key = __reduce_min_sync(0xffffffffu, (lane < NB && wmv == bvv) ? wkey : 0xffffffffu);
    int wk;
    float wx, wy, wz;
    {
      const int pl = __ffs(__ballot_sync(0xffffffffu, lane < NB && wmv == bvv && wkey == wk_key)) - 1;
      const int sl = __shfl_sync(0xffffffffu, wsrc, pl);
      const int o = pl * THREADS + warp * 32 + sl;
      wk = s_k[o]; wx = s_x[o]; wy = s_y[o]; wz = s_z[o];
    }
    if (NWARP > 1) {
      if (lane == 0) {
        s_v[par][warp] = bvv;
        s_key[par][warp] = wk_key;
        s_kk[par][warp] = wk;
        s_wx[par][warp] = wx; s_wy[par][warp] = wy; s_wz[par][warp] = wz;
      }
      __syncthreads();
      const int cv = lane < NWARP ? s_v[par][lane] : (int)0x80000000;
      const unsigned ckey = lane < NWARP ? s_key[par][lane] : 0xffffffffu;
      bvv = __reduce_max_sync(0xffffffffu, cv);
      wk_key = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
      const int src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == wk_key)) - 1;
      wk = s_kk[par][src];
      wx = s_wx[par][src]; wy = s_wy[par][src]; wz = s_wz[par][src];
    }
    if (CS > 1) {
      if (warp == 0 && lane < (int)CS) {
        unsigned dst, bar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(fpp_smem(&s_slot[par][rank])), "r"((unsigned)lane));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(fpp_smem(&s_bar[par])), "r"((unsigned)lane));
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(dst), "r"((unsigned)bvv), "r"(wk_key), "r"((unsigned)wk), "r"(0u), "r"(bar) : "memory");
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                     ::"r"(dst + 16), "r"(__float_as_uint(wx)), "r"(__float_as_uint(wy)), "r"(__float_as_uint(wz)), "r"(0u),
                     "r"(bar) : "memory");
      }
      {
        const unsigned parity = (unsigned)(((j - 1) >> 1) & 1);
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "FPP_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra FPP_DONE;\n"
            "bra FPP_WAIT;\n"
            "FPP_DONE:\n"
            "}\n" ::"r"(fpp_smem(&s_bar[par])), "r"(parity)
            : "memory");
      }
      const int cv = lane < (int)CS ? s_slot[par][lane].v : (int)0x80000000;
      const unsigned ckey = lane < (int)CS ? s_slot[par][lane].key : 0xffffffffu;
      bvv = __reduce_max_sync(0xffffffffu, cv);
      wk_key = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
      const int src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == wk_key)) - 1;
      wk = s_slot[par][src].k;
      wx = s_slot[par][src].x;
      wy = s_slot[par][src].y;
      wz = s_slot[par][src].z;
    }
    if (bvv < 0) {  // every candidate skipped: the reference's besti stays 0 everywhere
      wk = 0; wx = x0; wy = y0; wz = z0;
    }
    cx = wx; cy = wy; cz = wz;
    if (rank == 0 && tid == 0) out[j] = wk;  // :175-176
  }
  if (CS > 1) cluster.sync();
}

typedef void (*fpp_fn)(int, int, int, int, const float *, const int *, int32_t *);

// Opt-in (B200_FPS_PRUNE=1).  Measured on B200 at (B,N,m) = (8,40000,2048): 2.14 ms on 4 SMs per scene against 1.40 ms
// on 8 SMs per scene for fps.cu -- the iteration is a latency chain (block test -> block update -> three arg-max
// levels), and a 32-point block of this density has a 0.5 m sphere, wider than the 0.25 m sample spacing, so every
// warp still walks two or three blocks per sample.  Exact, so kept for denser / larger clouds.
bool fps_pruned_wanted(int B, int N, int m) {
  const char *e = getenv("B200_FPS_PRUNE");
  if (!e || atoi(e) == 0 || m < 2 || B > 65535) return false;
  const char *mn = getenv("B200_FPS_PRUNE_MIN_N");  // only clouds of at least this many points (default 256)
  return N >= (mn ? atoi(mn) : 256);
}

// returns 0 on success, -1 when the shape is not handled here (the caller falls back to fps.cu), >0 on error
int fps_pruned_launch(int B, int N, int m, int L, const float *xyz, int32_t *idx, cudaStream_t stream) {
  static int force_cs = -1, force_th = -1;
  if (force_cs < 0) {
    const char *e = getenv("B200_FPS_PRUNE_CLUSTER");
    force_cs = e ? atoi(e) : 0;
    e = getenv("B200_FPS_PRUNE_THREADS");
    force_th = e ? atoi(e) : 0;
  }
  // fewest CTAs per scene whose shared memory holds the scene (20 B per point) with at most 32 blocks per warp:
  // the pruned update is cheap, so SMs are better left to the feature kernels running on the other streams
  const size_t smem_cap = 220 * 1024;
  int cs = 0, th = 0, nb = 0;
  fpp_fn fn = nullptr;
  const int cs_list[5] = {1, 2, 4, 8, 16};
  for (int ci = 0; ci < 5 && !fn; ++ci) {
    const int c = cs_list[ci];
    if (force_cs > 0 && c != force_cs) continue;
    const int t = force_th == 256 ? 256 : 512;
    const int need = ceil_div(N, c * t);
    if (need > 32 || (size_t)need * t * 20 > smem_cap) continue;
    fpp_fn f = t == 256 ? fps_pruned_kernel<256> : fps_pruned_kernel<512>;
    if (c > 8 && cudaFuncSetAttribute((void *)f, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    fn = f; cs = c; th = t; nb = need;
  }
  if (!fn) return -1;
  const size_t dyn_smem = (size_t)nb * th * 20;
  B200_CUDA_OK(cudaFuncSetAttribute((void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));

  const int total = B * N;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (const int *)nullptr, (int *)nullptr, total, 0, 48);
  // workspace: keys[2][total] u64 | vals[2][total] i32 | bbox[2][B*3] u32 | cub temp
  const size_t off_vals = 2 * (size_t)total * 8;
  const size_t off_bbox = (off_vals + 2 * (size_t)total * 4 + 255) & ~(size_t)255;
  const size_t off_cub = (off_bbox + 6 * (size_t)B * 4 + 255) & ~(size_t)255;
  const size_t ws_bytes = off_cub + cub_bytes + 256;
  char *ws = nullptr;
  B200_CUDA_OK(scratch_alloc((void **)&ws, ws_bytes, stream));
  unsigned long long *keys_in = (unsigned long long *)ws, *keys_out = keys_in + total;
  int *vals_in = (int *)(ws + off_vals), *vals_out = vals_in + total;
  unsigned *bbox = (unsigned *)(ws + off_bbox);
  void *cub_temp = (void *)(ws + off_cub);

  B200_CUDA_OK(cudaMemsetAsync(bbox, 0xff, sizeof(unsigned) * 3 * (size_t)B, stream));            // minima: largest key
  B200_CUDA_OK(cudaMemsetAsync(bbox + 3 * (size_t)B, 0, sizeof(unsigned) * 3 * (size_t)B, stream));  // maxima: smallest
  const int bx = ceil_div(N, 256 * 8) > 0 ? ceil_div(N, 256 * 8) : 1;
  fpp_bbox_kernel<<<dim3(bx, B), 256, 0, stream>>>(N, B, xyz, bbox);
  B200_LAUNCH_OK("fpp_bbox_kernel");
  fpp_key_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(N, B, total, xyz, bbox, keys_in, vals_in);
  B200_LAUNCH_OK("fpp_key_kernel");
  int bbits = 0;
  while ((1 << bbits) < B) ++bbits;
  B200_CUDA_OK(cub::DeviceRadixSort::SortPairs(cub_temp, cub_bytes, keys_in, keys_out, vals_in, vals_out, total, 0,
                                               32 + bbits, stream));
  count_launch(2 + (32 + bbits + 7) / 8);  // radix sort: histogram, scan, one sweep per 8 key bits

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs, B, 1);
  cfg.blockDim = dim3(th, 1, 1);
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, N, m, L, nb, xyz, (const int *)vals_out, idx));
  B200_LAUNCH_OK("fps_pruned_kernel");
  B200_CUDA_OK(cudaFreeAsync(ws, stream));
  return 0;
}

}  // namespace b200
