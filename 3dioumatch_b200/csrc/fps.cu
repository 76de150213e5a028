// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel (pointnet2/_ext_src/src/sampling_gpu.cu:74-234), which runs ONE
// 512-thread block per scene and round-trips the running min-distance array through L2 on each of the
// m-1 serial iterations.
//
// B200 design: one thread-block CLUSTER per scene (up to 16 CTAs, chosen from the occupancy query).  Every
// point of the scene lives in registers (x, y, z, running min distance) for the whole kernel, so an
// iteration is: PPT fused distance updates per thread -> redux.sync arg-max in the warp -> one shared
// memory hop in the CTA -> one 32-byte DSMEM record per peer CTA -> one cluster barrier.  Nothing touches
// L2/HBM inside the chain except the 4-byte result store.
//
// Bit-exact tie order of the reference: thread t = k mod bs of the reference
// block keeps the first strict maximum over k = t, t+bs, ...; its shared-memory tree keeps the LEFT
// operand on ties, which orders equal maxima by the bit-reversed thread id.  Winner order is therefore
//     (min-dist desc, bitrev_L(k mod bs) asc, k asc),   bs = 2^L = opt_n_threads(N)
// which is reproduced here by a two-stage reduction key: value first, then the 31-bit tie key
//     (bitrev_L(k mod bs) << 22) | (k >> L).
#include <cooperative_groups.h>
#include <math.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {

// cuda_utils.h:18-24 of the reference; same double-precision expression, same libm.
static int ref_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

struct __align__(16) FpsRecord {  // what one CTA tells its peers each iteration
  int v;                          // float bits of the CTA's best min-distance (negative = no candidate)
  unsigned key;                   // tie key of that point
  int k;                          // its index
  int pad;
  float x, y, z, w;
};

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(int N, int m, int L, const float *__restrict__ xyz, int32_t *__restrict__ idx) {
  constexpr int NWARP = THREADS / 32;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CS = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (int)CS * THREADS;       // threads per scene; a multiple of bs, so k mod bs == g mod bs
  const int g = (int)rank * THREADS + tid;

  const float *pts = xyz + (size_t)b * N * 3;
  int32_t *out = idx + (size_t)b * m;

  __shared__ int s_v[2][NWARP];
  __shared__ unsigned s_key[2][NWARP];
  __shared__ int s_k[2][NWARP];
  __shared__ float s_x[2][NWARP], s_y[2][NWARP], s_z[2][NWARP];
  __shared__ FpsRecord s_slot[2][16];

  // ---- load this thread's points into registers -------------------------------------------
  float px[PPT], py[PPT], pz[PPT], pt[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int k = g + p * T;
    if (k < N) {
      px[p] = pts[(size_t)k * 3 + 0];
      py[p] = pts[(size_t)k * 3 + 1];
      pz[p] = pts[(size_t)k * 3 + 2];
      const float mag = sq3(px[p], py[p], pz[p]);  // sampling_gpu.cu:105
      // :106 `if (mag <= 1e-3) continue;` is a double compare; a skipped point never competes.
      // min-distance -1 makes fminf() pin it at -1, which can never beat the strict `>` against -1.
      pt[p] = ((double)mag <= 1e-3) ? -1.0f : 1e10f;  // sampling.cpp:78-80 temp = 1e10
    } else {
      px[p] = py[p] = pz[p] = 0.f;
      pt[p] = -1.0f;
    }
  }
  const unsigned bsmask = (1u << L) - 1u;
  const unsigned rev = L > 0 ? (__brev((unsigned)g & bsmask) >> (32 - L)) : 0u;
  const unsigned revshift = rev << 22;

  const float x0 = pts[0], y0 = pts[1], z0 = pts[2];
  float cx = x0, cy = y0, cz = z0;  // idx[0] = 0 (:89-92)
  if (rank == 0 && tid == 0) out[0] = 0;

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    // ---- distance update + per-thread strict arg-max (k ascending) ----------------------------
    float best = -1.0f;
    int bp = 0;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      const float d = sqdist3(px[p], py[p], pz[p], cx, cy, cz);  // :108-109 (x2 - x1)
      const float d2 = fminf(d, pt[p]);                           // :111
      pt[p] = d2;
      if (d2 > best) {  // :113-114
        best = d2;
        bp = p;
      }
    }
    float bx = px[0], by = py[0], bz = pz[0];
#pragma unroll
    for (int p = 1; p < PPT; ++p)
      if (bp == p) {
        bx = px[p];
        by = py[p];
        bz = pz[p];
      }
    const int bk = g + bp * T;
    const int v = __float_as_int(best);  // best >= +0 or == -1.0f: signed-int order == float order
    const unsigned key = revshift | ((unsigned)bk >> L);

    // ---- warp arg-max: value, then tie key ------------------------------------------------------
    const int wv = __reduce_max_sync(0xffffffffu, v);
    const unsigned wkey = __reduce_min_sync(0xffffffffu, v == wv ? key : 0xffffffffu);
    if (v == wv && key == wkey) {
      s_v[par][warp] = wv;
      s_key[par][warp] = wkey;
      s_k[par][warp] = bk;
      s_x[par][warp] = bx;
      s_y[par][warp] = by;
      s_z[par][warp] = bz;
    }
    __syncthreads();
    // every warp redundantly reduces the NWARP records (no second barrier needed)
    int cv = lane < NWARP ? s_v[par][lane] : (int)0x80000000;
    unsigned ckey = lane < NWARP ? s_key[par][lane] : 0xffffffffu;
    int bvv = __reduce_max_sync(0xffffffffu, cv);
    unsigned bkey = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
    int src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == bkey)) - 1;
    int wk = s_k[par][src];
    float wx = s_x[par][src], wy = s_y[par][src], wz = s_z[par][src];

    if (CS > 1) {
      // ---- one record per peer over distributed shared memory, then the cluster barrier ---------
      if (warp == 0 && lane < (int)CS) {
        FpsRecord *remote = cluster.map_shared_rank(&s_slot[par][rank], lane);
        FpsRecord r;
        r.v = bvv; r.key = bkey; r.k = wk; r.pad = 0;
        r.x = wx; r.y = wy; r.z = wz; r.w = 0.f;
        *remote = r;
      }
      cluster.sync();
      cv = lane < (int)CS ? s_slot[par][lane].v : (int)0x80000000;
      ckey = lane < (int)CS ? s_slot[par][lane].key : 0xffffffffu;
      bvv = __reduce_max_sync(0xffffffffu, cv);
      bkey = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
      src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == bkey)) - 1;
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
  if (CS > 1) cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

// ---- large-N fallback: min-distances in global scratch, one 1024-thread CTA per scene -----------
__global__ void __launch_bounds__(1024, 1)
fps_global_kernel(int N, int m, int L, const float *__restrict__ xyz, float *__restrict__ temp,
                  int32_t *__restrict__ idx) {
  constexpr int THREADS = 1024, NWARP = 32;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *pts = xyz + (size_t)b * N * 3;
  float *tmp = temp + (size_t)b * N;
  int32_t *out = idx + (size_t)b * m;
  __shared__ int s_v[2][NWARP];
  __shared__ unsigned s_key[2][NWARP];
  __shared__ int s_k[2][NWARP];
  for (int k = tid; k < N; k += THREADS) {
    const float mag = sq3(pts[(size_t)k * 3], pts[(size_t)k * 3 + 1], pts[(size_t)k * 3 + 2]);
    tmp[k] = ((double)mag <= 1e-3) ? -1.0f : 1e10f;
  }
  const unsigned bsmask = (1u << L) - 1u;
  const unsigned rev = L > 0 ? (__brev((unsigned)tid & bsmask) >> (32 - L)) : 0u;
  int old = 0;
  if (tid == 0) out[0] = 0;
  __syncthreads();
  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    const float cx = pts[(size_t)old * 3], cy = pts[(size_t)old * 3 + 1], cz = pts[(size_t)old * 3 + 2];
    float best = -1.0f;
    int bk = 0;
    for (int k = tid; k < N; k += THREADS) {
      const float d = sqdist3(pts[(size_t)k * 3], pts[(size_t)k * 3 + 1], pts[(size_t)k * 3 + 2], cx, cy, cz);
      const float d2 = fminf(d, tmp[k]);
      tmp[k] = d2;
      if (d2 > best) { best = d2; bk = k; }
    }
    const int v = __float_as_int(best);
    const unsigned key = (rev << 22) | ((unsigned)bk >> L);
    const int wv = __reduce_max_sync(0xffffffffu, v);
    const unsigned wkey = __reduce_min_sync(0xffffffffu, v == wv ? key : 0xffffffffu);
    if (v == wv && key == wkey) { s_v[par][warp] = wv; s_key[par][warp] = wkey; s_k[par][warp] = bk; }
    __syncthreads();
    const int cv = s_v[par][lane];
    const unsigned ckey = s_key[par][lane];
    const int bvv = __reduce_max_sync(0xffffffffu, cv);
    const unsigned bkey = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
    const int src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == bkey)) - 1;
    old = bvv < 0 ? 0 : s_k[par][src];
    if (tid == 0) out[j] = old;
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef void (*fps_fn)(int, int, int, const float *, int32_t *);

template <int THREADS>
static fps_fn pick_ppt(int ppt, int *ppt_out) {
#define B200_FPS_CASE(P)                 \
  if (ppt <= P) {                        \
    *ppt_out = P;                        \
    return fps_cluster_kernel<THREADS, P>; \
  }
  B200_FPS_CASE(1) B200_FPS_CASE(2) B200_FPS_CASE(3) B200_FPS_CASE(4) B200_FPS_CASE(5) B200_FPS_CASE(6)
  B200_FPS_CASE(8) B200_FPS_CASE(12) B200_FPS_CASE(16)
#undef B200_FPS_CASE
  *ppt_out = 0;
  return nullptr;
}

static int max_clusters(fps_fn fn, int threads, int cs) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (void *)fn, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace b200

using namespace b200;

extern "C" int b200pn2_furthest_point_sampling(int B, int N, int m, const float *xyz, int32_t *idx, float *scratch,
                                               b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(B >= 0 && N > 0 && m >= 0, "furthest_point_sampling: bad sizes B=%d N=%d m=%d", B, N, m);
  B200_CHECK_ARG(xyz && idx, "furthest_point_sampling: null pointer");
  if (B == 0 || m == 0) return 0;
  const int bs = ref_opt_n_threads(N);
  int L = 0;
  while ((1 << L) < bs) ++L;

  // thread count per CTA: a multiple of bs(<=512); 512 for small clouds, 1024 otherwise
  static int force_cs = -1, force_threads = -1;
  if (force_cs < 0) {
    const char *e = getenv("B200_FPS_CLUSTER");
    force_cs = e ? atoi(e) : 0;
    e = getenv("B200_FPS_THREADS");
    force_threads = e ? atoi(e) : 0;
  }
  int threads = (N <= 1024) ? 512 : 1024;
  if (force_threads == 512 || force_threads == 1024) threads = force_threads;

  // cluster size: the largest of {16,8,4,2,1} that (a) still lets all B scenes be co-resident (one wave) when
  // possible and (b) holds the cloud in registers (<= 16 points per thread); small clouds stay in one CTA.
  const int sms = num_sms();
  int best_cs = 0, best_ppt = 0;
  fps_fn best_fn = nullptr;
  double best_cost = 1e300;
  const int cs_list[5] = {16, 8, 4, 2, 1};
  for (int ci = 0; ci < 5; ++ci) {
    const int cs = cs_list[ci];
    if (force_cs > 0 && cs != force_cs) continue;
    if (cs > 1 && N < cs * threads) continue;  // do not spread fewer than one point per thread
    const int need = ceil_div(N, cs * threads);
    if (need > 16) continue;
    int ppt = 0;
    fps_fn fn = threads == 512 ? pick_ppt<512>(need, &ppt) : pick_ppt<1024>(need, &ppt);
    if (!fn) continue;
    if (cs > 8) {
      if (cudaFuncSetAttribute((void *)fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
    }
    int conc = cs == 1 ? sms : max_clusters(fn, threads, cs);
    if (conc <= 0) continue;
    const int waves = ceil_div(B, conc);
    // per-iteration cost model (cycles): update + block reduce (+ cluster exchange)
    const double iter = 14.0 * ppt * (threads / 128) + 260.0 + (cs > 1 ? 700.0 : 0.0);
    const double cost = waves * iter;
    if (cost < best_cost) {
      best_cost = cost; best_cs = cs; best_ppt = ppt; best_fn = fn;
    }
  }
  (void)best_ppt;

  if (!best_fn) {
    // cloud too large for the register-resident kernel
    B200_CHECK_ARG(scratch != nullptr, "furthest_point_sampling: N=%d needs a scratch buffer of B*N floats", N);
    fps_global_kernel<<<B, 1024, 0, stream>>>(N, m, L, xyz, scratch, idx);
    B200_LAUNCH_OK("fps_global_kernel");
    return 0;
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(best_cs, B, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = best_cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, best_fn, N, m, L, xyz, idx));
  B200_LAUNCH_OK("fps_cluster_kernel");
  return 0;
}
