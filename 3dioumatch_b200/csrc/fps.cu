// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel (pointnet2/_ext_src/src/sampling_gpu.cu:74-234), which runs ONE
// 512-thread block per scene and round-trips the running min-distance array through L2 on each of the
// m-1 serial iterations.
//
// B200 design: one thread-block CLUSTER per scene (up to 16 CTAs, chosen from the occupancy query).  Every
// point of the scene lives in registers (x, y, z, running min distance) for the whole kernel, so an
// iteration is: PPT fused distance updates per thread -> redux.sync arg-max in the warp -> one shared
// memory hop in the CTA -> one 32-byte DSMEM record per peer CTA, pushed with st.async and signalled through
// the peer's mbarrier (no cluster-wide barrier inside the loop).  Nothing touches L2/HBM inside the chain
// except the 4-byte result store.
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

#ifdef B200_FPS_PROFILE
__device__ unsigned long long g_fps_prof[8];
#define FPS_TICK(i)                                   \
  do {                                                \
    const long long _t = clock64();                   \
    prof[i] += (unsigned long long)(_t - tprev);      \
    tprev = _t;                                       \
  } while (0)
#else
#define FPS_TICK(i)
#endif

struct __align__(16) FpsRecord {  // what one CTA tells its peers each iteration (2 x 16 B st.async)
  int v;                          // float bits of the CTA's best min-distance (negative = no candidate)
  unsigned key;                   // tie key of that point
  int k;                          // its index
  int pad;
  float x, y, z, w;
};

// ---- mbarrier / DSMEM primitives (PTX ISA 8.x, sm_90+) ---------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "FPS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra FPS_DONE;\n"
      "bra FPS_WAIT;\n"
      "FPS_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ unsigned map_to_rank(unsigned local_smem_addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// 16-byte remote store that also completes 16 tx-bytes on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_v4(unsigned remote_addr, unsigned remote_bar, unsigned a, unsigned b,
                                            unsigned c, unsigned d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}

// One thread-block cluster per scene; each thread keeps PPT points (x, y, z, running min distance) in registers,
// and the CTA keeps a copy of its coordinates in shared memory so that a winner's xyz is one indexed LDS.
// Per iteration:  PPT branch-free distance updates -> warp arg-max (2 x redux.sync) -> CTA arg-max through shared
// memory -> [cluster] each CTA pushes its 32-byte record into every peer's shared memory with st.async, which also
// signals the peer's mbarrier (complete_tx); every thread waits on its own CTA's mbarrier.  No cluster-wide barrier.
// GROUPS = 2: the CTA is two independent 256-thread groups, each working on its OWN scene (own registers, coordinate
// table, records, mbarriers, and a named barrier instead of __syncthreads).  An iteration is half issue-bound update and
// half reduction/exchange latency; two scenes on one SM fill each other's latency, which two co-resident CTAs would too
// -- but the block scheduler spreads CTAs over free SMs first, so only a single CTA guarantees the sharing.
template <int THREADS, int PPT, int MINB = 1, int GROUPS = 1>
__global__ void __launch_bounds__(THREADS, MINB)
fps_cluster_kernel(int B, int N, int m, int L, const float *__restrict__ xyz, int32_t *__restrict__ idx) {
  constexpr int GT = THREADS / GROUPS;  // threads per scene in this CTA
  constexpr int NWARP = GT / 32;
  extern __shared__ float s_xyz_all[];  // [GROUPS][PPT][GT][3]
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CS = cluster.num_blocks();  // power of two
  const unsigned rank = cluster.block_rank();
  const int grp = GROUPS > 1 ? (int)threadIdx.x / GT : 0;
  const int b = blockIdx.y * GROUPS + grp;
  const bool live = b < B;  // an odd batch leaves the last CTA's second group without a scene
  const int tid = GROUPS > 1 ? (int)threadIdx.x % GT : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (int)CS * GT;  // threads per scene, a power of two
  const int log2T = 31 - __clz(T);
  const int g = (int)rank * GT + tid;
  float *s_xyz = s_xyz_all + (size_t)grp * PPT * GT * 3;

  const float *pts = xyz + (size_t)(live ? b : 0) * N * 3;
  int32_t *out = idx + (size_t)(live ? b : 0) * m;

  __shared__ int s_v_all[GROUPS][2][NWARP];
  __shared__ unsigned s_key_all[GROUPS][2][NWARP];
  __shared__ int s_k_all[GROUPS][2][NWARP];
  __shared__ FpsRecord s_slot_all[GROUPS][2][16];
  __shared__ __align__(8) unsigned long long s_bar_all[GROUPS][2];
  int (*s_v)[NWARP] = s_v_all[grp];
  unsigned (*s_key)[NWARP] = s_key_all[grp];
  int (*s_k)[NWARP] = s_k_all[grp];
  FpsRecord (*s_slot)[16] = s_slot_all[grp];
  unsigned long long *s_bar = s_bar_all[grp];
  auto group_sync = [&]() {
    if (GROUPS > 1)
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(GT) : "memory");
    else
      __syncthreads();
  };

  if (CS > 1) {
    if (tid == 0) {
      mbar_init(&s_bar[0], 1);
      mbar_init(&s_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();  // every CTA's barriers exist before any peer signals them
  }

  // ---- load this thread's points into registers (+ the shared-memory coordinate table) ----------------
  float px[PPT], py[PPT], pz[PPT], pt[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int k = g + p * T;
    if (k < N && live) {
      px[p] = pts[(size_t)k * 3 + 0];
      py[p] = pts[(size_t)k * 3 + 1];
      pz[p] = pts[(size_t)k * 3 + 2];
      const float mag = sq3(px[p], py[p], pz[p]);  // sampling_gpu.cu:105
      // :106 `if (mag <= 1e-3) continue;` is a double compare; a skipped point never competes.
      // min-distance -1 makes fminf() pin it at -1, which can never beat a real candidate (>= 0).
      pt[p] = ((double)mag <= 1e-3) ? -1.0f : 1e10f;  // sampling.cpp:78-80 temp = 1e10
    } else {
      px[p] = py[p] = pz[p] = 0.f;
      pt[p] = -1.0f;
    }
    float *sp = s_xyz + (size_t)(p * GT + tid) * 3;
    sp[0] = px[p]; sp[1] = py[p]; sp[2] = pz[p];
  }
  __syncthreads();
  const unsigned bsmask = (1u << L) - 1u;
  // tie key of point k: (bitrev_L(k mod bs) << 22) | (k >> L)   -- smaller key wins among equal distances
  auto tie_key = [&](int k) -> unsigned {
    const unsigned rev = L > 0 ? (__brev((unsigned)k & bsmask) >> (32 - L)) : 0u;
    return (rev << 22) | ((unsigned)k >> L);
  };
  // If T is a multiple of bs, all points of a thread share (k mod bs) and ascending slot == ascending key, so the
  // first strict maximum is already the reference's choice; otherwise exact ties inside a thread need the keys.
  const bool thread_ties = ((unsigned)T & bsmask) != 0u;
  auto coords_of = [&](int k, float &x, float &y, float &z) {  // k owned by this CTA
    const int p = k >> log2T, t = (k & (T - 1)) - (int)rank * GT;
    const float *sp = s_xyz + (size_t)(p * GT + t) * 3;
    x = sp[0]; y = sp[1]; z = sp[2];
  };

  const float x0 = pts[0], y0 = pts[1], z0 = pts[2];
  float cx = x0, cy = y0, cz = z0;  // idx[0] = 0 (:89-92)
  if (rank == 0 && tid == 0 && live) out[0] = 0;

#ifdef B200_FPS_PROFILE
  unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#endif
  for (int j = 1; j < m && live; ++j) {
    const int par = j & 1;
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[par], CS * (unsigned)sizeof(FpsRecord));
    // ---- distance update + per-thread arg-max (branch-free; first strict maximum in slot order) ----------
    // All PPT updates are independent; the arg-max is a tournament tree (depth log2 PPT instead of a PPT-long
    // dependent chain).  The left (lower-slot) entry survives ties, i.e. "first strict maximum in slot order".
    float tv[PPT];
    int ts[PPT];
    const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
#pragma unroll
    for (int p = 0; p + 1 < PPT; p += 2) {
      // two points per FADD2 / FMUL2 / FFMA2 (same per-element rounding as sqdist3): the update is issue-bound
      float d0, d1;
      unpack2(sqdist3_x2(pack2(px[p], px[p + 1]), pack2(py[p], py[p + 1]), pack2(pz[p], pz[p + 1]), cx2, cy2, cz2), d0, d1);
      pt[p] = fminf(d0, pt[p]);              // :108-111 (x2 - x1), min with the running distance
      pt[p + 1] = fminf(d1, pt[p + 1]);
      tv[p] = pt[p]; tv[p + 1] = pt[p + 1];
      ts[p] = p; ts[p + 1] = p + 1;
    }
    if (PPT & 1) {
      constexpr int p = PPT - 1;
      const float d = sqdist3(px[p], py[p], pz[p], cx, cy, cz);
      pt[p] = fminf(d, pt[p]);
      tv[p] = pt[p];
      ts[p] = p;
    }
#pragma unroll
    for (int s = 1; s < PPT; s *= 2) {
#pragma unroll
      for (int i = 0; i + s < PPT; i += 2 * s) {
        const bool gt = tv[i + s] > tv[i];  // :113-114 strict
        tv[i] = gt ? tv[i + s] : tv[i];
        ts[i] = gt ? ts[i + s] : ts[i];
      }
    }
    // a thread without candidates keeps the reference's (best = -1, besti = 0) state
    float best = tv[0] > -1.0f ? tv[0] : -1.0f;
    int bp = tv[0] > -1.0f ? ts[0] : 0;
    if (thread_ties) {  // uniform branch
      int same = 0;
#pragma unroll
      for (int p = 0; p < PPT; ++p) same += (pt[p] == best) ? 1 : 0;
      if (same > 1 && best >= 0.f) {  // rare: exact tie inside this thread -> smallest key
        unsigned bkey_t = 0xffffffffu;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
          const unsigned kp = tie_key(g + p * T);
          if (pt[p] == best && kp < bkey_t) {
            bkey_t = kp;
            bp = p;
          }
        }
      }
    }
    const int bk = g + bp * T;
    const int v = __float_as_int(best);  // best >= +0 or == -1.0f: signed-int order == float order
    const unsigned key = tie_key(bk);
    FPS_TICK(0);  // distance update

    // ---- warp arg-max: value, then tie key ------------------------------------------------------
    int bvv = __reduce_max_sync(0xffffffffu, v);
    unsigned bkey = __reduce_min_sync(0xffffffffu, v == bvv ? key : 0xffffffffu);
    int wk;
    FPS_TICK(1);  // warp arg-max
    if (NWARP > 1) {
      if (v == bvv && key == bkey) {
        s_v[par][warp] = bvv;
        s_key[par][warp] = bkey;
        s_k[par][warp] = bk;
      }
      group_sync();
      // every warp redundantly reduces the NWARP records (no second barrier needed)
      const int cv = lane < NWARP ? s_v[par][lane] : (int)0x80000000;
      const unsigned ckey = lane < NWARP ? s_key[par][lane] : 0xffffffffu;
      bvv = __reduce_max_sync(0xffffffffu, cv);
      bkey = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
      const int src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == bkey)) - 1;
      wk = s_k[par][src];
    } else {
      const int src = __ffs(__ballot_sync(0xffffffffu, v == bvv && key == bkey)) - 1;
      wk = __shfl_sync(0xffffffffu, bk, src);
    }
    float wx, wy, wz;
    FPS_TICK(2);  // CTA arg-max
    if (CS > 1) {
      // ---- one 32-byte record per peer over distributed shared memory, signalled through its mbarrier ----
      if (warp == 0) {
        coords_of(wk, wx, wy, wz);
        if (lane < (int)CS) {
          const unsigned dst = map_to_rank(smem_u32(&s_slot[par][rank]), (unsigned)lane);
          const unsigned bar = map_to_rank(smem_u32(&s_bar[par]), (unsigned)lane);
          st_async_v4(dst, bar, (unsigned)bvv, bkey, (unsigned)wk, 0u);
          st_async_v4(dst + 16, bar, __float_as_uint(wx), __float_as_uint(wy), __float_as_uint(wz), 0u);
        }
      }
      FPS_TICK(3);  // st.async issue
      mbar_wait(&s_bar[par], (unsigned)(((j - 1) >> 1) & 1));
      FPS_TICK(4);  // wait for the peers' records
      const int cv = lane < (int)CS ? s_slot[par][lane].v : (int)0x80000000;
      const unsigned ckey = lane < (int)CS ? s_slot[par][lane].key : 0xffffffffu;
      bvv = __reduce_max_sync(0xffffffffu, cv);
      bkey = __reduce_min_sync(0xffffffffu, cv == bvv ? ckey : 0xffffffffu);
      const int src = __ffs(__ballot_sync(0xffffffffu, cv == bvv && ckey == bkey)) - 1;
      wk = s_slot[par][src].k;
      wx = s_slot[par][src].x;
      wy = s_slot[par][src].y;
      wz = s_slot[par][src].z;
    } else {
      coords_of(wk, wx, wy, wz);
    }
    if (bvv < 0) {  // every candidate skipped: the reference's besti stays 0 everywhere
      wk = 0; wx = x0; wy = y0; wz = z0;
    }
    cx = wx; cy = wy; cz = wz;
    if (rank == 0 && tid == 0) out[j] = wk;  // :175-176
    FPS_TICK(5);  // cluster arg-max + bookkeeping
  }
#ifdef B200_FPS_PROFILE
  if (b == 0 && rank == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(&g_fps_prof[i], prof[i]);
#endif
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

#ifdef B200_DEV
// dev/fps_pruned.cu (developer library only): Morton-ordered clusters with exact spatial pruning -- a measured dead end
bool fps_pruned_wanted(int B, int N, int m);
int fps_pruned_launch(int B, int N, int m, int L, const float *xyz, int32_t *idx, cudaStream_t stream);
#endif

// ---- host side -----------------------------------------------------------------------------------
typedef void (*fps_fn)(int, int, int, int, const float *, int32_t *);

template <int THREADS, int MINB = 1>
static fps_fn pick_ppt(int ppt, int *ppt_out) {
  constexpr int MAXP = THREADS * MINB >= 512 ? 20 : 32;  // register budget: 4 registers per resident point
#define B200_FPS_CASE(P)                     \
  if (P <= MAXP && ppt <= P) {               \
    *ppt_out = P;                            \
    return fps_cluster_kernel<THREADS, (P <= MAXP ? P : 1), MINB>; \
  }
  B200_FPS_CASE(1) B200_FPS_CASE(2) B200_FPS_CASE(4) B200_FPS_CASE(6) B200_FPS_CASE(8) B200_FPS_CASE(10)
  B200_FPS_CASE(12) B200_FPS_CASE(16) B200_FPS_CASE(20) B200_FPS_CASE(24) B200_FPS_CASE(32)
#undef B200_FPS_CASE
  *ppt_out = 0;
  return nullptr;
}

// two scenes per 512-thread CTA (256 threads each; fps_cluster_kernel<512, P, 1, 2>)
static fps_fn pick_grouped(int ppt, int *ppt_out) {
#define B200_FPS_CASE(P)                      \
  if (ppt <= P) {                             \
    *ppt_out = P;                             \
    return fps_cluster_kernel<512, P, 1, 2>;  \
  }
  B200_FPS_CASE(8) B200_FPS_CASE(12) B200_FPS_CASE(16) B200_FPS_CASE(20)
#undef B200_FPS_CASE
  *ppt_out = 0;
  return nullptr;
}

static fps_fn pick_kernel(int threads, int ppt, int *ppt_out) {
  switch (threads) {
    case 32: return pick_ppt<32>(ppt, ppt_out);
    case 64: return pick_ppt<64>(ppt, ppt_out);
    case 128: return pick_ppt<128>(ppt, ppt_out);
    case 256: return pick_ppt<256>(ppt, ppt_out);
    case 512: return pick_ppt<512>(ppt, ppt_out);
  }
  *ppt_out = 0;
  return nullptr;
}

static int max_clusters(fps_fn fn, int threads, int cs, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
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

#ifdef B200_FPS_PROFILE
// developer hook (not part of the ABI): per-stage cycle totals of scene 0 / rank 0 / thread 0, then reset
extern "C" int b200_debug_fps_profile(unsigned long long *out8) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_fps_prof, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_fps_prof, z, sizeof(z));
  return 0;
}
#endif

static std::atomic<int> g_fps_policy{-1};  // -1: not initialised (B200_FPS_POLICY, default 0 = latency)
static int fps_policy() {
  int v = g_fps_policy.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = getenv("B200_FPS_POLICY");
    v = (e && atoi(e) == 1) ? 1 : 0;
    g_fps_policy.store(v, std::memory_order_relaxed);
  }
  return v;
}
extern "C" int b200pn2_fps_set_policy(int policy) {
  const int prev = fps_policy();
  g_fps_policy.store(policy == 1 ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

extern "C" int b200pn2_furthest_point_sampling(int B, int N, int m, const float *xyz, int32_t *idx, float *scratch,
                                               b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(B >= 0 && N > 0 && m >= 0, "furthest_point_sampling: bad sizes B=%d N=%d m=%d", B, N, m);
  B200_CHECK_ARG(xyz && idx, "furthest_point_sampling: null pointer");
  if (B == 0 || m == 0) return 0;
  const int bs = ref_opt_n_threads(N);
  int L = 0;
  while ((1 << L) < bs) ++L;

#ifdef B200_DEV
  if (fps_pruned_wanted(B, N, m)) {
    const int rc = fps_pruned_launch(B, N, m, L, xyz, idx, stream);
    if (rc >= 0) return rc;  // -1: shape not handled there
  }
#endif
  // Launch shape: (cluster size, threads per CTA, points per thread).  Candidates must hold the cloud in registers;
  // the cheapest per-iteration cost wins (model calibrated on B200, scripts/op_sweep.py): the update is issue-bound
  // (~9 cycles per point per warp sharing a scheduler), each level of the arg-max tree adds a fixed latency.
  static int env_cs = -1, env_threads = -1, env_min_n = 0, debug = 0, env_pack = 0, env_groups = -1;
  if (env_cs < 0) {
    const char *eg = getenv("B200_FPS_GROUPS");  // 1: allow two scenes per CTA (fps_cluster_kernel<512, P, 1, 2>)
    env_groups = eg ? atoi(eg) : 0;
    const char *ep = getenv("B200_FPS_PACK");  // 1: two CTAs per SM for the large-cloud shapes (pick_kernel2)
    env_pack = ep ? atoi(ep) : 0;
    const char *e = getenv("B200_FPS_CLUSTER");
    env_cs = e ? atoi(e) : 0;
    e = getenv("B200_FPS_THREADS");
    env_threads = e ? atoi(e) : 0;
    e = getenv("B200_FPS_FORCE_MIN_N");  // the two overrides above apply to clouds of at least this many points
    env_min_n = e ? atoi(e) : 0;
    debug = getenv("B200_FPS_DEBUG") != nullptr;
  }
  const int force_cs = N >= env_min_n ? env_cs : 0, force_threads = N >= env_min_n ? env_threads : 0;
  const int sms = num_sms();
  const bool throughput = fps_policy() == 1;
  int best_cs = 0, best_ppt = 0, threads = 0, best_groups = 1;
  fps_fn best_fn = nullptr;
  double best_cost = 1e300;
  const int cs_list[5] = {1, 2, 4, 8, 16};
  const int th_list[5] = {32, 64, 128, 256, 512};
  for (int ci = 0; ci < 5; ++ci) {
    const int cs = cs_list[ci];
    if (force_cs > 0 && cs != force_cs) continue;
    for (int ti = 0; ti < 5; ++ti) {
      const int th = th_list[ti];
      if (force_threads > 0 && th != force_threads) continue;
      if (cs > 1 && N < cs * th) continue;  // do not spread fewer than one point per thread
      // keep (threads per scene) a multiple of the reference block size whenever some candidate allows it: all points
      // of a thread then share k mod bs and the per-thread tie pass is skipped (measured ~2x cheaper update)
      if (((cs * th) % bs) != 0 && force_threads <= 0 && force_cs <= 0) continue;
      const int need = ceil_div(N, cs * th);
      int ppt = 0;
      fps_fn fn = pick_kernel(th, need, &ppt);
      if (!fn) continue;
      const size_t smem = sizeof(float) * 3 * (size_t)ppt * th;
      if (smem > 200 * 1024) continue;
      if (smem > 40 * 1024 &&
          cudaFuncSetAttribute((void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      if (cs > 8 &&
          cudaFuncSetAttribute((void *)fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      const int conc = cs == 1 ? sms : max_clusters(fn, th, cs, smem);
      if (conc <= 0) continue;
      const int waves = ceil_div(B, conc);
      const int warps_per_sched = th >= 128 ? th / 128 : 1;
      const bool ties = ((cs * th) % bs) != 0;
      const double iter = (ties ? 13.0 : 10.5) * ppt * warps_per_sched + 120.0 + (th > 32 ? 120.0 + 6.0 * (th / 32) : 0.0) +
                          (cs > 1 ? 620.0 : 0.0) + (cs > 8 ? 260.0 : 0.0);  // >8: non-portable size; also keeps half the SMs free for the feature kernels
      // latency policy: serial chain length; throughput policy: SM-cycles per scene (cs CTAs hold an SM each)
      const double cost = throughput ? waves * iter * cs : waves * iter;
      if (cost < best_cost) {
        best_cost = cost; best_cs = cs; best_ppt = ppt; best_fn = fn; threads = th; best_groups = 1;
      }
      // two scenes per CTA (256 threads each): the pair costs one CTA-iteration of ~1.25x the single-scene length
      // Measured (B=8, N=40000): 2.50 ms for the paired shape against 2.27 ms for one scene per 512-thread CTA on the
      // same 32 SMs -- at 16 warps the SM is issue-bound (~300 instructions per warp-iteration at IPC 0.6), not
      // latency-bound, so a second scene has no idle slots to fill.  Bit-exact, kept opt-in (B200_FPS_GROUPS=1).
      const bool groups_ok = env_groups == 1;
      if (groups_ok && th == 256 && B >= 2 && N >= 8192 && !ties) {
        int gppt = 0;
        fps_fn gfn = pick_grouped(need, &gppt);
        const size_t gsmem = 2 * sizeof(float) * 3 * (size_t)gppt * 256;
        if (gfn && gsmem <= 200 * 1024 &&
            cudaFuncSetAttribute((void *)gfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem) == cudaSuccess &&
            (cs <= 8 ||
             cudaFuncSetAttribute((void *)gfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess)) {
          const int gconc = cs == 1 ? sms : max_clusters(gfn, 512, cs, gsmem);
          if (gconc > 0) {
            const int gwaves = ceil_div(ceil_div(B, 2), gconc);
            const double giter = 1.25 * (10.5 * gppt * 2 + 120.0 + 120.0 + 6.0 * 8 + (cs > 1 ? 620.0 : 0.0) +
                                         (cs > 8 ? 260.0 : 0.0));
            const double gcost = throughput ? gwaves * giter * cs / 2.0 : gwaves * giter;
            if (gcost < best_cost) {
              best_cost = gcost; best_cs = cs; best_ppt = gppt; best_fn = gfn; threads = 512; best_groups = 2;
            }
          }
        } else {
          cudaGetLastError();
        }
      }
    }
  }
  if (debug)
    fprintf(stderr, "[b200 fps] B=%d N=%d m=%d -> cluster=%d threads=%d ppt=%d scenes/CTA=%d (model cost %.0f)\n", B, N, m,
            best_cs, threads, best_ppt, best_groups, best_cost);

  if (!best_fn) {
    // cloud too large for the register-resident kernel
    B200_CHECK_ARG(scratch != nullptr, "furthest_point_sampling: N=%d needs a scratch buffer of B*N floats", N);
    fps_global_kernel<<<B, 1024, 0, stream>>>(N, m, L, xyz, scratch, idx);
    B200_LAUNCH_OK("fps_global_kernel");
    return 0;
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(best_cs, ceil_div(B, best_groups), 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = sizeof(float) * 3 * (size_t)best_ppt * threads;  // both groups' coordinate tables
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = best_cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, best_fn, B, N, m, L, xyz, idx));
  B200_LAUNCH_OK("fps_cluster_kernel");
  return 0;
}
