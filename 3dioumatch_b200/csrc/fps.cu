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
#include <string.h>

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
fps_cluster_kernel(int B, int N, int m, int L, const float *__restrict__ xyz, int32_t *__restrict__ idx,
                   const int *__restrict__ redo) {
  // prefix speculation (fps_prefix_check_kernel below) already produced this scene's indices: nothing to do.  The test is
  // uniform over the cluster (one scene per cluster) and precedes every barrier.
  if (GROUPS == 1 && redo && redo[blockIdx.y] == 0) return;
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

// ---- fps_owner_kernel: value-only reduction, the owning lane identifies the winner -----------------------------------
// Second-generation register-resident kernel for the cluster shapes with many warps per SM.  fps_cluster_kernel above
// carries (value, tie key, index) through every level of the arg-max, which costs every thread a 3-instruction-per-
// point slot tournament and every warp the key arithmetic: ~160 update + ~190 reduction instructions per
// warp-iteration, issue- and ALU-pipe-bound at 16 warps per SM (profiles/r2_fps_cluster_kernel_warpstate.txt).  Here
// only the VALUE is reduced (20 FMNMX + 10 three-input FMNMX3 per thread, one redux per warp, one shared-memory hop
// per CTA); the point that owns the CTA maximum is looked up afterwards by the lanes whose running maximum equals it
// -- normally one lane of one warp; every other warp goes straight to the mbarrier wait.  Exact ties (duplicated
// points; the reference's bit-reversed tree order, see the file header) are resolved among the tied lanes / warps
// only: redux.min over their tie keys, and a named barrier that just the tied warps join.  The owning warp pushes the
// CTA's 32-byte record {value, key | x, y, z, index} to every peer with st.async (cluster) or stores it locally and
// arrives on the mbarrier (single CTA).  Measured (B=8, N=40000 -> 2048, 4 CTAs x 512 threads per scene): 2.27 -> 1.69 ms.
// Also measured and dropped: one flat exchange of per-WARP records (no CTA-level step; every thread reduces the 64
// records of the scene): 2.26 ms -- 128 st.async per CTA-iteration cost more than the barrier they replace.
// Same results as fps_cluster_kernel bit for bit (tests/test_gpu_pointops.py runs every shape of both).
struct __align__(16) FpsRecord2 {
  int v;         // float bits of the CTA's best min-distance (negative: the CTA has no candidate)
  unsigned key;  // tie key of that point
  int pad0, pad1;
  float x, y, z;
  int k;         // its index
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// maximum of N registers as a tree of three-input FMNMX3 (N = 20: 7 + 3 + 1 instructions)
template <int N>
__device__ __forceinline__ float max_tree(const float (&t)[N]) {
  if constexpr (N == 1) {
    return t[0];
  } else {
    constexpr int O = (N + 2) / 3;
    float o[O];
#pragma unroll
    for (int i = 0; i < O; ++i) {
      if (3 * i + 2 < N)
        o[i] = fmax3(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
      else if (3 * i + 1 < N)
        o[i] = fmaxf(t[3 * i], t[3 * i + 1]);
      else
        o[i] = t[3 * i];
    }
    return max_tree<O>(o);
  }
}

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_owner_kernel(int B, int N, int m, int L, const float *__restrict__ xyz, int32_t *__restrict__ idx,
                 const int *__restrict__ redo) {
  if (redo && redo[blockIdx.y] == 0) return;  // see fps_cluster_kernel
  constexpr int NWARP = THREADS / 32;
  extern __shared__ float4 s_tab[];  // [PPT][THREADS] coordinates of this CTA's points: the owner's one LDS.128
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned CS = cluster.num_blocks();  // power of two
  const unsigned rank = cluster.block_rank();
  const int b = blockIdx.y;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = (int)CS * THREADS;  // threads per scene, a power of two and a multiple of bs (the launcher checks)
  const int g = (int)rank * THREADS + tid;
  const float *pts = xyz + (size_t)b * N * 3;
  int32_t *out = idx + (size_t)b * m;

  __shared__ int s_v[2][NWARP];
  __shared__ unsigned s_key2[2][NWARP];
  __shared__ FpsRecord2 s_slot[2][16];
  __shared__ __align__(8) unsigned long long s_bar[2];

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float px[PPT], py[PPT], pz[PPT], pt[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int k = g + p * T;
    if (k < N) {
      px[p] = pts[(size_t)k * 3 + 0];
      py[p] = pts[(size_t)k * 3 + 1];
      pz[p] = pts[(size_t)k * 3 + 2];
      const float mag = sq3(px[p], py[p], pz[p]);     // sampling_gpu.cu:105
      pt[p] = ((double)mag <= 1e-3) ? -1.0f : 1e10f;  // :106 skipped points never compete; temp = 1e10 otherwise
    } else {
      px[p] = py[p] = pz[p] = 0.f;
      pt[p] = -1.0f;
    }
    s_tab[p * THREADS + tid] = make_float4(px[p], py[p], pz[p], 0.f);
  }
  if (CS > 1) cluster.sync(); else __syncthreads();  // barriers exist before any peer signals them; table visible

  const unsigned bsmask = (1u << L) - 1u;
  auto tie_key = [&](int k) -> unsigned {
    const unsigned rev = L > 0 ? (__brev((unsigned)k & bsmask) >> (32 - L)) : 0u;
    return (rev << 22) | ((unsigned)k >> L);
  };
  // remote addresses of this CTA's record slot and of the mbarrier in peer `lane`, for both parities (loop invariant)
  unsigned r_dst0 = 0u, r_dst1 = 0u, r_bar0 = 0u, r_bar1 = 0u;
  if (CS > 1 && lane < (int)CS) {
    r_dst0 = map_to_rank(smem_u32(&s_slot[0][rank]), (unsigned)lane);
    r_dst1 = map_to_rank(smem_u32(&s_slot[1][rank]), (unsigned)lane);
    r_bar0 = map_to_rank(smem_u32(&s_bar[0]), (unsigned)lane);
    r_bar1 = map_to_rank(smem_u32(&s_bar[1]), (unsigned)lane);
  }
  const float x0 = pts[0], y0 = pts[1], z0 = pts[2];
  float cx = x0, cy = y0, cz = z0;  // idx[0] = 0 (:89-92)
  if (rank == 0 && tid == 0) out[0] = 0;

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    const unsigned r_dst = par ? r_dst1 : r_dst0, r_bar = par ? r_bar1 : r_bar0;
    if (CS > 1 && tid == 0) mbar_arrive_expect_tx(&s_bar[par], CS * (unsigned)sizeof(FpsRecord2));
    // ---- distance update (packed pairs, the reference's rounding sequence) + value-only maximum ----------------------
    const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
#pragma unroll
    for (int p = 0; p + 1 < PPT; p += 2) {
      float d0, d1;
      unpack2(sqdist3_x2(pack2(px[p], px[p + 1]), pack2(py[p], py[p + 1]), pack2(pz[p], pz[p + 1]), cx2, cy2, cz2), d0, d1);
      pt[p] = fminf(d0, pt[p]);  // :108-111
      pt[p + 1] = fminf(d1, pt[p + 1]);
    }
    if (PPT & 1) {
      constexpr int p = PPT - 1;
      pt[p] = fminf(sqdist3(px[p], py[p], pz[p], cx, cy, cz), pt[p]);
    }
    const float tmax = max_tree<PPT>(pt);  // >= +0, or -1.0f when this thread has no candidate
    const int v = __float_as_int(tmax);    // signed-int order == float order on that range
    const int Vw = __reduce_max_sync(0xffffffffu, v);
    int Vc = Vw, cv = Vw;
    if (NWARP > 1) {
      if (lane == 0) s_v[par][warp] = Vw;
      __syncthreads();
      cv = lane < NWARP ? s_v[par][lane] : (int)0x80000000;
      Vc = __reduce_max_sync(0xffffffffu, cv);
    }
    // ---- the CTA's record: identified and sent by the warp that owns the maximum --------------------------------------
    if (Vc < 0) {  // no candidate in this CTA: the reference's (best = -1, besti = 0) state
      if (warp == 0) {
        if (CS > 1) {
          if (lane < (int)CS) {
            st_async_v4(r_dst, r_bar, (unsigned)Vc, 0xffffffffu, 0u, 0u);
            st_async_v4(r_dst + 16, r_bar, 0u, 0u, 0u, 0u);
          }
        } else if (lane == 0) {
          s_slot[par][0].v = Vc;
          mbar_arrive(&s_bar[par]);
        }
      }
    } else if (Vw == Vc) {
      const bool mine = (v == Vc);
      unsigned key = 0xffffffffu;
      int bk = 0;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mine) {  // ascending slot == ascending key inside a thread (T is a multiple of bs): the FIRST matching slot
        int slot = 0;
#pragma unroll
        for (int p = PPT - 1; p >= 0; --p) slot = (pt[p] == tmax) ? p : slot;
        bk = g + slot * T;
        key = tie_key(bk);
        c = s_tab[slot * THREADS + tid];
      }
      const unsigned tied = __ballot_sync(0xffffffffu, mine);
      int src = __ffs(tied) - 1;
      unsigned wkey;
      if (tied & (tied - 1u)) {  // rare: several lanes hold the maximum -> smallest key
        wkey = __reduce_min_sync(0xffffffffu, key);
        src = __ffs(__ballot_sync(0xffffffffu, key == wkey)) - 1;
      } else {
        wkey = __shfl_sync(0xffffffffu, key, src);
      }
      bool win = true;
      if (NWARP > 1) {
        const unsigned cand = __ballot_sync(0xffffffffu, cv == Vc);  // warps tied at the CTA maximum
        if (cand & (cand - 1u)) {  // rare: only the tied warps meet, on their own named barrier
          if (lane == 0) s_key2[par][warp] = wkey;
          asm volatile("bar.sync 1, %0;" ::"r"(32 * __popc(cand)) : "memory");
          const unsigned ck = ((cand >> lane) & 1u) ? s_key2[par][lane] : 0xffffffffu;
          win = (wkey == __reduce_min_sync(0xffffffffu, ck));  // keys are unique per point
        }
      }
      if (win) {
        bk = __shfl_sync(0xffffffffu, bk, src);
        c.x = __shfl_sync(0xffffffffu, c.x, src);
        c.y = __shfl_sync(0xffffffffu, c.y, src);
        c.z = __shfl_sync(0xffffffffu, c.z, src);
        if (CS > 1) {
          if (lane < (int)CS) {
            st_async_v4(r_dst, r_bar, (unsigned)Vc, wkey, 0u, 0u);
            st_async_v4(r_dst + 16, r_bar, __float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z),
                        (unsigned)bk);
          }
        } else if (lane == 0) {
          s_slot[par][0].v = Vc;
          *reinterpret_cast<float4 *>(&s_slot[par][0].x) = make_float4(c.x, c.y, c.z, __int_as_float(bk));
          mbar_arrive(&s_bar[par]);
        }
      }
    }
    mbar_wait(&s_bar[par], (unsigned)(((j - 1) >> 1) & 1));
    // ---- arg-max over the CS records (value; tie key only when values tie) -------------------------------------------
    int bvv;
    float4 w;
    if (CS > 1) {
      uint2 t = make_uint2(0x80000000u, 0xffffffffu);
      if (lane < (int)CS) t = *reinterpret_cast<const uint2 *>(&s_slot[par][lane]);
      bvv = __reduce_max_sync(0xffffffffu, (int)t.x);
      unsigned cand = __ballot_sync(0xffffffffu, (int)t.x == bvv);
      if (cand & (cand - 1u)) {
        const unsigned mk = __reduce_min_sync(0xffffffffu, (int)t.x == bvv ? t.y : 0xffffffffu);
        cand = __ballot_sync(0xffffffffu, (int)t.x == bvv && t.y == mk);
      }
      w = *reinterpret_cast<const float4 *>(&s_slot[par][__ffs(cand) - 1].x);
    } else {
      bvv = s_slot[par][0].v;
      w = *reinterpret_cast<const float4 *>(&s_slot[par][0].x);
    }
    int wk = __float_as_int(w.w);
    if (bvv < 0) {  // every candidate skipped: the reference's besti stays 0 everywhere
      wk = 0; w.x = x0; w.y = y0; w.z = z0;
    }
    cx = w.x; cy = w.y; cz = w.z;
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

#ifdef B200_DEV
// dev/fps_pruned.cu (developer library only): Morton-ordered clusters with exact spatial pruning -- a measured dead end
bool fps_pruned_wanted(int B, int N, int m);
int fps_pruned_launch(int B, int N, int m, int L, const float *xyz, int32_t *idx, cudaStream_t stream);
#endif

// ---- prefix speculation: is the input already in furthest-point order? -------------------------------------------------
// PointNet++ runs FPS hierarchically: level l+1 samples from the points level l selected, IN THE ORDER level l selected
// them (pointnet2_modules.py:238-247, backbone_module.py:97-121; the proposal module's seed FPS the same,
// proposal_module.py:98-104).  Furthest-point samples are nested: the j-th pick of level l maximises the min-distance
// over the whole cloud, it is a member of the subset, so it also maximises it over the subset -- same operands, same
// rounding sequence, same floats.  FPS of that subset therefore returns 0, 1, 2, ..., m-1 unless an exact tie or the
// origin-skip rule intervenes (ties are ordered by the level's own thread ids, which differ between levels).
// Instead of assuming it, three fully parallel kernels VERIFY it with the reference's exact semantics:
//   fps_prefix_head_kernel   : the first 128 columns, self-contained -- refutes an ordinary cloud in microseconds;
//   fps_prefix_values_kernel : V[j] = min-distance of point j to points 0..j-1 (what the reference's temp[j] holds when
//                              step j picks), -1 for a point inside the skip sphere;
//   fps_prefix_check_kernel  : every point k walks the columns j = 1..m-1 with its running min-distance and raises
//                              redo[b] if at some step it would beat point j under (value desc, tie key asc).
// No flag raised <=> at every step the reference's arg-max is point j  <=> idx = 0..m-1 exactly.  The serial kernel is
// launched right behind with the same flag and returns immediately for verified scenes; a scene that fails (any input
// not in furthest-point order fails at its first column, after a few microseconds) runs it in full.
// m x N independent distance evaluations (2 M for 2048 -> 1024) instead of a 1023-step dependent chain.
constexpr int FPS_PFX_THREADS = 128;
constexpr int FPS_PFX_MAX_N = 4096;

// One tile of columns for one point k.  `run` is temp[k] as step j0 + c sees it BEFORE pick j0 + c is included.
// Fast walk (branch-free, ~11 instructions per column): the speculated pick jj can only lose to k if run >= V[jj], which
// in a cloud that is in furthest-point order happens for exact ties only -- so the walk records "some column had
// run >= V" and nothing else.  A tile with such a column is replayed from the saved running value with the reference's
// full rule (value desc, tie key asc, pick 0 fixed): sampling_gpu.cu:64-70,113-114.  s_c[c].w of column 0 is staged as
// +inf (pick 0 is never contested), `own` = k - j0 is k's own column (its V equals `run` there by construction).
__device__ __forceinline__ bool fps_prefix_walk(const float4 *__restrict__ s_c, const unsigned *__restrict__ s_key, int n,
                                                int j0, int k, unsigned key_k, float px, float py, float pz, float &run) {
  const float run0 = run;
  const int own = k - j0;
  int ge_seen = 0;
#pragma unroll 8
  for (int c = 0; c < n; ++c) {
    const float4 s = s_c[c];
    ge_seen |= (int)(run >= s.w) & (int)(c != own);
    run = fminf(sqdist3(px, py, pz, s.x, s.y, s.z), run);  // :108-111
  }
  if (!ge_seen) return false;
  bool bad = false;
  float r = run0;
  for (int c = 0; c < n; ++c) {  // rare
    const float4 s = s_c[c];
    const int jj = j0 + c;
    const bool beats = (r > s.w) || (r == s.w && key_k < s_key[c]);
    bad = bad || (beats && jj >= 1 && k != jj);
    r = fminf(sqdist3(px, py, pz, s.x, s.y, s.z), r);
  }
  return bad;
}

__global__ void __launch_bounds__(FPS_PFX_THREADS)
fps_prefix_values_kernel(int N, int m, const float *__restrict__ xyz, float *__restrict__ V,
                         const int *__restrict__ redo) {
  __shared__ float4 s_p[2][FPS_PFX_THREADS];
  const int b = blockIdx.y, tid = threadIdx.x;
  if (redo[b]) return;  // already refuted by fps_prefix_head_kernel (uniform over the CTA)
  const int j = blockIdx.x * FPS_PFX_THREADS + tid;
  const float *pts = xyz + (size_t)b * N * 3;
  float px = 0.f, py = 0.f, pz = 0.f, run = -1.0f;
  if (j < m) {
    px = pts[(size_t)j * 3]; py = pts[(size_t)j * 3 + 1]; pz = pts[(size_t)j * 3 + 2];
    run = ((double)sq3(px, py, pz) <= 1e-3) ? -1.0f : 1e10f;  // sampling_gpu.cu:105-106, sampling.cpp:78-80
  }
  const int j_end = min(m, (blockIdx.x + 1) * FPS_PFX_THREADS);  // columns of this CTA need picks 0 .. j_end-2
  const int ntile = (j_end - 1 + FPS_PFX_THREADS - 1) / FPS_PFX_THREADS;
  auto fetch = [&](int i0) -> float4 {  // pick i0 + tid (i0 + tid < j_end - 1 <= N for every pick that is read)
    const int i = i0 + tid;
    return i < N ? make_float4(pts[(size_t)i * 3], pts[(size_t)i * 3 + 1], pts[(size_t)i * 3 + 2], 0.f)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (ntile > 0) s_p[0][tid] = fetch(0);
  __syncthreads();
  for (int t = 0; t < ntile; ++t) {
    const int i0 = t * FPS_PFX_THREADS;
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t + 1 < ntile) nxt = fetch(i0 + FPS_PFX_THREADS);  // in flight under this tile's walk
    const float4 *sp = s_p[t & 1];
    const int lim = min(FPS_PFX_THREADS, j - i0);  // picks i < j only
#pragma unroll 8
    for (int i = 0; i < lim; ++i) {
      const float4 s = sp[i];
      run = fminf(sqdist3(px, py, pz, s.x, s.y, s.z), run);  // :108-111 (x2 - x1)
    }
    if (t + 1 < ntile) s_p[(t + 1) & 1][tid] = nxt;
    __syncthreads();
  }
  if (j < m) V[(size_t)b * m + j] = run;
}

// First 128 columns only, self-contained (each CTA recomputes V[0..127] from the staged picks): an input that is NOT in
// furthest-point order -- any ordinary cloud -- is refuted here within a few microseconds, and the two kernels below
// return at once for that scene.
__global__ void __launch_bounds__(FPS_PFX_THREADS)
fps_prefix_head_kernel(int N, int m, int L, const float *__restrict__ xyz, int *__restrict__ redo) {
  __shared__ float4 s_c[FPS_PFX_THREADS];
  __shared__ unsigned s_key[FPS_PFX_THREADS];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int k = blockIdx.x * FPS_PFX_THREADS + tid;
  const float *pts = xyz + (size_t)b * N * 3;
  const unsigned bsmask = (1u << L) - 1u;
  auto tie_key = [&](int q) -> unsigned {
    const unsigned rev = L > 0 ? (__brev((unsigned)q & bsmask) >> (32 - L)) : 0u;
    return (rev << 22) | ((unsigned)q >> L);
  };
  const int n = min(FPS_PFX_THREADS, m);
  if (tid < n) {
    s_c[tid] = make_float4(pts[(size_t)tid * 3], pts[(size_t)tid * 3 + 1], pts[(size_t)tid * 3 + 2], 0.f);
    s_key[tid] = tie_key(tid);
  }
  __syncthreads();
  if (tid < n) {  // V[tid]: min-distance of pick tid to the picks before it
    const float4 me = s_c[tid];
    float run = ((double)sq3(me.x, me.y, me.z) <= 1e-3) ? -1.0f : 1e10f;
#pragma unroll 4
    for (int i = 0; i < tid; ++i) {
      const float4 s = s_c[i];
      run = fminf(sqdist3(me.x, me.y, me.z, s.x, s.y, s.z), run);
    }
    // no hazard: the loop above reads x, y, z only; .w is read after the barrier below.  Pick 0 is never contested.
    s_c[tid].w = tid == 0 ? __int_as_float(0x7f800000) : run;
  }
  __syncthreads();
  if (k >= N) return;
  const float px = pts[(size_t)k * 3], py = pts[(size_t)k * 3 + 1], pz = pts[(size_t)k * 3 + 2];
  float run = ((double)sq3(px, py, pz) <= 1e-3) ? -1.0f : 1e10f;
  if (fps_prefix_walk(s_c, s_key, n, 0, k, tie_key(k), px, py, pz, run)) redo[b] = 1;
}

__global__ void __launch_bounds__(FPS_PFX_THREADS)
fps_prefix_check_kernel(int N, int m, int L, const float *__restrict__ xyz, const float *__restrict__ V,
                        int32_t *__restrict__ idx, int *__restrict__ redo) {
  __shared__ float4 s_c[2][FPS_PFX_THREADS];  // pick j: x, y, z, V[j] (+inf for pick 0)
  __shared__ unsigned s_key[2][FPS_PFX_THREADS];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int k = blockIdx.x * FPS_PFX_THREADS + tid;
  const float *pts = xyz + (size_t)b * N * 3;
  const unsigned bsmask = (1u << L) - 1u;
  auto tie_key = [&](int q) -> unsigned {
    const unsigned rev = L > 0 ? (__brev((unsigned)q & bsmask) >> (32 - L)) : 0u;
    return (rev << 22) | ((unsigned)q >> L);
  };
  float px = 0.f, py = 0.f, pz = 0.f, run = -1.0f;
  const bool have = k < N;
  if (have) {
    px = pts[(size_t)k * 3]; py = pts[(size_t)k * 3 + 1]; pz = pts[(size_t)k * 3 + 2];
    run = ((double)sq3(px, py, pz) <= 1e-3) ? -1.0f : 1e10f;
  }
  const unsigned key_k = tie_key(k);
  if (k < m) idx[(size_t)b * m + k] = k;  // the speculated answer; the serial kernel overwrites it if the check fails
  const float inf = __int_as_float(0x7f800000);
  auto fetch = [&](int j0) -> float4 {  // column j0 + tid of the next tile: coordinates | V
    const int j = j0 + tid;
    if (j >= m) return make_float4(0.f, 0.f, 0.f, inf);
    return make_float4(pts[(size_t)j * 3], pts[(size_t)j * 3 + 1], pts[(size_t)j * 3 + 2],
                       j == 0 ? inf : V[(size_t)b * m + j]);
  };
  s_c[0][tid] = fetch(0);
  s_key[0][tid] = tie_key(tid);
  int stop = *(volatile int *)&redo[b];  // someone already refuted the speculation: everybody leaves
  int t = 0;
  for (int j0 = 0; j0 < m; j0 += FPS_PFX_THREADS, ++t) {
    if (__syncthreads_or(stop)) break;  // tile t staged (and tile t-1 consumed); uniform exit
    const int nj0 = j0 + FPS_PFX_THREADS;
    float4 nxt = make_float4(0.f, 0.f, 0.f, inf);
    if (nj0 < m) {  // next tile in flight under this tile's walk
      nxt = fetch(nj0);
      stop = *(volatile int *)&redo[b];
    }
    const int n = min(FPS_PFX_THREADS, m - j0);
    if (have && fps_prefix_walk(s_c[t & 1], s_key[t & 1], n, j0, k, key_k, px, py, pz, run)) {
      redo[b] = 1;  // benign race: every writer stores 1
      stop = 1;
    }
    if (nj0 < m) {
      s_c[(t + 1) & 1][tid] = nxt;
      s_key[(t + 1) & 1][tid] = tie_key(nj0 + tid);
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef void (*fps_fn)(int, int, int, int, const float *, int32_t *, const int *);

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

template <int THREADS>
static fps_fn pick_owner_ppt(int ppt, int *ppt_out) {
  constexpr int MAXP = THREADS >= 512 ? 20 : 32;
#define B200_FPS_CASE(P)                     \
  if (P <= MAXP && ppt <= P) {               \
    *ppt_out = P;                            \
    return fps_owner_kernel<THREADS, (P <= MAXP ? P : 1)>; \
  }
  B200_FPS_CASE(1) B200_FPS_CASE(2) B200_FPS_CASE(4) B200_FPS_CASE(6) B200_FPS_CASE(8) B200_FPS_CASE(10)
  B200_FPS_CASE(12) B200_FPS_CASE(16) B200_FPS_CASE(20) B200_FPS_CASE(24) B200_FPS_CASE(32)
#undef B200_FPS_CASE
  *ppt_out = 0;
  return nullptr;
}

static fps_fn pick_owner(int threads, int ppt, int *ppt_out) {
  switch (threads) {
    case 32: return pick_owner_ppt<32>(ppt, ppt_out);
    case 64: return pick_owner_ppt<64>(ppt, ppt_out);
    case 128: return pick_owner_ppt<128>(ppt, ppt_out);
    case 256: return pick_owner_ppt<256>(ppt, ppt_out);
    case 512: return pick_owner_ppt<512>(ppt, ppt_out);
  }
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

// tuning / test hook: force the kernel generation (0 fps_owner_kernel wherever it applies, 2 fps_cluster_kernel;
// -1 = B200_FPS_KERNEL / default: owner for cluster shapes, fps_cluster_kernel for single-CTA shapes) and the launch
// shape (0 = cost model).  Results are identical for every choice.
static std::atomic<int> g_fps_force[4] = {{-1}, {0}, {0}, {-1}};  // kernel, cluster, threads, prefix speculation
extern "C" int b200pn2_fps_force_shape(int kernel, int cluster, int threads) {
  g_fps_force[0].store(kernel, std::memory_order_relaxed);
  g_fps_force[1].store(cluster, std::memory_order_relaxed);
  g_fps_force[2].store(threads, std::memory_order_relaxed);
  return 0;
}
// prefix speculation for clouds of at most 4096 points (see fps_prefix_check_kernel): 1 on, 0 off, -1 = B200_FPS_PREFIX /
// default (on).  Results are identical either way.
extern "C" int b200pn2_fps_set_prefix_speculation(int mode) {
  const int prev = g_fps_force[3].load(std::memory_order_relaxed);
  g_fps_force[3].store(mode, std::memory_order_relaxed);
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
  // the cheapest per-iteration cost wins (model calibrated on B200, scripts/op_sweep.py fps_shapes): the update is
  // issue-bound (cycles per point per warp sharing a scheduler), each level of the arg-max adds a fixed latency.
  // B200_FPS_KERNEL = auto (default: fps_owner_kernel for cluster shapes) | owner | v1 (fps_cluster_kernel everywhere).
  static int env_cs = -1, env_threads = -1, env_min_n = 0, debug = 0, env_kernel = -1;
  if (env_cs < 0) {
    const char *ek = getenv("B200_FPS_KERNEL");
    env_kernel = !ek ? -1 : (!strcmp(ek, "v1") ? 2 : (!strcmp(ek, "owner") ? 0 : -1));
    const char *e = getenv("B200_FPS_THREADS");
    env_threads = e ? atoi(e) : 0;
    e = getenv("B200_FPS_FORCE_MIN_N");  // the two shape overrides apply to clouds of at least this many points
    env_min_n = e ? atoi(e) : 0;
    debug = getenv("B200_FPS_DEBUG") != nullptr;
    e = getenv("B200_FPS_CLUSTER");
    env_cs = e ? atoi(e) : 0;
  }
  const int f_kernel = g_fps_force[0].load(std::memory_order_relaxed), f_cs = g_fps_force[1].load(std::memory_order_relaxed),
            f_th = g_fps_force[2].load(std::memory_order_relaxed);
  const int kernel_sel = f_kernel >= 0 ? f_kernel : env_kernel;
  const int force_cs = f_cs > 0 ? f_cs : (N >= env_min_n ? env_cs : 0);
  const int force_threads = f_th > 0 ? f_th : (N >= env_min_n ? env_threads : 0);
  const int sms = num_sms();
  const bool throughput = fps_policy() == 1;
  // kernel_sel: -1 auto, 0 owner wherever it applies, 2 v1
  int best_cs = 0, best_ppt = 0, threads = 0;
  bool best_own = false;
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
      // of a thread then share k mod bs and no per-thread tie pass is needed
      if (((cs * th) % bs) != 0 && force_threads <= 0 && force_cs <= 0) continue;
      const int need = ceil_div(N, cs * th);
      int ppt = 0;
      const bool ties = ((cs * th) % bs) != 0;
      // fps_owner_kernel: threads per scene a multiple of bs (exact ties inside a thread need the keyed tournament of
      // fps_cluster_kernel); by default only for the issue-bound shape it was built for -- clusters of 512-thread CTAs
      // (measured, B=8 N=40000: 4x512 2.27 -> 1.69 ms, 8x256 1.35 -> 1.32 ms; single-CTA and 128-thread shapes are
      // latency chains and stay faster on fps_cluster_kernel, which has no slot search on the critical path)
      const bool own = kernel_sel != 2 && !ties && (kernel_sel == 0 || (cs > 1 && th >= 256));
      fps_fn fn = !own ? pick_kernel(th, need, &ppt) : pick_owner(th, need, &ppt);
      if (!fn) continue;
      const size_t smem = (own ? sizeof(float4) : 3 * sizeof(float)) * (size_t)ppt * th;
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
      double iter;
      if (own)  // value-only update; the owner's slot search, the barrier and the exchange are the fixed part
        iter = 6.5 * ppt * warps_per_sched + 60.0 + (th > 32 ? 130.0 + 6.0 * (th / 32) : 0.0) + (cs > 1 ? 780.0 : 330.0) +
               (cs > 8 ? 260.0 : 0.0);
      else
        iter = (ties ? 13.0 : 10.5) * ppt * warps_per_sched + 120.0 + (th > 32 ? 120.0 + 6.0 * (th / 32) : 0.0) +
               (cs > 1 ? 620.0 : 0.0) + (cs > 8 ? 60.0 : 0.0);  // >8: non-portable size (measured 16x128: 1160 cycles)
      // latency policy: serial chain length; throughput policy: SM-cycles per scene (cs CTAs hold an SM each)
      const double cost = throughput ? waves * iter * cs : waves * iter;
      if (cost < best_cost) {
        best_cost = cost; best_cs = cs; best_ppt = ppt; best_fn = fn; threads = th; best_own = own;
      }
    }
  }
  if (debug)
    fprintf(stderr, "[b200 fps] B=%d N=%d m=%d -> %s cluster=%d threads=%d ppt=%d (model cost %.0f)\n", B, N, m,
            best_own ? "owner" : "v1", best_cs, threads, best_ppt, best_cost);

  if (!best_fn) {
    // cloud too large for the register-resident kernel
    B200_CHECK_ARG(scratch != nullptr, "furthest_point_sampling: N=%d needs a scratch buffer of B*N floats", N);
    fps_global_kernel<<<B, 1024, 0, stream>>>(N, m, L, xyz, scratch, idx);
    B200_LAUNCH_OK("fps_global_kernel");
    return 0;
  }

  // ---- prefix speculation for the small (hierarchical) levels ------------------------------------------------------
  static int env_prefix = -1;
  if (env_prefix < 0) {
    const char *e = getenv("B200_FPS_PREFIX");
    env_prefix = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int f_prefix = g_fps_force[3].load(std::memory_order_relaxed);
  const bool speculate = (f_prefix >= 0 ? f_prefix != 0 : env_prefix != 0) && N <= FPS_PFX_MAX_N && m >= 2 && m <= N &&
                         best_fn != nullptr;
  ScratchGuard pfx;
  const int *redo = nullptr;
  if (speculate) {
    const size_t v_off = ((size_t)B * sizeof(int) + 255) & ~(size_t)255;
    B200_CUDA_OK(pfx.alloc(v_off + (size_t)B * m * sizeof(float), stream));
    int *redo_w = (int *)pfx.ptr;
    float *V = (float *)((char *)pfx.ptr + v_off);
    B200_CUDA_OK(cudaMemsetAsync(redo_w, 0, (size_t)B * sizeof(int), stream));
    fps_prefix_head_kernel<<<dim3(ceil_div(N, FPS_PFX_THREADS), B), FPS_PFX_THREADS, 0, stream>>>(N, m, L, xyz, redo_w);
    B200_LAUNCH_OK("fps_prefix_head_kernel");
    fps_prefix_values_kernel<<<dim3(ceil_div(m, FPS_PFX_THREADS), B), FPS_PFX_THREADS, 0, stream>>>(N, m, xyz, V, redo_w);
    B200_LAUNCH_OK("fps_prefix_values_kernel");
    fps_prefix_check_kernel<<<dim3(ceil_div(N, FPS_PFX_THREADS), B), FPS_PFX_THREADS, 0, stream>>>(N, m, L, xyz, V, idx, redo_w);
    B200_LAUNCH_OK("fps_prefix_check_kernel");
    redo = redo_w;
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(best_cs, B, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = (best_own ? sizeof(float4) : 3 * sizeof(float)) * (size_t)best_ppt * threads;  // coordinate table
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = best_cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, best_fn, B, N, m, L, xyz, idx, redo));
  B200_LAUNCH_OK(best_own ? "fps_owner_kernel" : "fps_cluster_kernel");
  return 0;
}
