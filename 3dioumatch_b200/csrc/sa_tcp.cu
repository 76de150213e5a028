// sa_tcp.cu -- persistent, warp-specialised fused set-abstraction forward on tcgen05 tensor cores (sm_100a).
//
// Same contract and numerics as sa_tc_kernel (sa_tc.cu: gather grouped rows -> SharedMLP as split-precision TF32 GEMMs
// with fp32 accumulation in TMEM -> max over nsample; pointnet2_modules.py:215-277, pytorch_utils.py:14-39), different
// schedule.  One CTA per SM for the whole launch; tiles (128 grouped rows) come from an atomic counter, so a CTA that
// becomes resident late (the SMs are shared with the long-running FPS cluster of the next step) finds the queue
// empty instead of holding up the launch.  Sixteen warps (fourteen working; see the register reallocation in the kernel):
//   warps 0-7   epilogue : two warpgroups, alternate 32-column chunks.  TMEM -> scale/shift/ReLU -> hi/lo split ->
//                          swizzled shared memory (hidden layers); final layer: max over nsample by a butterfly
//                          transpose-reduce in registers (warp shuffles; no slab, no CTA barrier per chunk)
//   warps 8-11  producers: layer-1 gather of the NEXT tile -- index -> row -> feature global loads are in flight while
//                          the current tile computes; two k-blocks are held in registers before R1 becomes free
//   warp  12    MMA      : runs converged; one elected lane issues tcgen05.mma / tcgen05.commit, so descriptors live in
//                          uniform registers (the one-thread branch of sa_tc_kernel compiled to a 7-R2UR waterfall loop
//                          per MMA, ~150 cycles of issue per 64-cycle MMA: measured with scripts/tcp_profile.py)
//   warp  13    loader   : tile scheduler (atomicAdd + shared-memory tile ring, one tile ahead) and the weight stream
//                          (cp.async.bulk ring of up to 8 half stages, continuous across layers and tiles)
//   warps 14-15 idle     : complete the last warpgroup, whose registers (setmaxnreg.dec) go to the producers
// The kernel is compiled per row source and first-layer mode (template <MODE, PRE>): ball-query lists, three-neighbour
// blends (GridConv sampler) or plain rows (the per-point GEMM of a factorised first layer, sa_tc.cuh).
// TMEM (512 columns) is split in two 256-column regions that swap roles every tile: one holds the accumulators
// (products in [0,128), split-precision corrections in [128,256)), the other the hidden activations X_hi | X_lo, which
// the epilogue writes with tcgen05.st and the next layer's MMAs read as their A operand straight from tensor memory
// (measured: 70 cycles per 128x128x8 TF32 MMA with A in TMEM vs 86 from shared memory, scripts/tc_rate.py).  Hidden
// activations therefore never touch shared memory: the 128 KB they used to occupy is a 3-stage ring for the layer-1
// operand plus a deeper weight ring, so the gather of tile i+1 runs completely under the MMAs of tile i; and because the regions swap, the final
// epilogue of tile i drains its accumulators while tile i+1's layer-1 MMAs already fill the other region.
#include <stdlib.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"
#include "tc_common.cuh"
#include "sa_tc.cuh"

namespace b200 {

constexpr int TP_THREADS = 512;   // 16 warps: the last warpgroup is {MMA, loader, 2 idle warps that only donate registers}
constexpr int TP_EPI = 256;     // epilogue threads (warps 0-7)
constexpr int TP_PROD0 = 256;   // first producer thread (warps 8-11)
constexpr int TP_MMA_WARP = 12, TP_LOAD_WARP = 13;
constexpr int TP_TQ = 4;        // depth of the tile-id ring
constexpr int TP_MAXSLOTS = 8;  // weight ring slots (p.nslots = 3..8)
constexpr int TP_ASTAGES = 4;   // layer-1 operand ring: at most this many stages (A_hi | A_lo of one k-block each)

// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (128 lanes x 8 columns of TF32) is read from tensor memory
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
// 32 lanes x 32 columns: thread i of the warp writes lane (base lane + i), columns [col, col+32)
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

// Butterfly transpose-reduce: every lane enters with 32 column values of its own row; after STEPS exchange steps over
// the lane bits below min(nsample, 32) each lane holds the maxima, over the nsample-lane group it belongs to, of
// 32 >> STEPS columns:  column = ((lane & (min(ns,32) - 1)) << (5 - STEPS)) | i.
template <int STEPS>
__device__ __forceinline__ void transpose_max(float (&v)[32], int lane, int top_off) {
#pragma unroll
  for (int s = 0; s < STEPS; ++s) {
    const int n = 16 >> s;
    const int off = top_off >> s;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < n) {
        const float keep = upper ? v[i + n] : v[i];
        const float send = upper ? v[i] : v[i + n];
        v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
      }
    }
  }
}


// Same butterfly with addition over all 32 lanes: lane j ends with the sum over the warp's rows of column j (in v[0]).
// The order of additions is fixed by the butterfly, so the result is deterministic.
__device__ __forceinline__ void transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int n = 16 >> s;
    const int off = 16 >> s;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < n) {
        const float keep = upper ? v[i + n] : v[i];
        const float send = upper ? v[i] : v[i + n];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
}

// ---- compacted tiles: which 16-slot units of each centre hold distinct neighbours -----------------------------------
// The reference's ball query pads a neighbour list shorter than nsample with copies of its first hit
// (ball_query_gpu.cu:40-46), and max() over duplicated rows is the max over the distinct ones, so only the slots up
// to the last one that differs from slot 0 need to go through the MLP (true for any index list, padded or not).
// The unit list (unit -> centre * 8 + unit-in-centre) is built by ATOMIC APPEND: the ball-query kernels append a centre's
// units the moment its neighbour row is final (ball_query.cu, ball_query_grid.cu), sa_unit_append_kernel does the same
// for caller-provided index lists.  The order of the list -- hence which units share a tile -- varies from run to run;
// the results do not: every output row of the MLP depends on its own input row only, and a centre's units meet through
// an order-independent integer max.  total[0] = number of units (zeroed in stream order by the launcher).
__global__ void __launch_bounds__(256) sa_unit_append_kernel(int centres, int ns, const int32_t *__restrict__ idx,
                                                            int *__restrict__ unit_list, int *__restrict__ total) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;  // one warp per centre, coalesced row
  if (c >= centres) return;
  const int32_t *row = idx + (size_t)c * ns;
  const int32_t first = row[0];
  int last_diff = 0;
  for (int t0 = 0; t0 < ns; t0 += 32) {
    const int t = t0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, t < ns && row[t] != first);
    if (m) last_diff = t0 + 31 - __clz(m);
  }
  if (lane == 0) {
    const int units = (last_diff >> 4) + 1;
    const int base = atomicAdd(total, units);
    for (int u = 0; u < units; ++u) unit_list[base + u] = c * 8 + u;
  }
}

// Tensor-pipe work actually issued by sa_tcp_kernel since the last read (b200pn2_sa_tensor_work): bench.py turns it into
// executed TF32 FLOP/s for the roofline.
__device__ unsigned long long g_sa_mma_cols;

#ifdef B200_TC_PROFILE
__device__ unsigned long long g_tcp_prof[48];
#define TPW(cat, bar, par)                                \
  do {                                                    \
    const long long _w0 = clock64();                      \
    mbar_wait_wd(bar, par);                               \
    tp_acc[cat] += (unsigned long long)(clock64() - _w0); \
  } while (0)
#define TP_BEGIN() long long _r0 = clock64()
#define TP_END(cat) tp_acc[cat] += (unsigned long long)(clock64() - _r0)
#define TP_HBEGIN() long long _h0 = clock64()
#define TP_HLAP(cat)                                       \
  do {                                                     \
    const long long _n = clock64();                        \
    tp_acc[cat] += (unsigned long long)(_n - _h0);         \
    _h0 = _n;                                              \
  } while (0)
#define TP_LAP(cat)                                        \
  do {                                                     \
    const long long _n = clock64();                        \
    tp_acc[cat] += (unsigned long long)(_n - _r0);         \
    _r0 = _n;                                              \
  } while (0)
#else
#define TPW(cat, bar, par) mbar_wait_wd(bar, par)
#define TP_BEGIN()
#define TP_END(cat)
#define TP_LAP(cat)
#define TP_HBEGIN()
#define TP_HLAP(cat)
#endif

// Fused ball query (QUERY kernels): one producer warp finds, for its CPW consecutive centres, the first `ns` points of
// the staged scene (shared memory, [N][3]) inside the ball, in ascending point index -- the contract of
// query_ball_point_kernel (ball_query_gpu.cu:27-46) and the predicate of ball_query.cu (same fp32 rounding sequence).
// Four 32-point chunks are loaded per iteration (12 independent LDS) and tested against every centre; ballot + popc rank
// the hits.  list[cw * ns + t] = t-th hit of centre cw, count[cw] = min(hits, ns).
template <int CPW>
__device__ __forceinline__ void query_scan(const float *__restrict__ s_xyz, int N, int ns, float radius2,
                                           const float *__restrict__ centres, int live, int lane, int *list, int *count) {
  float qx[CPW], qy[CPW], qz[CPW];
  int cnt[CPW];
#pragma unroll
  for (int cw = 0; cw < CPW; ++cw) {
    const bool on = cw < live;
    qx[cw] = on ? centres[cw * 3 + 0] : 0.f;
    qy[cw] = on ? centres[cw * 3 + 1] : 0.f;
    qz[cw] = on ? centres[cw * 3 + 2] : 0.f;
    cnt[cw] = on ? 0 : ns;  // nothing to find
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int k0 = 0; k0 < N; k0 += 128) {
    bool full = true;
#pragma unroll
    for (int cw = 0; cw < CPW; ++cw) full = full && (cnt[cw] >= ns);
    if (full) break;
    float x[4], y[4], z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 32 + lane;
      const bool in = k < N;
      x[u] = in ? s_xyz[k * 3 + 0] : 0.f;
      y[u] = in ? s_xyz[k * 3 + 1] : 0.f;
      z[u] = in ? s_xyz[k * 3 + 2] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * 32 + lane;
#pragma unroll
      for (int cw = 0; cw < CPW; ++cw) {
        const float d2 = sqdist3(qx[cw], qy[cw], qz[cw], x[u], y[u], z[u]);  // :36-37 (new - x)
        const bool hit = k < N && d2 < radius2 && cnt[cw] < ns;             // :38 strict; :32 stop after ns hits
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt[cw] + __popc(mask & lt_mask);
        if (hit && pos < ns) list[cw * ns + pos] = k;
        cnt[cw] += __popc(mask);
      }
    }
  }
#pragma unroll
  for (int cw = 0; cw < CPW; ++cw)
    if (lane == 0) count[cw] = min(cnt[cw], ns);
}

// MODE: row source (0 ball-query lists, 1 three-neighbour blends, 2 plain row GEMM with row output); PRE: factorised first
// layer (rows of P + xyz FMAs + ReLU in the producers).  Compile-time so that each variant's producer only carries its own
// registers -- the gather is register-bound (a runtime-mode kernel with one more row source measured 20 % slower).
// ROWOUT: the final epilogue stores every row's output (point-major and/or channel-major, optional ReLU) instead of the
// max over nsample rows -- the per-point GEMM of a factorised first layer, the feature-propagation MLP and the 1x1-conv
// heads (row MLPs).
// TRAIN: one layer of a training-mode stack (plain rows in, raw conv rows out): affine + ReLU of the previous layer's
// batch-statistics BatchNorm applied to the input rows in the producers, per-tile column sums of the output (the next
// BatchNorm's statistics) in the epilogue.
// QUERY (mode 0, uncompacted tiles, nsample <= 32, N <= TC_QUERY_MAX_N): no neighbour lists come in.  The producer warps
// stage the tile's scene coordinates in shared memory with ONE cp.async.bulk (mbarrier completion) and run the radius
// search for their own rows' centres there -- first nsample hits in ascending point index, padded with the first hit,
// zeros when the ball is empty (ball_query_gpu.cu:27-46) -- straight into the row gather: ball query and grouping are
// one kernel, the (B, M, nsample) index tensor exists only if the caller asks for it (p.idx_out).
template <int MODE, int PRE, int ROWOUT, int TRAIN = 0, int QUERY = 0>
__global__ void __launch_bounds__(TP_THREADS, 1) sa_tcp_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle, by OFFSET from the __shared__ array: a round trip through uintptr_t
  // loses the address space and turns every access below into a generic LD.E / ST.E with a descriptor R2UR pair
  uint8_t *base = smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u);
  uint8_t *R1 = base;                                                        // layer-1 operand ring: TP_ASTAGES x 32 KB
  uint8_t *R2 = base + p.r1_bytes;                                           // weight ring: nslots x wslot_bytes
  float *s_scale = reinterpret_cast<float *>(R2 + p.nslots * p.wslot_bytes);  // [TC_MAXL][256]
  float *s_shift = s_scale + TC_MAXL * 256;
  float *s_partial = s_shift + TC_MAXL * 256;                                // [2][4][256] per-warp maxima (nsample > 32)
  float *s_wx = s_partial + 2 * 4 * 256;                                     // [3][128] scale1 * W1x (factorised layer 1)
  float *s_in = s_wx + 3 * 128;                                              // [4][256] input coefficients, [4][256] output (TRAIN)
  float *s_out = s_in + 4 * 256;
  float *s_qxyz = s_out + 4 * 256;                                           // [N][3] staged scene coordinates (QUERY)
  __shared__ uint64_t q_bar;
  __shared__ int s_qlist[4][32], s_qcnt[4][4];                               // per producer warp: hit lists of its centres

  __shared__ uint64_t full_a[TP_ASTAGES], empty_a[TP_ASTAGES], full_w[TP_MAXSLOTS], empty_w[TP_MAXSLOTS];
  // accum_half is indexed by (accumulator region, half): a barrier advances once per TWO tiles, and the MMA warp may not
  // start a tile before every epilogue thread has finished the tile that used its region (acc_free / the hidden hand-off),
  // so no waiter can be lapped (a per-tile barrier could complete two phases -- K = 4 layers take ~200 cycles per tile --
  // before a slow epilogue warp polled the first: its parity wait would then block forever)
  __shared__ uint64_t accum_full, x_ready, accum_half[4], d_free, acc_free[2], tq_full[TP_TQ], tq_empty[TP_TQ];
  __shared__ int tq_tile[TP_TQ];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ns = p.ns, nl = p.nl, G = p.G;
  const int nslots = p.nslots;
  const int nhalf_last = p.L[nl - 1].nhalf;

  if (warp == TP_MMA_WARP) tc::tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int s = 0; s < TP_ASTAGES; ++s) {
      tc::mbar_init(&full_a[s], 128);
      tc::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < 4; ++s) tc::mbar_init(&accum_half[s], 1);
    tc::mbar_init(&d_free, TP_EPI);
    tc::mbar_init(&acc_free[0], TP_EPI);
    tc::mbar_init(&acc_free[1], TP_EPI);
    for (int s = 0; s < TP_MAXSLOTS; ++s) {
      tc::mbar_init(&full_w[s], 1);
      tc::mbar_init(&empty_w[s], 1);
    }
    for (int s = 0; s < TP_TQ; ++s) {
      tc::mbar_init(&tq_full[s], 1);
      tc::mbar_init(&tq_empty[s], TP_EPI + 128 + 1);  // epilogue + producer threads + one lane of the MMA warp
    }
    tc::mbar_init(&accum_full, 1);
    tc::mbar_init(&q_bar, 1);
    tc::mbar_init(&x_ready, TP_EPI);
    tc::mbar_fence_init();
  }
  for (int e = tid; e < nl * 256; e += TP_THREADS) {
    const int l = e >> 8, c = e & 255;
    s_scale[e] = c < p.L[l].cout ? p.L[l].scale[c] : 0.f;
    s_shift[e] = c < p.L[l].cout ? p.L[l].shift[c] : 0.f;
  }
  for (int e = tid; e < 3 * 128; e += TP_THREADS) s_wx[e] = PRE ? p.wx[e] : 0.f;
  if (TRAIN) {
    const int co = p.L[0].cout;
    for (int e = tid; e < 256; e += TP_THREADS) {
      s_in[e] = (p.in_scale && e < p.C) ? p.in_scale[e] : 1.f;
      s_in[256 + e] = (p.in_shift && e < p.C) ? p.in_shift[e] : 0.f;
      s_in[512 + e] = (p.dz_b && e < p.C) ? p.dz_b[e] : 0.f;
      s_in[768 + e] = (p.dz_c && e < p.C) ? p.dz_c[e] : 0.f;
      s_out[e] = (p.out_scale && e < co) ? p.out_scale[e] : 0.f;
      s_out[256 + e] = (p.out_shift && e < co) ? p.out_shift[e] : 0.f;
      s_out[512 + e] = (p.out_mean && e < co) ? p.out_mean[e] : 0.f;
      s_out[768 + e] = (p.out_invstd && e < co) ? p.out_invstd[e] : 0.f;
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  // Register reallocation between warpgroups (the SM's register file is split per scheduler: 4 warps x 128 at launch).
  // The MMA / loader warpgroup needs few registers; the producers (gather double buffer: 64 registers of loads in flight
  // per buffer) take what it gives up.  Per scheduler: 2 epilogue warps x 128 + 1 producer x 200 + 1 x 56 = 512.
  // (each setmaxnreg sits inside its role's branch, so that only that role's code is compiled against the new budget)
#ifdef B200_TC_PROFILE
  unsigned long long tp_acc[32] = {0};
  const int tp_tile_cat = warp == TP_MMA_WARP ? 2 : (warp >= 8 ? 8 : 12);
  const long long tp_start = clock64();
  unsigned long long tp_tiles = 0;
#endif

  // every consumer role walks the tile ring with its own counter; `arrive` = this thread is one of the ring's
  // registered consumers (all epilogue and producer threads, one lane of the MMA warp)
  int n_get = 0;
  auto next_tile = [&](bool arrive) -> int {
    const int slot = n_get % TP_TQ;
    TPW(tp_tile_cat, &tq_full[slot], (uint32_t)((n_get / TP_TQ) & 1));
    const int t = tq_tile[slot];
    if (arrive) tc::mbar_arrive(&tq_empty[slot]);
    ++n_get;
    return t;
  };

  if (warp >= 12) {
  // last warpgroup: loader, MMA issuer and two idle warps; their registers go to the producers
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
  if (warp == TP_LOAD_WARP) {
    // ================= tile scheduler + weight loader (converged warp, elected lane issues) =================
    int n_pub = 0;
    const int total_tiles = p.units ? (__ldg(p.total_units) + 7) >> 3 : p.total_tiles;
    auto publish = [&]() -> int {
      const int slot = n_pub % TP_TQ;
      TPW(0, &tq_empty[slot], (uint32_t)(((n_pub / TP_TQ) & 1) ^ 1));
      int t = 0;
      if (lane == 0) {
        {
          t = atomicAdd(p.tile_counter, 1);
          if (t >= total_tiles) t = -1;
        }
        tq_tile[slot] = t;
        tc::mbar_arrive(&tq_full[slot]);  // release: the tile id is visible to the waiters
      }
      t = __shfl_sync(0xffffffffu, t, 0);
      ++n_pub;
      return t;
    };
    uint32_t sw = 0, pw = 0;  // weight-ring slot and phase
    int cur = publish();
    while (cur >= 0) {
      const int nxt = publish();  // one tile ahead: the producers prefetch its rows while this one computes
      for (int l = 0; l < nl; ++l) {
        const int nst = p.L[l].nhalf * p.L[l].nkb;
        const uint8_t *src = p.packed + p.L[l].packed_off;
        const uint32_t part_bytes = (uint32_t)p.L[l].rows * 128u;  // W_hi or W_lo of one k-block
        for (int s = 0; s < 2 * nst; ++s) {                       // packed order: hi(0), lo(0), hi(1), lo(1), ...
          TPW(1, &empty_w[sw], pw ^ 1u);
          if (elect_one()) {
            tc::mbar_arrive_expect_tx(&full_w[sw], part_bytes);
            tc::bulk_g2s(R2 + sw * p.wslot_bytes, src + (size_t)s * part_bytes, part_bytes, &full_w[sw]);
          }
          __syncwarp();
          if (++sw == (uint32_t)nslots) { sw = 0; pw ^= 1u; }
        }
      }
      cur = nxt;
    }
  } else if (warp == TP_MMA_WARP) {
    // ================= MMA issuer (converged warp, elected lane issues) =================
    uint32_t sw = 0, pw = 0, sa = 0, pa = 0, n_xr = 0, tl = 0;
    unsigned long long mma_cols = 0;  // sum over issued MMAs of their N (each MMA = 128 x N x 8 TF32 multiply-adds)
    const uint32_t r1_addr = tc::smem_addr(R1), r2_addr = tc::smem_addr(R2);
    for (;;) {
      const int t = next_tile(lane == 0);
      if (t < 0) break;
      const uint32_t d_big = tmem_d + ((tl & 1u) ? 256u : 0u);  // this tile's accumulator region
      const uint32_t d_small = d_big + 128u;
      const uint32_t x_hi = tmem_d + ((tl & 1u) ? 0u : 256u);   // this tile's activation region (A operand, layers >= 2)
      const uint32_t x_lo = x_hi + 128u;
      if (nl == 1 && tl >= 2u) {
        // single-layer stack: no hidden hand-off orders this tile's MMAs after the epilogue of the tile that used this
        // accumulator region two tiles ago -- wait for that epilogue's last TMEM read explicitly
        TPW(4, &acc_free[tl & 1u], ((tl >> 1) - 1u) & 1u);
        tc::tc_fence_after_sync();
      }
      for (int l = 0; l < nl; ++l) {
        const bool last = l == nl - 1;
        if (l > 0) {
          TPW(3, &x_ready, n_xr & 1u);  // layer l-1's activations are in tensor memory, its accumulators were read
          ++n_xr;
          tc::tc_fence_after_sync();
        }
        const int nkb = p.L[l].nkb;
        const uint32_t idesc = tc::make_idesc_tf32(128, p.L[l].rows);
        for (int h = 0; h < p.L[l].nhalf; ++h) {
          if (last && h == 1) {  // the second half reuses the accumulator columns the epilogue is draining
            TPW(4, &d_free, tl & 1u);
            tc::tc_fence_after_sync();
          }
          for (int kb = 0; kb < nkb; ++kb) {
            uint64_t da_hi = 0, da_lo = 0;
            int nks = 4;  // UMMA_K = 8 columns per MMA; a partially filled last layer-1 k-block needs fewer steps
            if (l == 0) {
              TPW(5, &full_a[sa], pa);
              const uint32_t a_hi = r1_addr + sa * 2u * TC_KB_BYTES;
              da_hi = tc::make_desc_sw128(a_hi);
              da_lo = tc::make_desc_sw128(a_hi + TC_KB_BYTES);
              if (kb == nkb - 1) nks = (p.L[0].cin - kb * 32 + 7) >> 3;
            }
            const uint32_t xc = (uint32_t)(kb * 32);
            // both halves of the k-block's weights (W_hi in slot sw, W_lo in the next slot), then ONE issue burst:
            //   A_hi*W_hi -> products, A_lo*W_hi -> corrections, A_hi*W_lo -> corrections
            uint32_t sw2 = sw + 1, pw2 = pw;
            if (sw2 == (uint32_t)nslots) { sw2 = 0; pw2 ^= 1u; }
            TPW(6, &full_w[sw], pw);
            TPW(6, &full_w[sw2], pw2);
            tc::tc_fence_after_sync();
            TP_BEGIN();
            const uint64_t dwh = tc::make_desc_sw128(r2_addr + sw * (uint32_t)p.wslot_bytes);
            const uint64_t dwl = tc::make_desc_sw128(r2_addr + sw2 * (uint32_t)p.wslot_bytes);
            if (elect_one()) {
              if (l == 0) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  if (ks < nks) {
                    const uint64_t adv = (uint64_t)(ks * 2);  // 8 floats = 32 B = 2 x 16 B
                    tc::mma_tf32(d_big, da_hi + adv, dwh + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                    tc::mma_tf32(d_small, da_lo + adv, dwh + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                  }
                }
                tc::mma_commit(&empty_w[sw]);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (ks < nks) tc::mma_tf32(d_small, da_hi + (uint64_t)(ks * 2), dwl + (uint64_t)(ks * 2), idesc, 1u);
                tc::mma_commit(&empty_a[sa]);
                tc::mma_commit(&empty_w[sw2]);
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t adv = (uint64_t)(ks * 2);
                  mma_tf32_ts(d_big, x_hi + xc + (uint32_t)(ks * 8), dwh + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                  mma_tf32_ts(d_small, x_lo + xc + (uint32_t)(ks * 8), dwh + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                }
                tc::mma_commit(&empty_w[sw]);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  mma_tf32_ts(d_small, x_hi + xc + (uint32_t)(ks * 8), dwl + (uint64_t)(ks * 2), idesc, 1u);
                tc::mma_commit(&empty_w[sw2]);
              }
            }
            __syncwarp();
            TP_END(7);
            mma_cols += (unsigned long long)(3 * nks) * (unsigned)p.L[l].rows;
            if (l == 0 && ++sa == (uint32_t)p.a_stages) { sa = 0; pa ^= 1u; }
            sw = sw2 + 1; pw = pw2;
            if (sw == (uint32_t)nslots) { sw = 0; pw ^= 1u; }
          }
          if (last) {
            if (elect_one()) tc::mma_commit(&accum_half[(tl & 1u) * 2u + h]);
            __syncwarp();
          }
        }
        if (!last) {
          if (elect_one()) tc::mma_commit(&accum_full);
          __syncwarp();
        }
      }
      ++tl;
    }
    if (lane == 0 && mma_cols) atomicAdd(&g_sa_mma_cols, mma_cols);
  }
  } else if (warp >= 8) {
    // ================= producers: row (tid - 256) of the tile, layer-1 A operand =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;" ::: "memory");
    const int row = tid - TP_PROD0;
    const int g = row / ns;
    const int C = p.C;
    uint32_t sa = 0, pa = 0;
    int q_scene = -1;        // QUERY: scene whose coordinates are staged in s_qxyz
    uint32_t q_phase = 0;
    for (;;) {
      const int t = next_tile(true);
      if (t < 0) break;
      int b, m0, g_here, slot, gi;
      bool valid;
      if (MODE == 2) {
        // plain row GEMM (pass 1 of the factorised layer 1): tile row = point t * 128 + row of the flattened (B*N, C) input
        b = 0; m0 = 0; g_here = 1; gi = 0; slot = 0;
        valid = t * TC_ROWS + row < p.rows_total;
      } else if (p.units) {
        // compacted tile: 8 units of 16 slots, each from the centre the unit list names
        const int u = t * 8 + (row >> 4);
        valid = u < __ldg(p.total_units);
        const int e = valid ? __ldg(p.unit_list + u) : 0;
        const int c = e >> 3;
        b = c / p.M;
        m0 = c - b * p.M;
        gi = 0;
        g_here = 1;
        slot = (e & 7) * 16 + (row & 15);
      } else {
        b = t / p.tiles_per_scene;
        m0 = (t - b * p.tiles_per_scene) * G;
        g_here = min(G, p.M - m0);
        valid = g < g_here;
        gi = g;
        slot = row - g * ns;
      }
      int src_idx = 0;
      float ctr[3] = {0.f, 0.f, 0.f};
      if (valid && MODE == 0) {
        if (!QUERY) src_idx = p.idx[((size_t)b * p.M + m0 + gi) * ns + slot];
        const float *c = p.new_xyz + ((size_t)b * p.M + m0 + gi) * 3;
        ctr[0] = c[0]; ctr[1] = c[1]; ctr[2] = c[2];
      }
      if (QUERY && MODE == 0) {
        // ---- fused ball query: this warp's rows are the nsample slots of 32 / ns consecutive centres -------------------
        if (b != q_scene) {  // uniform over the producers: they walk the same tile sequence
          asm volatile("bar.sync 2, 128;" ::: "memory");  // every producer warp is done with the scene staged before
          if (tid == TP_PROD0) {
            tc::mbar_arrive_expect_tx(&q_bar, (uint32_t)p.N * 12u);
            tc::bulk_g2s(s_qxyz, p.xyz + (size_t)b * p.N * 3, (uint32_t)p.N * 12u, &q_bar);
          }
          TPW(10, &q_bar, q_phase);
          q_phase ^= 1u;
          q_scene = b;
        }
        const int pw = warp - 8, cpw = 32 / ns;  // centres per warp: 4, 2 or 1
        __syncwarp();  // the previous tile's lists have been read
        const float *cbase = p.new_xyz + ((size_t)b * p.M + m0 + pw * cpw) * 3;
        const int live = max(0, min(cpw, g_here - pw * cpw));  // warp-uniform
        if (cpw == 2)
          query_scan<2>(s_qxyz, p.N, ns, p.radius2, cbase, live, lane, s_qlist[pw], s_qcnt[pw]);
        else if (cpw == 4)
          query_scan<4>(s_qxyz, p.N, ns, p.radius2, cbase, live, lane, s_qlist[pw], s_qcnt[pw]);
        else
          query_scan<1>(s_qxyz, p.N, ns, p.radius2, cbase, live, lane, s_qlist[pw], s_qcnt[pw]);
        __syncwarp();
        if (valid) {
          const int cw = lane / ns, have = s_qcnt[pw][cw];
          // :39-46 slot t holds the t-th hit, the first hit fills the rest, an empty ball keeps the zero row
          src_idx = slot < have ? s_qlist[pw][cw * ns + slot] : (have > 0 ? s_qlist[pw][cw * ns] : 0);
          if (p.idx_out) p.idx_out[((size_t)b * p.M + m0 + gi) * ns + slot] = src_idx;
        }
      }
      const float *frow = (valid && C > 0 && MODE == 0) ? p.feat_pm + ((size_t)b * p.N + src_idx) * C : nullptr;
      if (valid && MODE == 2) frow = p.feat_pm + ((size_t)t * TC_ROWS + row) * p.ld;
      size_t cm_stride = 0;  // channel-major source (MODE 2): element (row, ch) = frow[ch * cm_stride]
      if (MODE == 2 && !TRAIN && p.cm_in && valid) {
        const size_t r = (size_t)t * TC_ROWS + row, sb = r / (size_t)p.rows_per_scene;
        cm_stride = (size_t)p.rows_per_scene;
        frow = p.feat_pm + sb * (size_t)C * cm_stride + (r - sb * cm_stride);
      }
      // mode 1, second source: this row's own (skip) features follow the blended channels (PointnetFPModule's
      // torch.cat([interpolated, unknow_feats]), pointnet2_modules.py:413-418)
      const float *f2row = (MODE == 1 && valid && p.C2 > 0)
                               ? p.feat2_pm + (((size_t)b * p.M + m0 + gi) * ns + slot) * p.C2 : nullptr;
      float rel[3] = {0.f, 0.f, 0.f};
      const float *f3[3] = {nullptr, nullptr, nullptr};
      float wt[3] = {0.f, 0.f, 0.f};
      if (MODE == 1 && valid) {
        const size_t q = ((size_t)b * p.M + m0 + gi) * ns + slot;
        for (int u = 0; u < 3; ++u) {
          f3[u] = p.feat_pm + ((size_t)b * p.N + p.idx3[q * 3 + u]) * C;
          wt[u] = p.w3[q * 3 + u];
        }
        if (p.rel3 && p.use_xyz) {
          rel[0] = p.rel3[q * 3 + 0]; rel[1] = p.rel3[q * 3 + 1]; rel[2] = p.rel3[q * 3 + 2];
        }
      }
      if (valid && p.use_xyz && MODE == 0) {
        const float *q = p.xyz + ((size_t)b * p.N + src_idx) * 3;
        // pointnet2_utils.py:351-353: grouped_xyz -= new_xyz ; /= radius  (x * fp32(1/r) on CUDA)
        rel[0] = __fmul_rn(__fsub_rn(q[0], ctr[0]), p.inv_r);
        rel[1] = __fmul_rn(__fsub_rn(q[1], ctr[1]), p.inv_r);
        rel[2] = __fmul_rn(__fsub_rn(q[2], ctr[2]), p.inv_r);
      }
      // layer-1 A operand: [features (C) | rel xyz (3) | 0 ...]
      // TRAIN backward: the row's upstream gradient (dense rows, or the pooled gradient of the row's centre)
      const float *g_row = nullptr, *go_row = nullptr;
      const int32_t *arg_row = nullptr;
      int my_slot = 0;
      if (TRAIN && valid && p.train_in == 2) g_row = p.g_rows + ((size_t)t * TC_ROWS + row) * p.ld;
      if (TRAIN && valid && p.train_in == 3) {
        const long long rr = (long long)t * TC_ROWS + row, grp = rr / p.pool_ns;
        my_slot = (int)(rr - grp * p.pool_ns);
        go_row = p.gout_pm + grp * C;
        arg_row = p.arg_pm + grp * C;
      }
      auto load_kb = [&](int kb, float4 (&v)[8]) {
        if (TRAIN && p.train_in >= 2) {
          // dz = in_scale * g + dz_b + dz_c * z  (all loads of the k-block first, then the arithmetic)
          float4 zv[8], gv[8];
          int4 av[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int ch = kb * 32 + c * 4;
            const bool in = valid && ch + 3 < C;
            zv[c] = in ? __ldg(reinterpret_cast<const float4 *>(frow + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.train_in == 2) {
              gv[c] = in ? __ldg(reinterpret_cast<const float4 *>(g_row + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              gv[c] = in ? __ldg(reinterpret_cast<const float4 *>(go_row + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
              av[c] = in ? __ldg(reinterpret_cast<const int4 *>(arg_row + ch)) : make_int4(-1, -1, -1, -1);
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int ch = kb * 32 + c * 4;
            const bool in = valid && ch + 3 < C;
            const float4 a4 = *reinterpret_cast<const float4 *>(s_in + ch);
            const float4 h4 = *reinterpret_cast<const float4 *>(s_in + 256 + ch);
            const float4 b4 = *reinterpret_cast<const float4 *>(s_in + 512 + ch);
            const float4 c4 = *reinterpret_cast<const float4 *>(s_in + 768 + ch);
            float4 g = gv[c];
            if (p.train_in == 3) {
              g.x = (av[c].x == my_slot && fmaf(zv[c].x, a4.x, h4.x) > 0.f) ? g.x : 0.f;
              g.y = (av[c].y == my_slot && fmaf(zv[c].y, a4.y, h4.y) > 0.f) ? g.y : 0.f;
              g.z = (av[c].z == my_slot && fmaf(zv[c].z, a4.z, h4.z) > 0.f) ? g.z : 0.f;
              g.w = (av[c].w == my_slot && fmaf(zv[c].w, a4.w, h4.w) > 0.f) ? g.w : 0.f;
            }
            v[c].x = in ? fmaf(c4.x, zv[c].x, fmaf(a4.x, g.x, b4.x)) : 0.f;
            v[c].y = in ? fmaf(c4.y, zv[c].y, fmaf(a4.y, g.y, b4.y)) : 0.f;
            v[c].z = in ? fmaf(c4.z, zv[c].z, fmaf(a4.z, g.z, b4.z)) : 0.f;
            v[c].w = in ? fmaf(c4.w, zv[c].w, fmaf(a4.w, g.w, b4.w)) : 0.f;
          }
          return;
        }
        if (MODE == 2 && !TRAIN && p.cm_in) {  // uniform: 32 coalesced scalar loads per k-block (lanes = consecutive rows)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int ch = kb * 32 + c * 4;
            v[c].x = (valid && ch + 0 < C) ? __ldg(frow + (size_t)(ch + 0) * cm_stride) : 0.f;
            v[c].y = (valid && ch + 1 < C) ? __ldg(frow + (size_t)(ch + 1) * cm_stride) : 0.f;
            v[c].z = (valid && ch + 2 < C) ? __ldg(frow + (size_t)(ch + 2) * cm_stride) : 0.f;
            v[c].w = (valid && ch + 3 < C) ? __ldg(frow + (size_t)(ch + 3) * cm_stride) : 0.f;
          }
          return;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int ch = kb * 32 + c * 4;
          if (MODE == 1 && !PRE && f2row && ch >= C) {  // skip features (C and C2 are multiples of 4 here)
            v[c] = ch < C + p.C2 ? __ldg(reinterpret_cast<const float4 *>(f2row + (ch - C))) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (valid && MODE == 1 && (PRE || ch + 3 < C)) {  // blend of the three neighbours (three_interpolate + concat)
            const float4 a0 = __ldg(reinterpret_cast<const float4 *>(f3[0] + ch));
            const float4 a1 = __ldg(reinterpret_cast<const float4 *>(f3[1] + ch));
            const float4 a2 = __ldg(reinterpret_cast<const float4 *>(f3[2] + ch));
            // interpolate_gpu.cu:100-104 rounding: fma(p3,w3, fma(p1,w1, p2*w2))
            v[c].x = __fmaf_rn(a2.x, wt[2], __fmaf_rn(a0.x, wt[0], __fmul_rn(a1.x, wt[1])));
            v[c].y = __fmaf_rn(a2.y, wt[2], __fmaf_rn(a0.y, wt[0], __fmul_rn(a1.y, wt[1])));
            v[c].z = __fmaf_rn(a2.z, wt[2], __fmaf_rn(a0.z, wt[0], __fmul_rn(a1.z, wt[1])));
            v[c].w = __fmaf_rn(a2.w, wt[2], __fmaf_rn(a0.w, wt[0], __fmul_rn(a1.w, wt[1])));
          } else if (valid && MODE != 1 && (PRE || (p.vec_gather && ch + 3 < C))) {
            v[c] = __ldg(reinterpret_cast<const float4 *>(frow + ch));
          } else if (PRE) {
            v[c] = make_float4(0.f, 0.f, 0.f, 0.f);  // rows of P are whole 128-byte k-blocks: no scalar tail
          } else {
            float e4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int k = ch + e;
              float x = 0.f;
              if (valid) {
                if (k < C) {
                  x = MODE != 1 ? frow[k]
                                  : __fmaf_rn(f3[2][k], wt[2], __fmaf_rn(f3[0][k], wt[0], __fmul_rn(f3[1][k], wt[1])));
                } else if (p.use_xyz && k < C + 3) {
                  x = k == C ? rel[0] : (k == C + 1 ? rel[1] : rel[2]);
                }
              }
              e4[e] = x;
            }
            v[c] = make_float4(e4[0], e4[1], e4[2], e4[3]);
          }
        }
      };
      auto store_kb = [&](int kb, float4 (&v)[8]) {
        if (PRE) {
          // factorised layer 1, applied when the k-block is consumed (the loads stay in flight until here): v holds
          // scale1 * (W1f * f) + shift1 of the source point (or the blend of three); add the relative-xyz part of the
          // layer and apply its ReLU (rel = 0 and v = 0 for rows past the end)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int ch = kb * 32 + c * 4;
            const float4 w0 = *reinterpret_cast<const float4 *>(s_wx + ch);
            const float4 w1 = *reinterpret_cast<const float4 *>(s_wx + 128 + ch);
            const float4 w2 = *reinterpret_cast<const float4 *>(s_wx + 256 + ch);
            v[c].x = fmaxf(__fmaf_rn(w2.x, rel[2], __fmaf_rn(w1.x, rel[1], __fmaf_rn(w0.x, rel[0], v[c].x))), 0.f);
            v[c].y = fmaxf(__fmaf_rn(w2.y, rel[2], __fmaf_rn(w1.y, rel[1], __fmaf_rn(w0.y, rel[0], v[c].y))), 0.f);
            v[c].z = fmaxf(__fmaf_rn(w2.z, rel[2], __fmaf_rn(w1.z, rel[1], __fmaf_rn(w0.z, rel[0], v[c].z))), 0.f);
            v[c].w = fmaxf(__fmaf_rn(w2.w, rel[2], __fmaf_rn(w1.w, rel[1], __fmaf_rn(w0.w, rel[0], v[c].w))), 0.f);
          }
        }
        if (TRAIN && p.in_scale && p.train_in < 2) {
          // the previous layer's BatchNorm (batch statistics folded to an affine) + ReLU; rows past the end stay zero
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int ch = kb * 32 + c * 4;
            const float4 s4 = *reinterpret_cast<const float4 *>(s_in + ch);
            const float4 h4 = *reinterpret_cast<const float4 *>(s_in + 256 + ch);
            v[c].x = valid ? fmaxf(fmaf(v[c].x, s4.x, h4.x), 0.f) : 0.f;
            v[c].y = valid ? fmaxf(fmaf(v[c].y, s4.y, h4.y), 0.f) : 0.f;
            v[c].z = valid ? fmaxf(fmaf(v[c].z, s4.z, h4.z), 0.f) : 0.f;
            v[c].w = valid ? fmaxf(fmaf(v[c].w, s4.w, h4.w), 0.f) : 0.f;
          }
        }
        TPW(9, &empty_a[sa], pa ^ 1u);
        uint8_t *a_hi = R1 + sa * 2 * TC_KB_BYTES, *a_lo = a_hi + TC_KB_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 h, l;
          tc::split_tf32(v[c].x, h.x, l.x); tc::split_tf32(v[c].y, h.y, l.y);
          tc::split_tf32(v[c].z, h.z, l.z); tc::split_tf32(v[c].w, h.w, l.w);
          const uint32_t off = tc::sw128_offset(row, c);
          *reinterpret_cast<float4 *>(a_hi + off) = h;
          *reinterpret_cast<float4 *>(a_lo + off) = l;
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive(&full_a[sa]);
        if (++sa == (uint32_t)p.a_stages) { sa = 0; pa ^= 1u; }
      };
      // register double buffer: the loads of k-block kb+1 are in flight while kb is split and stored; the operand ring
      // is private to layer 1, so the whole gather runs ahead of the MMAs (up to TP_ASTAGES k-blocks)
      // (a single-layer stack with two output halves consumes the operand once per half: emit it twice)
      const int nkb1 = p.L[0].nkb;
      const int reps = nl == 1 ? p.L[0].nhalf : 1;
      for (int rep = 0; rep < reps; ++rep) {
        float4 va[8], vb[8];
        load_kb(0, va);
        for (int kb = 0; kb < nkb1; kb += 2) {
          if (kb + 1 < nkb1) load_kb(kb + 1, vb);
          store_kb(kb, va);
          if (kb + 1 < nkb1) {
            if (kb + 2 < nkb1) load_kb(kb + 2, va);
            store_kb(kb + 1, vb);
          }
        }
      }
    }
  } else {
    // ================= epilogue: warpgroup wg takes the 32-column chunks wg, wg+2, ...; row = TMEM lane =================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    uint32_t n_acc = 0, tl = 0;
    const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    for (;;) {
      const int t = next_tile(true);
      if (t < 0) break;
      int b = 0, m0 = 0, g_here = 0;
      int u_centre = -1;  // compacted mode: the centre of this lane's 16-row unit (-1: past the end)
      if (p.units) {
        const int u = t * 8 + (row >> 4);
        if (u < __ldg(p.total_units)) u_centre = __ldg(p.unit_list + u) >> 3;
      } else {
        b = t / p.tiles_per_scene;
        m0 = (t - b * p.tiles_per_scene) * G;
        g_here = min(G, p.M - m0);
      }
      const uint32_t d_addr = lane_base + ((tl & 1u) ? 256u : 0u);  // accumulators: products | corrections (+128)
      const uint32_t x_addr = lane_base + ((tl & 1u) ? 0u : 256u);  // activations:  X_hi | X_lo (+128)
      for (int l = 0; l < nl; ++l) {
        const float *sc = s_scale + l * 256, *sh = s_shift + l * 256;
        if (l + 1 < nl) {
          TPW(13, &accum_full, n_acc & 1u);
          ++n_acc;
          tc::tc_fence_after_sync();
          TP_BEGIN();
          // hidden layer: X = relu(scale*acc+shift) -> hi/lo split -> tensor memory (the next layer's A operand)
          const int H = p.L[l].cout;  // hidden width, multiple of 32
          for (int c0 = wg * 32; c0 < H; c0 += 64) {
            uint32_t r[32], r2[32];
            TP_HBEGIN();
            tc::tmem_ld_32x32(d_addr + (uint32_t)c0, r);
            tc::tmem_ld_32x32(d_addr + 128u + (uint32_t)c0, r2);
            tc::tmem_ld_wait();
            TP_HLAP(20);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 s4 = *reinterpret_cast<const float4 *>(sc + c0 + c * 4);
              const float4 h4 = *reinterpret_cast<const float4 *>(sh + c0 + c * 4);
              const float scv[4] = {s4.x, s4.y, s4.z, s4.w}, shv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float acc = __uint_as_float(r[c * 4 + e]) + __uint_as_float(r2[c * 4 + e]);
                const float y = fmaxf(fmaf(acc, scv[e], shv[e]), 0.f);
                float hh, lo;
                tc::split_tf32(y, hh, lo);
                r[c * 4 + e] = __float_as_uint(hh);
                r2[c * 4 + e] = __float_as_uint(lo);
              }
            }
            TP_HLAP(21);
            tmem_st_32x32(x_addr + (uint32_t)c0, r);
            tmem_st_32x32(x_addr + 128u + (uint32_t)c0, r2);
          }
          tmem_st_wait();
          tc::tc_fence_before_sync();
          tc::mbar_arrive(&x_ready);
          TP_END(15);
        } else {
          // last layer: relu(scale*acc+shift), then the max over each centre's nsample rows.  Rows are TMEM lanes =
          // lanes of this warp (and of its neighbours when nsample > 32): reduce in registers with warp shuffles.
          const int cout = p.L[l].cout;
          const int rows_l = p.L[l].rows;
          float *part = s_partial + (tl & 1u) * (4 * 256);
          for (int h = 0; h < nhalf_last; ++h) {
            TPW(14, &accum_half[(tl & 1u) * 2u + h], (tl >> 1) & 1u);
            tc::tc_fence_after_sync();
            const bool release = nhalf_last == 2 && h == 0;  // half 1 accumulates into the same columns
            if (release && wg * 32 >= rows_l) {              // no chunk for this warpgroup: nothing to drain
              tc::tc_fence_before_sync();
              tc::mbar_arrive(&d_free);
            }
            for (int cc0 = wg * 32; cc0 < rows_l; cc0 += 64) {
              const int c0 = h * rows_l + cc0;  // output channel of the chunk's first column
              uint32_t r[32], r2[32];
              TP_BEGIN();
              tc::tmem_ld_32x32(d_addr + (uint32_t)cc0, r);
              tc::tmem_ld_32x32(d_addr + 128u + (uint32_t)cc0, r2);
              tc::tmem_ld_wait();
              TP_LAP(16);
              if (release && cc0 + 64 >= rows_l) {  // this thread's last read of half 0
                tc::tc_fence_before_sync();
                tc::mbar_arrive(&d_free);
              }
              float v[32];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 s4 = *reinterpret_cast<const float4 *>(sc + c0 + c * 4);
                const float4 h4 = *reinterpret_cast<const float4 *>(sh + c0 + c * 4);
                v[c * 4 + 0] = fmaxf(fmaf(__uint_as_float(r[c * 4 + 0]) + __uint_as_float(r2[c * 4 + 0]), s4.x, h4.x), 0.f);
                v[c * 4 + 1] = fmaxf(fmaf(__uint_as_float(r[c * 4 + 1]) + __uint_as_float(r2[c * 4 + 1]), s4.y, h4.y), 0.f);
                v[c * 4 + 2] = fmaxf(fmaf(__uint_as_float(r[c * 4 + 2]) + __uint_as_float(r2[c * 4 + 2]), s4.z, h4.z), 0.f);
                v[c * 4 + 3] = fmaxf(fmaf(__uint_as_float(r[c * 4 + 3]) + __uint_as_float(r2[c * 4 + 3]), s4.w, h4.w), 0.f);
              }
              if (ROWOUT) {
                // row output: affine (+ ReLU when asked), this row's 32 columns.  Point-major: one 128-byte run per row
                // (width a multiple of 4); channel-major (B, cout, rows_per_scene): lanes = consecutive rows, coalesced
                // per column.  Columns past cout (stacks padded to a multiple of 32) are dropped.
                size_t grow;
                bool rvalid;
                if (MODE == 2) {
                  grow = (size_t)t * TC_ROWS + row;
                  rvalid = grow < (size_t)p.rows_total;
                } else {
                  const int gg = row / ns;
                  rvalid = gg < g_here;
                  grow = ((size_t)b * p.M + m0 + gg) * ns + (row - gg * ns);
                }
                if (TRAIN) {
                  // training passes: raw rows out (forward: z_l; backward: g_{l-1} = da * ReLU mask of the layer below) and
                  // per-tile column sums over this warp's 32 rows (forward: sum z, sum z^2; backward: sum g, sum g * xhat).
                  // Rows past the end are exact zeros (zero operand rows, unit scale, zero shift).
                  float sv[32], sq[32];
                  const size_t rbase = grow * (size_t)cout + c0;
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    float x[4] = {__uint_as_float(r[c * 4 + 0]) + __uint_as_float(r2[c * 4 + 0]),
                                  __uint_as_float(r[c * 4 + 1]) + __uint_as_float(r2[c * 4 + 1]),
                                  __uint_as_float(r[c * 4 + 2]) + __uint_as_float(r2[c * 4 + 2]),
                                  __uint_as_float(r[c * 4 + 3]) + __uint_as_float(r2[c * 4 + 3])};
                    float y[4];
                    const bool in = rvalid && c0 + c * 4 + 3 < cout;
                    if (p.train_out == 1) {
                      const float4 zp = in ? __ldg(reinterpret_cast<const float4 *>(p.zprev + rbase + c * 4))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                      const float zz[4] = {zp.x, zp.y, zp.z, zp.w};
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const int cc = c0 + c * 4 + e;
                        x[e] = (in && fmaf(zz[e], s_out[cc], s_out[256 + cc]) > 0.f) ? x[e] : 0.f;
                        y[e] = x[e] * ((zz[e] - s_out[512 + cc]) * s_out[768 + cc]);
                      }
                    } else {
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        x[e] = in ? x[e] : 0.f;
                        y[e] = x[e] * x[e];
                      }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      sv[c * 4 + e] = x[e];
                      sq[c * 4 + e] = y[e];
                    }
                    if (in) *reinterpret_cast<float4 *>(p.out_pm + rbase + c * 4) = make_float4(x[0], x[1], x[2], x[3]);
                  }
                  if (p.stats) {
                    transpose_sum(sv, lane);
                    transpose_sum(sq, lane);
                    float *dst = p.stats + ((size_t)(t * 4 + (warp & 3)) * 2) * 256 + c0 + lane;
                    dst[0] = sv[0];
                    dst[256] = sq[0];
                  }
                  TP_LAP(18);
                  continue;
                }
                if (rvalid) {
                  float o[32];
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 s4 = *reinterpret_cast<const float4 *>(sc + c0 + c * 4);
                    const float4 h4 = *reinterpret_cast<const float4 *>(sh + c0 + c * 4);
                    o[c * 4 + 0] = fmaf(__uint_as_float(r[c * 4 + 0]) + __uint_as_float(r2[c * 4 + 0]), s4.x, h4.x);
                    o[c * 4 + 1] = fmaf(__uint_as_float(r[c * 4 + 1]) + __uint_as_float(r2[c * 4 + 1]), s4.y, h4.y);
                    o[c * 4 + 2] = fmaf(__uint_as_float(r[c * 4 + 2]) + __uint_as_float(r2[c * 4 + 2]), s4.z, h4.z);
                    o[c * 4 + 3] = fmaf(__uint_as_float(r[c * 4 + 3]) + __uint_as_float(r2[c * 4 + 3]), s4.w, h4.w);
                  }
                  if (p.final_relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = fmaxf(o[i], 0.f);
                  }
                  if (p.out_pm) {
                    if (c0 + 32 <= cout && (cout & 3) == 0) {
                      float4 *dst = reinterpret_cast<float4 *>(p.out_pm + grow * cout + c0);
#pragma unroll
                      for (int c = 0; c < 8; ++c) dst[c] = make_float4(o[c * 4], o[c * 4 + 1], o[c * 4 + 2], o[c * 4 + 3]);
                    } else {
#pragma unroll
                      for (int i = 0; i < 32; ++i)
                        if (c0 + i < cout) p.out_pm[grow * cout + c0 + i] = o[i];
                    }
                  }
                  if (p.out) {
                    const size_t sb = grow / (size_t)p.rows_per_scene, si = grow - sb * (size_t)p.rows_per_scene;
                    float *dst = p.out + (sb * cout + c0) * (size_t)p.rows_per_scene + si;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                      if (c0 + i < cout) dst[(size_t)i * p.rows_per_scene] = o[i];
                  }
                }
                TP_LAP(18);
                continue;
              }
              if (p.units) {
                // 16-row units: lanes 0-15 / 16-31 reduce separately; a centre's units meet in global memory through an
                // integer atomicMax (outputs are >= 0 after the ReLU, so the int order is the float order; the
                // launcher zero-fills the outputs, and max is order-independent -> still bit-exact)
                transpose_max<4>(v, lane, 8);
                if (u_centre >= 0) {
                  const int ub = u_centre / p.M, um = u_centre - ub * p.M;
                  const int cb = (lane & 15) * 2;
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    const int cc = c0 + cb + i;
                    atomicMax(reinterpret_cast<int *>(p.out + ((size_t)ub * cout + cc) * p.M + um), __float_as_int(v[i]));
                    if (p.out_pm)
                      atomicMax(reinterpret_cast<int *>(p.out_pm + ((size_t)ub * p.M + um) * cout + cc), __float_as_int(v[i]));
                  }
                }
                TP_LAP(18);
                continue;
              }
              if (ns >= 32) {
                transpose_max<5>(v, lane, 16);
              } else if (ns == 16) {
                transpose_max<4>(v, lane, 8);
              } else {
                transpose_max<3>(v, lane, 4);
              }
              TP_LAP(17);
              if (ns > 32) {
                part[(warp & 3) * 256 + c0 + lane] = v[0];
              } else {
                const int gg = row / ns;
                const int nv = ns == 32 ? 1 : (ns == 16 ? 2 : 4);
                const int cb = (lane & (ns - 1)) * nv;
                if (gg < g_here) {
                  const int m = m0 + gg;
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    if (i < nv) {
                      const int cc = c0 + cb + i;
                      p.out[((size_t)b * cout + cc) * p.M + m] = v[i];
                      if (p.out_pm) p.out_pm[((size_t)b * p.M + m) * cout + cc] = v[i];
                    }
                  }
                }
              }
              TP_LAP(18);
            }
          }
          if (ns > 32 && !p.units && !ROWOUT) {
            TP_BEGIN();
            // a centre spans nsample / 32 warps: combine their maxima (double-buffered by tile, one barrier per tile)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int wpc = ns >> 5;
            for (int o = tid; o < G * cout; o += TP_EPI) {
              const int gg = o / cout, cc = o - gg * cout;
              float mx = part[(gg * wpc) * 256 + cc];
              for (int w = 1; w < wpc; ++w) mx = fmaxf(mx, part[(gg * wpc + w) * 256 + cc]);
              if (gg < g_here) {
                const int m = m0 + gg;
                p.out[((size_t)b * cout + cc) * p.M + m] = mx;
                if (p.out_pm) p.out_pm[((size_t)b * p.M + m) * cout + cc] = mx;
              }
            }
            TP_END(19);
          }
        }
      }
      if (nl == 1) {  // this tile's accumulators are fully read: the MMA warp may reuse the region (see acc_free there)
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&acc_free[tl & 1u]);
      }
      ++tl;
#ifdef B200_TC_PROFILE
      ++tp_tiles;
#endif
    }
  }
#ifdef B200_TC_PROFILE
  if (blockIdx.x == 0 && (tid == 0 || tid == TP_PROD0 || tid == TP_MMA_WARP * 32 || tid == TP_LOAD_WARP * 32)) {
    const int role = tid == 0 ? 3 : (tid == TP_PROD0 ? 2 : (tid == TP_MMA_WARP * 32 ? 1 : 0));
    for (int c = 0; c < 32; ++c)
      if (tp_acc[c]) atomicAdd(&g_tcp_prof[c < 16 ? c : c + 8], tp_acc[c]);
    atomicAdd(&g_tcp_prof[16 + role], (unsigned long long)(clock64() - tp_start));
    if (tid == 0) atomicAdd(&g_tcp_prof[21], tp_tiles);
  }
#endif
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == TP_MMA_WARP) tc::tmem_dealloc<512>(tmem_d);
}

#ifdef B200_TC_PROFILE
extern "C" int b200_debug_tcp_profile(unsigned long long *out32) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, g_tcp_prof, sizeof(unsigned long long) * 48);
  unsigned long long z[48] = {0};
  cudaMemcpyToSymbol(g_tcp_prof, z, sizeof(z));
  return 0;
}
#endif

// Geometry + launch.  `p` carries the layer table, packed weights and row sources prepared by sa_launch.cu.
size_t sa_tcp_unit_list_bytes(int B, int M, int nsample) {
  return ((size_t)B * M * (size_t)(nsample / 16 + 1) + 64) * sizeof(int);
}

// compacted tiles: neighbour lists, nsample 32..128 in whole 16-slot units (<= 8 units per centre), max-pooled output
bool sa_tcp_units_wanted(int mode, int rowout, int nsample, int B, int M) {
  const char *e = getenv("B200_SA_TC_UNITS");  // read per call: the parity suite runs both ways in one process
  if (e && atoi(e) == 0) return false;
  return mode == 0 && !rowout && nsample >= 32 && (nsample & 15) == 0 && nsample <= 128 && (long long)B * M < (1ll << 27);
}

int sa_tcp_units_from_idx(int B, int M, int nsample, const int32_t *idx, int *unit_list, int *total, cudaStream_t stream) {
  const int centres = B * M;
  sa_unit_append_kernel<<<ceil_div(centres, 8), 256, 0, stream>>>(centres, nsample, idx, unit_list, total);
  B200_LAUNCH_OK("sa_unit_append_kernel");
  return 0;
}

bool sa_tcp_query_fusable(int B, int N, int M, int nsample, const float *xyz) {
  const char *e = getenv("B200_SA_TC_QUERY");  // read per call: the parity suite runs both ways in one process
  if (e && atoi(e) == 0) return false;
  // one 16-byte-granular bulk copy per scene (N * 12 bytes from a 16-byte-aligned base), rows of one warp = whole centres
  return N >= 1 && N <= TC_QUERY_MAX_N && (N & 3) == 0 && ((((uintptr_t)xyz) & 15) == 0) && nsample >= 8 && nsample <= 32 &&
         (32 % nsample) == 0 && (128 % nsample) == 0 && B >= 1 && M >= 1;
}

template <int MODE, int PRE, int ROWOUT, int TRAIN = 0, int QUERY = 0>
static int launch_variant(const TcParams &p, int grid, size_t smem, cudaStream_t stream) {
  static DynSmemOptIn optin;  // one per kernel instantiation, per device inside
  B200_CUDA_OK(optin.ensure(sa_tcp_kernel<MODE, PRE, ROWOUT, TRAIN, QUERY>, smem));
  sa_tcp_kernel<MODE, PRE, ROWOUT, TRAIN, QUERY><<<grid, TP_THREADS, smem, stream>>>(p);
  B200_LAUNCH_OK("sa_tcp_kernel");
  return 0;
}

int sa_tcp_launch(TcParams &p, int *tile_counter, cudaStream_t stream) {
  static int force_slots = -1, force_astages = -1;
  if (force_slots < 0) {
    const char *e = getenv("B200_SA_TC_SLOTS");
    force_slots = e ? atoi(e) : 0;
    e = getenv("B200_SA_TC_ASTAGES");
    force_astages = e ? atoi(e) : 0;
  }
  if (p.units) {  // a centre's units meet through atomicMax on zero-initialised outputs
    const int cout = p.L[p.nl - 1].cout;
    B200_CUDA_OK(cudaMemsetAsync(p.out, 0, (size_t)p.B * cout * p.M * sizeof(float), stream));
    if (p.out_pm) B200_CUDA_OK(cudaMemsetAsync(p.out_pm, 0, (size_t)p.B * cout * p.M * sizeof(float), stream));
  }
  p.tiles_per_scene = ceil_div(p.M, p.G);
  p.total_tiles = p.mode == 2 ? ceil_div(p.rows_total, TC_ROWS) : p.B * p.tiles_per_scene;
  p.tile_counter = tile_counter;
  p.final_shfl = 1;
  if (p.query) B200_CHECK_ARG(p.mode == 0 && !p.units && !p.rowout && p.N <= TC_QUERY_MAX_N && (p.N & 3) == 0,
                              "sa_forward(tc): fused ball query needs uncompacted tiles and N <= %d", TC_QUERY_MAX_N);
  const size_t fixed = 1024 + 2 * TC_MAXL * 256 * sizeof(float) + 2 * 4 * 256 * sizeof(float) + 3 * 128 * sizeof(float) +
                       8 * 256 * sizeof(float) + (p.query ? (((size_t)p.N * 12 + 127) & ~(size_t)127) : 0);
  p.a_stages = (force_astages >= 2 && force_astages <= TP_ASTAGES) ? force_astages : 3;
  p.r1_bytes = p.a_stages * 2 * (int)TC_KB_BYTES;  // layer-1 operand ring only: hidden activations live in TMEM
  const size_t rest = fixed + (size_t)p.r1_bytes;
  const size_t budget = 227 * 1024 - 1024;         // dynamic + the kernel's static shared memory (<= 944 B)
  int slots = (int)((budget - rest) / (size_t)p.wslot_bytes);
  p.nslots = slots > TP_MAXSLOTS ? TP_MAXSLOTS : slots;
  if (force_slots >= 2 && force_slots < p.nslots) p.nslots = force_slots;
  B200_CHECK_ARG(p.nslots >= 2, "sa_forward(tc): weight ring does not fit (%d-byte slots)", p.wslot_bytes);
  const size_t smem = rest + (size_t)p.nslots * p.wslot_bytes;
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (grid <= 0) return 0;
  if (p.in_scale || p.stats || p.train_in || p.train_out) {
    B200_CHECK_ARG(p.mode == 2 && p.rowout && !p.pre && p.nl == 1, "training pass: one plain-row layer per launch");
    return launch_variant<2, 0, 1, 1>(p, grid, smem, stream);
  }
  const int key = p.mode * 4 + (p.pre ? 2 : 0) + (p.rowout ? 1 : 0);
  if (p.query) {
    if (key == 0) return launch_variant<0, 0, 0, 0, 1>(p, grid, smem, stream);  // fused ball query -> gather -> MLP -> max
    if (key == 2) return launch_variant<0, 1, 0, 0, 1>(p, grid, smem, stream);  //   with a factorised first layer
  }
  switch (key) {
    case 0: return launch_variant<0, 0, 0>(p, grid, smem, stream);   // ball-query lists -> MLP -> max
    case 2: return launch_variant<0, 1, 0>(p, grid, smem, stream);   //   with a factorised first layer
    case 4: return launch_variant<1, 0, 0>(p, grid, smem, stream);   // three-neighbour blends -> MLP -> max (GridConv sampler)
    case 6: return launch_variant<1, 1, 0>(p, grid, smem, stream);
    case 5: return launch_variant<1, 0, 1>(p, grid, smem, stream);   // blends | skip features -> row output (feature propagation)
    case 9: return launch_variant<2, 0, 1>(p, grid, smem, stream);   // plain rows -> row output (per-point GEMM, 1x1-conv heads)
    default: break;
  }
  set_error("sa_forward(tc): no kernel variant for mode %d pre %d rowout %d", p.mode, p.pre, p.rowout);
  return 1;
}

}  // namespace b200

extern "C" int b200pn2_sa_tensor_work(unsigned long long *mma_n_columns, int reset) {
  using namespace b200;
  B200_CHECK_ARG(mma_n_columns != nullptr, "sa_tensor_work: null pointer");
  B200_CUDA_OK(cudaDeviceSynchronize());
  B200_CUDA_OK(cudaMemcpyFromSymbol(mma_n_columns, g_sa_mma_cols, sizeof(unsigned long long)));
  if (reset) {
    const unsigned long long z = 0;
    B200_CUDA_OK(cudaMemcpyToSymbol(g_sa_mma_cols, &z, sizeof(z)));
  }
  return 0;
}
