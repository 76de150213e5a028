// sa_launch.cu -- host side of the tensor-core SharedMLP kernels (sa_tcp.cu): geometry of a layer stack, weight packing
// (once per module through a caller-owned "plan", or per call into stream-ordered scratch), and the launch sequences
//   * fused set abstraction        ball-query lists -> SharedMLP -> max            (pointnet2_modules.py:215-277)
//   * GridConv sampler             three-neighbour blends -> SharedMLP -> max      (grid_conv_module.py:87-113)
//   * feature propagation          blends | skip features -> SharedMLP rows        (pointnet2_modules.py:377-422)
//   * row MLP                      plain rows -> SharedMLP rows: the per-point GEMM of a factorised first layer and the
//                                  1x1-conv heads (voting_module.py:38-65, proposal_module.py:98-123)
//
// Precision: the reference is fp32 (parity bar 1e-5), plain TF32 is ~1e-3.  Every GEMM therefore runs as THREE
// kind::tf32 MMAs on operands split as x = hi + lo (hi = x rounded to TF32, lo = x - hi, exact in fp32):
//     D_big += A_hi*W_hi ;  D_small += A_lo*W_hi + A_hi*W_lo ;  D = D_big + D_small   (fp32 accumulation in TMEM)
// The 2^-11-sized correction terms get their own TMEM accumulator columns so that their accumulation rounding is
// negligible; measured error is ~3x an fp32 FFMA GEMM (scripts/tc_precision.py); the lo*lo term (~2^-22) is dropped.
// Weights are split/packed (tc_pack_weights_kernel) into the exact shared-memory image of a pipeline stage
// ([half][k-block][hi|lo][rows x 128 B, 128-byte swizzle]), so a stage is ONE cp.async.bulk with mbarrier completion.
#include <stdlib.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"
#include "tc_common.cuh"
#include "sa_tc.cuh"

namespace b200 {

// ---- weight packing: (cout, cin) fp32 -> [half][kb][hi|lo][128 rows x 128 B, 128-byte swizzle] -----------------------
struct PackParams {
  const float *w[TC_MAXL];
  int cin[TC_MAXL], cout[TC_MAXL], nkb[TC_MAXL], nhalf[TC_MAXL], rows[TC_MAXL];
  int ld[TC_MAXL];  // row stride of the source matrix (== cin unless a column range of a wider matrix is packed)
  int transposed;   // layer 0: element (n, k) of the packed matrix is source[k * ld + n] (backward passes: W^T)
  size_t off[TC_MAXL];
  // grid row nl: for a factorised first layer, wx[k][c] = scale1[c] * W1[c][k] for its three relative-xyz input
  // columns (zero without xyz channels)
  float *wx;
  const float *wx_w, *wx_scale;
  int wx_ld, wx_cout, wx_xyz;
  int nl, perm_c;  // perm_c >= 0: layer 0 column k reads source channel (k < perm_c ? 3 + k : k - perm_c)
};

__global__ void __launch_bounds__(256) tc_pack_weights_kernel(PackParams p, uint8_t *__restrict__ packed) {
  const int l = blockIdx.y;
  if (l >= p.nl) {
    if (blockIdx.x == 0) {
      if (p.wx_w && threadIdx.x < 128) {
        const int c = threadIdx.x;
        for (int k = 0; k < 3; ++k)
          p.wx[k * 128 + c] = (k < p.wx_xyz && c < p.wx_cout) ? p.wx_scale[c] * p.wx_w[(size_t)c * p.wx_ld + k] : 0.f;
      }
    }
    return;
  }
  const int rows = p.rows[l];
  const int items = p.nhalf[l] * p.nkb[l] * rows * 8;  // (half, kb, row, chunk)
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int chunk = it & 7, row = (it >> 3) % rows, rest = (it >> 3) / rows;
    const int kb = rest % p.nkb[l], half = rest / p.nkb[l];
    const int n = half * rows + row;
    float v[4], hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kb * 32 + chunk * 4 + e;
      int src = k;
      if (l == 0 && p.perm_c >= 0) src = k < p.perm_c ? 3 + k : k - p.perm_c;
      v[e] = (n < p.cout[l] && k < p.cin[l])
                 ? ((l == 0 && p.transposed) ? p.w[l][(size_t)src * p.ld[l] + n] : p.w[l][(size_t)n * p.ld[l] + src])
                 : 0.f;
      tc::split_tf32(v[e], hi[e], lo[e]);
    }
    const size_t stage_bytes = (size_t)rows * 256;  // hi rows | lo rows, 128 B each
    uint8_t *stage = packed + p.off[l] + (size_t)(half * p.nkb[l] + kb) * stage_bytes;
    const uint32_t off = tc::sw128_offset(row, chunk);
    *reinterpret_cast<float4 *>(stage + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4 *>(stage + (size_t)rows * 128 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}


static bool env_off(const char *name) {
  const char *e = getenv(name);  // read per call: the parity tests toggle these in-process
  return e && atoi(e) == 0;
}

// ---- geometry of a layer stack: a pure function of the dimensions, shared by the plan builder and every launch --------
struct StackGeom {
  int nl;              // layers the fused / pass-2 kernel runs (all of them, or all but a factorised first one)
  int factor;          // first layer factorised: per-point row GEMM (pass 1) + xyz FMAs in the gather (pass 2)
  int first;           // index of the kernel's first layer in the caller's layer array (factor ? 1 : 0)
  int nkb[TC_MAXL], nhalf[TC_MAXL], rows[TC_MAXL];
  size_t off[TC_MAXL];
  int wslot_bytes;     // weight-ring slot: W_hi or W_lo of one k-block of the widest stage
  int nkb1, rows1, wslot1_bytes;
  size_t off1;         // pass-1 weights (feature columns of the first layer)
  size_t wx_off;       // float [3][128]: scale1 * W1x
  size_t plan_bytes;
};

// rowout: the last layer may have any width (padded to a multiple of 32, <= 256); otherwise (max-pooled stacks) widths
// are multiples of 32, hidden <= 128, last <= 128 or == 256.
static bool stack_geom(int C_feat, int use_xyz, int num_layers, const b200_mlp_layer *layers, int rowout, bool allow_factor,
                       StackGeom &g) {
  if (num_layers < 1 || num_layers > TC_MAXL) return false;
  if (layers[0].cin != C_feat + (use_xyz ? 3 : 0)) return false;
  for (int l = 0; l < num_layers; ++l) {
    const bool last = l == num_layers - 1;
    const int co = layers[l].cout;
    if (co < 1) return false;
    if (l > 0 && layers[l].cin != layers[l - 1].cout) return false;
    if (last && rowout) {
      if (co > 256) return false;
    } else {
      if (co % 32 != 0) return false;
      if (last ? !(co <= 128 || co == 256) : co > 128) return false;
    }
  }
  const int H1 = layers[0].cout;
  g.factor = (allow_factor && !rowout && !env_off("B200_SA_TC_FACTOR") && num_layers >= 3 && C_feat >= 32 && (C_feat & 3) == 0 && H1 <= 128) ? 1 : 0;
  g.first = g.factor;
  g.nl = num_layers - g.first;
  if (g.factor && g.nl > TC_MAXL - 1) return false;
  size_t off = 0;
  int slot = 0;
  for (int i = 0; i < g.nl; ++i) {
    const b200_mlp_layer &L = layers[g.first + i];
    const bool last = i == g.nl - 1;
    const int cin = (i == 0 && g.factor) ? H1 : L.cin;
    g.nkb[i] = (cin + 31) / 32;
    if (last) {
      g.nhalf[i] = L.cout > 128 ? 2 : 1;
      g.rows[i] = (((L.cout + g.nhalf[i] - 1) / g.nhalf[i]) + 31) & ~31;
    } else {
      g.nhalf[i] = 1;
      g.rows[i] = L.cout;
    }
    g.off[i] = off;
    off += (size_t)g.nhalf[i] * g.nkb[i] * (size_t)g.rows[i] * 256;
    slot = g.rows[i] > slot ? g.rows[i] : slot;
  }
  g.wslot_bytes = slot * 128;
  g.nkb1 = g.rows1 = g.wslot1_bytes = 0;
  g.off1 = off;
  if (g.factor) {
    g.nkb1 = (C_feat + 31) / 32;
    g.rows1 = H1;
    g.wslot1_bytes = H1 * 128;
    off += (size_t)g.nkb1 * H1 * 256;
  }
  g.wx_off = (off + 255) & ~(size_t)255;
  g.plan_bytes = g.wx_off + 3 * 128 * sizeof(float);
  return true;
}

// Can the tensor-core kernel take this max-pooled stage?  (nsample divides 128, >= 2 layers, aligned point-major features)
bool sa_tc_supported(int C, int nsample, int use_xyz, int num_layers, const b200_mlp_layer *layers, const float *feat_pm) {
  if (env_off("B200_SA_TC")) return false;
  if (nsample < 8 || nsample > 128 || (128 % nsample) != 0) return false;
  if (num_layers < 2) return false;
  if (C > 0 && !feat_pm) return false;
  StackGeom g;
  return stack_geom(C, use_xyz, num_layers, layers, 0, true, g);
}

// ---- plan: the packed weights of a stack in device memory the CALLER owns (one per module, rebuilt when weights change) --
static int pack_into(const StackGeom &g, int C_feat, int use_xyz, int perm_xyz_last, const b200_mlp_layer *layers,
                     uint8_t *plan, cudaStream_t stream, int w_transposed = 0, int w_ld = 0) {
  PackParams pk = {};
  pk.transposed = w_transposed;
  pk.nl = g.nl;
  // the kernel's layer-1 operand is [features (C) | rel xyz (3)] while the reference concatenates [xyz | features]
  // (pointnet2_utils.py:358-360): permute the first layer's columns while packing
  pk.perm_c = (use_xyz && !g.factor && perm_xyz_last) ? C_feat : -1;
  for (int i = 0; i < g.nl; ++i) {
    const b200_mlp_layer &L = layers[g.first + i];
    pk.w[i] = L.weight;
    pk.cin[i] = (i == 0 && g.factor) ? layers[0].cout : L.cin;
    pk.ld[i] = L.cin;
    pk.cout[i] = L.cout;
    pk.nkb[i] = g.nkb[i]; pk.nhalf[i] = g.nhalf[i]; pk.rows[i] = g.rows[i]; pk.off[i] = g.off[i];
    if (i == 0 && w_ld > 0) pk.ld[i] = w_ld;
  }
  pk.wx = reinterpret_cast<float *>(plan + g.wx_off);
  if (g.factor) {
    const int l = g.nl, cin0 = layers[0].cin;
    pk.nl = g.nl + 1;
    pk.w[l] = layers[0].weight + (cin0 - C_feat);  // skip the xyz columns
    pk.cin[l] = C_feat; pk.ld[l] = cin0; pk.cout[l] = layers[0].cout; pk.nkb[l] = g.nkb1; pk.nhalf[l] = 1;
    pk.rows[l] = g.rows1; pk.off[l] = g.off1;
    pk.wx_w = layers[0].weight; pk.wx_scale = layers[0].scale; pk.wx_ld = cin0; pk.wx_cout = layers[0].cout;
    pk.wx_xyz = cin0 - C_feat;
  }
  tc_pack_weights_kernel<<<dim3(32, pk.nl + 1), 256, 0, stream>>>(pk, plan);  // last grid row: wx
  B200_LAUNCH_OK("tc_pack_weights_kernel");
  return 0;
}

// ---- one stack, one row source, one launch sequence (TcCall: sa_tc.cuh) -------------------------------------------------
int sa_tc_run(const TcCall &c, cudaStream_t stream) {
  StackGeom g;
  const int C_stack = c.C + c.C2;
  // a first layer can only be factorised over aligned point-major feature rows (and never for plain-row stacks)
  const bool can_factor = c.mode != 2 && c.C2 == 0 && c.feat_pm && (((uintptr_t)c.feat_pm) & 15) == 0 &&
                          (long long)c.B * c.N < (1ll << 30);
  B200_CHECK_ARG(stack_geom(C_stack, c.use_xyz, c.num_layers, c.layers, c.rowout, can_factor, g),
                 "tensor-core MLP: layer widths not supported");
  const void *plan_in = c.plan;
  if (plan_in) {
    // the plan was packed for the geometry its builder assumed (factorised whenever the widths allow it); if this call
    // cannot run that geometry the plan does not apply and the weights are packed per call
    StackGeom gp;
    if (!stack_geom(C_stack, c.use_xyz, c.num_layers, c.layers, c.rowout, c.mode != 2, gp) || gp.factor != g.factor ||
        c.plan_bytes < g.plan_bytes)
      plan_in = nullptr;
  }
  const int H1 = c.layers[0].cout;
  const int rows1 = g.factor ? c.B * c.N : 0;

  // scratch: [tile counters (256 B) | packed weights + wx (no plan) | P (factorised first layer)]
  const size_t plan_off = 256;
  const size_t p_off = plan_off + (plan_in ? 0 : ((g.plan_bytes + 255) & ~(size_t)255));
  const size_t p_bytes = g.factor ? (((size_t)rows1 * H1 * sizeof(float) + 255) & ~(size_t)255) : 0;
  ScratchGuard scratch;
  B200_CUDA_OK(scratch.alloc(p_off + p_bytes, stream));
  uint8_t *sc = (uint8_t *)scratch.ptr;
  int *counters = reinterpret_cast<int *>(sc);  // [0]: fused / pass 2, [1]: pass 1
  B200_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(int), stream));
  const uint8_t *plan = (const uint8_t *)plan_in;
  if (!plan) {
    const int rc = pack_into(g, C_stack, c.use_xyz, c.mode != 2, c.layers, sc + plan_off, stream, c.w_transposed, c.w_ld);
    if (rc) return rc;
    plan = sc + plan_off;
  }

  TcParams p = {};
  p.mode = c.mode; p.pre = g.factor; p.rowout = c.rowout; p.final_relu = c.final_relu;
  p.B = c.B; p.N = c.N; p.M = c.M; p.ns = c.ns; p.G = TC_ROWS / c.ns; p.use_xyz = c.use_xyz ? 1 : 0;
  p.C = g.factor ? H1 : c.C;
  p.C2 = c.C2; p.feat2_pm = c.feat2_pm;
  p.nl = g.nl;
  p.inv_r = c.normalize_xyz ? (float)(1.0 / (double)c.radius) : 1.0f;
  p.xyz = c.xyz; p.feat_pm = c.feat_pm; p.new_xyz = c.new_xyz; p.idx = c.idx;
  p.idx3 = c.idx3; p.w3 = c.w3; p.rel3 = c.rel3;
  p.out = c.out; p.out_pm = c.out_pm;
  p.rows_total = c.rows_total; p.rows_per_scene = c.rows_per_scene > 0 ? c.rows_per_scene : 1;
  p.ld = c.ld > 0 ? c.ld : c.C;
  p.cm_in = (c.mode == 2 && c.cm_in) ? 1 : 0;
  p.in_scale = c.in_scale; p.in_shift = c.in_shift; p.stats = c.stats;
  p.train_in = c.train_in; p.train_out = c.train_out; p.pool_ns = c.pool_ns > 0 ? c.pool_ns : 1;
  p.g_rows = c.g_rows; p.gout_pm = c.gout_pm; p.dz_b = c.dz_b; p.dz_c = c.dz_c; p.arg_pm = c.arg_pm;
  p.zprev = c.zprev; p.out_scale = c.out_scale; p.out_shift = c.out_shift; p.out_mean = c.out_mean; p.out_invstd = c.out_invstd;
  // 16-byte loads of whole 4-channel groups: aligned base and row stride; a ragged tail (C % 4) goes through scalar loads
  p.vec_gather = g.factor ? 1 : ((c.C >= 4 && (p.ld & 3) == 0 && (c.mode == 2 || (c.C & 3) == 0) &&
                                  ((((uintptr_t)c.feat_pm) & 15) == 0)) ? 1 : 0);
  if (p.cm_in) p.vec_gather = 0;
  p.packed = plan;
  p.wslot_bytes = g.wslot_bytes;
  p.small_off = 128;
  p.wx = reinterpret_cast<const float *>(plan + g.wx_off);
  for (int i = 0; i < g.nl; ++i) {
    const b200_mlp_layer &L = c.layers[g.first + i];
    TcLayer &t = p.L[i];
    t.scale = L.scale; t.shift = L.shift;
    t.cin = (i == 0 && g.factor) ? H1 : L.cin;
    t.cout = L.cout; t.nkb = g.nkb[i]; t.nhalf = g.nhalf[i]; t.rows = g.rows[i]; t.packed_off = g.off[i];
  }
  if (g.factor) {
    // pass 1: P = scale1 * (W1f * f) + shift1, one row per source point, no ReLU
    float *P = reinterpret_cast<float *>(sc + p_off);
    TcParams q = {};
    q.mode = 2; q.rowout = 1; q.final_relu = 0; q.rows_total = rows1; q.rows_per_scene = rows1;
    q.B = 1; q.N = rows1; q.M = rows1; q.C = c.C; q.ns = 32; q.G = TC_ROWS / 32; q.nl = 1;
    q.inv_r = 1.0f; q.feat_pm = c.feat_pm; q.vec_gather = 1; q.out_pm = P; q.ld = c.C;
    q.packed = plan; q.wslot_bytes = g.wslot1_bytes; q.small_off = 128;
    q.L[0].scale = c.layers[0].scale; q.L[0].shift = c.layers[0].shift; q.L[0].cin = c.C; q.L[0].cout = H1;
    q.L[0].nkb = g.nkb1; q.L[0].nhalf = 1; q.L[0].rows = g.rows1; q.L[0].packed_off = g.off1;
    const int rc1 = sa_tcp_launch(q, counters + 1, stream);
    if (rc1) return rc1;
    p.feat_pm = P;
  }
  p.query = (c.mode == 0 && c.query) ? 1 : 0;
  p.radius2 = c.radius * c.radius;  // ball_query_gpu.cu:27, one fp32 multiply
  p.idx_out = c.idx_out;
  p.units = 0;
  if (c.unit_list && c.unit_total) {
    p.units = 1; p.unit_list = c.unit_list; p.total_units = c.unit_total;
  }
  return sa_tcp_launch(p, counters, stream);
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200pn2_mlp_plan_bytes(int C_feat, int use_xyz, int num_layers, const b200_mlp_layer *layers, int row_output,
                                         int plain_rows) {
  StackGeom g;
  if (!layers || !stack_geom(C_feat, use_xyz, num_layers, layers, row_output, !plain_rows, g)) return 0;
  return g.plan_bytes;
}

extern "C" int b200pn2_mlp_plan_build(int C_feat, int use_xyz, int num_layers, const b200_mlp_layer *layers, int row_output,
                                      int plain_rows, void *plan, size_t plan_bytes, b200_stream_t stream_) {
  StackGeom g;
  B200_CHECK_ARG(layers && plan, "mlp_plan_build: null pointer");
  B200_CHECK_ARG(stack_geom(C_feat, use_xyz, num_layers, layers, row_output, !plain_rows, g),
                 "mlp_plan_build: layer widths not supported");
  B200_CHECK_ARG(plan_bytes >= g.plan_bytes, "mlp_plan_build: plan buffer too small (%zu < %zu bytes)", plan_bytes, g.plan_bytes);
  return pack_into(g, C_feat, use_xyz, plain_rows ? 0 : 1, layers, (uint8_t *)plan, (cudaStream_t)stream_);
}
