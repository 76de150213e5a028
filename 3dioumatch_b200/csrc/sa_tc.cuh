// sa_tc.cuh -- parameter block shared by the tensor-core set-abstraction kernels (sa_tc.cu: one tile per CTA;
// sa_tcp.cu: persistent, warp-specialised).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200_pointnet2.h"

namespace b200 {

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_MAXL = 4;
constexpr uint32_t TC_KB_BYTES = 128 * 128;          // one operand k-block: 128 rows x 128 B
constexpr uint32_t TC_WSTAGE_BYTES = 2 * TC_KB_BYTES;  // W_hi | W_lo

struct TcLayer {
  const float *scale, *shift;
  int cin, cout, nkb, nhalf;
  int rows;           // weight rows per stage = MMA N (cout for hidden layers, <= 128 per half for the last layer)
  size_t packed_off;  // byte offset of this layer's stages in the packed weight buffer
};

struct TcParams {
  int B, N, M, C, ns, G, use_xyz, nl;
  float inv_r;
  const float *xyz, *feat_pm, *new_xyz;
  const int32_t *idx;
  float *out, *out_pm;
  const uint8_t *packed;
  int vec_gather;  // feature rows are 16-byte aligned runs of a multiple of 4 floats
  // mode 1 (feature-propagation style rows, models/grid_conv_module.py:87-108): row (centre g, sample s) is the
  // inverse-distance blend of three source rows, channels [rel xyz (3) | sum_t w_t * feat[idx_t] (C)]
  // shared-memory / TMEM geometry chosen by the launcher
  int r1_bytes;      // activation region: layer-1 A stages, later X_hi | X_lo
  int x_lo_off;      // byte offset of X_lo inside R1 (= hidden k-blocks * 16 KB)
  int wslot_bytes;   // size of one weight stage slot in R2
  int small_off;     // TMEM column offset of the correction-term accumulators (128 or 256)
  int compact;       // 1: the final epilogue's slab aliases R2 (all MMAs finished first) -> ~105 KB, 2 CTAs per SM
  int cluster;       // 2: CTA pairs share every weight stage through one multicast bulk copy (half the L2 reads)
  int mode;
  // persistent kernel (sa_tcp_kernel): weight-ring depth, tile queue
  int nslots, a_stages, total_tiles, tiles_per_scene;
  int final_shfl;  // 1: final max-reduce by warp shuffles (no slab, no CTA barriers per chunk)
  int *tile_counter;
  // compacted mode (sa_tcp.cu): tiles are 8 units of 16 neighbour slots; a centre only contributes the units that
  // hold distinct neighbours (the ball query pads a short list with copies of its first hit: ball_query_gpu.cu:40-46)
  const int *unit_list;    // unit u -> centre * 8 + first slot / 16
  const int *total_units;  // device scalar written by sa_units_kernel
  int units;               // 1: compacted mode
  // fused ball query (QUERY kernels, sa_tcp.cu): idx == NULL, the producers stage the scene's coordinates in shared memory
  // with one cp.async.bulk and find each centre's first-nsample ascending hits themselves (ball_query_gpu.cu:27-46);
  // idx_out (optional) receives the lists
  int query;
  float radius2;
  int32_t *idx_out;
  const int32_t *idx3;   // (B, M*ns, 3)
  const float *w3;       // (B, M*ns, 3)
  const float *rel3;     // (B, M*ns, 3) or NULL
  // factorised first layer (sa_tcp.cu): W1 * [rel xyz | f] = W1x * rel + W1f * f, and W1f * f depends on the source point
  // only.  Pass 1 (mode 2, rowout): plain row GEMM P = scale1 * (W1f * f) + shift1 over all B*N points, rows written
  // point-major without ReLU.  Pass 2 (pre = 1): the producers gather rows of P (or blend three of them in mode 1),
  // add wx[k][c] * rel[k] (wx = scale1 * W1x, [3][128]) and apply the ReLU on their way into the operand ring; the
  // kernel then runs layers 2.. only.
  int pre, rowout, rows_total;
  const float *wx;
  // row output (ROWOUT kernels): ReLU on the last layer or not; channel-major destination `out` is (S, cout,
  // rows_per_scene) with S = rows_total / rows_per_scene; point-major destination `out_pm` is (rows_total, cout)
  int final_relu, rows_per_scene;
  // mode 1, optional second source: row q's own C2 features follow the C blended channels (feature propagation)
  const float *feat2_pm;
  int C2;
  int ld;  // mode 2: row stride of feat_pm in floats (>= C; rows padded to a multiple of 4 floats stay vector-loadable)
  // mode 2, channel-major source: feat_pm is (S, C, rows_per_scene) -- the layout torch's conv stacks hand over
  // (B, C, H, W); row r of scene s reads channel c at feat_pm[(s * C + c) * rows_per_scene + r].  The 32 lanes of a
  // producer warp are 32 consecutive rows, so every channel is one coalesced 128-byte request: no transpose pass.
  int cm_in;
  // training-mode layer passes (TRAIN kernels; sa_train.cu): input rows are the previous layer's raw conv output and
  // become relu(in_scale * z + in_shift) on their way into the operand ring (BatchNorm with batch statistics + ReLU,
  // pytorch_utils.py:42-61); the epilogue also emits per-tile column sums of the output and of its square,
  // stats[((tile * 4 + row quarter) * 2 + {0: sum, 1: sum of squares}) * 256 + column]
  const float *in_scale, *in_shift;
  float *stats;
  // backward passes of the same kernels (da = dz W): the input rows are dz_l, built in the producers from the saved raw
  // conv output z (feat_pm) and the upstream gradient,   dz = in_scale * g + dz_b + dz_c * z   (BatchNorm backward folded to
  // three per-channel coefficients);  train_in 2: g = g_rows (R, ld);  train_in 3 (top layer): g = grad_out of the pooled
  // output where this row is its centre's arg-max slot and the pooled value is positive (in_shift = the layer's shift).
  // train_out 1: the epilogue multiplies by the ReLU mask of the layer below (zprev, out_scale, out_shift) and the column
  // sums become sum g, sum g * xhat (out_mean, out_invstd) -- the next BatchNorm backward's statistics.
  int train_in, train_out, pool_ns;
  const float *g_rows, *gout_pm, *dz_b, *dz_c;
  const int32_t *arg_pm;
  const float *zprev, *out_scale, *out_shift, *out_mean, *out_invstd;
  TcLayer L[TC_MAXL];
};


// sa_tcp.cu: launches the persistent kernel on a fully prepared parameter block (weights already packed, tile counter
// zeroed in stream order, unit list + total already built when p.units is set)
int sa_tcp_launch(TcParams &p, int *tile_counter, cudaStream_t stream);
// compacted-tile bookkeeping: appends the 16-slot units of every centre's neighbour list (sa_tcp.cu)
size_t sa_tcp_unit_list_bytes(int B, int M, int nsample);
bool sa_tcp_units_wanted(int mode, int rowout, int nsample, int B, int M);
int sa_tcp_units_from_idx(int B, int M, int nsample, const int32_t *idx, int *unit_list, int *total, cudaStream_t stream);

// sa_launch.cu: one SharedMLP stack on the tensor-core kernel, from one row source
struct TcCall {
  int mode = 0;  // 0 ball-query lists, 1 three-neighbour blends (+ optional skip rows), 2 plain rows
  int B = 0, N = 0, M = 0, C = 0, ns = 1, use_xyz = 0, normalize_xyz = 0;
  float radius = 1.f;
  const float *xyz = nullptr, *feat_pm = nullptr, *new_xyz = nullptr;
  const int32_t *idx = nullptr;
  const int32_t *idx3 = nullptr;
  const float *w3 = nullptr, *rel3 = nullptr;
  const float *feat2_pm = nullptr;
  int C2 = 0;
  int ld = 0;  // mode 2: row stride (0: C)
  int cm_in = 0;  // mode 2: feat_pm is channel-major (S, C, rows_per_scene), see TcParams::cm_in
  const float *in_scale = nullptr, *in_shift = nullptr;  // training passes: affine + ReLU applied to the input rows
  float *stats = nullptr;                                 // training passes: per-tile column sums (see TcParams)
  int train_in = 0, train_out = 0, pool_ns = 1;           // backward passes (see TcParams)
  const float *g_rows = nullptr, *gout_pm = nullptr, *dz_b = nullptr, *dz_c = nullptr;
  const int32_t *arg_pm = nullptr;
  const float *zprev = nullptr, *out_scale = nullptr, *out_shift = nullptr, *out_mean = nullptr, *out_invstd = nullptr;
  int w_transposed = 0, w_ld = 0;  // single-layer calls: the weight matrix is read transposed (cin x cout view of a (cout, w_ld) matrix)
  int rowout = 0, final_relu = 1, rows_total = 0, rows_per_scene = 0;
  float *out = nullptr, *out_pm = nullptr;
  int num_layers = 0;
  const b200_mlp_layer *layers = nullptr;
  const void *plan = nullptr;      // packed by b200pn2_mlp_plan_build for exactly this stack, or NULL
  size_t plan_bytes = 0;
  int *unit_list = nullptr, *unit_total = nullptr;  // compacted tiles: built by the ball query (or sa_tcp_units_from_idx)
  int query = 0;                 // mode 0 without idx: the kernel's producers run the ball query (see TcParams::query)
  int32_t *idx_out = nullptr;    // optional (B, M, ns) destination of the fused query's lists
};
// can the persistent kernel's producers run the ball query of this stage themselves?  (uncompacted tiles, a scene's
// coordinates fit the shared-memory staging buffer and can be copied with one 16-byte-granular cp.async.bulk)
bool sa_tcp_query_fusable(int B, int N, int M, int nsample, const float *xyz);
constexpr int TC_QUERY_MAX_N = 2048;
bool sa_tc_supported(int C, int nsample, int use_xyz, int num_layers, const b200_mlp_layer *layers, const float *feat_pm);
int sa_tc_run(const TcCall &c, cudaStream_t stream);

}  // namespace b200
