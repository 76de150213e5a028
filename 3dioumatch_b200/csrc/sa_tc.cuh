// sa_tc.cuh -- parameter block shared by the tensor-core set-abstraction kernels (sa_tc.cu: one tile per CTA;
// sa_tcp.cu: persistent, warp-specialised).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
  TcLayer L[TC_MAXL];
};


// sa_tcp.cu: launches the persistent kernel on a fully prepared parameter block (weights already packed)
int sa_tcp_launch(TcParams &p, int *tile_counter, int *unit_scratch, cudaStream_t stream);
size_t sa_tcp_unit_scratch_bytes(int B, int M, int nsample);

}  // namespace b200
