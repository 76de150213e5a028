/*
 * b200_pointnet2.h -- C ABI of the B200-native PointNet++ set-abstraction / feature-propagation
 * operators (libb200pc.so).  Plain device pointers + sizes + a CUDA stream; no torch types.
 *
 * Each entry replaces one function of the reference's pybind module `pointnet2._ext`
 * (pointnet2/_ext_src/src/bindings.cpp:11-24); the line cited at each prototype is the
 * reference host wrapper whose behaviour it reproduces.  All tensors are contiguous, fp32
 * data / int32 indices, resident on the current CUDA device.
 *
 * Return value: 0 on success, non-zero on error (b200_last_error() gives the message; the
 * library never calls exit(), unlike cuda_utils.h:35-44 of the reference).
 * Launches are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 * default stream), like the reference which launches on the current torch stream.
 */
#ifndef B200_POINTNET2_H
#define B200_POINTNET2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *b200_stream_t; /* cudaStream_t */

/* library identification / error reporting */
int b200_abi_version(void);
const char *b200_last_error(void);
/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
unsigned long long b200_launch_count(void);

/* Launch-shape policy of furthest_point_sampling (no reference counterpart: the reference always runs one 512-thread
 * block per scene, sampling_gpu.cu:180-230).  0 = latency (default): the cluster shape with the shortest serial chain,
 * right for a call that runs alone on its stream.  1 = throughput: the shape with the least SM-time (fewer, fuller
 * CTAs per scene), right when other kernels of a pipelined step share the GPU.  Results are identical (bit-exact) for
 * every shape.  Returns the previous policy; the B200_FPS_POLICY environment variable sets the initial value.      */
int b200pn2_fps_set_policy(int policy);
/* Tuning / test hook (no reference counterpart): force the FPS kernel generation (0 = fps_owner_kernel wherever it applies,
 * 2 = fps_cluster_kernel of round 1; -1 = default: owner for cluster shapes, cluster kernel for one CTA) and the launch shape
 * (cluster size, threads per CTA; 0 = cost model).  A shape that cannot hold the cloud in registers makes the next call
 * fail with an argument error.  Results are bit-identical for every choice (tests/test_gpu_pointops.py).            */
int b200pn2_fps_force_shape(int kernel, int cluster, int threads);
/* Prefix speculation of furthest_point_sampling for clouds of at most 4096 points (no reference counterpart).  The
 * reference samples hierarchically from points that are already in furthest-point order (pointnet2_modules.py:238-247
 * on the previous level's new_xyz), for which the answer is 0..m-1 unless a tie or the origin-skip rule intervenes; two
 * parallel kernels verify exactly that with the reference's semantics (sampling_gpu.cu:105-114, 64-70) and the serial
 * kernel behind them returns at once for verified scenes, runs in full for the others.  mode 1 on, 0 off, -1 default
 * (on, or B200_FPS_PREFIX).  Returns the previous mode.  Results are bit-identical either way.                        */
int b200pn2_fps_set_prefix_speculation(int mode);

/* furthest_point_sampling(points (B,N,3), nsamples) -> idx (B,m) int32
 * reference: sampling.cpp:70-91 + sampling_gpu.cu:74-234.
 * idx[b][0] = 0; tie order = (min-dist desc, bit-reversed (k mod bs) asc, k asc) with
 * bs = opt_n_threads(N) (cuda_utils.h:18-24); points with x^2+y^2+z^2 <= 1e-3 never compete.
 * `scratch` may be NULL; otherwise >= B*N floats used only by the large-N fallback
 * (the reference's `tmp` tensor).                                                          */
int b200pn2_furthest_point_sampling(int B, int N, int m, const float *xyz, int32_t *idx, float *scratch,
                                    b200_stream_t stream);

/* gather_points(points (B,C,N), idx (B,m)) -> out (B,C,m)        reference: sampling.cpp:20-44 */
int b200pn2_gather_points(int B, int C, int N, int m, const float *points, const int32_t *idx, float *out,
                          b200_stream_t stream);
/* gather_points_grad(grad_out (B,C,m), idx, n) -> grad_points (B,C,N); the callee zero-fills
 * grad_points first.                                             reference: sampling.cpp:46-69 */
int b200pn2_gather_points_grad(int B, int C, int N, int m, const float *grad_out, const int32_t *idx,
                               float *grad_points, b200_stream_t stream);

/* ball_query(new_xyz (B,M,3), xyz (B,N,3), radius, nsample) -> idx (B,M,nsample) int32
 * first `nsample` hits in ascending point index, padded with the first hit, zeros if empty.
 * reference: ball_query.cpp:13-37 + ball_query_gpu.cu:14-59                                */
int b200pn2_ball_query(int B, int N, int M, float radius, int nsample, const float *new_xyz, const float *xyz,
                       int32_t *idx, b200_stream_t stream);

/* group_points(points (B,C,N), idx (B,M,ns)) -> out (B,C,M,ns)   reference: group_points.cpp:17-39 */
int b200pn2_group_points(int B, int C, int N, int M, int ns, const float *points, const int32_t *idx, float *out,
                         b200_stream_t stream);
/* group_points_grad(grad_out (B,C,M,ns), idx, n) -> grad_points (B,C,N), zero-filled by the callee.
 * reference: group_points.cpp:41-65                                                          */
int b200pn2_group_points_grad(int B, int C, int N, int M, int ns, const float *grad_out, const int32_t *idx,
                              float *grad_points, b200_stream_t stream);

/* three_nn(unknown (B,n,3), known (B,m,3)) -> dist2 (B,n,3) f32 ascending, idx (B,n,3) int32
 * reference: interpolate.cpp:19-45 + interpolate_gpu.cu:14-73 (squared distances; the Python
 * layer takes the sqrt, pointnet2_utils.py:143)                                               */
int b200pn2_three_nn(int B, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx,
                     b200_stream_t stream);

/* three_interpolate(points (B,C,m), idx (B,n,3), weight (B,n,3)) -> out (B,C,n)
 * reference: interpolate.cpp:47-74 + interpolate_gpu.cu:77-116                                */
int b200pn2_three_interpolate(int B, int C, int m, int n, const float *points, const int32_t *idx,
                              const float *weight, float *out, b200_stream_t stream);
/* three_interpolate_grad(grad_out (B,C,n), idx, weight, m) -> grad_points (B,C,m), zero-filled by
 * the callee.  Implements the scatter-add gradient of interpolate_gpu.cu:121-148 (the reference host
 * wrapper interpolate.cpp:95 launches the forward kernel by mistake; deliberate divergence).   */
int b200pn2_three_interpolate_grad(int B, int C, int n, int m, const float *grad_out, const int32_t *idx,
                                   const float *weight, float *grad_points, b200_stream_t stream);

/* ---- deterministic gradients (next-row n4 of SURVEY.md section 8f) -------------------------------------------------
 * Same results as the three *_grad entries above up to fp32 summation order, but the order is fixed: every target
 * point sums the entries that feed it in ascending entry position (the order of a sequential loop), so the output is
 * bit-reproducible -- the reference's atomicAdd kernels (sampling_gpu.cu:39-52, group_points_gpu.cu:48-68,
 * interpolate_gpu.cu:121-148) are not.  `workspace` >= b200pn2_scatter_det_workspace(B, targets per scene, entries per
 * scene) bytes of device memory: targets = N (gather, group) or m (interpolate); entries = m (gather), M*ns (group),
 * 3*n (interpolate).                                                                                              */
size_t b200pn2_scatter_det_workspace(int B, int n_targets, int entries_per_scene);
int b200pn2_gather_points_grad_det(int B, int C, int N, int m, const float *grad_out, const int32_t *idx,
                                   float *grad_points, void *workspace, size_t workspace_bytes, b200_stream_t stream);
int b200pn2_group_points_grad_det(int B, int C, int N, int M, int ns, const float *grad_out, const int32_t *idx,
                                  float *grad_points, void *workspace, size_t workspace_bytes, b200_stream_t stream);
int b200pn2_three_interpolate_grad_det(int B, int C, int n, int m, const float *grad_out, const int32_t *idx,
                                       const float *weight, float *grad_points, void *workspace,
                                       size_t workspace_bytes, b200_stream_t stream);

/* ---- fused set-abstraction forward (the wide entry; no single reference counterpart) ------------
 * Replaces the sequence  ball_query -> group_points(xyz) -> (-centre, *1/r) -> group_points(features)
 * -> concat -> SharedMLP (1x1 conv + eval-mode BN + ReLU, up to 3 layers) -> max over nsample
 * of PointnetSAModuleVotes.forward (pointnet2_modules.py:215-277, pointnet2_utils.py:318-377,
 * pytorch_utils.py:14-39) with grouped tensors kept on chip.
 *
 *   xyz      (B,N,3)   features (B,C,N) or NULL (C=0)   new_xyz (B,M,3)
 *   weights  layer l: W_l (cout_l, cin_l) row-major fp32 (nn.Conv2d.weight viewed 2-D),
 *            scale_l / shift_l (cout_l): y = relu(scale * (W x) + shift)  (BN eval folded to an
 *            affine; scale=1, shift=bias for bn=False)
 *   cin_0 = (use_xyz ? 3 : 0) + C ; channel order [dx,dy,dz, features...] as torch.cat in
 *            pointnet2_utils.py:358-360
 *   normalize_xyz: multiply relative xyz by fp32 (1/radius)                                      */
typedef struct {
  int cin;
  int cout;
  const float *weight; /* (cout, cin) */
  const float *scale;  /* (cout) */
  const float *shift;  /* (cout) */
} b200_mlp_layer;

/*   features     (B,C,N) channel-major (the reference layout) or NULL
 *   features_pm  (B,N,C) point-major copy of the same data or NULL; when NULL and C % 4 == 0 the library
 *                transposes `features` into `workspace` so that one grouped row is one aligned run
 *   idx_in       (B,M,nsample) neighbour indices or NULL (NULL: the library runs the ball query)
 *   out          (B,cout_last,M);  out_pm (B,M,cout_last) or NULL;  idx_out (B,M,nsample) or NULL
 *   workspace    device scratch of at least b200pn2_sa_forward_workspace(...) bytes (may be NULL if 0)
 * ReLU follows every layer (pytorch_utils.py:21, activation=nn.ReLU).                                   */
size_t b200pn2_sa_forward_workspace(int B, int N, int M, int C, int nsample, int have_features_pm, int have_idx);

int b200pn2_sa_forward(int B, int N, int M, int C, float radius, int nsample, int use_xyz, int normalize_xyz,
                       const float *xyz, const float *features, const float *features_pm, const float *new_xyz,
                       const int32_t *idx_in, int num_layers, const b200_mlp_layer *layers, float *out,
                       float *out_pm, int32_t *idx_out, void *workspace, size_t workspace_bytes,
                       b200_stream_t stream);

/* ---- fused feature-propagation rows -> SharedMLP -> max (next-row n1 of SURVEY.md section 8f) -------------------
 * Replaces, for the IoU branch (models/grid_conv_module.py:87-113):
 *     three_interpolate-style blend of 3 neighbours -> cat([relative grid xyz, blended feats]) -> SharedMLP -> max over
 *     the nsample grid points of each box
 * Row (centre g, sample s) = [rel_xyz (3) | sum_t weight3[.,t] * known_feats[:, idx3[.,t]] (C)], q = g*nsample + s.
 *   known_feats (B,C,m) channel-major and/or known_feats_pm (B,m,C) point-major (one may be NULL; with only the
 *   channel-major tensor the library transposes into `workspace`, >= B*m*C*4 bytes rounded up to 256)
 *   idx3 (B, M*nsample, 3) int32, weight3 (B, M*nsample, 3), rel_xyz (B, M*nsample, 3) or NULL (no xyz channels)
 *   out (B, cout_last, M).  Runs on the tcgen05 kernel; widths must be multiples of 32 (hidden <= 128, last <= 128 or 256),
 *   nsample a divisor of 128 (>= 8), C a multiple of 4.                                                               */
int b200pn2_interp_mlp_forward(int B, int m_known, int M, int nsample, int C, const float *known_feats,
                               const float *known_feats_pm, const int32_t *idx3, const float *weight3,
                               const float *rel_xyz, int num_layers, const b200_mlp_layer *layers, float *out,
                               void *workspace, size_t workspace_bytes, b200_stream_t stream);


/* ---- packed-weight plans -------------------------------------------------------------------------------------------
 * The tensor-core kernels consume SharedMLP weights split into TF32 hi/lo parts and laid out as the shared-memory image
 * of their pipeline stages.  Packing is a function of the weights only, so a module packs ONCE into device memory it
 * owns (a "plan") and passes it to the *_planned / row entries below; with plan == NULL the library packs into
 * stream-ordered scratch on every call.  A plan is valid for exactly the (C_feat, use_xyz, layers, row_output,
 * plain_rows) it was built for and until the weights / folded BN affine change (the first layer's scale is baked in
 * when that layer is factorised).
 *   row_output  0: max-pooled stacks (sa_forward, interp_mlp_forward)   1: row stacks (fp_rows_forward, row_mlp_forward)
 *   plain_rows  1: the stack reads plain rows (row_mlp_forward): no xyz column permutation, never factorised
 * b200pn2_mlp_plan_bytes returns 0 when the tensor-core kernel does not take the stack.                              */
size_t b200pn2_mlp_plan_bytes(int C_feat, int use_xyz, int num_layers, const b200_mlp_layer *layers, int row_output,
                              int plain_rows);
int b200pn2_mlp_plan_build(int C_feat, int use_xyz, int num_layers, const b200_mlp_layer *layers, int row_output,
                           int plain_rows, void *plan, size_t plan_bytes, b200_stream_t stream);

/* b200pn2_sa_forward / b200pn2_interp_mlp_forward with a caller-owned plan (NULL, 0: pack per call).                 */
int b200pn2_sa_forward_planned(int B, int N, int M, int C, float radius, int nsample, int use_xyz, int normalize_xyz,
                               const float *xyz, const float *features, const float *features_pm, const float *new_xyz,
                               const int32_t *idx_in, int num_layers, const b200_mlp_layer *layers, float *out,
                               float *out_pm, int32_t *idx_out, void *workspace, size_t workspace_bytes,
                               const void *plan, size_t plan_bytes, b200_stream_t stream);
int b200pn2_interp_mlp_forward_planned(int B, int m_known, int M, int nsample, int C, const float *known_feats,
                                       const float *known_feats_pm, const int32_t *idx3, const float *weight3,
                                       const float *rel_xyz, int num_layers, const b200_mlp_layer *layers, float *out,
                                       void *workspace, size_t workspace_bytes, const void *plan, size_t plan_bytes,
                                       b200_stream_t stream);

/* ---- feature propagation rows -> SharedMLP rows (PointnetFPModule.forward, pointnet2_modules.py:377-422) -------------
 * Replaces three_interpolate -> torch.cat([interpolated, unknow_feats]) -> SharedMLP (1x1 conv + BN + ReLU):
 * row q of scene b = [sum_t weight3[b,q,t] * known_feats_pm[b, idx3[b,q,t], :] (C2) | skip_feats_pm[b,q,:] (C1)]
 * runs through the stack; one output row per q.  Blend rounding = three_interpolate's (interpolate_gpu.cu:100-104).
 *   known_feats_pm (B,m_known,C2), skip_feats_pm (B,n,C1) or NULL (C1 = 0), idx3 / weight3 (B,n,3)
 *   out (B,cout,n) channel-major and/or out_pm (B,n,cout) point-major (either may be NULL)
 *   hidden widths multiples of 32 and <= 128, last width <= 256 (wider stacks: one call per layer, chained through
 *   b200pn2_row_mlp_forward);  relu_last: ReLU after the last layer of THIS call.                                    */
int b200pn2_fp_rows_forward(int B, int n, int m_known, int C2, int C1, const float *known_feats_pm,
                            const float *skip_feats_pm, const int32_t *idx3, const float *weight3, int num_layers,
                            const b200_mlp_layer *layers, int relu_last, float *out, float *out_pm, const void *plan,
                            size_t plan_bytes, b200_stream_t stream);

/* ---- row MLP: 1x1-conv stacks on plain rows (FP layers 2.., voting_module.py:38-65, proposal_module.py:98-123,
 * grid_conv_module.py:108-115) ------------------------------------------------------------------------------------
 *   x_pm (S*R, ld) point-major rows of C channels (ld = 0: C; rows whose stride is a multiple of 4 floats on a
 *   16-byte-aligned base are read with vector loads), S scenes of R rows;  out (S,cout,R) channel-major and/or
 *   out_pm (S*R,cout).                                                                                              */
int b200pn2_row_mlp_forward(int S, int R, int C, int ld, const float *x_pm, int num_layers,
                            const b200_mlp_layer *layers, int relu_last, float *out, float *out_pm, const void *plan,
                            size_t plan_bytes, b200_stream_t stream);
/* The same stack on channel-major input x_cm (S, C, R) -- the (B, C, H, W) tensor an eval-mode SharedMLP receives
 * (pytorch_utils.py:14-39), read in place by the kernel's producers: no transpose pass.                             */
int b200pn2_row_mlp_forward_cm(int S, int R, int C, const float *x_cm, int num_layers, const b200_mlp_layer *layers,
                               int relu_last, float *out, float *out_pm, const void *plan, size_t plan_bytes,
                               b200_stream_t stream);

/* (B,C,N) channel-major -> (B,N,out_ld) point-major (the layout the fused kernels gather from); out_ld = 0: C,
 * otherwise >= C with zero-filled padding columns.                                                                  */
int b200pn2_transpose_cn(int B, int C, int N, const float *in_cm, float *out_pm, int out_ld, b200_stream_t stream);

/* Tensor-pipe work issued by the fused kernels since the last reset: sum over tcgen05.mma instructions of their N
 * (each is a 128 x N x 8 TF32 multiply-add block = 2048*N FLOP).  Synchronises the device.  bench.py: executed TF32
 * FLOP/s of the roofline.                                                                                          */
int b200pn2_sa_tensor_work(unsigned long long *mma_n_columns, int reset);

/* ---- training-mode set-abstraction MLP: conv1x1 -> BatchNorm2d on BATCH statistics -> ReLU per layer,
 * max over nsample, and the backward of all of it (pointnet2/pytorch_utils.py:14-61,70-123; pointnet2_modules.py:256-262;
 * input-feature gradient = group_points_grad, group_points_gpu.cu:48-68).  Layer at a time; only the raw conv outputs of
 * each layer are kept for backward (`saved`), BatchNorm / ReLU / max-pool are fused into the loads and stores of the GEMMs.
 *   weight (cout, cin), gamma / beta (cout); running_mean / running_var (cout) are updated in place with torch's rule
 *   (momentum, unbiased variance) or may be NULL.  Widths: multiples of 4, <= 256; cin <= 320; 1..4 layers.
 *   Gather mode  (x_rows == NULL): rows = grouped [rel xyz * 1/r (3) | features (C)] of idx (B,M,nsample), features_pm
 *                (B,N,C) point-major; backward returns grad_features (B,C,N) (xyz is treated as a constant).
 *   Rows mode    (x_rows != NULL): the stack on given rows (B*M*nsample, C), C a multiple of 4; backward returns grad_rows.
 *   out (B, cout_last, M).  `saved` >= b200pn2_sa_train_saved_bytes(...), `workspace` >= ..._workspace_bytes(..., backward).
 * Weight / gamma / beta gradients and the statistics are reduced in a fixed order (bit-reproducible).               */
typedef struct {
  int cin;
  int cout;
  const float *weight;
  const float *gamma;
  const float *beta;
  float *running_mean;
  float *running_var;
} b200_bn_layer;

size_t b200pn2_sa_train_saved_bytes(int B, int M, int nsample, int C, int use_xyz, int num_layers,
                                    const b200_bn_layer *layers, int rows_mode);
size_t b200pn2_sa_train_workspace_bytes(int B, int M, int nsample, int C, int use_xyz, int num_layers,
                                        const b200_bn_layer *layers, int rows_mode, int backward);
int b200pn2_sa_train_forward(int B, int N, int M, int C, float radius, int nsample, int use_xyz, int normalize_xyz,
                             const float *xyz, const float *features_pm, const float *new_xyz, const int32_t *idx,
                             const float *x_rows, int num_layers, const b200_bn_layer *layers, float eps, float momentum,
                             float *out, void *saved, size_t saved_bytes, void *workspace, size_t workspace_bytes,
                             b200_stream_t stream);
int b200pn2_sa_train_backward(int B, int N, int M, int C, int nsample, int use_xyz, const int32_t *idx,
                              const float *x_rows, int num_layers, const b200_bn_layer *layers, const float *grad_out,
                              const void *saved, size_t saved_bytes, float *grad_features, float *grad_rows,
                              float *const *grad_weight, float *const *grad_gamma, float *const *grad_beta,
                              void *workspace, size_t workspace_bytes, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_POINTNET2_H */
