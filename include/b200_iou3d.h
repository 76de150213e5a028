/*
 * b200_iou3d.h -- C ABI of the B200-native rotated 3D-IoU / NMS operator (libb200pc.so).
 *
 * Each entry replaces one function of the reference's pybind module
 * `pcdet.ops.iou3d_nms.iou3d_nms_cuda` (OpenPCDet/pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17).
 * Boxes are (n,7) fp32 rows [x, y, z, dx, dy, dz, heading], contiguous.
 * Return value: 0 on success, non-zero on error (b200_last_error()); never exit()s, unlike
 * iou3d_nms.cpp:14-38.  Launches go to `stream` (the reference uses the legacy default stream).
 */
#ifndef B200_IOU3D_H
#define B200_IOU3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *b200_stream_t; /* cudaStream_t */

/* boxes_overlap_bev_gpu(boxes_a (N,7), boxes_b (M,7), ans_overlap (N,M))
 * reference: iou3d_nms.cpp:49-68 + iou3d_nms_kernel.cu:105-226,249-262 (rotated BEV overlap area) */
int b200iou_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_overlap,
                              b200_stream_t stream);

/* boxes_iou_bev_gpu: overlap / max(sa + sb - overlap, 1e-8)
 * reference: iou3d_nms.cpp:70-88 + iou3d_nms_kernel.cu:228-235,264-278                            */
int b200iou_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou,
                          b200_stream_t stream);

/* boxes_iou3d_gpu in one launch: BEV overlap x height overlap / clamp(va + vb - ov3d, 1e-6)
 * reference: iou3d_nms_utils.py:48-81 (kernel + 10 torch elementwise launches)                    */
int b200iou_boxes_iou3d(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou,
                        b200_stream_t stream);

/* Block-diagonal variant for the IoU-label computation (models/loss_helper_iou.py:95-111):
 * boxes_a (S,K,7), boxes_b (S,G,7) -> iou (S,K,G); only same-scene pairs are evaluated.          */
int b200iou_boxes_iou3d_batched(int S, int K, const float *boxes_a, int G, const float *boxes_b, float *ans_iou,
                                b200_stream_t stream);

/* nms_gpu(boxes (N,7) device, sorted by score; keep; thresh) -> num_to_keep
 * reference: iou3d_nms.cpp:90-138 + iou3d_nms_kernel.cu:237-247,280-324 (3D IoU > thresh).
 * Mask AND greedy sweep run on the device.  `keep_dev` (N int32, device) receives kept positions,
 * `num_dev` (1 int32, device) the count; `workspace` is >= N*ceil(N/64) uint64 on the device.
 * Fully asynchronous.  mode 0 = rotated 3D IoU (nms_gpu), mode 1 = axis-aligned BEV (nms_normal_gpu,
 * iou3d_nms.cpp:141-190 + iou3d_nms_kernel.cu:327-385).                                           */
int b200iou_nms_device(int n, const float *boxes, float thresh, int mode, unsigned long long *workspace,
                       int32_t *keep_dev, int32_t *num_dev, b200_stream_t stream);

/* Reference-shaped blocking form: `keep_host` (N int32, host) is filled, *num_out = num_to_keep.
 * Allocates nothing per call beyond a cached workspace; synchronises `stream` once.              */
int b200iou_nms(int n, const float *boxes, float thresh, int mode, int32_t *keep_host, int *num_out,
                b200_stream_t stream);

/* boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou): host-memory entry of the reference API
 * (iou3d_cpu.cpp:232-252); same geometry code compiled for the host.                             */
int b200iou_boxes_iou_bev_cpu(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou);

#ifdef __cplusplus
}
#endif
#endif /* B200_IOU3D_H */
