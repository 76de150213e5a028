/*
 * b200_nms.h -- C ABI of the device-side axis-aligned box suppression (libb200pc.so).
 *
 * the pseudo-label filter of the SSL step and the test-time NMS run, in the reference,
 * as host numpy loops after a device->host copy of every head output:
 *   utils/nms.py:52-81 nms_2d_faster, :84-122 nms_3d_faster, :125-165 nms_3d_faster_samecls,
 *   :168-215 lhs_3d_faster_samecls; callers models/ap_helper.py:139-202, models/loss_helper_unlabeled.py:441-492;
 *   box corners: models/ap_helper.py:76-93, utils/box_util.py:266-272,335-358.
 * These entries keep that work on the device (no host synchronisation) and return what the reference returns.
 * Return value: 0 on success, non-zero on error (b200_last_error()).  All pointers are device pointers.
 */
#ifndef B200_NMS_H
#define B200_NMS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* boxes (B,K,8) float64 rows [x1, y1, z1, x2, y2, z2, score, class] (the reference's `boxes_3d_with_prob`,
 * utils/nms.py:126-133; for the 2-D variant pass z1 = 0, z2 = 1, which leaves every product unchanged);
 * valid (B,K) u8 or NULL = the reference's `nonempty_box_mask` (rows with 0 are left out before sorting,
 * models/ap_helper.py:198-199).
 *   use_cls  : multiply the overlap by (class_i == class_j)            (utils/nms.py:160)
 *   lhs      : after each pick also keep the better half of the boxes it suppresses, best first, and add 1e-8 to the
 *              volumes (utils/nms.py:177,202-209)
 *   old_type : overlap = intersection / volume_j instead of IoU        (utils/nms.py:113-117)
 * Arithmetic is float64 in numpy's operation order, so every `overlap > thresh` decision is the reference's.  Scores
 * are ranked ascending with ties broken by index (numpy.argsort's default sort leaves the order of equal scores
 * unspecified; NaN scores rank last as in numpy).
 * Outputs: pick (B,K) int32 = the reference's `pick` list in order, padded with -1; num_pick (B) int32;
 * picked_mask (B,K) u8 = 1 for rows in `pick` (the reference's pred_mask, models/ap_helper.py:201).
 * K <= 1024 (shared-memory bit matrix). */
int b200nms_aabb_suppress(int B, int K, int use_cls, int lhs, int old_type, double thresh, const double *boxes,
                          const unsigned char *valid, int32_t *pick, int32_t *num_pick, unsigned char *picked_mask,
                          void *stream);

/* predictions2corners3d + the per-box min/max loops (models/ap_helper.py:76-93,187-197):
 * center (B,K,3) f32 in upright-depth coordinates, size (B,K,3) f64 = class2size(...) (l, w, h),
 * heading (B,K) f64 = class2angle(...).  corners (B,K,8,3) f32 in upright-camera coordinates (may be NULL),
 * extents (B,K,6) f32 = [min x, min y, min z, max x, max y, max z] over the corners.  float64 math, float32 storage,
 * like the reference's arrays. */
int b200nms_box_extents(int B, int K, const float *center, const double *size, const double *heading, float *corners,
                        float *extents, void *stream);

#ifdef __cplusplus
}
#endif
#endif
