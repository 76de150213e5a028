/*
 * oracle/iou3d_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32) of the reference's rotated-box overlap / IoU /
 * NMS operator (module `pcdet.ops.iou3d_nms.iou3d_nms_cuda` of yezhen17/3DIoUMatch).
 * Follows OpenPCDet/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu (the CUDA path the
 * losses call), not iou3d_cpu.cpp; the two differ only in the type of EPS.
 * Parity for this half is float parity (<= 1e-5), not bit parity: libm vs
 * libdevice sinf/cosf/atan2f and nvcc's FMA contraction differ in the last ulp.
 *
 * Pinned against: (i) the reference's own CPU entry boxes_iou_bev_cpu compiled
 * from /root/reference (oracle/build_ref.py, tests/test_oracle_pinned.py,
 * tests/golden/iou_bev_cpu_*.npz), (ii) analytic known answers, (iii) the
 * reference CUDA extension on the GPU box (tests/test_ref_cuda.py).
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y; } pt2;

static inline float fmin2(float a, float b) { return a < b ? a : b; } /* CUDA min/max on floats */
static inline float fmax2(float a, float b) { return a > b ? a : b; }

/* iou3d_nms_kernel.cu:40-42 */
static inline float cross3(pt2 p1, pt2 p2, pt2 p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
/* :36-38 */
static inline float cross2(pt2 a, pt2 b) { return a.x * b.y - a.y * b.x; }

/* :44-50 */
static inline int check_rect_cross(pt2 p1, pt2 p2, pt2 q1, pt2 q2) {
  return fmin2(p1.x, p2.x) <= fmax2(q1.x, q2.x) && fmin2(q1.x, q2.x) <= fmax2(p1.x, p2.x) &&
         fmin2(p1.y, p2.y) <= fmax2(q1.y, q2.y) && fmin2(q1.y, q2.y) <= fmax2(p1.y, p2.y);
}

/* :52-62 */
static inline int check_in_box2d(const float *box, pt2 p) {
  const float MARGIN = 1e-2f;
  const float center_x = box[0], center_y = box[1];
  const float angle_cos = cosf(-box[6]), angle_sin = sinf(-box[6]);
  const float rot_x = (p.x - center_x) * angle_cos + (p.y - center_y) * (-angle_sin);
  const float rot_y = (p.x - center_x) * angle_sin + (p.y - center_y) * angle_cos;
  return (fabsf(rot_x) < box[3] / 2 + MARGIN && fabsf(rot_y) < box[4] / 2 + MARGIN);
}

/* :64-93 */
static inline int intersection(pt2 p1, pt2 p0, pt2 q1, pt2 q0, pt2 *ans) {
  if (check_rect_cross(p0, p1, q0, q1) == 0) return 0;
  const float s1 = cross3(q0, p1, p0);
  const float s2 = cross3(p1, q1, p0);
  const float s3 = cross3(p0, q1, q0);
  const float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  const float s5 = cross3(q1, p1, p0);
  if ((double)fabsf(s5 - s1) > 1e-8) { /* EPS is a double macro in the .cu (:15) */
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / D;
    ans->y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

/* :95-99 */
static inline pt2 rotate_around_center(pt2 center, float angle_cos, float angle_sin, pt2 p) {
  pt2 r;
  r.x = (p.x - center.x) * angle_cos + (p.y - center.y) * (-angle_sin) + center.x;
  r.y = (p.x - center.x) * angle_sin + (p.y - center.y) * angle_cos + center.y;
  return r;
}

/* :101-103 */
static inline int point_cmp(pt2 a, pt2 b, pt2 center) {
  return atan2f(a.y - center.y, a.x - center.x) > atan2f(b.y - center.y, b.x - center.x);
}

/* :105-226.  box = [x, y, z, dx, dy, dz, heading] */
float orc_box_overlap(const float *box_a, const float *box_b) {
  const float a_angle = box_a[6], b_angle = box_b[6];
  const float a_dx_half = box_a[3] / 2, b_dx_half = box_b[3] / 2, a_dy_half = box_a[4] / 2, b_dy_half = box_b[4] / 2;
  const float a_x1 = box_a[0] - a_dx_half, a_y1 = box_a[1] - a_dy_half;
  const float a_x2 = box_a[0] + a_dx_half, a_y2 = box_a[1] + a_dy_half;
  const float b_x1 = box_b[0] - b_dx_half, b_y1 = box_b[1] - b_dy_half;
  const float b_x2 = box_b[0] + b_dx_half, b_y2 = box_b[1] + b_dy_half;
  pt2 center_a = {box_a[0], box_a[1]}, center_b = {box_b[0], box_b[1]};

  pt2 ca[5] = {{a_x1, a_y1}, {a_x2, a_y1}, {a_x2, a_y2}, {a_x1, a_y2}, {0, 0}};
  pt2 cb[5] = {{b_x1, b_y1}, {b_x2, b_y1}, {b_x2, b_y2}, {b_x1, b_y2}, {0, 0}};
  const float a_cos = cosf(a_angle), a_sin = sinf(a_angle);
  const float b_cos = cosf(b_angle), b_sin = sinf(b_angle);
  for (int k = 0; k < 4; k++) {
    ca[k] = rotate_around_center(center_a, a_cos, a_sin, ca[k]);
    cb[k] = rotate_around_center(center_b, b_cos, b_sin, cb[k]);
  }
  ca[4] = ca[0];
  cb[4] = cb[0];

  pt2 cross_points[24]; /* the reference declares 16 (:156) and can overrun; 24 is the true maximum */
  pt2 poly_center = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      pt2 ans;
      if (intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &ans)) {
        cross_points[cnt] = ans;
        poly_center.x = poly_center.x + ans.x;
        poly_center.y = poly_center.y + ans.y;
        cnt++;
      }
    }
  for (int k = 0; k < 4; k++) {
    if (check_in_box2d(box_a, cb[k])) {
      poly_center.x = poly_center.x + cb[k].x;
      poly_center.y = poly_center.y + cb[k].y;
      cross_points[cnt++] = cb[k];
    }
    if (check_in_box2d(box_b, ca[k])) {
      poly_center.x = poly_center.x + ca[k].x;
      poly_center.y = poly_center.y + ca[k].y;
      cross_points[cnt++] = ca[k];
    }
  }
  poly_center.x /= cnt; /* cnt==0 -> NaN, unused */
  poly_center.y /= cnt;

  for (int j = 0; j < cnt - 1; j++)
    for (int i = 0; i < cnt - j - 1; i++)
      if (point_cmp(cross_points[i], cross_points[i + 1], poly_center)) {
        pt2 t = cross_points[i];
        cross_points[i] = cross_points[i + 1];
        cross_points[i + 1] = t;
      }

  float area = 0;
  for (int k = 0; k < cnt - 1; k++) {
    pt2 u = {cross_points[k].x - cross_points[0].x, cross_points[k].y - cross_points[0].y};
    pt2 v = {cross_points[k + 1].x - cross_points[0].x, cross_points[k + 1].y - cross_points[0].y};
    area += cross2(u, v);
  }
  return (float)(fabsf(area) / 2.0);
}

/* :228-235 */
float orc_iou_bev(const float *a, const float *b) {
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  const float s_overlap = orc_box_overlap(a, b);
  return s_overlap / fmaxf(sa + sb - s_overlap, (float)1e-8);
}

/* :237-247 -- the 3DIoUMatch modification: NMS uses 3D IoU */
float orc_iou_bev_3D(const float *a, const float *b) {
  const float sa = a[3] * a[4] * a[5], sb = b[3] * b[4] * b[5];
  const float top = fmaxf(a[2] - a[5] / 2, b[2] - b[5] / 2);
  const float bottom = fminf(a[2] + a[5] / 2, b[2] + b[5] / 2);
  const float height = fmaxf(bottom - top, 0.f);
  const float s_overlap = orc_box_overlap(a, b) * height;
  return s_overlap / fmaxf(sa + sb - s_overlap, (float)1e-8);
}

/* :327-338 */
float orc_iou_normal(const float *a, const float *b) {
  const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  const float interS = width * height;
  const float Sa = a[3] * a[4], Sb = b[3] * b[4];
  return interS / fmaxf(Sa + Sb - interS, (float)1e-8);
}

/* boxes_overlap_kernel :249-262 */
void orc_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) ans[(size_t)i * nb + j] = orc_box_overlap(boxes_a + i * 7, boxes_b + j * 7);
}

/* boxes_iou_bev_kernel :264-278 */
void orc_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j) ans[(size_t)i * nb + j] = orc_iou_bev(boxes_a + i * 7, boxes_b + j * 7);
}

/* iou3d_nms_utils.py:48-81 boxes_iou3d_gpu: torch epilogue around the BEV overlap,
 * every step rounded to fp32 like the torch elementwise kernels.                  */
void orc_boxes_iou3d(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < na; ++i) {
    const float *a = boxes_a + i * 7;
    const float a_max = a[2] + a[5] / 2, a_min = a[2] - a[5] / 2;
    const float vol_a = a[3] * a[4] * a[5];
    for (int j = 0; j < nb; ++j) {
      const float *b = boxes_b + j * 7;
      const float b_max = b[2] + b[5] / 2, b_min = b[2] - b[5] / 2;
      const float ov_bev = orc_box_overlap(a, b);
      const float max_of_min = fmaxf(a_min, b_min), min_of_max = fminf(a_max, b_max);
      float ov_h = min_of_max - max_of_min;
      if (ov_h < 0.f) ov_h = 0.f; /* clamp(min=0) */
      const float ov3d = ov_bev * ov_h;
      const float vol_b = b[3] * b[4] * b[5];
      float den = vol_a + vol_b - ov3d;
      if (den < 1e-6f) den = 1e-6f; /* clamp(min=1e-6) */
      ans[(size_t)i * nb + j] = ov3d / den;
    }
  }
}

/* nms_kernel :280-324 + host sweep iou3d_nms.cpp:121-137 (and the _normal twins
 * :341-385 / iou3d_nms.cpp:172-188).  boxes are already sorted by score.
 * mode 0: iou_bev_3D (rotated, 3D)   mode 1: iou_normal (axis-aligned BEV).
 * Returns num_to_keep, writes keep[0..num).                                       */
int orc_nms(int n, const float *boxes, float thresh, int mode, int32_t *keep) {
  const int col_blocks = (n + 63) / 64;
  uint64_t *mask = (uint64_t *)calloc((size_t)(n > 0 ? n : 1) * (size_t)(col_blocks > 0 ? col_blocks : 1), sizeof(uint64_t));
#pragma omp parallel for schedule(dynamic, 8)
  for (int i = 0; i < n; ++i) {
    const int row_start = i / 64;
    for (int cb = 0; cb < col_blocks; ++cb) {
      const int col_size = (n - cb * 64) < 64 ? (n - cb * 64) : 64;
      uint64_t t = 0;
      const int start = (row_start == cb) ? (i % 64) + 1 : 0;
      for (int jj = start; jj < col_size; ++jj) {
        const float *bj = boxes + (size_t)(cb * 64 + jj) * 7;
        const float v = mode == 0 ? orc_iou_bev_3D(boxes + (size_t)i * 7, bj) : orc_iou_normal(boxes + (size_t)i * 7, bj);
        if (v > thresh) t |= 1ULL << jj;
      }
      mask[(size_t)i * col_blocks + cb] = t;
    }
  }
  uint64_t *remv = (uint64_t *)calloc((size_t)(col_blocks > 0 ? col_blocks : 1), sizeof(uint64_t));
  int num_to_keep = 0;
  for (int i = 0; i < n; ++i) {
    const int nblock = i / 64, inblock = i % 64;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep[num_to_keep++] = i;
      const uint64_t *p = mask + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; ++j) remv[j] |= p[j];
    }
  }
  free(mask);
  free(remv);
  return num_to_keep;
}
