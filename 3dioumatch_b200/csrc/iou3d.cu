// iou3d.cu -- rotated BEV overlap / BEV IoU / 3D IoU / NMS for sm_100a.
//
// Replaces OpenPCDet/pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu (one thread per box pair, 24-entry local
// arrays -> 208 B of local-memory stack per thread, atan2 inside an O(n^2) bubble sort, heavy divergence)
// and the host round trip of iou3d_nms.cpp:90-138 (cudaMalloc + blocking D2H + CPU sweep + cudaFree).
//
// B200 design: a warp owns a strip (one box a x 32 boxes b).  Lane l first runs a conservative bounding-circle
// reject for its own pair (most pairs of a scene are disjoint -> exact 0, as the reference computes); surviving
// pairs are then clipped by the WHOLE warp, one pair at a time:
//   lanes 0-15 : the 16 edge-edge intersection tests          (iou3d_nms_kernel.cu:161-175, :64-93)
//   lanes 16-23: the 8 corner-in-box tests with the 1e-2 margin (:178-195, :52-62)
//   ballot -> vertex count; centroid by an ordered shuffle sum (same summation order as the reference);
//   every vertex lane computes ONE atan2f; the reference's bubble sort is a stable ascending sort, reproduced
//   as a rank computed with 24 shuffles; the fan area is summed in the reference's order.
// Everything lives in registers + 192 B of shared memory per warp; no local memory.
// The same vertex-collection rule (not an exact polygon clip) is kept on purpose: float parity <= 1e-5 with the
// reference on near-degenerate boxes depends on it (SURVEY.md section 7 "IoU parity").
#include <math.h>

#include <vector>

#include "../../include/b200_iou3d.h"
#include "common.cuh"

namespace b200 {

constexpr int IOU_WARPS = 8;
constexpr float IOU_MARGIN = 1e-2f;

struct Box {  // [x, y, z, dx, dy, dz, heading] + cos/sin(heading)
  float x, y, z, dx, dy, dz, h, c, s;
};

__device__ __forceinline__ Box load_box(const float *p) {
  Box b;
  b.x = p[0]; b.y = p[1]; b.z = p[2]; b.dx = p[3]; b.dy = p[4]; b.dz = p[5]; b.h = p[6];
  b.c = cosf(b.h);
  b.s = sinf(b.h);
  return b;
}
__device__ __forceinline__ Box shfl_box(const Box &b, int src) {
  Box r;
  r.x = __shfl_sync(0xffffffffu, b.x, src); r.y = __shfl_sync(0xffffffffu, b.y, src);
  r.z = __shfl_sync(0xffffffffu, b.z, src); r.dx = __shfl_sync(0xffffffffu, b.dx, src);
  r.dy = __shfl_sync(0xffffffffu, b.dy, src); r.dz = __shfl_sync(0xffffffffu, b.dz, src);
  r.h = __shfl_sync(0xffffffffu, b.h, src); r.c = __shfl_sync(0xffffffffu, b.c, src);
  r.s = __shfl_sync(0xffffffffu, b.s, src);
  return r;
}

// corner k of the rotated rectangle, following iou3d_nms_kernel.cu:111-150 literally
__device__ __forceinline__ void box_corner(const Box &b, int k, float &ox, float &oy) {
  const float hx = b.dx / 2, hy = b.dy / 2;
  const float px = (k == 1 || k == 2) ? b.x + hx : b.x - hx;
  const float py = (k >= 2) ? b.y + hy : b.y - hy;
  ox = (px - b.x) * b.c + (py - b.y) * (-b.s) + b.x;  // rotate_around_center :95-99
  oy = (px - b.x) * b.s + (py - b.y) * b.c + b.y;
}

__device__ __forceinline__ float cross3(float p1x, float p1y, float p2x, float p2y, float p0x, float p0y) {
  return (p1x - p0x) * (p2y - p0y) - (p2x - p0x) * (p1y - p0y);  // :40-42
}

// :52-62 ; cos(-h) = cos(h), sin(-h) = -sin(h)
__device__ __forceinline__ bool in_box2d(const Box &b, float px, float py) {
  const float angle_cos = b.c, angle_sin = -b.s;
  const float rot_x = (px - b.x) * angle_cos + (py - b.y) * (-angle_sin);
  const float rot_y = (px - b.x) * angle_sin + (py - b.y) * angle_cos;
  return fabsf(rot_x) < b.dx / 2 + IOU_MARGIN && fabsf(rot_y) < b.dy / 2 + IOU_MARGIN;
}

// true when the pair can contribute no vertex at all (disjoint even with the corner margin): overlap == 0 exactly
__device__ __forceinline__ bool surely_disjoint(const Box &a, const Box &b) {
  const float ra = 0.5f * sqrtf(a.dx * a.dx + a.dy * a.dy), rb = 0.5f * sqrtf(b.dx * b.dx + b.dy * b.dy);
  const float ddx = a.x - b.x, ddy = a.y - b.y;
  const float t = ra + rb + 0.05f + 1e-5f * (fabsf(a.x) + fabsf(a.y) + fabsf(b.x) + fabsf(b.y));
  return ddx * ddx + ddy * ddy > t * t;  // NaN -> false -> full path
}

// Whole-warp rotated-rectangle overlap area of (a, b); every lane returns the same value.
// s_poly: 24 float2 of shared memory private to this warp.
__device__ float warp_box_overlap(const Box &a, const Box &b, float2 *s_poly, int lane) {
  bool flag = false;
  float vx = 0.f, vy = 0.f;
  if (lane < 16) {
    const int i = lane >> 2, j = lane & 3;
    float p0x, p0y, p1x, p1y, q0x, q0y, q1x, q1y;
    box_corner(a, i, p0x, p0y);
    box_corner(a, (i + 1) & 3, p1x, p1y);
    box_corner(b, j, q0x, q0y);
    box_corner(b, (j + 1) & 3, q1x, q1y);
    // intersection(p1, p0, q1, q0) :64-93
    const bool rect = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                      fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
    if (rect) {
      const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
      const float s2 = cross3(p1x, p1y, q1x, q1y, p0x, p0y);
      const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
      const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
      if (s1 * s2 > 0 && s3 * s4 > 0) {
        const float s5 = cross3(q1x, q1y, p1x, p1y, p0x, p0y);
        if (fabsf(s5 - s1) > 1e-8f) {  // EPS: the fp32 and the double compare select the same floats
          vx = (s5 * q0x - s1 * q1x) / (s5 - s1);
          vy = (s5 * q0y - s1 * q1y) / (s5 - s1);
        } else {
          const float a0 = p0y - p1y, b0 = p1x - p0x, c0 = p0x * p1y - p1x * p0y;
          const float a1 = q0y - q1y, b1 = q1x - q0x, c1 = q0x * q1y - q1x * q0y;
          const float D = a0 * b1 - a1 * b0;
          vx = (b0 * c1 - b1 * c0) / D;
          vy = (a1 * c0 - a0 * c1) / D;
        }
        flag = true;
      }
    }
  } else if (lane < 24) {
    const int t = lane - 16, k = t >> 1;
    if ((t & 1) == 0) {  // corner k of b inside a   (:179-187)
      box_corner(b, k, vx, vy);
      flag = in_box2d(a, vx, vy);
    } else {  // corner k of a inside b             (:188-195)
      box_corner(a, k, vx, vy);
      flag = in_box2d(b, vx, vy);
    }
  }
  const unsigned mask = __ballot_sync(0xffffffffu, flag);
  const int cnt = __popc(mask);
  if (cnt < 3) return 0.f;  // the fan sum of <3 vertices is exactly 0

  // centroid: ordered sum over the vertices in collection order (:166,181,189), then / cnt (:198-199)
  float sx = 0.f, sy = 0.f;
  for (unsigned mm = mask; mm; mm &= mm - 1) {
    const int l = __ffs(mm) - 1;
    sx = sx + __shfl_sync(0xffffffffu, vx, l);
    sy = sy + __shfl_sync(0xffffffffu, vy, l);
  }
  const float cxm = sx / cnt, cym = sy / cnt;
  const float ang = flag ? atan2f(vy - cym, vx - cxm) : 0.f;  // point_cmp :101-103
  // stable ascending rank == result of the reference's bubble sort (:202-210)
  int rank = 0;
  for (unsigned mm = mask; mm; mm &= mm - 1) {
    const int l = __ffs(mm) - 1;
    const float al = __shfl_sync(0xffffffffu, ang, l);
    rank += (al < ang || (al == ang && l < lane)) ? 1 : 0;
  }
  __syncwarp();
  if (flag) s_poly[rank] = make_float2(vx, vy);
  __syncwarp();
  // fan area (:220-225): area += cross(P[k]-P[0], P[k+1]-P[0]), k = 0..cnt-2, summed in order
  float term = 0.f;
  if (lane < cnt - 1) {
    const float2 p0 = s_poly[0], pk = s_poly[lane], pn = s_poly[lane + 1];
    const float ux = pk.x - p0.x, uy = pk.y - p0.y, wx = pn.x - p0.x, wy = pn.y - p0.y;
    term = ux * wy - uy * wx;
  }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) area += __shfl_sync(0xffffffffu, term, k);
  return fabsf(area) * 0.5f;
}

enum { MODE_OVERLAP = 0, MODE_IOU_BEV = 1, MODE_IOU3D = 2 };

__device__ __forceinline__ float finish_pair(int mode, const Box &a, const Box &b, float ov) {
  if (mode == MODE_OVERLAP) return ov;
  if (mode == MODE_IOU_BEV) {  // iou_bev :228-235
    const float sa = a.dx * a.dy, sb = b.dx * b.dy;
    return ov / fmaxf(sa + sb - ov, 1e-8f);
  }
  // boxes_iou3d_gpu, iou3d_nms_utils.py:60-79 (each torch elementwise step rounds to fp32)
  const float a_max = a.z + a.dz / 2, a_min = a.z - a.dz / 2;
  const float b_max = b.z + b.dz / 2, b_min = b.z - b.dz / 2;
  const float ov_h = fmaxf(fminf(a_max, b_max) - fmaxf(a_min, b_min), 0.f);
  const float ov3d = ov * ov_h;
  const float vol_a = a.dx * a.dy * a.dz, vol_b = b.dx * b.dy * b.dz;
  return ov3d / fmaxf(vol_a + vol_b - ov3d, 1e-6f);
}

// boxes_a (S, K, 7), boxes_b (S, G, 7) -> ans (S, K, G); S = 1 is the reference's all-pairs call.
__global__ void __launch_bounds__(IOU_WARPS * 32)
pair_kernel(int mode, int S, int K, int G, const float *__restrict__ boxes_a, const float *__restrict__ boxes_b,
            float *__restrict__ ans) {
  __shared__ float2 s_poly[IOU_WARPS][24];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strips_per_row = ceil_div(G, 32);
  const long long nstrips = (long long)S * K * strips_per_row;
  for (long long st = (long long)blockIdx.x * IOU_WARPS + warp; st < nstrips; st += (long long)gridDim.x * IOU_WARPS) {
    const int sb = (int)(st % strips_per_row);
    const long long row = st / strips_per_row;  // s*K + ia
    const int s = (int)(row / K);
    const Box a = load_box(boxes_a + row * 7);
    const int jb = sb * 32 + lane;
    const bool have = jb < G;
    Box b = {};
    if (have) b = load_box(boxes_b + ((long long)s * G + jb) * 7);
    float res = 0.f;
    unsigned todo = __ballot_sync(0xffffffffu, have && !surely_disjoint(a, b));
    for (; todo; todo &= todo - 1) {
      const int l = __ffs(todo) - 1;
      const Box bl = shfl_box(b, l);
      const float ov = warp_box_overlap(a, bl, s_poly[warp], lane);
      if (lane == l) res = ov;
    }
    if (have) ans[row * G + jb] = finish_pair(mode, a, b, res);
  }
}

// iou_bev_3D :237-247 (3DIoUMatch's NMS criterion) and iou_normal :327-338
__device__ __forceinline__ float iou_bev_3d(const Box &a, const Box &b, float ov_bev) {
  const float sa = a.dx * a.dy * a.dz, sb = b.dx * b.dy * b.dz;
  const float top = fmaxf(a.z - a.dz / 2, b.z - b.dz / 2);
  const float bottom = fminf(a.z + a.dz / 2, b.z + b.dz / 2);
  const float height = fmaxf(bottom - top, 0.f);
  const float s_overlap = ov_bev * height;
  return s_overlap / fmaxf(sa + sb - s_overlap, 1e-8f);
}
__device__ __forceinline__ float iou_normal(const Box &a, const Box &b) {
  const float left = fmaxf(a.x - a.dx / 2, b.x - b.dx / 2), right = fminf(a.x + a.dx / 2, b.x + b.dx / 2);
  const float top = fmaxf(a.y - a.dy / 2, b.y - b.dy / 2), bottom = fminf(a.y + a.dy / 2, b.y + b.dy / 2);
  const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  const float interS = width * height;
  const float Sa = a.dx * a.dy, Sb = b.dx * b.dy;
  return interS / fmaxf(Sa + Sb - interS, 1e-8f);
}

// nms_kernel :280-324 / nms_normal_kernel :341-385: mask[i][cb] bit j <=> IoU(box i, box cb*64+j) > thresh,
// only j > i inside the diagonal block.  One warp per (row i, column block cb): two 32-wide strips.
__global__ void __launch_bounds__(IOU_WARPS * 32)
nms_mask_kernel(int n, int col_blocks, float thresh, int mode, const float *__restrict__ boxes,
                unsigned long long *__restrict__ mask) {
  __shared__ float2 s_poly[IOU_WARPS][24];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nitems = (long long)n * col_blocks;
  for (long long it = (long long)blockIdx.x * IOU_WARPS + warp; it < nitems; it += (long long)gridDim.x * IOU_WARPS) {
    const int i = (int)(it / col_blocks), cb = (int)(it % col_blocks);
    unsigned long long word = 0ull;
    if (cb >= i / 64) {  // the sweep never reads blocks left of the diagonal (iou3d_nms.cpp:130)
      const Box a = load_box(boxes + (size_t)i * 7);
      for (int half = 0; half < 2; ++half) {
        const int j = cb * 64 + half * 32 + lane;
        const bool have = j < n && j > i;  // j > i: diagonal-block rule (:311-313); other blocks satisfy it anyway
        Box b = {};
        if (have) b = load_box(boxes + (size_t)j * 7);
        bool over = false;
        if (mode == 1) {
          over = have && iou_normal(a, b) > thresh;
        } else {
          float res = 0.f;
          unsigned todo = __ballot_sync(0xffffffffu, have && !surely_disjoint(a, b));
          for (; todo; todo &= todo - 1) {
            const int l = __ffs(todo) - 1;
            const Box bl = shfl_box(b, l);
            const float ov = warp_box_overlap(a, bl, s_poly[warp], lane);
            if (lane == l) res = ov;
          }
          over = have && iou_bev_3d(a, b, res) > thresh;
        }
        const unsigned bits = __ballot_sync(0xffffffffu, over);
        word |= (unsigned long long)bits << (32 * half);
      }
    }
    if (lane == 0) mask[it] = word;
  }
}

// Greedy sweep of iou3d_nms.cpp:121-137 on the device: one CTA, 64 boxes per step.
__global__ void __launch_bounds__(1024)
nms_sweep_kernel(int n, int col_blocks, const unsigned long long *__restrict__ mask, int32_t *__restrict__ keep,
                 int32_t *__restrict__ num_out) {
  extern __shared__ unsigned long long s_remv[];  // col_blocks words
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_kept;
  __shared__ int s_count;
  const int tid = threadIdx.x;
  for (int j = tid; j < col_blocks; j += blockDim.x) s_remv[j] = 0ull;
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int blk = 0; blk < col_blocks; ++blk) {
    const int base = blk * 64;
    const int nb = min(64, n - base);
    if (tid < 64) s_diag[tid] = tid < nb ? mask[(size_t)(base + tid) * col_blocks + blk] : 0ull;
    __syncthreads();
    if (tid == 0) {
      unsigned long long r = s_remv[blk], kept = 0ull;
      for (int t = 0; t < nb; ++t) {
        if (!((r >> t) & 1ull)) {
          kept |= 1ull << t;
          r |= s_diag[t];
        }
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    const int count = s_count;
    if (tid < 64 && ((kept >> tid) & 1ull))
      keep[count + __popcll(kept & ((1ull << tid) - 1ull))] = base + tid;
    for (int j = blk + 1 + tid; j < col_blocks; j += blockDim.x) {
      unsigned long long r = s_remv[j];
      for (unsigned long long kk = kept; kk; kk &= kk - 1ull) {
        const int t = __ffsll((long long)kk) - 1;
        r |= mask[(size_t)(base + t) * col_blocks + j];
      }
      s_remv[j] = r;
    }
    __syncthreads();
    if (tid == 0) s_count = count + __popcll(kept);
    __syncthreads();
  }
  if (tid == 0) *num_out = s_count;
}

static int pair_launch(int mode, int S, int K, int G, const float *a, const float *b, float *ans, cudaStream_t st) {
  if (S <= 0 || K <= 0 || G <= 0) return 0;
  B200_CHECK_ARG(a && b && ans, "iou3d: null pointer");
  const long long nstrips = (long long)S * K * ceil_div(G, 32);
  long long blocks = (nstrips + IOU_WARPS - 1) / IOU_WARPS;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pair_kernel<<<(int)blocks, IOU_WARPS * 32, 0, st>>>(mode, S, K, G, a, b, ans);
  B200_LAUNCH_OK("pair_kernel");
  return 0;
}

// ---- host-memory entry (iou3d_cpu.cpp:232-252): scalar form of the same vertex-collection rule ------
struct HPt { float x, y; };
static inline float hcross3(HPt p1, HPt p2, HPt p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }
static inline HPt hcorner(const float *bx, int k, float c, float s) {
  const float hx = bx[3] / 2, hy = bx[4] / 2;
  const float px = (k == 1 || k == 2) ? bx[0] + hx : bx[0] - hx;
  const float py = (k >= 2) ? bx[1] + hy : bx[1] - hy;
  HPt r;
  r.x = (px - bx[0]) * c + (py - bx[1]) * (-s) + bx[0];
  r.y = (px - bx[0]) * s + (py - bx[1]) * c + bx[1];
  return r;
}
static inline bool hin_box(const float *bx, float c, float s, HPt p) {
  const float ac = c, as = -s;
  const float rx = (p.x - bx[0]) * ac + (p.y - bx[1]) * (-as);
  const float ry = (p.x - bx[0]) * as + (p.y - bx[1]) * ac;
  return fabsf(rx) < bx[3] / 2 + IOU_MARGIN && fabsf(ry) < bx[4] / 2 + IOU_MARGIN;
}
static float host_box_overlap(const float *A, const float *Bx) {
  const float ca = cosf(A[6]), sa = sinf(A[6]), cb = cosf(Bx[6]), sb = sinf(Bx[6]);
  HPt pa[4], pb[4];
  for (int k = 0; k < 4; ++k) { pa[k] = hcorner(A, k, ca, sa); pb[k] = hcorner(Bx, k, cb, sb); }
  HPt v[24];
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const HPt p0 = pa[i], p1 = pa[(i + 1) & 3], q0 = pb[j], q1 = pb[(j + 1) & 3];
      const bool rect = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
                        fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
      if (!rect) continue;
      const float s1 = hcross3(q0, p1, p0), s2 = hcross3(p1, q1, p0), s3 = hcross3(p0, q1, q0), s4 = hcross3(q1, p1, q0);
      if (!(s1 * s2 > 0 && s3 * s4 > 0)) continue;
      const float s5 = hcross3(q1, p1, p0);
      HPt r;
      if (fabsf(s5 - s1) > 1e-8f) {
        r.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        r.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
      } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        r.x = (b0 * c1 - b1 * c0) / D;
        r.y = (a1 * c0 - a0 * c1) / D;
      }
      v[cnt++] = r;
    }
  for (int k = 0; k < 4; ++k) {
    if (hin_box(A, ca, sa, pb[k])) v[cnt++] = pb[k];
    if (hin_box(Bx, cb, sb, pa[k])) v[cnt++] = pa[k];
  }
  if (cnt < 3) return 0.f;
  float sx = 0.f, sy = 0.f;
  for (int k = 0; k < cnt; ++k) { sx = sx + v[k].x; sy = sy + v[k].y; }
  const float cx = sx / cnt, cy = sy / cnt;
  float ang[24];
  for (int k = 0; k < cnt; ++k) ang[k] = atan2f(v[k].y - cy, v[k].x - cx);
  for (int k = 1; k < cnt; ++k) {  // stable insertion sort == the reference's bubble sort result
    const HPt pv = v[k];
    const float av = ang[k];
    int q = k - 1;
    while (q >= 0 && ang[q] > av) { v[q + 1] = v[q]; ang[q + 1] = ang[q]; --q; }
    v[q + 1] = pv;
    ang[q + 1] = av;
  }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ux = v[k].x - v[0].x, uy = v[k].y - v[0].y, wx = v[k + 1].x - v[0].x, wy = v[k + 1].y - v[0].y;
    area += ux * wy - uy * wx;
  }
  return fabsf(area) * 0.5f;
}

}  // namespace b200

using namespace b200;

extern "C" int b200iou_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b,
                                         float *ans, b200_stream_t s) {
  return pair_launch(MODE_OVERLAP, 1, num_a, num_b, boxes_a, boxes_b, ans, (cudaStream_t)s);
}
extern "C" int b200iou_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans,
                                     b200_stream_t s) {
  return pair_launch(MODE_IOU_BEV, 1, num_a, num_b, boxes_a, boxes_b, ans, (cudaStream_t)s);
}
extern "C" int b200iou_boxes_iou3d(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans,
                                   b200_stream_t s) {
  return pair_launch(MODE_IOU3D, 1, num_a, num_b, boxes_a, boxes_b, ans, (cudaStream_t)s);
}
extern "C" int b200iou_boxes_iou3d_batched(int S, int K, const float *boxes_a, int G, const float *boxes_b,
                                           float *ans, b200_stream_t s) {
  return pair_launch(MODE_IOU3D, S, K, G, boxes_a, boxes_b, ans, (cudaStream_t)s);
}

extern "C" int b200iou_nms_device(int n, const float *boxes, float thresh, int mode, unsigned long long *workspace,
                                  int32_t *keep_dev, int32_t *num_dev, b200_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  B200_CHECK_ARG(n >= 0 && (mode == 0 || mode == 1), "nms: bad arguments n=%d mode=%d", n, mode);
  B200_CHECK_ARG(num_dev, "nms: null pointer");
  if (n == 0) {
    B200_CUDA_OK(cudaMemsetAsync(num_dev, 0, sizeof(int32_t), st));
    return 0;
  }
  B200_CHECK_ARG(boxes && workspace && keep_dev, "nms: null pointer");
  const int col_blocks = ceil_div(n, 64);
  B200_CHECK_ARG((size_t)col_blocks * 8 <= 200 * 1024, "nms: n=%d too large for the single-CTA sweep", n);
  const long long nitems = (long long)n * col_blocks;
  long long blocks = (nitems + IOU_WARPS - 1) / IOU_WARPS;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  nms_mask_kernel<<<(int)blocks, IOU_WARPS * 32, 0, st>>>(n, col_blocks, thresh, mode, boxes, workspace);
  B200_LAUNCH_OK("nms_mask_kernel");
  const size_t smem = sizeof(unsigned long long) * (size_t)col_blocks;
  if (smem > 48 * 1024) {
    static DynSmemOptIn optin;
    B200_CUDA_OK(optin.ensure(nms_sweep_kernel, smem));
  }
  const int threads = col_blocks <= 64 ? 64 : (col_blocks <= 256 ? 256 : 1024);
  nms_sweep_kernel<<<1, threads, smem, st>>>(n, col_blocks, workspace, keep_dev, num_dev);
  B200_LAUNCH_OK("nms_sweep_kernel");
  return 0;
}

extern "C" int b200iou_nms(int n, const float *boxes, float thresh, int mode, int32_t *keep_host, int *num_out,
                           b200_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  B200_CHECK_ARG(n >= 0 && num_out && (keep_host || n == 0), "nms: bad arguments");
  *num_out = 0;
  if (n == 0) return 0;
  // stream-ordered device workspace + a per-call pinned staging buffer owned by the caller's thread: no state shared
  // between devices or threads (the reference cudaMalloc/cudaFree's per call, iou3d_nms.cpp:102-114)
  const size_t cb = ((size_t)n + 63) / 64;
  const size_t ws_bytes = (sizeof(unsigned long long) * (size_t)n * cb + 255) & ~(size_t)255;
  ScratchGuard ws;
  B200_CUDA_OK(ws.alloc(ws_bytes + sizeof(int32_t) * ((size_t)n + 1), st));
  unsigned long long *d_ws = (unsigned long long *)ws.ptr;
  int32_t *d_keep = (int32_t *)((char *)ws.ptr + ws_bytes);
  thread_local int32_t *h_pinned = nullptr;  // pinned host memory is not tied to a device
  thread_local size_t cap_n = 0;
  if ((size_t)n > cap_n) {
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    cap_n = 0;
    B200_CUDA_OK(cudaMallocHost(&h_pinned, sizeof(int32_t) * ((size_t)n + 1025)));
    cap_n = (size_t)n + 1024;
  }
  const int rc = b200iou_nms_device(n, boxes, thresh, mode, d_ws, d_keep + 1, d_keep, s);
  if (rc) return rc;
  B200_CUDA_OK(cudaMemcpyAsync(h_pinned, d_keep, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
  B200_CUDA_OK(cudaStreamSynchronize(st));
  const int num = h_pinned[0];
  for (int i = 0; i < num; ++i) keep_host[i] = h_pinned[1 + i];
  *num_out = num;
  return 0;
}

extern "C" int b200iou_boxes_iou_bev_cpu(int num_a, const float *boxes_a, int num_b, const float *boxes_b,
                                         float *ans_iou) {
  B200_CHECK_ARG(num_a >= 0 && num_b >= 0, "boxes_iou_bev_cpu: negative size");
  if (num_a == 0 || num_b == 0) return 0;
  B200_CHECK_ARG(boxes_a && boxes_b && ans_iou, "boxes_iou_bev_cpu: null pointer");
  for (int i = 0; i < num_a; ++i)
    for (int j = 0; j < num_b; ++j) {
      const float *a = boxes_a + (size_t)i * 7, *b = boxes_b + (size_t)j * 7;
      const float sa = a[3] * a[4], sb = b[3] * b[4];
      const float ov = host_box_overlap(a, b);
      ans_iou[(size_t)i * num_b + j] = ov / fmaxf(sa + sb - ov, 1e-8f);  // iou3d_cpu.cpp:222-229
    }
  return 0;
}
