// aabb_nms.cu -- device-side axis-aligned box suppression (sm_100a): the reference's host loops
//   utils/nms.py:52-81    nms_2d_faster            (packed as 3-D boxes with z in [0,1])
//   utils/nms.py:84-122   nms_3d_faster
//   utils/nms.py:125-165  nms_3d_faster_samecls    (suppress only inside the same class)
//   utils/nms.py:168-215  lhs_3d_faster_samecls    (keep the picked box AND the better half of what it suppresses)
// and the per-box corner loops that feed them
//   models/ap_helper.py:76-93 (predictions2corners3d), utils/box_util.py:266-272,335-358 (roty, get_3d_box),
//   models/ap_helper.py:187-197 / models/loss_helper_unlabeled.py:470-483 (min / max over the 8 corners).
// The reference does this on the host in float64 numpy after a device->host copy of every head output, once per scene
// and per box in Python.  Here one CTA handles one scene: scores are ranked, the K x K "would suppress" relation is
// evaluated once into a bit matrix (float64, same operation order as numpy -> identical overlap decisions; boxes with
// EQUAL scores are ordered by index, one of the orders numpy's default non-stable argsort may produce -- pick lists are
// guaranteed identical to the reference only for distinct scores, see tests/test_oracle_aabb_nms.py), and a single warp
// sweeps it in score order.  No host synchronisation; the pick list comes back in the reference's order.
#include <math.h>

#include "../../include/b200_nms.h"
#include "common.cuh"

namespace b200 {

constexpr int NMS_THREADS = 256;
constexpr int NMS_MAX_K = 2048;

// numpy.maximum(0, d): d if d > 0, NaN if d is NaN, else 0
__device__ __forceinline__ double np_max0(double d) { return (d > 0.0 || d != d) ? d : 0.0; }
// numpy.maximum / minimum propagate NaN
__device__ __forceinline__ double np_max(double a, double b) { return (a != a || b != b) ? (a + b) : (a > b ? a : b); }
__device__ __forceinline__ double np_min(double a, double b) { return (a != a || b != b) ? (a + b) : (a < b ? a : b); }

// Shared-memory plan (dynamic): double box[K][9] (x1,y1,z1,x2,y2,z2,area,score,class) | u64 mat[K][W] |
//   int order[K] (ascending-score position -> box) | u8 ok[K]
__global__ void __launch_bounds__(NMS_THREADS)
aabb_suppress_kernel(int K, int W, int use_cls, int lhs, int old_type, double thresh, const double *__restrict__ boxes,
                     const unsigned char *__restrict__ valid, int32_t *__restrict__ pick, int32_t *__restrict__ num_pick,
                     unsigned char *__restrict__ picked_mask) {
  extern __shared__ __align__(16) unsigned char nms_dyn[];
  constexpr int BS = 9;                                                        // doubles per box record
  double *s_box = (double *)nms_dyn;                                           // [K][BS]
  unsigned long long *s_mat = (unsigned long long *)(s_box + (size_t)K * BS);  // [K][W]
  int *s_order = (int *)(s_mat + (size_t)K * W);
  unsigned char *s_ok = (unsigned char *)(s_order + K);
  __shared__ int s_n;

  const int b = blockIdx.x, tid = threadIdx.x;
  const double *bx = boxes + (size_t)b * K * 8;
  if (tid == 0) s_n = 0;
  for (int j = tid; j < K; j += NMS_THREADS) {
    const double x1 = bx[j * 8 + 0], y1 = bx[j * 8 + 1], z1 = bx[j * 8 + 2];
    const double x2 = bx[j * 8 + 3], y2 = bx[j * 8 + 4], z2 = bx[j * 8 + 5];
    double area = __dmul_rn(__dmul_rn(x2 - x1, y2 - y1), z2 - z1);     // nms.py:92 / :134 / :177
    if (lhs) area = area + 1e-8;                                        // :177
    double *o = s_box + (size_t)j * BS;
    o[0] = x1; o[1] = y1; o[2] = z1; o[3] = x2; o[4] = y2; o[5] = z2; o[6] = area;
    o[7] = bx[j * 8 + 6];
    o[8] = bx[j * 8 + 7];
    s_ok[j] = valid ? (valid[(size_t)b * K + j] != 0) : 1;
    pick[(size_t)b * K + j] = -1;
    picked_mask[(size_t)b * K + j] = 0;
  }
  __syncthreads();

  // ---- ascending stable rank by score (NaN last, like numpy.argsort); invalid boxes are left out ------------------
  for (int j = tid; j < K; j += NMS_THREADS) {
    if (!s_ok[j]) continue;
    const double sj = s_box[(size_t)j * BS + 7];
    const bool nj = sj != sj;
    int r = 0;
    for (int i = 0; i < K; ++i) {
      if (!s_ok[i]) continue;
      const double si = s_box[(size_t)i * BS + 7];
      const bool ni = si != si;
      const bool before = nj ? (!ni || i < j) : (!ni && (si < sj || (si == sj && i < j)));
      r += before;
    }
    s_order[r] = j;
    atomicAdd(&s_n, 1);
  }
  __syncthreads();
  const int n = s_n;

  // ---- bit matrix over score positions: bit q of row p (q < p) <=> box order[p] suppresses box order[q] -----------
  for (int e = tid; e < n * W; e += NMS_THREADS) {
    const int p = e / W, w = e - p * W;
    const double *bi = s_box + (size_t)s_order[p] * BS;
    unsigned long long bits = 0ull;
    const int q0 = w * 64;
    for (int t = 0; t < 64; ++t) {
      const int q = q0 + t;
      if (q >= p) break;
      const double *bj = s_box + (size_t)s_order[q] * BS;
      const double l = np_max0(np_min(bi[3], bj[3]) - np_max(bi[0], bj[0]));   // :104-111
      const double wd = np_max0(np_min(bi[4], bj[4]) - np_max(bi[1], bj[1]));
      const double h = np_max0(np_min(bi[5], bj[5]) - np_max(bi[2], bj[2]));
      const double inter = __dmul_rn(__dmul_rn(l, wd), h);
      double o = old_type ? inter / bj[6] : inter / ((bi[6] + bj[6]) - inter);  // :113-117
      if (use_cls) o = o * (bi[8] == bj[8] ? 1.0 : 0.0);                        // :160
      if (o > thresh) bits |= 1ull << t;
    }
    s_mat[(size_t)p * W + w] = bits;
  }
  __syncthreads();

  // ---- sweep in descending score order (one warp; lane w owns word w, w + 32, ... of the alive set) ---------------
  if (tid < 32) {
    const int lane = tid;
    constexpr int WPL = NMS_MAX_K / 64 / 32;  // words per lane
    unsigned long long alive[WPL];
#pragma unroll
    for (int u = 0; u < WPL; ++u) {
      const int w = lane + u * 32;
      const int lo = w * 64;
      alive[u] = (lo >= n) ? 0ull : ((n - lo >= 64) ? ~0ull : ((1ull << (n - lo)) - 1ull));
    }
    int np = 0;
    for (int p = n - 1; p >= 0; --p) {
      const int pw = p >> 6, pu = pw >> 5, pl = pw & 31;
      unsigned long long word = 0ull;
#pragma unroll
      for (int u = 0; u < WPL; ++u)
        if (u == pu) word = alive[u];
      word = __shfl_sync(0xffffffffu, word, pl);
      if (!((word >> (p & 63)) & 1ull)) continue;  // already removed
      const int box = s_order[p];
      if (lane == 0) {
        pick[(size_t)b * K + np] = box;
        picked_mask[(size_t)b * K + box] = 1;
      }
      ++np;
      // the boxes this one removes
      unsigned long long hit[WPL];
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < WPL; ++u) {
        const int w = lane + u * 32;
        hit[u] = (w < W) ? (s_mat[(size_t)p * W + w] & alive[u]) : 0ull;
        cnt += __popcll(hit[u]);
        alive[u] &= ~hit[u];
        if (u == pu && lane == pl) alive[u] &= ~(1ull << (p & 63));
      }
      if (lhs) {  // keep the better half of the suppressed boxes, best first (:205-209)
        const int total = __reduce_add_sync(0xffffffffu, cnt);
        int want = total >> 1;
        for (int w = W - 1; w >= 0 && want > 0; --w) {
          unsigned long long hw = 0ull;
#pragma unroll
          for (int u = 0; u < WPL; ++u)
            if (u == (w >> 5)) hw = hit[u];
          hw = __shfl_sync(0xffffffffu, hw, w & 31);
          while (hw && want > 0) {
            const int t = 63 - __clzll(hw);
            hw &= ~(1ull << t);
            const int bq = s_order[w * 64 + t];
            if (lane == 0) {
              pick[(size_t)b * K + np] = bq;
              picked_mask[(size_t)b * K + bq] = 1;
            }
            ++np;
            --want;
          }
        }
      }
    }
    if (lane == 0) num_pick[b] = np;
  }
}

// One thread per box: upright-camera corners (float64 math, stored as float32 like the reference's array) and their
// axis-aligned extents.
__global__ void __launch_bounds__(256)
box_extents_kernel(int total, const float *__restrict__ center, const double *__restrict__ size,
                   const double *__restrict__ heading, float *__restrict__ corners, float *__restrict__ extents) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  // flip_axis_to_camera (ap_helper.py:28-35): cam (x, y, z) = depth (x, -z, y), in float32
  const float cx = center[i * 3 + 0], cy = -center[i * 3 + 2], cz = center[i * 3 + 1];
  const double l = size[i * 3 + 0], w = size[i * 3 + 1], h = size[i * 3 + 2];
  const double t = heading[i];
  const double c = cos(t), s = sin(t);  // roty (box_util.py:266-272)
  const double sx[8] = {1, 1, -1, -1, 1, 1, -1, -1};
  const double sy[8] = {1, 1, 1, 1, -1, -1, -1, -1};
  const double sz[8] = {1, -1, -1, 1, 1, -1, -1, 1};
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool nan_seen[3] = {false, false, false};
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const double xc = sx[v] * (l / 2), yc = sy[v] * (h / 2), zc = sz[v] * (w / 2);  // box_util.py:350-352
    // R @ corner, row by row in the order a 3-term dot product accumulates (box_util.py:353)
    const double rx = __dadd_rn(__dadd_rn(__dmul_rn(c, xc), __dmul_rn(0.0, yc)), __dmul_rn(s, zc));
    const double ry = __dadd_rn(__dadd_rn(__dmul_rn(0.0, xc), __dmul_rn(1.0, yc)), __dmul_rn(0.0, zc));
    const double rz = __dadd_rn(__dadd_rn(__dmul_rn(-s, xc), __dmul_rn(0.0, yc)), __dmul_rn(c, zc));
    const float p[3] = {(float)(rx + (double)cx), (float)(ry + (double)cy), (float)(rz + (double)cz)};  // :354-356
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (corners) corners[((size_t)i * 8 + v) * 3 + d] = p[d];
      nan_seen[d] |= p[d] != p[d];
      mn[d] = fminf(mn[d], p[d]);
      mx[d] = fmaxf(mx[d], p[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {  // numpy.min / numpy.max return NaN when any corner is NaN
    extents[(size_t)i * 6 + d] = nan_seen[d] ? NAN : mn[d];
    extents[(size_t)i * 6 + 3 + d] = nan_seen[d] ? NAN : mx[d];
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200nms_aabb_suppress(int B, int K, int use_cls, int lhs, int old_type, double thresh, const double *boxes,
                                     const unsigned char *valid, int32_t *pick, int32_t *num_pick,
                                     unsigned char *picked_mask, void *stream) {
  B200_CHECK_ARG(B >= 0 && K >= 0, "aabb_suppress: negative size (B=%d, K=%d)", B, K);
  B200_CHECK_ARG(K <= NMS_MAX_K, "aabb_suppress: K=%d exceeds the supported %d boxes per scene", K, NMS_MAX_K);
  if (B == 0) return 0;
  B200_CHECK_ARG(num_pick, "aabb_suppress: null num_pick");
  if (K == 0) {
    B200_CUDA_OK(cudaMemsetAsync(num_pick, 0, sizeof(int32_t) * (size_t)B, (cudaStream_t)stream));
    return 0;
  }
  B200_CHECK_ARG(boxes && pick && picked_mask, "aabb_suppress: null pointer");
  const int W = ceil_div(K, 64);
  const size_t smem = (size_t)K * 9 * 8 + (size_t)K * W * 8 + (size_t)K * (4 + 1) + 16;
  B200_CHECK_ARG(smem <= 220 * 1024, "aabb_suppress: K=%d needs %zu B of shared memory", K, smem);
  B200_CUDA_OK(cudaFuncSetAttribute((void *)aabb_suppress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  aabb_suppress_kernel<<<B, NMS_THREADS, smem, (cudaStream_t)stream>>>(K, W, use_cls != 0, lhs != 0, old_type != 0,
                                                                       thresh, boxes, valid, pick, num_pick, picked_mask);
  B200_LAUNCH_OK("aabb_suppress_kernel");
  return 0;
}

extern "C" int b200nms_box_extents(int B, int K, const float *center, const double *size, const double *heading,
                                   float *corners, float *extents, void *stream) {
  B200_CHECK_ARG(B >= 0 && K >= 0, "box_extents: negative size (B=%d, K=%d)", B, K);
  const long long total = (long long)B * K;
  if (total == 0) return 0;
  B200_CHECK_ARG(total < (1ll << 28), "box_extents: too many boxes (%lld)", total);
  B200_CHECK_ARG(center && size && heading && extents, "box_extents: null pointer");
  box_extents_kernel<<<ceil_div((int)total, 256), 256, 0, (cudaStream_t)stream>>>((int)total, center, size, heading,
                                                                                   corners, extents);
  B200_LAUNCH_OK("box_extents_kernel");
  return 0;
}
