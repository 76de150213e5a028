// sa_train.cu -- training-mode set-abstraction MLP: conv1x1 -> BatchNorm2d with BATCH statistics -> ReLU per layer, max over
// nsample, and its backward (pointnet2/pytorch_utils.py:14-61,70-123,
// pointnet2_modules.py:256-262, group_points_gpu.cu:48-68 for the input gradient).
//
// Layer at a time, everything between two GEMMs fused into their loads and stores:
//   forward   a0 = grouped rows [rel xyz | features]                          group_rows_kernel (one gather, point-major rows)
//             z_l = a_{l-1} W_l^T, a_{l-1} = relu(scale_{l-1} z_{l-1} + shift_{l-1}) applied while loading,
//                   per-tile column sums of z_l and z_l^2 in the epilogue        sa_tcp_kernel<2,0,1,1> (tcgen05, split TF32)
//             (mean, var) -> scale_l, shift_l, running statistics                bn_finalize_kernel (fp64 combination, fixed order)
//             out = max_s relu(scale_L z_L + shift_L), arg-max slot              bn_relu_maxpool_kernel
//   backward  S1_L = sum g, S2_L = sum g xhat over the arg-max rows             pool_bwd_stats_kernel
//             per layer l = L..1:  dz_l = scale_l (g_l - S1/R - xhat_l S2/R)   built on the fly by both kernels below
//               dW_l = dz_l^T a_{l-1}                                           bwd_dw_kernel (fp32 FFMA, split over rows, fixed-order reduce)
//               g_{l-1} = (dz_l W_l) * [a_{l-1} > 0] + column sums for BN_{l-1}  bwd_da_kernel (fp32 FFMA)
//             d features = scatter-add of the feature columns of da_0           rows_scatter_kernel
// Only the raw conv outputs z_l (and g_l in backward) exist in HBM -- one row tensor per layer instead of the
// conv / BN / ReLU / max-pool round trips (and their saved copies) of the op-by-op path.  Every reduction runs in a fixed
// order: weight, gamma, beta gradients and the running statistics are bit-reproducible; the input-feature gradient is a
// float atomicAdd scatter like the reference's (group_points_gpu.cu:65).
#include <stdlib.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"
#include "sa_tc.cuh"
#include "sa_train.cuh"

namespace b200 {

constexpr int TRN_MAXL = 4;

// ---- a0: grouped rows, reference channel order [dx,dy,dz | features] (pointnet2_utils.py:350-360) -------------------
__global__ void __launch_bounds__(256)
group_rows_kernel(int N, int M, int ns, int C, int ld, int use_xyz, float inv_r, const float *__restrict__ xyz,
                  const float *__restrict__ new_xyz, const float *__restrict__ feat_pm, const int32_t *__restrict__ idx,
                  float *__restrict__ rows, long long R) {
  // one warp per row: lanes stride over the row's channels (coalesced reads of the point-major source row)
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const long long g = r / ns;           // (scene, centre)
  const int b = (int)(g / M);
  const int src = idx[r];
  float *dst = rows + r * ld;
  const int c0 = use_xyz ? 3 : 0;
  if (use_xyz && lane < 3) {
    const float q = xyz[((size_t)b * N + src) * 3 + lane], c = new_xyz[g * 3 + lane];
    dst[lane] = __fmul_rn(__fsub_rn(q, c), inv_r);
  }
  const float *f = feat_pm + ((size_t)b * N + src) * C;
  for (int k = lane; k < C; k += 32) dst[c0 + k] = f[k];
  for (int k = c0 + C + lane; k < ld; k += 32) dst[k] = 0.f;
}

// ---- BatchNorm statistics: per-tile partials -> mean / var -> affine, running statistics ------------------------------
// partial layout (sa_tcp_kernel TRAIN epilogue): [(tile * 4 + quarter) * 2 + {sum, sumsq}][256]
__global__ void __launch_bounds__(256)
bn_finalize_kernel(int entries, long long R, int cout, float eps, float momentum, const float *__restrict__ partial,
                   const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ running_mean,
                   float *__restrict__ running_var, float *__restrict__ mean_out, float *__restrict__ invstd_out,
                   float *__restrict__ scale_out, float *__restrict__ shift_out) {
  __shared__ double s1[256], s2[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int e = tid; e < entries; e += 256) {  // fixed assignment + fixed tree below: deterministic
    a += (double)partial[((size_t)e * 2) * 256 + c];
    b += (double)partial[((size_t)e * 2 + 1) * 256 + c];
  }
  s1[tid] = a; s2[tid] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    const double m = s1[0] / (double)R;
    double v = s2[0] / (double)R - m * m;  // biased variance: what BatchNorm normalises with
    if (v < 0.0) v = 0.0;
    const float mean = (float)m, var = (float)v;
    const float invstd = 1.0f / sqrtf(var + eps);
    const float sc = gamma[c] * invstd;
    mean_out[c] = mean; invstd_out[c] = invstd; scale_out[c] = sc; shift_out[c] = beta[c] - mean * sc;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) {
      const double unbiased = v * ((double)R / (double)(R > 1 ? R - 1 : 1));  // torch: momentum update uses the unbiased variance
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// generic fixed-order reduction of `entries` partial vectors of `width` floats (stride `stride`) -> out[width] (+ optional scale)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(int entries, int width, size_t stride, const float *__restrict__ partial, float *__restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= width) return;
  double a = 0.0;
  for (int e = 0; e < entries; ++e) a += (double)partial[(size_t)e * stride + i];
  out[i] = (float)a;
}

// ---- out = max over nsample of relu(scale z_L + shift), arg-max slot (first maximum wins, like max_pool2d) -----------
__global__ void __launch_bounds__(256)
bn_relu_maxpool_kernel(long long G, int M, int ns, int CL, const float *__restrict__ z, const float *__restrict__ scale,
                       const float *__restrict__ shift, float *__restrict__ out, int32_t *__restrict__ arg_pm) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;  // (centre g, channel c), c fastest: coalesced row reads
  if (e >= G * CL) return;
  const long long g = e / CL;
  const int c = (int)(e - g * CL);
  const float sc = scale[c], sh = shift[c];
  const float *col = z + (g * ns) * CL + c;
  float best = -1.f;
  int bi = 0;
  for (int s = 0; s < ns; ++s) {
    const float y = fmaxf(fmaf(col[(size_t)s * CL], sc, sh), 0.f);
    if (y > best) { best = y; bi = s; }
  }
  const int b = (int)(g / M), m = (int)(g - (long long)b * M);
  out[((size_t)b * CL + c) * M + m] = best;
  arg_pm[e] = bi;
}

// ---- top layer: S1 = sum g, S2 = sum g * xhat over the arg-max rows (g = grad_out where the pooled value is > 0) -----
__global__ void __launch_bounds__(256)
pool_bwd_stats_kernel(long long G, int ns, int CL, const float *__restrict__ z, const float *__restrict__ gout_pm,
                      const int32_t *__restrict__ arg_pm, const float *__restrict__ mean, const float *__restrict__ invstd,
                      const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ S1,
                      float *__restrict__ S2) {
  __shared__ double r1[256], r2[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  const float mu = mean[c], is = invstd[c], sc = scale[c], sh = shift[c];
  double a = 0.0, b = 0.0;
  for (long long g = tid; g < G; g += 256) {
    const int s = arg_pm[g * CL + c];
    const float zz = z[(g * ns + s) * CL + c];
    if (fmaf(zz, sc, sh) > 0.f) {
      const float go = gout_pm[g * CL + c];
      a += (double)go;
      b += (double)(go * ((zz - mu) * is));
    }
  }
  r1[tid] = a; r2[tid] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { r1[tid] += r1[tid + o]; r2[tid] += r2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) { S1[c] = (float)r1[0]; S2[c] = (float)r2[0]; }
}

// BatchNorm backward folded to per-channel coefficients:  dz = scale * g + b + c * z  with
//   b = scale * (m2 * invstd * mean - m1),  c = -scale * m2 * invstd,  m1 = S1 / R,  m2 = S2 / R
__global__ void __launch_bounds__(256)
dz_coeff_kernel(int C, float inv_R, const float *__restrict__ mean, const float *__restrict__ invstd,
                const float *__restrict__ scale, const float *__restrict__ S1, const float *__restrict__ S2,
                float *__restrict__ coef_b, float *__restrict__ coef_c) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float m1 = S1[c] * inv_R, m2 = S2[c] * inv_R;
  coef_b[c] = scale[c] * (m2 * invstd[c] * mean[c] - m1);
  coef_c[c] = -scale[c] * m2 * invstd[c];
}

// per-tile column sums of the tensor-core epilogue ([(entry) * 2 + which][256]) -> two vectors, fixed order, fp64
__global__ void __launch_bounds__(256)
stats_reduce_kernel(int entries, const float *__restrict__ partial, float *__restrict__ out0, float *__restrict__ out1) {
  __shared__ double s1[256], s2[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int e = tid; e < entries; e += 256) {
    a += (double)partial[((size_t)e * 2) * 256 + c];
    b += (double)partial[((size_t)e * 2 + 1) * 256 + c];
  }
  s1[tid] = a; s2[tid] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) { out0[c] = (float)s1[0]; out1[c] = (float)s2[0]; }
}

// ---- dz_l built on the fly ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float dz_at(const DzSrc &d, long long r, int c) {
  const float zz = d.z[r * d.C + c];
  float g;
  if (d.g) {
    g = d.g[r * d.C + c];
  } else {
    const long long grp = r / d.ns;
    const int slot = (int)(r - grp * d.ns);
    g = (d.arg_pm[grp * d.C + c] == slot && fmaf(zz, d.scale[c], d.shift[c]) > 0.f) ? d.gout_pm[grp * d.C + c] : 0.f;
  }
  const float xhat = (zz - d.mean[c]) * d.invstd[c];
  return d.scale[c] * (g - d.S1[c] * d.inv_R - xhat * (d.S2[c] * d.inv_R));
}

// ---- g_{l-1} = (dz_l W_l) * [a_{l-1} > 0]  (+ column sums of g and g * xhat for the BatchNorm below) ------------------
// Tile: 64 rows x all cin columns; 256 threads = 16 (row groups of 4) x 16 (column groups of 4 within each 64-column block).
constexpr int DA_TR = 64, DA_NJ = 5;  // cin <= 64 * DA_NJ
__global__ void __launch_bounds__(256, 1)
bwd_da_kernel(long long R, int cin, int cout, int cin_s, DzSrc dzs, const float *__restrict__ W, ActSrc act,
              float *__restrict__ g_out, int g_ld, float *__restrict__ stats_partial) {
  extern __shared__ float smem[];
  float *Ws = smem;                               // [cout][cin_s], zero padded
  float *Dz = Ws + (size_t)cout * cin_s;          // [DA_TR][cout + 1]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int dzld = cout + 1;
  for (int e = tid; e < cout * cin_s; e += 256) {
    const int k = e / cin_s, n = e - k * cin_s;
    Ws[e] = n < cin ? W[(size_t)k * cin + n] : 0.f;
  }
  float s1[DA_NJ][4], s2[DA_NJ][4];
#pragma unroll
  for (int j = 0; j < DA_NJ; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) s1[j][e] = s2[j][e] = 0.f;
  const int nj = cin_s >> 6;
  const long long tiles = (R + DA_TR - 1) / DA_TR;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {  // static assignment: fixed summation order per CTA
    __syncthreads();
    const long long r0 = t * DA_TR;
    for (int e = tid; e < DA_TR * cout; e += 256) {
      const int rr = e / cout, c = e - rr * cout;
      Dz[rr * dzld + c] = (r0 + rr < R) ? dz_at(dzs, r0 + rr, c) : 0.f;
    }
    __syncthreads();
    float acc[4][DA_NJ][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < DA_NJ; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
    for (int k = 0; k < cout; ++k) {
      float a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Dz[(ty * 4 + i) * dzld + k];
#pragma unroll
      for (int j = 0; j < DA_NJ; ++j) {
        if (j < nj) {
          const float4 w = *reinterpret_cast<const float4 *>(Ws + (size_t)k * cin_s + j * 64 + tx * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][j][0] = fmaf(a[i], w.x, acc[i][j][0]);
            acc[i][j][1] = fmaf(a[i], w.y, acc[i][j][1]);
            acc[i][j][2] = fmaf(a[i], w.z, acc[i][j][2]);
            acc[i][j][3] = fmaf(a[i], w.w, acc[i][j][3]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = r0 + ty * 4 + i;
      if (r >= R) continue;
#pragma unroll
      for (int j = 0; j < DA_NJ; ++j) {
        if (j >= nj) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = j * 64 + tx * 4 + e;
          if (n >= cin) continue;
          float g = acc[i][j][e];
          if (act.scale) {  // ReLU mask of a_{l-1} and the sums its BatchNorm's gradient needs
            const float zz = act.rows[r * act.ld + n];
            g = fmaf(zz, act.scale[n], act.shift[n]) > 0.f ? g : 0.f;
            s1[j][e] += g;
            s2[j][e] += g * ((zz - act.mean[n]) * act.invstd[n]);
          }
          g_out[r * g_ld + n] = g;
        }
      }
    }
  }
  if (stats_partial) {
    // combine the 16 row groups in a fixed order: [ty][cin_s] through shared memory (reuses the Dz region when it fits)
    __syncthreads();
    float *red = smem;  // Ws is no longer needed
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int j = 0; j < DA_NJ; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j < nj) red[ty * cin_s + j * 64 + tx * 4 + e] = pass == 0 ? s1[j][e] : s2[j][e];
      __syncthreads();
      for (int n = tid; n < cin; n += 256) {
        float a = 0.f;
        for (int y = 0; y < 16; ++y) a += red[y * cin_s + n];
        stats_partial[((size_t)blockIdx.x * 2 + pass) * cin_s + n] = a;
      }
      __syncthreads();
    }
  }
}

// ---- dW_l partials = dz_l^T a_{l-1} over a strided share of the rows ---------------------------------------------------
// CTA: output tile 64 (cout) x 64 (cin); 256 threads = 16 x 16, 4 x 4 outputs each; rows in chunks of 32.
__global__ void __launch_bounds__(256)
bwd_dw_kernel(long long R, int cin, int cout, DzSrc dzs, ActSrc act, float *__restrict__ partial) {
  __shared__ __align__(16) float Dz[32][64];
  __shared__ __align__(16) float A[32][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ct = blockIdx.y / ((cin + 63) / 64), nt = blockIdx.y - ct * ((cin + 63) / 64);
  const int c0 = ct * 64, n0 = nt * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long chunks = (R + 31) / 32;
  for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const long long r0 = ch * 32;
    __syncthreads();
    for (int e = tid; e < 32 * 64; e += 256) {
      const int rr = e >> 6, k = e & 63;
      const long long r = r0 + rr;
      const int c = c0 + k, n = n0 + k;
      Dz[rr][k] = (r < R && c < cout) ? dz_at(dzs, r, c) : 0.f;
      float a = 0.f;
      if (r < R && n < cin) {
        a = act.rows[r * act.ld + n];
        if (act.scale) a = fmaxf(fmaf(a, act.scale[n], act.shift[n]), 0.f);
      }
      A[rr][k] = a;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const float4 d = *reinterpret_cast<const float4 *>(&Dz[rr][ty * 4]);
      const float4 a = *reinterpret_cast<const float4 *>(&A[rr][tx * 4]);
      const float dv[4] = {d.x, d.y, d.z, d.w}, av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], av[j], acc[i][j]);
    }
  }
  float *dst = partial + (size_t)blockIdx.x * cout * cin;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (c < cout && n < cin) dst[(size_t)c * cin + n] = acc[i][j];
    }
}

// ---- d features: scatter-add of the feature columns of da_0 (group_points_gpu.cu:48-68 semantics) ---------------------
__global__ void __launch_bounds__(256)
rows_scatter_kernel(long long R, int N, int M, int ns, int C, int ld, int col0, const float *__restrict__ da0,
                    const int32_t *__restrict__ idx, float *__restrict__ grad_feat) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const int b = (int)(r / ((long long)M * ns));
  const int src = idx[r];
  for (int c = lane; c < C; c += 32) atomicAdd(&grad_feat[((size_t)b * C + c) * N + src], da0[r * ld + col0 + c]);
}

__global__ void __launch_bounds__(256) fill_kernel(float *p, float v, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- host side ---------------------------------------------------------------------------------------------------------
static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TrainLayout {
  long long R, G;
  int L, ld0, C0;
  int cout[TRN_MAXL], cin[TRN_MAXL];
  size_t off_a0, off_z[TRN_MAXL], off_par[TRN_MAXL], off_arg, saved_bytes;  // par: mean | invstd | scale | shift (4 x 256 floats)
  long long tiles;
  size_t ws_fwd, ws_bwd;
};

static bool train_layout(int rows_mode, long long R, int ns, int C0, int L, const b200_bn_layer *layers, TrainLayout &t) {
  if (L < 1 || L > TRN_MAXL || R <= 0 || ns <= 0 || (R % ns) != 0) return false;
  t.R = R; t.G = R / ns; t.L = L; t.C0 = C0;
  t.ld0 = (C0 + 3) & ~3;
  size_t off = 0;
  t.off_a0 = off;
  if (!rows_mode) off += a256(sizeof(float) * (size_t)R * t.ld0);
  for (int l = 0; l < L; ++l) {
    t.cin[l] = layers[l].cin; t.cout[l] = layers[l].cout;
    if (t.cout[l] < 1 || t.cout[l] > 256 || (t.cout[l] & 3) != 0) return false;
    if (t.cin[l] != (l == 0 ? C0 : layers[l - 1].cout) || t.cin[l] > 64 * DA_NJ) return false;
    t.off_z[l] = off; off += a256(sizeof(float) * (size_t)R * t.cout[l]);
    t.off_par[l] = off; off += 4 * 256 * sizeof(float);
  }
  t.off_arg = off; off += a256(sizeof(int32_t) * (size_t)t.G * t.cout[L - 1]);
  t.saved_bytes = off;
  t.tiles = (R + TC_ROWS - 1) / TC_ROWS;
  t.ws_fwd = a256((size_t)t.tiles * 4 * 2 * 256 * sizeof(float));
  // backward: grad_out point-major | S1,S2 per layer | g buffers (two, ping-pong, widest hidden) | da_0 | partials
  int wmax = 4;
  for (int l = 0; l + 1 < L; ++l) wmax = t.cout[l] > wmax ? t.cout[l] : wmax;
  size_t wpart = 0;
  for (int l = 0; l < L; ++l) wpart = (size_t)t.cout[l] * t.cin[l] > wpart ? (size_t)t.cout[l] * t.cin[l] : wpart;
  t.ws_bwd = a256(sizeof(float) * (size_t)t.G * t.cout[L - 1]) + (size_t)L * 2 * 256 * sizeof(float) +
             2 * a256(sizeof(float) * (size_t)R * wmax) + a256(sizeof(float) * (size_t)R * t.ld0) +
             a256(sizeof(float) * 320 * wpart) + a256(sizeof(float) * 2 * 160 * 320) + t.ws_fwd + 4 * 256 * sizeof(float);
  return true;
}

static int train_forward_rows(const TrainLayout &t, const float *a0, int ns, int M, const b200_bn_layer *layers, float eps,
                              float momentum, float *out, uint8_t *saved, uint8_t *ws, cudaStream_t stream) {
  const float *in_rows = a0;
  int in_ld = t.ld0, in_C = t.C0;
  const float *in_scale = nullptr, *in_shift = nullptr;
  float *stats = reinterpret_cast<float *>(ws);
  ScratchGuard ones_guard;
  B200_CUDA_OK(ones_guard.alloc(2 * 256 * sizeof(float), stream));
  float *ones = (float *)ones_guard.ptr, *zeros = ones + 256;
  fill_kernel<<<1, 256, 0, stream>>>(ones, 1.f, 256);
  B200_LAUNCH_OK("fill_kernel");
  B200_CUDA_OK(cudaMemsetAsync(zeros, 0, 256 * sizeof(float), stream));
  for (int l = 0; l < t.L; ++l) {
    float *z = reinterpret_cast<float *>(saved + t.off_z[l]);
    float *par = reinterpret_cast<float *>(saved + t.off_par[l]);
    b200_mlp_layer raw;  // raw conv output: unit scale, zero shift, no ReLU
    raw.cin = t.cin[l]; raw.cout = t.cout[l]; raw.weight = layers[l].weight; raw.scale = ones; raw.shift = zeros;
    TcCall c;
    c.mode = 2; c.B = 1; c.N = (int)t.R; c.M = (int)t.R; c.C = in_C; c.ld = in_ld; c.ns = 32; c.use_xyz = 0;
    c.feat_pm = in_rows; c.rowout = 1; c.final_relu = 0; c.rows_total = (int)t.R; c.rows_per_scene = (int)t.R;
    c.out_pm = z; c.num_layers = 1; c.layers = &raw;
    c.in_scale = in_scale; c.in_shift = in_shift; c.stats = stats;
    const int rc = sa_tc_run(c, stream);
    if (rc) return rc;
    bn_finalize_kernel<<<t.cout[l], 256, 0, stream>>>((int)(t.tiles * 4), t.R, t.cout[l], eps, momentum, stats, layers[l].gamma,
                                                     layers[l].beta, layers[l].running_mean, layers[l].running_var, par,
                                                     par + 256, par + 512, par + 768);
    B200_LAUNCH_OK("bn_finalize_kernel");
    in_rows = z; in_ld = t.cout[l]; in_C = t.cout[l];
    in_scale = par + 512; in_shift = par + 768;
  }
  const int CL = t.cout[t.L - 1];
  const float *par = reinterpret_cast<const float *>(saved + t.off_par[t.L - 1]);
  const long long elems = t.G * CL;
  bn_relu_maxpool_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, stream>>>(
      t.G, M, ns, CL, reinterpret_cast<const float *>(saved + t.off_z[t.L - 1]), par + 512, par + 768, out,
      reinterpret_cast<int32_t *>(saved + t.off_arg));
  B200_LAUNCH_OK("bn_relu_maxpool_kernel");
  return 0;
}

// da0_out: (R, da0_ld) rows receiving the gradient of the first layer's input columns [da0_col0, cin_0) -- packed from
// column 0 when the tensor-core path ran (*da0_is_full = 0), or the full-width rows of the FFMA fallback (*da0_is_full = 1)
static int train_backward_rows(const TrainLayout &t, const float *a0, int ns, int B, int M, const b200_bn_layer *layers,
                               const float *grad_out, const uint8_t *saved, float *const *grad_weight,
                               float *const *grad_gamma, float *const *grad_beta, float *da0_out, int da0_col0, int da0_ld,
                               int *da0_is_full, uint8_t *ws, cudaStream_t stream) {
  *da0_is_full = 0;
  const int L = t.L, CL = t.cout[L - 1];
  size_t off = 0;
  float *gout_pm = reinterpret_cast<float *>(ws + off); off += a256(sizeof(float) * (size_t)t.G * CL);
  float *S = reinterpret_cast<float *>(ws + off); off += (size_t)L * 2 * 256 * sizeof(float);
  int wmax = 4;
  for (int l = 0; l + 1 < L; ++l) wmax = t.cout[l] > wmax ? t.cout[l] : wmax;
  float *gbuf[2];
  gbuf[0] = reinterpret_cast<float *>(ws + off); off += a256(sizeof(float) * (size_t)t.R * wmax);
  gbuf[1] = reinterpret_cast<float *>(ws + off); off += a256(sizeof(float) * (size_t)t.R * wmax);
  float *da0 = da0_out ? da0_out : reinterpret_cast<float *>(ws + off);
  off += a256(sizeof(float) * (size_t)t.R * t.ld0);
  size_t wpart = 0;
  for (int l = 0; l < L; ++l) wpart = (size_t)t.cout[l] * t.cin[l] > wpart ? (size_t)t.cout[l] * t.cin[l] : wpart;
  float *dw_partial = reinterpret_cast<float *>(ws + off); off += a256(sizeof(float) * 320 * wpart);
  float *st_partial = reinterpret_cast<float *>(ws + off); off += a256(sizeof(float) * 2 * 160 * 320);
  float *tc_stats = reinterpret_cast<float *>(ws + off); off += t.ws_fwd;
  float *coef = reinterpret_cast<float *>(ws + off);  // [b | c | ones | zeros] x 256
  fill_kernel<<<1, 256, 0, stream>>>(coef + 512, 1.f, 256);
  B200_LAUNCH_OK("fill_kernel");
  B200_CUDA_OK(cudaMemsetAsync(coef + 768, 0, 256 * sizeof(float), stream));
  // grad_out (B, CL, M) -> point-major (G, CL)
  {
    const int rc = b200pn2_transpose_cn(B, CL, M, grad_out, gout_pm, 0, (b200_stream_t)stream);
    if (rc) return rc;
  }
  const float *z_top = reinterpret_cast<const float *>(saved + t.off_z[L - 1]);
  const float *par_top = reinterpret_cast<const float *>(saved + t.off_par[L - 1]);
  const int32_t *arg = reinterpret_cast<const int32_t *>(saved + t.off_arg);
  pool_bwd_stats_kernel<<<CL, 256, 0, stream>>>(t.G, ns, CL, z_top, gout_pm, arg, par_top, par_top + 256, par_top + 512,
                                                par_top + 768, S + (size_t)(L - 1) * 512, S + (size_t)(L - 1) * 512 + 256);
  B200_LAUNCH_OK("pool_bwd_stats_kernel");
  const int sms = num_sms();
  const float *g_cur = nullptr;  // top layer: sparse
  for (int l = L - 1; l >= 0; --l) {
    const float *par = reinterpret_cast<const float *>(saved + t.off_par[l]);
    float *S1 = S + (size_t)l * 512, *S2 = S1 + 256;
    if (grad_beta && grad_beta[l]) B200_CUDA_OK(cudaMemcpyAsync(grad_beta[l], S1, sizeof(float) * t.cout[l], cudaMemcpyDeviceToDevice, stream));
    if (grad_gamma && grad_gamma[l]) B200_CUDA_OK(cudaMemcpyAsync(grad_gamma[l], S2, sizeof(float) * t.cout[l], cudaMemcpyDeviceToDevice, stream));
    DzSrc dz;
    dz.z = reinterpret_cast<const float *>(saved + t.off_z[l]);
    dz.g = g_cur; dz.gout_pm = gout_pm; dz.arg_pm = arg;
    dz.mean = par; dz.invstd = par + 256; dz.scale = par + 512; dz.shift = par + 768; dz.S1 = S1; dz.S2 = S2;
    dz.C = t.cout[l]; dz.ns = ns; dz.inv_R = (float)(1.0 / (double)t.R);
    ActSrc act;
    if (l == 0) {
      act.rows = a0; act.ld = t.ld0; act.C = t.C0; act.scale = act.shift = act.mean = act.invstd = nullptr;
    } else {
      const float *pp = reinterpret_cast<const float *>(saved + t.off_par[l - 1]);
      act.rows = reinterpret_cast<const float *>(saved + t.off_z[l - 1]); act.ld = t.cout[l - 1]; act.C = t.cout[l - 1];
      act.mean = pp; act.invstd = pp + 256; act.scale = pp + 512; act.shift = pp + 768;
    }
    const int cin = t.cin[l], cout = t.cout[l];
    // dW_l
    static int use_tc_dw = -1;
    if (use_tc_dw < 0) {
      const char *e = getenv("B200_SA_TRAIN_TC");
      use_tc_dw = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (grad_weight && grad_weight[l] && use_tc_dw && (cout & 3) == 0) {
      // tensor cores (sa_train_dw.cu): both operands built and transposed on the fly
      dz_coeff_kernel<<<1, 256, 0, stream>>>(cout, dz.inv_R, dz.mean, dz.invstd, dz.scale, S1, S2, coef, coef + 256);
      B200_LAUNCH_OK("dz_coeff_kernel");
      DwParams dp;
      dp.R = t.R; dp.cin = cin; dp.cout = cout; dp.dz = dz; dp.act = act; dp.coef_b = coef; dp.coef_c = coef + 256;
      dp.partial = dw_partial;
      dp.vec_act = ((act.ld & 3) == 0 && (((uintptr_t)act.rows) & 15) == 0) ? 1 : 0;
      const int otiles = ((cout + 127) / 128) * ((cin + 127) / 128);
      const long long nkb = (t.R + 31) / 32;
      int splits = (sms + otiles - 1) / otiles;
      if (splits > 320) splits = 320;
      if (splits > nkb) splits = (int)nkb;
      const int rc = dw_tc_launch(dp, splits, stream);
      if (rc) return rc;
      reduce_partials_kernel<<<(cout * cin + 255) / 256, 256, 0, stream>>>(splits, cout * cin, (size_t)cout * cin, dw_partial,
                                                                          grad_weight[l]);
      B200_LAUNCH_OK("reduce_partials_kernel");
    } else if (grad_weight && grad_weight[l]) {
      const int otiles = ((cout + 63) / 64) * ((cin + 63) / 64);
      long long chunks = (t.R + 31) / 32;
      int splits = (2 * sms + otiles - 1) / otiles;
      if (splits > 320) splits = 320;
      if (splits > chunks) splits = (int)chunks;
      bwd_dw_kernel<<<dim3(splits, otiles), 256, 0, stream>>>(t.R, cin, cout, dz, act, dw_partial);
      B200_LAUNCH_OK("bwd_dw_kernel");
      reduce_partials_kernel<<<(cout * cin + 255) / 256, 256, 0, stream>>>(splits, cout * cin, (size_t)cout * cin, dw_partial,
                                                                          grad_weight[l]);
      B200_LAUNCH_OK("reduce_partials_kernel");
    }
    // g_{l-1} (or da_0: only the columns of the caller's features, da0_col0 onwards)
    const bool need_da = l > 0 || da0_out != nullptr;
    if (need_da) {
      const int col0 = l == 0 ? da0_col0 : 0;
      const int n_out = cin - col0;
      float *g_next = l == 0 ? da0 : gbuf[l & 1];
      float *n1 = l > 0 ? S + (size_t)(l - 1) * 512 : nullptr, *n2 = l > 0 ? n1 + 256 : nullptr;
      static int use_tc = -1;
      if (use_tc < 0) {
        const char *e = getenv("B200_SA_TRAIN_TC");
        use_tc = (e && atoi(e) == 0) ? 0 : 1;
      }
      if (use_tc && (cout & 3) == 0 && (n_out & 3) == 0 && n_out <= 256) {
        // tensor cores: the forward's row-GEMM kernel on W_l^T, dz built in its producers, ReLU mask + statistics of
        // the layer below in its epilogue
        dz_coeff_kernel<<<1, 256, 0, stream>>>(cout, dz.inv_R, dz.mean, dz.invstd, dz.scale, S1, S2, coef, coef + 256);
        B200_LAUNCH_OK("dz_coeff_kernel");
        b200_mlp_layer lay;
        lay.cin = cout; lay.cout = n_out; lay.weight = layers[l].weight + col0; lay.scale = coef + 512; lay.shift = coef + 768;
        TcCall c;
        c.mode = 2; c.B = 1; c.N = (int)t.R; c.M = (int)t.R; c.C = cout; c.ld = cout; c.ns = 32; c.use_xyz = 0;
        c.feat_pm = dz.z; c.rowout = 1; c.final_relu = 0; c.rows_total = (int)t.R; c.rows_per_scene = (int)t.R;
        c.out_pm = g_next; c.num_layers = 1; c.layers = &lay; c.w_transposed = 1; c.w_ld = cin;
        c.train_in = dz.g ? 2 : 3; c.in_scale = dz.scale; c.in_shift = dz.shift; c.dz_b = coef; c.dz_c = coef + 256;
        c.g_rows = dz.g; c.gout_pm = dz.gout_pm; c.arg_pm = dz.arg_pm; c.pool_ns = ns;
        if (l > 0) {
          c.train_out = 1; c.zprev = act.rows; c.out_scale = act.scale; c.out_shift = act.shift; c.out_mean = act.mean;
          c.out_invstd = act.invstd; c.stats = tc_stats;
        }
        const int rc = sa_tc_run(c, stream);
        if (rc) return rc;
        if (l > 0) {
          stats_reduce_kernel<<<cin, 256, 0, stream>>>((int)(t.tiles * 4), tc_stats, n1, n2);
          B200_LAUNCH_OK("stats_reduce_kernel");
        }
      } else {
        // fp32 FFMA fallback (widths that are no multiple of 4): full-width rows, then the caller's columns are read with col0
        B200_CHECK_ARG(l > 0 || col0 == 0 || da0_ld == t.ld0, "sa_train_backward: internal layout error");
        const int cin_s = ((cin + 63) / 64) * 64;
        const size_t smem = sizeof(float) * ((size_t)cout * cin_s + (size_t)DA_TR * (cout + 1));
        const size_t smem_red = sizeof(float) * 16 * (size_t)cin_s;
        const size_t smem_all = smem > smem_red ? smem : smem_red;
        B200_CHECK_ARG(smem_all <= 227 * 1024, "sa_train_backward: layer %d (%d -> %d) needs %zu B of shared memory", l, cin, cout, smem_all);
        static DynSmemOptIn optin;
        B200_CUDA_OK(optin.ensure(bwd_da_kernel, smem_all));
        const long long tiles = (t.R + DA_TR - 1) / DA_TR;
        int grid = tiles < sms ? (int)tiles : sms;
        if (grid > 160) grid = 160;
        bwd_da_kernel<<<grid, 256, smem_all, stream>>>(t.R, cin, cout, cin_s, dz, layers[l].weight, act, g_next,
                                                      l == 0 ? t.ld0 : cin, l == 0 ? nullptr : st_partial);
        B200_LAUNCH_OK("bwd_da_kernel");
        if (l > 0) {
          reduce_partials_kernel<<<(cin + 255) / 256, 256, 0, stream>>>(grid, cin, (size_t)2 * cin_s, st_partial, n1);
          B200_LAUNCH_OK("reduce_partials_kernel");
          reduce_partials_kernel<<<(cin + 255) / 256, 256, 0, stream>>>(grid, cin, (size_t)2 * cin_s, st_partial + cin_s, n2);
          B200_LAUNCH_OK("reduce_partials_kernel");
        }
        if (l == 0) *da0_is_full = 1;
      }
      g_cur = g_next;
    }
  }
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200pn2_sa_train_saved_bytes(int B, int M, int nsample, int C, int use_xyz, int num_layers,
                                               const b200_bn_layer *layers, int rows_mode) {
  TrainLayout t;
  const int C0 = rows_mode ? C : C + (use_xyz ? 3 : 0);
  if (!layers || !train_layout(rows_mode, (long long)B * M * nsample, nsample, C0, num_layers, layers, t)) return 0;
  return t.saved_bytes;
}

extern "C" size_t b200pn2_sa_train_workspace_bytes(int B, int M, int nsample, int C, int use_xyz, int num_layers,
                                                   const b200_bn_layer *layers, int rows_mode, int backward) {
  TrainLayout t;
  const int C0 = rows_mode ? C : C + (use_xyz ? 3 : 0);
  if (!layers || !train_layout(rows_mode, (long long)B * M * nsample, nsample, C0, num_layers, layers, t)) return 0;
  return backward ? t.ws_bwd : t.ws_fwd;
}

// rows mode (x_rows != NULL): the stack on given rows (B * M * nsample, C), no gather -- the oracle's interface.
extern "C" int b200pn2_sa_train_forward(int B, int N, int M, int C, float radius, int nsample, int use_xyz, int normalize_xyz,
                                        const float *xyz, const float *features_pm, const float *new_xyz, const int32_t *idx,
                                        const float *x_rows, int num_layers, const b200_bn_layer *layers, float eps,
                                        float momentum, float *out, void *saved, size_t saved_bytes, void *workspace,
                                        size_t workspace_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rows_mode = x_rows ? 1 : 0;
  const int C0 = rows_mode ? C : C + (use_xyz ? 3 : 0);
  TrainLayout t;
  B200_CHECK_ARG(layers && out && saved, "sa_train_forward: null pointer");
  B200_CHECK_ARG(train_layout(rows_mode, (long long)B * M * nsample, nsample, C0, num_layers, layers, t),
                 "sa_train_forward: unsupported stack (1..4 layers, widths multiples of 4 and <= 256, cin <= 320)");
  B200_CHECK_ARG(t.R < (1ll << 31) - 256, "sa_train_forward: too many rows");
  B200_CHECK_ARG(saved_bytes >= t.saved_bytes && workspace && workspace_bytes >= t.ws_fwd, "sa_train_forward: buffers too small");
  const float *a0 = x_rows;
  if (!rows_mode) {
    B200_CHECK_ARG(xyz && new_xyz && idx && (C == 0 || features_pm), "sa_train_forward: null pointer");
    float *rows = reinterpret_cast<float *>((uint8_t *)saved + t.off_a0);
    const float inv_r = normalize_xyz ? (float)(1.0 / (double)radius) : 1.0f;
    group_rows_kernel<<<(unsigned)((t.R + 7) / 8), 256, 0, stream>>>(N, M, nsample, C, t.ld0, use_xyz, inv_r, xyz, new_xyz,
                                                                     features_pm, idx, rows, t.R);
    B200_LAUNCH_OK("group_rows_kernel");
    a0 = rows;
  } else {
    B200_CHECK_ARG((C & 3) == 0, "sa_train_forward(rows): C must be a multiple of 4");
  }
  return train_forward_rows(t, a0, nsample, M, layers, eps, momentum, out, (uint8_t *)saved, (uint8_t *)workspace, stream);
}

extern "C" int b200pn2_sa_train_backward(int B, int N, int M, int C, int nsample, int use_xyz, const int32_t *idx,
                                         const float *x_rows, int num_layers, const b200_bn_layer *layers,
                                         const float *grad_out, const void *saved, size_t saved_bytes,
                                         float *grad_features, float *grad_rows, float *const *grad_weight,
                                         float *const *grad_gamma, float *const *grad_beta, void *workspace,
                                         size_t workspace_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rows_mode = x_rows ? 1 : 0;
  const int C0 = rows_mode ? C : C + (use_xyz ? 3 : 0);
  TrainLayout t;
  B200_CHECK_ARG(layers && grad_out && saved, "sa_train_backward: null pointer");
  B200_CHECK_ARG(train_layout(rows_mode, (long long)B * M * nsample, nsample, C0, num_layers, layers, t),
                 "sa_train_backward: unsupported stack");
  B200_CHECK_ARG(saved_bytes >= t.saved_bytes && workspace && workspace_bytes >= t.ws_bwd, "sa_train_backward: buffers too small");
  const float *a0 = rows_mode ? x_rows : reinterpret_cast<const float *>((const uint8_t *)saved + t.off_a0);
  float *da0 = nullptr;
  ScratchGuard da_guard;
  const int col0 = (!rows_mode && use_xyz) ? 3 : 0;
  if (rows_mode && grad_rows) {
    da0 = grad_rows;  // (R, C) with ld0 == C in rows mode
  } else if (!rows_mode && grad_features && C > 0) {
    B200_CUDA_OK(da_guard.alloc(sizeof(float) * (size_t)t.R * t.ld0, stream));
    da0 = (float *)da_guard.ptr;
  }
  int full = 0;
  const int rc = train_backward_rows(t, a0, nsample, B, M, layers, grad_out, (const uint8_t *)saved, grad_weight, grad_gamma,
                                     grad_beta, da0, col0, t.ld0, &full, (uint8_t *)workspace, stream);
  if (rc) return rc;
  if (!rows_mode && grad_features && C > 0) {
    B200_CHECK_ARG(idx, "sa_train_backward: idx needed for the feature gradient");
    B200_CUDA_OK(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)B * C * N, stream));
    rows_scatter_kernel<<<(unsigned)((t.R + 7) / 8), 256, 0, stream>>>(t.R, N, M, nsample, C, full ? t.ld0 : C, full ? col0 : 0,
                                                                       da0, idx, grad_features);
    B200_LAUNCH_OK("rows_scatter_kernel");
  }
  return 0;
}
