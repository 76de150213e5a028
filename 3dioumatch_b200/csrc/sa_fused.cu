// sa_fused.cu -- fused PointNet++ set-abstraction forward for sm_100a (fp32 FFMA version).
//
// Replaces, for PointnetSAModuleVotes.forward (pointnet2/pointnet2_modules.py:215-277):
//     ball_query -> group_points(xyz) -> "-= centre", "/= radius" -> group_points(features) -> torch.cat
//     -> SharedMLP (1x1 Conv2d + eval BatchNorm2d + ReLU, pytorch_utils.py:14-39) x L -> max_pool2d over nsample
// which in the reference materialises (B, C, npoint, nsample) tensors in HBM five to ten times per layer
// (1.47 GB of traffic per scene).
//
// Here one CTA owns a tile of SA_R = G*nsample grouped rows (G centres).  The grouped rows are gathered
// straight into shared memory (cp.async, 16 B per request, from point-major features), the whole MLP chain
// runs shared-memory -> registers -> shared-memory as register-tiled fp32 GEMMs (8x8 outputs per thread,
// weights streamed through a double-buffered shared tile), and the last layer's epilogue max-reduces over the
// nsample rows of each centre on chip.  HBM sees: indices in, gathered rows in (L2-resident), (B,Cout,M) out.
//
// Arithmetic: fp32 FFMA with fp32 accumulation (parity target 1e-5 vs the fp32 reference rules out plain
// TF32/BF16 tensor-core math; see DESIGN.md for the 3xTF32 tcgen05 plan).
#include <limits.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"
#include "sa_tc.cuh"

namespace b200 {

constexpr int SA_THREADS = 256;
constexpr int SA_R = 128;    // grouped rows per CTA
constexpr int SA_KC = 8;     // k-chunk of the streamed weight tile
constexpr int SA_MAXL = 4;   // layers supported in one launch
constexpr int SA_MAXG = 8;   // centres per CTA
constexpr int SA_WLD = 128 + 4;

struct SaLayer {
  const float *w, *scale, *shift;
  int cin, cout;
};

struct SaParams {
  int B, N, M, C, ns, G, use_xyz, nl;
  float inv_r;  // 1/radius or 1
  const float *xyz, *feat_cm, *feat_pm, *new_xyz;
  const int32_t *idx;
  float *out, *out_pm;
  SaLayer L[SA_MAXL];
  int ldA, ldB;  // row strides (floats) of the two activation buffers
  int vec_gather;  // features_pm is usable with 16-byte cp.async
};

__device__ __forceinline__ int round8(int x) { return (x + 7) & ~7; }

// monotone float -> int key, so that atomicMax on ints is a float max for any sign
__device__ __forceinline__ int f2ord(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// One MLP layer on the CTA's SA_R x cin activation tile.
//   As  : [SA_R][lda] row-major activations in shared memory, columns [cin, round8(cin)) are zero
//   Wg  : (cout, cin) row-major weights in global memory (nn.Conv2d.weight viewed 2-D)
//   perm_c >= 0 : first layer; activation column k holds source channel (k < perm_c ? 3 + k : k - perm_c)
//                 (features first, relative xyz last, so that feature rows land 16-byte aligned)
//   hidden layer: Out[r][c] = relu(scale*acc + shift), zero for c in [cout, round8(cout))
//   last layer  : maxbuf[g][c] = max over the nsample rows of centre g
template <int TN>
__device__ __forceinline__ void mlp_layer(const float *__restrict__ As, int lda, const SaLayer &ly, int perm_c,
                                          float *__restrict__ Ws, float *__restrict__ Out, int ldo,
                                          int *__restrict__ maxbuf, int ns, int valid_rows, bool last) {
  constexpr int NT = 64 * TN;
  constexpr int WPT = SA_KC * NT / SA_THREADS;  // weight elements staged per thread per chunk
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int cin = ly.cin, cout = ly.cout;
  const int kpad = round8(cin);
  const int nchunks = kpad / SA_KC;
  const float *__restrict__ Wg = ly.w;

  for (int n0 = 0; n0 < cout; n0 += NT) {
    float acc[8][4 * TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4 * TN; ++j) acc[i][j] = 0.f;

    float wreg[WPT];
    auto load_chunk = [&](int kc) {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const int e = tid + i * SA_THREADS;
        const int kk = e & (SA_KC - 1), n = e / SA_KC;
        const int k = kc * SA_KC + kk;
        float v = 0.f;
        if (k < cin && n0 + n < cout) {
          const int src = perm_c < 0 ? k : (k < perm_c ? 3 + k : k - perm_c);
          v = Wg[(size_t)(n0 + n) * cin + src];
        }
        wreg[i] = v;
      }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const int e = tid + i * SA_THREADS;
        const int kk = e & (SA_KC - 1), n = e / SA_KC;
        Ws[buf * SA_KC * SA_WLD + kk * SA_WLD + n] = wreg[i];
      }
    };

    __syncthreads();  // previous users of Ws / As writers are done
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    int cur = 0;
    for (int kc = 0; kc < nchunks; ++kc) {
      if (kc + 1 < nchunks) load_chunk(kc + 1);
      const float *wb = Ws + cur * SA_KC * SA_WLD;
#pragma unroll
      for (int k4 = 0; k4 < SA_KC; k4 += 4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
          a[i] = *reinterpret_cast<const float4 *>(As + (size_t)row * lda + kc * SA_KC + k4);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float bv[4 * TN];
#pragma unroll
          for (int jq = 0; jq < TN; ++jq) {
            const float4 t = *reinterpret_cast<const float4 *>(wb + (k4 + kk) * SA_WLD + jq * 64 + tx * 4);
            bv[jq * 4 + 0] = t.x; bv[jq * 4 + 1] = t.y; bv[jq * 4 + 2] = t.z; bv[jq * 4 + 3] = t.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
#pragma unroll
            for (int j = 0; j < 4 * TN; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
          }
        }
      }
      if (kc + 1 < nchunks) store_chunk(cur ^ 1);
      __syncthreads();
      cur ^= 1;
    }

    // ---- epilogue ------------------------------------------------------------------------------
    float sc[4 * TN], sh[4 * TN];
    bool cvalid[4 * TN];
#pragma unroll
    for (int j = 0; j < 4 * TN; ++j) {
      const int col = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      cvalid[j] = col < cout;
      sc[j] = cvalid[j] ? ly.scale[col] : 0.f;
      sh[j] = cvalid[j] ? ly.shift[col] : 0.f;
    }
    if (!last) {
      const int kpad_next = round8(cout);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
#pragma unroll
        for (int jq = 0; jq < TN; ++jq) {
          const int col0 = n0 + jq * 64 + tx * 4;
          if (col0 < kpad_next) {  // kpad_next is a multiple of 8, col0 of 4: the float4 stays inside the row
            float4 v;
            v.x = cvalid[jq * 4 + 0] ? fmaxf(fmaf(acc[i][jq * 4 + 0], sc[jq * 4 + 0], sh[jq * 4 + 0]), 0.f) : 0.f;
            v.y = cvalid[jq * 4 + 1] ? fmaxf(fmaf(acc[i][jq * 4 + 1], sc[jq * 4 + 1], sh[jq * 4 + 1]), 0.f) : 0.f;
            v.z = cvalid[jq * 4 + 2] ? fmaxf(fmaf(acc[i][jq * 4 + 2], sc[jq * 4 + 2], sh[jq * 4 + 2]), 0.f) : 0.f;
            v.w = cvalid[jq * 4 + 3] ? fmaxf(fmaf(acc[i][jq * 4 + 3], sc[jq * 4 + 3], sh[jq * 4 + 3]), 0.f) : 0.f;
            *reinterpret_cast<float4 *>(Out + (size_t)row * ldo + col0) = v;
          }
        }
      }
    } else {
      int rg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
        rg[i] = row < valid_rows ? row / ns : -1;
      }
#pragma unroll
      for (int j = 0; j < 4 * TN; ++j) {
        if (!cvalid[j]) continue;
        const int col = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
        int curg = -1;
        float curm = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (rg[i] < 0) continue;
          const float v = fmaxf(fmaf(acc[i][j], sc[j], sh[j]), 0.f);
          const int g = rg[i];
          if (g != curg) {
            if (curg >= 0) atomicMax(maxbuf + curg * cout + col, f2ord(curm));
            curg = g;
            curm = v;
          } else {
            curm = fmaxf(curm, v);
          }
        }
        if (curg >= 0) atomicMax(maxbuf + curg * cout + col, f2ord(curm));
      }
    }
  }
}

__global__ void __launch_bounds__(SA_THREADS, 1) sa_mlp_max_kernel(const SaParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int m0 = blockIdx.x * p.G;
  const int ns = p.ns;
  const int g_here = min(p.G, p.M - m0);
  const int valid_rows = g_here * ns;
  const int cout_last = p.L[p.nl - 1].cout;

  float *bufA = smem;
  float *bufB = bufA + (size_t)SA_R * p.ldA;
  float *Ws = bufB + (size_t)SA_R * p.ldB;
  int *maxbuf = reinterpret_cast<int *>(Ws + 2 * SA_KC * SA_WLD);
  int *s_idx = maxbuf + SA_MAXG * cout_last;
  float *s_ctr = reinterpret_cast<float *>(s_idx + SA_R);

  // ---- neighbour indices and centres of this tile ------------------------------------------------
  if (tid < SA_R) {
    int v = -1;
    if (tid < valid_rows) {
      const int g = tid / ns, s = tid - g * ns;
      v = p.idx[((size_t)b * p.M + m0 + g) * ns + s];
    }
    s_idx[tid] = v;
  }
  if (tid < g_here * 3) s_ctr[tid] = p.new_xyz[((size_t)b * p.M + m0) * 3 + tid];
  for (int e = tid; e < SA_MAXG * cout_last; e += SA_THREADS) maxbuf[e] = INT_MIN;
  __syncthreads();

  // ---- gather the grouped rows into bufA: [features (C) | relative xyz (3) | zero pad] ---------------
  const int C = p.C, lda = p.ldA;
  const int cin0 = p.L[0].cin;
  const int kpad0 = round8(cin0);
  if (C > 0) {
    if (p.vec_gather) {
      const int nq = C >> 2;
      for (int e = tid; e < SA_R * nq; e += SA_THREADS) {
        const int r = e / nq, q = e - r * nq;
        const int i = s_idx[r];
        float *dst = bufA + (size_t)r * lda + q * 4;
        if (i >= 0) {
          cp_async16(dst, p.feat_pm + ((size_t)b * p.N + i) * C + q * 4);
        } else {
          *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
      for (int e = tid; e < SA_R * C; e += SA_THREADS) {
        const int c = e / SA_R, r = e - c * SA_R;
        const int i = s_idx[r];
        float v = 0.f;
        if (i >= 0)
          v = p.feat_cm ? p.feat_cm[((size_t)b * C + c) * p.N + i] : p.feat_pm[((size_t)b * p.N + i) * C + c];
        bufA[(size_t)r * lda + c] = v;
      }
    }
  }
  const int tailw = kpad0 - C;  // xyz (if used) + zero padding columns
  for (int e = tid; e < SA_R * tailw; e += SA_THREADS) {
    const int r = e / tailw, d = e - r * tailw;
    const int i = s_idx[r];
    float v = 0.f;
    if (p.use_xyz && d < 3 && i >= 0) {
      const int g = r / ns;
      // pointnet2_utils.py:351-353: grouped_xyz -= new_xyz ; grouped_xyz /= radius (x * fp32(1/r) on CUDA)
      v = __fmul_rn(__fsub_rn(p.xyz[((size_t)b * p.N + i) * 3 + d], s_ctr[g * 3 + d]), p.inv_r);
    }
    bufA[(size_t)r * lda + C + d] = v;
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- MLP chain --------------------------------------------------------------------------------------
  for (int l = 0; l < p.nl; ++l) {
    const bool last = (l == p.nl - 1);
    const float *As = (l & 1) ? bufB : bufA;
    const int la = (l & 1) ? p.ldB : p.ldA;
    float *Out = (l & 1) ? bufA : bufB;
    const int lo = (l & 1) ? p.ldA : p.ldB;
    const int perm_c = (l == 0 && p.use_xyz) ? C : -1;
    if (p.L[l].cout <= 64)
      mlp_layer<1>(As, la, p.L[l], perm_c, Ws, Out, lo, maxbuf, ns, valid_rows, last);
    else
      mlp_layer<2>(As, la, p.L[l], perm_c, Ws, Out, lo, maxbuf, ns, valid_rows, last);
  }
  __syncthreads();

  // ---- write the pooled features ------------------------------------------------------------------------
  const int G = p.G;
  for (int e = tid; e < G * cout_last; e += SA_THREADS) {
    const int c = e / G, g = e - c * G;
    if (g < g_here) p.out[((size_t)b * cout_last + c) * p.M + m0 + g] = ord2f(maxbuf[g * cout_last + c]);
  }
  if (p.out_pm != nullptr) {
    for (int e = tid; e < g_here * cout_last; e += SA_THREADS) {
      const int g = e / cout_last, c = e - g * cout_last;
      p.out_pm[((size_t)b * p.M + m0 + g) * cout_last + c] = ord2f(maxbuf[e]);
    }
  }
}

// (B,C,N) -> (B,N,C) so that one grouped row is one contiguous, 16-byte aligned run
// (B,C,N) -> (B,N,ld) with ld >= C; the padding columns [C, ld) are zero-filled
__global__ void __launch_bounds__(256) transpose_cn_kernel(int C, int N, int ld, const float *__restrict__ in,
                                                           float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, n = n0 + tx;
    tile[i][tx] = (c < C && n < N) ? in[((size_t)b * C + c) * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int n = n0 + i, c = c0 + tx;
    if (n < N && c < ld) out[((size_t)b * N + n) * ld + c] = c < C ? tile[tx][i] : 0.f;
  }
}

int ball_query_launch(int B, int N, int M, float radius, int nsample, const float *new_xyz, const float *xyz,
                      int32_t *idx, cudaStream_t stream, int *unit_list, int *unit_total);
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace b200

using namespace b200;

// Feature-propagation rows through the same fused MLP + max kernel: row (centre g, sample s) =
// [rel_xyz (3) | sum_t weight_t * known_feats[idx_t] (C)]  (models/grid_conv_module.py:87-113 of the reference:
// three_nn -> inverse-distance weights -> gather/blend -> cat(relative grid) -> SharedMLP -> max over the 64 grid points)
extern "C" int b200pn2_transpose_cn(int B, int C, int N, const float *in_cm, float *out_pm, int out_ld,
                                    b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(B >= 0 && C > 0 && N > 0 && in_cm && out_pm && (out_ld == 0 || out_ld >= C), "transpose_cn: bad arguments");
  B200_CHECK_ARG(B <= 65535, "transpose_cn: B=%d exceeds grid.z", B);
  if (B == 0) return 0;
  const int ld = out_ld > 0 ? out_ld : C;
  dim3 grid(ceil_div(N, 32), ceil_div(ld, 32), B);
  transpose_cn_kernel<<<grid, 256, 0, stream>>>(C, N, ld, in_cm, out_pm);
  B200_LAUNCH_OK("transpose_cn_kernel");
  return 0;
}

extern "C" int b200pn2_interp_mlp_forward_planned(int B, int m_known, int M, int nsample, int C, const float *known_feats,
                                                  const float *known_feats_pm, const int32_t *idx3, const float *weight3,
                                                  const float *rel_xyz, int num_layers, const b200_mlp_layer *layers,
                                                  float *out, void *workspace, size_t workspace_bytes, const void *plan,
                                                  size_t plan_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(B >= 0 && m_known > 0 && M >= 0 && C > 0 && (C & 3) == 0, "interp_mlp_forward: bad sizes");
  B200_CHECK_ARG(idx3 && weight3 && out && layers && (known_feats || known_feats_pm), "interp_mlp_forward: null pointer");
  B200_CHECK_ARG(B <= 65535, "interp_mlp_forward: B=%d exceeds grid.y", B);
  if (B == 0 || M == 0) return 0;
  const int use_xyz = rel_xyz ? 1 : 0;
  const float *fpm = known_feats_pm;
  if (!fpm) {
    const size_t need = align256(sizeof(float) * (size_t)B * m_known * C);
    B200_CHECK_ARG(workspace && need <= workspace_bytes, "interp_mlp_forward: workspace too small");
    float *t = (float *)workspace;
    const int rc = b200pn2_transpose_cn(B, C, m_known, known_feats, t, 0, stream_);
    if (rc) return rc;
    fpm = t;
  }
  B200_CHECK_ARG((((uintptr_t)fpm) & 15) == 0, "interp_mlp_forward: point-major features must be 16-byte aligned");
  B200_CHECK_ARG(sa_tc_supported(C, nsample, use_xyz, num_layers, layers, fpm),
                 "interp_mlp_forward: layer widths / nsample not supported by the tensor-core kernel");
  TcCall c;
  c.mode = 1; c.B = B; c.N = m_known; c.M = M; c.C = C; c.ns = nsample; c.use_xyz = use_xyz;
  c.feat_pm = fpm; c.idx3 = idx3; c.w3 = weight3; c.rel3 = rel_xyz; c.out = out;
  c.num_layers = num_layers; c.layers = layers; c.plan = plan; c.plan_bytes = plan_bytes;
  return sa_tc_run(c, stream);
}

extern "C" int b200pn2_interp_mlp_forward(int B, int m_known, int M, int nsample, int C, const float *known_feats,
                                          const float *known_feats_pm, const int32_t *idx3, const float *weight3,
                                          const float *rel_xyz, int num_layers, const b200_mlp_layer *layers,
                                          float *out, void *workspace, size_t workspace_bytes, b200_stream_t stream_) {
  return b200pn2_interp_mlp_forward_planned(B, m_known, M, nsample, C, known_feats, known_feats_pm, idx3, weight3, rel_xyz,
                                            num_layers, layers, out, workspace, workspace_bytes, nullptr, 0, stream_);
}

// Feature propagation rows (pointnet2_modules.py:399-420): row q of scene b = [sum_t weight3[q,t] * known[idx3[q,t]] (C2) |
// skip[q] (C1)] -> SharedMLP stack -> one output row per q (ReLU after every layer; `relu_last` for the final one).
extern "C" int b200pn2_fp_rows_forward(int B, int n, int m_known, int C2, int C1, const float *known_feats_pm,
                                       const float *skip_feats_pm, const int32_t *idx3, const float *weight3,
                                       int num_layers, const b200_mlp_layer *layers, int relu_last, float *out,
                                       float *out_pm, const void *plan, size_t plan_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(B >= 0 && n >= 0 && m_known > 0 && C2 > 0 && (C2 & 3) == 0 && C1 >= 0 && (C1 & 3) == 0,
                 "fp_rows_forward: bad sizes (C2 and C1 must be multiples of 4)");
  B200_CHECK_ARG(known_feats_pm && idx3 && weight3 && layers && (out || out_pm) && (C1 == 0 || skip_feats_pm),
                 "fp_rows_forward: null pointer");
  B200_CHECK_ARG(((((uintptr_t)known_feats_pm) | ((uintptr_t)skip_feats_pm)) & 15) == 0,
                 "fp_rows_forward: point-major features must be 16-byte aligned");
  B200_CHECK_ARG(num_layers >= 1 && layers[0].cin == C2 + C1, "fp_rows_forward: layer 0 expects cin=%d", C2 + C1);
  if (B == 0 || n == 0) return 0;
  TcCall c;
  c.mode = 1; c.B = B; c.N = m_known; c.M = n; c.C = C2; c.ns = 1; c.use_xyz = 0;
  c.feat_pm = known_feats_pm; c.idx3 = idx3; c.w3 = weight3; c.feat2_pm = skip_feats_pm; c.C2 = C1;
  c.rowout = 1; c.final_relu = relu_last ? 1 : 0; c.rows_total = B * n; c.rows_per_scene = n;
  c.out = out; c.out_pm = out_pm; c.num_layers = num_layers; c.layers = layers; c.plan = plan; c.plan_bytes = plan_bytes;
  return sa_tc_run(c, stream);
}

// Row MLP: S scenes x R rows of C channels (point-major) -> SharedMLP stack -> rows; the 1x1-conv blocks outside the SA
// layers (FP layers 2.., voting_module.py:38-65, proposal_module.py:98-123, grid_conv_module.py:108-115).
extern "C" int b200pn2_row_mlp_forward(int S, int R, int C, int ld, const float *x_pm, int num_layers,
                                       const b200_mlp_layer *layers, int relu_last, float *out, float *out_pm,
                                       const void *plan, size_t plan_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(S >= 0 && R >= 0 && C > 0 && (ld == 0 || ld >= C), "row_mlp_forward: bad sizes");
  B200_CHECK_ARG(x_pm && layers && (out || out_pm), "row_mlp_forward: null pointer");
  B200_CHECK_ARG((long long)S * R < (1ll << 30), "row_mlp_forward: too many rows");
  B200_CHECK_ARG(num_layers >= 1 && layers[0].cin == C, "row_mlp_forward: layer 0 expects cin=%d", C);
  if (S == 0 || R == 0) return 0;
  TcCall c;
  c.mode = 2; c.B = 1; c.N = S * R; c.M = S * R; c.C = C; c.ns = 32; c.use_xyz = 0;
  c.feat_pm = x_pm; c.ld = ld;
  c.rowout = 1; c.final_relu = relu_last ? 1 : 0; c.rows_total = S * R; c.rows_per_scene = R;
  c.out = out; c.out_pm = out_pm; c.num_layers = num_layers; c.layers = layers; c.plan = plan; c.plan_bytes = plan_bytes;
  return sa_tc_run(c, stream);
}

// The same stack on CHANNEL-MAJOR rows: x_cm (S, C, R) is what torch's conv stacks hold ((B, C, H, W) viewed (B, C, H*W));
// the producers read it in place (one coalesced request per channel and warp), so no transpose pass precedes the GEMMs.
extern "C" int b200pn2_row_mlp_forward_cm(int S, int R, int C, const float *x_cm, int num_layers,
                                          const b200_mlp_layer *layers, int relu_last, float *out, float *out_pm,
                                          const void *plan, size_t plan_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(S >= 0 && R >= 0 && C > 0, "row_mlp_forward_cm: bad sizes");
  B200_CHECK_ARG(x_cm && layers && (out || out_pm), "row_mlp_forward_cm: null pointer");
  B200_CHECK_ARG((long long)S * R < (1ll << 30), "row_mlp_forward_cm: too many rows");
  B200_CHECK_ARG(num_layers >= 1 && layers[0].cin == C, "row_mlp_forward_cm: layer 0 expects cin=%d", C);
  if (S == 0 || R == 0) return 0;
  TcCall c;
  c.mode = 2; c.B = 1; c.N = S * R; c.M = S * R; c.C = C; c.ns = 32; c.use_xyz = 0;
  c.feat_pm = x_cm; c.ld = C; c.cm_in = 1;
  c.rowout = 1; c.final_relu = relu_last ? 1 : 0; c.rows_total = S * R; c.rows_per_scene = R;
  c.out = out; c.out_pm = out_pm; c.num_layers = num_layers; c.layers = layers; c.plan = plan; c.plan_bytes = plan_bytes;
  return sa_tc_run(c, stream);
}

extern "C" size_t b200pn2_sa_forward_workspace(int B, int N, int M, int C, int nsample, int have_features_pm,
                                               int have_idx) {
  size_t bytes = 0;
  if (!have_idx) bytes += align256(sizeof(int32_t) * (size_t)B * M * nsample);
  if (!have_features_pm && C > 1) bytes += align256(sizeof(float) * (size_t)B * N * C);
  return bytes;
}

extern "C" int b200pn2_sa_forward(int B, int N, int M, int C, float radius, int nsample, int use_xyz,
                                  int normalize_xyz, const float *xyz, const float *features,
                                  const float *features_pm, const float *new_xyz, const int32_t *idx_in,
                                  int num_layers, const b200_mlp_layer *layers, float *out, float *out_pm,
                                  int32_t *idx_out, void *workspace, size_t workspace_bytes, b200_stream_t stream_) {
  return b200pn2_sa_forward_planned(B, N, M, C, radius, nsample, use_xyz, normalize_xyz, xyz, features, features_pm,
                                    new_xyz, idx_in, num_layers, layers, out, out_pm, idx_out, workspace, workspace_bytes,
                                    nullptr, 0, stream_);
}

extern "C" int b200pn2_sa_forward_planned(int B, int N, int M, int C, float radius, int nsample, int use_xyz,
                                          int normalize_xyz, const float *xyz, const float *features,
                                          const float *features_pm, const float *new_xyz, const int32_t *idx_in,
                                          int num_layers, const b200_mlp_layer *layers, float *out, float *out_pm,
                                          int32_t *idx_out, void *workspace, size_t workspace_bytes, const void *plan,
                                          size_t plan_bytes, b200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK_ARG(B >= 0 && N > 0 && M >= 0 && C >= 0, "sa_forward: bad sizes B=%d N=%d M=%d C=%d", B, N, M, C);
  B200_CHECK_ARG(num_layers >= 1 && num_layers <= SA_MAXL, "sa_forward: 1..%d layers supported, got %d", SA_MAXL,
                 num_layers);
  B200_CHECK_ARG(nsample >= 1 && nsample <= SA_R, "sa_forward: nsample=%d unsupported (1..%d)", nsample, SA_R);
  B200_CHECK_ARG(B <= 65535, "sa_forward: B=%d exceeds grid.y", B);
  B200_CHECK_ARG(xyz && new_xyz && out && layers, "sa_forward: null pointer");
  B200_CHECK_ARG(C == 0 || features || features_pm, "sa_forward: C=%d but no features", C);
  B200_CHECK_ARG(use_xyz || C > 0, "sa_forward: neither xyz nor features as input");
  if (B == 0 || M == 0) return 0;
  const int cin0 = (use_xyz ? 3 : 0) + C;
  B200_CHECK_ARG(layers[0].cin == cin0, "sa_forward: layer 0 expects cin=%d, got %d", cin0, layers[0].cin);
  for (int l = 0; l < num_layers; ++l) {
    B200_CHECK_ARG(layers[l].cin > 0 && layers[l].cout > 0 && layers[l].weight && layers[l].scale && layers[l].shift,
                   "sa_forward: layer %d malformed", l);
    if (l > 0) B200_CHECK_ARG(layers[l].cin == layers[l - 1].cout, "sa_forward: layer %d cin != previous cout", l);
  }

  // ---- workspace carving -----------------------------------------------------------------------------
  char *ws = (char *)workspace;
  size_t off = 0;
  // point-major features first: whether the tensor-core kernel (and its compacted tiles) will run decides what the ball
  // query has to produce
  const float *fpm = features_pm;
  if (C == 1 && !fpm) fpm = features;  // (B,1,N) and (B,N,1) are the same memory
  // fused ball query (north-star shape of the op): for the small levels the tensor-core kernel's producers stage the
  // scene in shared memory and run the radius search themselves -- no query launch, no idx round trip
  const bool tc_maybe = sa_tc_supported(C, nsample, use_xyz, num_layers, layers, fpm ? fpm : (C > 1 ? features : nullptr));
  const bool fuse_query = !idx_in && tc_maybe && !sa_tcp_units_wanted(0, 0, nsample, B, M) &&
                          sa_tcp_query_fusable(B, N, M, nsample, xyz);
  int32_t *idx_buf = nullptr;
  if (!idx_in && !fuse_query) {
    idx_buf = idx_out;
    if (!idx_buf) {
      const size_t need = align256(sizeof(int32_t) * (size_t)B * M * nsample);
      B200_CHECK_ARG(ws && off + need <= workspace_bytes, "sa_forward: workspace too small (idx)");
      idx_buf = (int32_t *)(ws + off);
      off += need;
    }
  }
  if (C > 1 && !fpm) {
    const size_t need = align256(sizeof(float) * (size_t)B * N * C);
    B200_CHECK_ARG(ws && off + need <= workspace_bytes, "sa_forward: workspace too small (features_pm)");
    float *t = (float *)(ws + off);
    off += need;
    const int rc = b200pn2_transpose_cn(B, C, N, features, t, 0, stream_);
    if (rc) return rc;
    fpm = t;
  }
  const bool use_tc = sa_tc_supported(C, nsample, use_xyz, num_layers, layers, fpm);
  // compacted tiles (sa_tcp.cu): the unit list is appended to by the ball query itself (or by one pass over a given idx)
  ScratchGuard unit_scratch;
  int *unit_total = nullptr, *unit_list = nullptr;
  if (use_tc && sa_tcp_units_wanted(0, 0, nsample, B, M)) {
    B200_CUDA_OK(unit_scratch.alloc(256 + sa_tcp_unit_list_bytes(B, M, nsample), stream));
    unit_total = (int *)unit_scratch.ptr;
    unit_list = (int *)((char *)unit_scratch.ptr + 256);
    B200_CUDA_OK(cudaMemsetAsync(unit_total, 0, sizeof(int), stream));
  }
  const int32_t *idx = idx_in;
  if (!idx && !fuse_query) {
    const int rc = ball_query_launch(B, N, M, radius, nsample, new_xyz, xyz, idx_buf, stream, unit_list, unit_total);
    if (rc) return rc;
    idx = idx_buf;
  } else if (idx) {
    if (idx_out && idx_out != idx_in)
      B200_CUDA_OK(cudaMemcpyAsync(idx_out, idx_in, sizeof(int32_t) * (size_t)B * M * nsample, cudaMemcpyDeviceToDevice,
                                   stream));
    if (unit_list) {
      const int rc = sa_tcp_units_from_idx(B, M, nsample, idx, unit_list, unit_total, stream);
      if (rc) return rc;
    }
  }
  const bool vec_ok = fpm && (C & 3) == 0 && ((((uintptr_t)fpm) & 15) == 0);
  if (use_tc) {
    TcCall c;
    c.mode = 0; c.B = B; c.N = N; c.M = M; c.C = C; c.ns = nsample; c.use_xyz = use_xyz; c.normalize_xyz = normalize_xyz;
    c.radius = radius; c.xyz = xyz; c.feat_pm = fpm; c.new_xyz = new_xyz; c.idx = idx; c.out = out; c.out_pm = out_pm;
    c.num_layers = num_layers; c.layers = layers; c.plan = plan; c.plan_bytes = plan_bytes;
    c.unit_list = unit_list; c.unit_total = unit_total;
    c.query = fuse_query ? 1 : 0; c.idx_out = fuse_query ? idx_out : nullptr;
    return sa_tc_run(c, stream);
  }

  SaParams p;
  p.B = B; p.N = N; p.M = M; p.C = C; p.ns = nsample; p.use_xyz = use_xyz ? 1 : 0; p.nl = num_layers;
  p.G = SA_R / nsample;
  if (p.G > SA_MAXG) p.G = SA_MAXG;
  if (p.G < 1) p.G = 1;
  p.inv_r = normalize_xyz ? (float)(1.0 / (double)radius) : 1.0f;
  p.xyz = xyz; p.feat_cm = features; p.feat_pm = fpm; p.new_xyz = new_xyz; p.idx = idx;
  p.out = out; p.out_pm = out_pm; p.vec_gather = vec_ok ? 1 : 0;
  int wA = 0, wB = 0;  // widest input held by each activation buffer
  for (int l = 0; l < num_layers; ++l) {
    p.L[l].w = layers[l].weight; p.L[l].scale = layers[l].scale; p.L[l].shift = layers[l].shift;
    p.L[l].cin = layers[l].cin; p.L[l].cout = layers[l].cout;
    const int w = (layers[l].cin + 7) & ~7;
    if (l & 1) wB = w > wB ? w : wB; else wA = w > wA ? w : wA;
  }
  p.ldA = wA + 4;
  p.ldB = (wB > 0 ? wB : 0) + 4;
  const int cout_last = layers[num_layers - 1].cout;
  const size_t smem = sizeof(float) * ((size_t)SA_R * p.ldA + (size_t)SA_R * p.ldB + 2 * SA_KC * SA_WLD) +
                      sizeof(int) * ((size_t)SA_MAXG * cout_last + SA_R) + sizeof(float) * (SA_MAXG * 3 + 4);
  B200_CHECK_ARG(smem <= 227 * 1024, "sa_forward: channel widths need %zu B of shared memory (> 227 KB)", smem);
  static DynSmemOptIn optin;
  B200_CUDA_OK(optin.ensure(sa_mlp_max_kernel, smem));
  dim3 grid(ceil_div(M, p.G), B);
  sa_mlp_max_kernel<<<grid, SA_THREADS, smem, stream>>>(p);
  B200_LAUNCH_OK("sa_mlp_max_kernel");
  return 0;
}
