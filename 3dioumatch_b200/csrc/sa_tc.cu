// sa_tc.cu -- fused set-abstraction forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as sa_mlp_max_kernel (sa_fused.cu): gather grouped rows -> SharedMLP (1x1 conv + folded BN + ReLU) x L
// -> max over nsample, grouped tensors never in HBM.  Used when the channel widths make the MLP a real dense
// contraction (VoteNet SA2/SA3/SA4/vote aggregation: cin 131..259 -> 128 -> 128 -> 128|256, backbone_module.py:44-69,
// proposal_module.py:72-79).
//
// Precision: the reference is fp32 (parity bar 1e-5), plain TF32 is ~1e-3.  Every GEMM therefore runs as THREE
// kind::tf32 MMAs on operands split as x = hi + lo (hi = x rounded to TF32, lo = x - hi, exact in fp32):
//     D_big += A_hi*W_hi ;  D_small += A_lo*W_hi + A_hi*W_lo ;  D = D_big + D_small   (fp32 accumulation in TMEM)
// The 2^-11-sized correction terms get their own TMEM accumulator (columns 256..511) so that their accumulation
// rounding is negligible; measured error is ~3x an fp32 FFMA GEMM (scripts/tc_precision.py); the lo*lo term (~2^-22)
// is dropped.
// Weights are split/packed once per call (tc_pack_weights_kernel) into the exact shared-memory image of a pipeline
// stage, so a stage is ONE 32 KB cp.async.bulk (TMA engine) with mbarrier completion.  Activations are split by the
// CUDA cores on their way into shared memory (gather for layer 1, TMEM epilogue for layers 2+).
//
// CTA = 128 grouped rows (UMMA M = 128), 192 threads:
//   warps 0-3  workers : layer-1 gather producer (LDG.128 -> split -> swizzled STS), TMEM epilogues
//   warp  4    MMA     : one thread issues tcgen05.mma / tcgen05.commit; also owns the TMEM allocation
//   warp  5    loader  : one thread streams packed weight stages with cp.async.bulk
// Shared memory: R1 = 128 KB (layer-1 A stages, later X_hi | X_lo of the hidden activations), R2 = 2 x 32 KB weight
// stages.  Operands are K-major with the 128-byte swizzle (tc_common.cuh).
#include <stdlib.h>

#include "../../include/b200_pointnet2.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace b200 {

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_MAXL = 4;
constexpr uint32_t TC_KB_BYTES = 128 * 128;          // one operand k-block: 128 rows x 128 B
constexpr uint32_t TC_WSTAGE_BYTES = 2 * TC_KB_BYTES;  // W_hi | W_lo

struct TcLayer {
  const float *scale, *shift;
  int cin, cout, nkb, nhalf;
  size_t packed_off;  // byte offset of this layer's stages in the packed weight buffer
};

struct TcParams {
  int B, N, M, C, ns, G, use_xyz, nl;
  float inv_r;
  const float *xyz, *feat_pm, *new_xyz;
  const int32_t *idx;
  float *out, *out_pm;
  const uint8_t *packed;
  TcLayer L[TC_MAXL];
};

// ---- weight packing: (cout, cin) fp32 -> [half][kb][hi|lo][128 rows x 128 B, 128-byte swizzle] -----------------------
struct PackParams {
  const float *w[TC_MAXL];
  int cin[TC_MAXL], cout[TC_MAXL], nkb[TC_MAXL], nhalf[TC_MAXL];
  size_t off[TC_MAXL];
  int nl, perm_c;  // perm_c >= 0: layer 0 column k reads source channel (k < perm_c ? 3 + k : k - perm_c)
};

__global__ void __launch_bounds__(256) tc_pack_weights_kernel(PackParams p, uint8_t *__restrict__ packed) {
  const int l = blockIdx.y;
  if (l >= p.nl) return;
  const int items = p.nhalf[l] * p.nkb[l] * 128 * 8;  // (half, kb, row, chunk)
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int chunk = it & 7, row = (it >> 3) & 127, rest = it >> 10;
    const int kb = rest % p.nkb[l], half = rest / p.nkb[l];
    const int n = half * 128 + row;
    float v[4], hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kb * 32 + chunk * 4 + e;
      int src = k;
      if (l == 0 && p.perm_c >= 0) src = k < p.perm_c ? 3 + k : k - p.perm_c;
      v[e] = (n < p.cout[l] && k < p.cin[l]) ? p.w[l][(size_t)n * p.cin[l] + src] : 0.f;
      tc::split_tf32(v[e], hi[e], lo[e]);
    }
    uint8_t *stage = packed + p.off[l] + (size_t)(half * p.nkb[l] + kb) * TC_WSTAGE_BYTES;
    const uint32_t off = tc::sw128_offset(row, chunk);
    *reinterpret_cast<float4 *>(stage + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4 *>(stage + TC_KB_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- the fused kernel -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) sa_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *R1 = base;                       // 128 KB
  uint8_t *R2 = base + 8 * TC_KB_BYTES;     // 2 x 32 KB
  float *s_scale = reinterpret_cast<float *>(R2 + 2 * TC_WSTAGE_BYTES);  // [TC_MAXL][256]
  float *s_shift = s_scale + TC_MAXL * 256;

  __shared__ uint64_t full_a[2], empty_a[2], full_w[2], empty_w[2], accum_full, x_ready;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int m0 = blockIdx.x * p.G;
  const int ns = p.ns;
  const int g_here = min(p.G, p.M - m0);
  const int nl = p.nl;

  if (warp == 4) tc::tmem_alloc<512>(&tmem_base_s);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&full_a[s], 128);
      tc::mbar_init(&empty_a[s], 1);
      tc::mbar_init(&full_w[s], 1);
      tc::mbar_init(&empty_w[s], 1);
    }
    tc::mbar_init(&accum_full, 1);
    tc::mbar_init(&x_ready, 128);
    tc::mbar_fence_init();
  }
  for (int e = tid; e < nl * 256; e += TC_THREADS) {
    const int l = e >> 8, c = e & 255;
    s_scale[e] = c < p.L[l].cout ? p.L[l].scale[c] : 0.f;
    s_shift[e] = c < p.L[l].cout ? p.L[l].shift[c] : 0.f;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;

  const int nkb1 = p.L[0].nkb;

  if (warp == 5) {
    // ================= weight loader =================
    if (lane == 0) {
      int i = 0;
      for (int l = 0; l < nl; ++l) {
        const int nst = p.L[l].nhalf * p.L[l].nkb;
        const uint8_t *src = p.packed + p.L[l].packed_off;
        for (int s = 0; s < nst; ++s, ++i) {
          const int st = i & 1;
          tc::mbar_wait(&empty_w[st], (uint32_t)(((i >> 1) & 1) ^ 1));
          tc::mbar_arrive_expect_tx(&full_w[st], TC_WSTAGE_BYTES);
          tc::bulk_g2s(R2 + st * TC_WSTAGE_BYTES, src + (size_t)s * TC_WSTAGE_BYTES, TC_WSTAGE_BYTES, &full_w[st]);
        }
      }
    }
  } else if (warp == 4) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, 128);
      int i = 0;  // flat weight-stage counter (same order as the loader)
      int acc_use = 0, xr_use = 0;
      for (int l = 0; l < nl; ++l) {
        if (l > 0) {
          tc::mbar_wait(&x_ready, (uint32_t)(xr_use & 1));  // hidden activations of layer l-1 are in R1
          ++xr_use;
          tc::tc_fence_after_sync();
        }
        const int nkb = p.L[l].nkb;
        for (int h = 0; h < p.L[l].nhalf; ++h) {
          const uint32_t d_addr = tmem_d + (uint32_t)(h * 128);
          for (int kb = 0; kb < nkb; ++kb, ++i) {
            const int ws = i & 1;
            uint32_t a_hi, a_lo;
            if (l == 0) {
              const int as = kb & 1;
              tc::mbar_wait(&full_a[as], (uint32_t)((kb >> 1) & 1));
              a_hi = tc::smem_addr(R1 + as * 2 * TC_KB_BYTES);
              a_lo = a_hi + TC_KB_BYTES;
            } else {
              a_hi = tc::smem_addr(R1 + kb * TC_KB_BYTES);
              a_lo = a_hi + 4 * TC_KB_BYTES;
            }
            tc::mbar_wait(&full_w[ws], (uint32_t)((i >> 1) & 1));
            tc::tc_fence_after_sync();
            const uint32_t w_hi = tc::smem_addr(R2 + ws * TC_WSTAGE_BYTES), w_lo = w_hi + TC_KB_BYTES;
            const uint64_t da_hi = tc::make_desc_sw128(a_hi), da_lo = tc::make_desc_sw128(a_lo);
            const uint64_t dw_hi = tc::make_desc_sw128(w_hi), dw_lo = tc::make_desc_sw128(w_lo);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t adv = (uint64_t)(ks * 2);  // 8 floats = 32 B = 2 x 16 B
              tc::mma_tf32(d_addr, da_hi + adv, dw_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
              tc::mma_tf32(d_addr + 256u, da_lo + adv, dw_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
              tc::mma_tf32(d_addr + 256u, da_hi + adv, dw_lo + adv, idesc, 1u);
            }
            if (l == 0) tc::mma_commit(&empty_a[kb & 1]);
            tc::mma_commit(&empty_w[ws]);
          }
        }
        tc::mma_commit(&accum_full);
        ++acc_use;
      }
      (void)acc_use;
    }
  } else {
    // ================= workers: row `tid` of the tile =================
    const int row = tid;
    const int g = row / ns;
    const bool valid = g < g_here;
    int src_idx = -1;
    float ctr[3] = {0.f, 0.f, 0.f};
    if (valid) {
      src_idx = p.idx[((size_t)b * p.M + m0 + g) * ns + (row - g * ns)];
      const float *c = p.new_xyz + ((size_t)b * p.M + m0 + g) * 3;
      ctr[0] = c[0]; ctr[1] = c[1]; ctr[2] = c[2];
    }
    const int C = p.C;
    const float *frow = valid ? p.feat_pm + ((size_t)b * p.N + src_idx) * C : nullptr;
    float rel[3] = {0.f, 0.f, 0.f};
    if (valid && p.use_xyz) {
      const float *q = p.xyz + ((size_t)b * p.N + src_idx) * 3;
      // pointnet2_utils.py:351-353: grouped_xyz -= new_xyz ; /= radius  (x * fp32(1/r) on CUDA)
      rel[0] = __fmul_rn(__fsub_rn(q[0], ctr[0]), p.inv_r);
      rel[1] = __fmul_rn(__fsub_rn(q[1], ctr[1]), p.inv_r);
      rel[2] = __fmul_rn(__fsub_rn(q[2], ctr[2]), p.inv_r);
    }
    // ---- layer-1 A operand: [features (C) | rel xyz (3) | 0 ...], two stages of one k-block each ----
    for (int kb = 0; kb < nkb1; ++kb) {
      const int as = kb & 1;
      tc::mbar_wait(&empty_a[as], (uint32_t)(((kb >> 1) & 1) ^ 1));
      uint8_t *a_hi = R1 + as * 2 * TC_KB_BYTES, *a_lo = a_hi + TC_KB_BYTES;
      float4 v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ch = kb * 32 + c * 4;
        if (valid && ch + 3 < C) {
          v[c] = *reinterpret_cast<const float4 *>(frow + ch);
        } else {
          float t[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = ch + e;
            float x = 0.f;
            if (valid) {
              if (k < C) x = frow[k];
              else if (p.use_xyz && k < C + 3) x = rel[k - C];
            }
            t[e] = x;
          }
          v[c] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 h, l;
        tc::split_tf32(v[c].x, h.x, l.x); tc::split_tf32(v[c].y, h.y, l.y);
        tc::split_tf32(v[c].z, h.z, l.z); tc::split_tf32(v[c].w, h.w, l.w);
        const uint32_t off = tc::sw128_offset(row, c);
        *reinterpret_cast<float4 *>(a_hi + off) = h;
        *reinterpret_cast<float4 *>(a_lo + off) = l;
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full_a[as]);
    }
    // ---- epilogues ----
    for (int l = 0; l < nl; ++l) {
      tc::mbar_wait(&accum_full, (uint32_t)(l & 1));
      tc::tc_fence_after_sync();
      const float *sc = s_scale + l * 256, *sh = s_shift + l * 256;
      const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
      if (l + 1 < nl) {
        // hidden layer: X = relu(scale*acc+shift) -> split -> R1 as the next layer's K-major operand
        uint8_t *x_hi = R1, *x_lo = R1 + 4 * TC_KB_BYTES;
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32], r2[32];
          tc::tmem_ld_32x32(lane_addr + (uint32_t)c0, r);
          tc::tmem_ld_32x32(lane_addr + 256u + (uint32_t)c0, r2);
          tc::tmem_ld_wait();
          uint8_t *kb_hi = x_hi + (c0 >> 5) * TC_KB_BYTES, *kb_lo = x_lo + (c0 >> 5) * TC_KB_BYTES;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float y[4], h[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = c0 + c * 4 + e;
              const float acc = __uint_as_float(r[c * 4 + e]) + __uint_as_float(r2[c * 4 + e]);
              y[e] = fmaxf(fmaf(acc, sc[col], sh[col]), 0.f);
              tc::split_tf32(y[e], h[e], lo[e]);
            }
            const uint32_t off = tc::sw128_offset(row, c);
            *reinterpret_cast<float4 *>(kb_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(kb_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&x_ready);
      } else {
        // last layer: relu(scale*acc+shift), max over the nsample rows of each centre, write (B,cout,M) [+ (B,M,cout)]
        const int cout = p.L[l].cout;
        const unsigned gmask = ns >= 32 ? 0xffffffffu : (lane < 16 ? 0x0000ffffu : 0xffff0000u);
        const int gl = ns >= 32 ? warp : warp * 2 + (lane >> 4);  // centre (within the tile) of this lane's rows
        for (int c0 = 0; c0 < cout; c0 += 32) {
          uint32_t r[32], r2[32];
          tc::tmem_ld_32x32(lane_addr + (uint32_t)c0, r);
          tc::tmem_ld_32x32(lane_addr + 256u + (uint32_t)c0, r2);
          tc::tmem_ld_wait();
          float keep0 = 0.f, keep1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c0 + j;
            const float acc = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            const float y = fmaxf(fmaf(acc, sc[col], sh[col]), 0.f);  // >= 0: int order == float order
            const int mx = __reduce_max_sync(gmask, __float_as_int(y));
            if (ns >= 32) {
              if (lane == j) keep0 = __int_as_float(mx);
            } else {
              if ((lane & 15) == (j & 15)) {
                if (j < 16) keep0 = __int_as_float(mx); else keep1 = __int_as_float(mx);
              }
            }
          }
          if (gl < g_here) {
            const int m = m0 + gl;
            if (ns >= 32) {
              const int col = c0 + lane;
              if (col < cout) {
                p.out[((size_t)b * cout + col) * p.M + m] = keep0;
                if (p.out_pm) p.out_pm[((size_t)b * p.M + m) * cout + col] = keep0;
              }
            } else {
              const int cA = c0 + (lane & 15), cB = cA + 16;
              if (cA < cout) {
                p.out[((size_t)b * cout + cA) * p.M + m] = keep0;
                if (p.out_pm) p.out_pm[((size_t)b * p.M + m) * cout + cA] = keep0;
              }
              if (cB < cout) {
                p.out[((size_t)b * cout + cB) * p.M + m] = keep1;
                if (p.out_pm) p.out_pm[((size_t)b * p.M + m) * cout + cB] = keep1;
              }
            }
          }
        }
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<512>(tmem_d);
}

// Can the tensor-core kernel take this stage?  (hidden widths 128, last 128|256, nsample 16|32, aligned point-major features)
bool sa_tc_supported(int C, int nsample, int use_xyz, int num_layers, const b200_mlp_layer *layers, const float *feat_pm) {
  static int enabled = -1;
  if (enabled < 0) {
    const char *e = getenv("B200_SA_TC");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!enabled) return false;
  if (!(nsample == 16 || nsample == 32)) return false;
  if (num_layers < 2 || num_layers > TC_MAXL) return false;
  if (C < 32 || (C & 3) || !feat_pm || (((uintptr_t)feat_pm) & 15)) return false;
  if (layers[0].cin != C + (use_xyz ? 3 : 0)) return false;
  for (int l = 0; l < num_layers; ++l) {
    const bool last = l == num_layers - 1;
    if (last ? !(layers[l].cout == 128 || layers[l].cout == 256) : layers[l].cout != 128) return false;
    if (l > 0 && layers[l].cin != 128) return false;
  }
  return true;
}

int sa_tc_launch(int B, int N, int M, int C, float radius, int nsample, int use_xyz, int normalize_xyz, const float *xyz,
                 const float *feat_pm, const float *new_xyz, const int32_t *idx, int num_layers,
                 const b200_mlp_layer *layers, float *out, float *out_pm, cudaStream_t stream) {
  TcParams p;
  PackParams pk;
  p.B = B; p.N = N; p.M = M; p.C = C; p.ns = nsample; p.G = TC_ROWS / nsample; p.use_xyz = use_xyz ? 1 : 0;
  p.nl = num_layers;
  p.inv_r = normalize_xyz ? (float)(1.0 / (double)radius) : 1.0f;
  p.xyz = xyz; p.feat_pm = feat_pm; p.new_xyz = new_xyz; p.idx = idx; p.out = out; p.out_pm = out_pm;
  size_t off = 0;
  pk.nl = num_layers;
  pk.perm_c = use_xyz ? C : -1;
  for (int l = 0; l < num_layers; ++l) {
    TcLayer &t = p.L[l];
    t.scale = layers[l].scale; t.shift = layers[l].shift; t.cin = layers[l].cin; t.cout = layers[l].cout;
    t.nkb = (layers[l].cin + 31) / 32;
    t.nhalf = (layers[l].cout + 127) / 128;
    t.packed_off = off;
    pk.w[l] = layers[l].weight; pk.cin[l] = t.cin; pk.cout[l] = t.cout; pk.nkb[l] = t.nkb; pk.nhalf[l] = t.nhalf;
    pk.off[l] = off;
    off += (size_t)t.nhalf * t.nkb * TC_WSTAGE_BYTES;
  }
  uint8_t *packed = nullptr;
  B200_CUDA_OK(cudaMallocAsync((void **)&packed, off, stream));
  tc_pack_weights_kernel<<<dim3(32, num_layers), 256, 0, stream>>>(pk, packed);
  B200_LAUNCH_OK("tc_pack_weights_kernel");
  p.packed = packed;
  const size_t smem = 1024 + 8 * TC_KB_BYTES + 2 * TC_WSTAGE_BYTES + 2 * TC_MAXL * 256 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(sa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(M, p.G), B);
  sa_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
  B200_LAUNCH_OK("sa_tc_kernel");
  B200_CUDA_OK(cudaFreeAsync(packed, stream));
  return 0;
}

}  // namespace b200
