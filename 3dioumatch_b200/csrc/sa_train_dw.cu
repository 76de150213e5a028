// sa_train_dw.cu -- weight gradient of one training-mode layer on the tensor cores:  dW_l = dz_l^T a_{l-1}   (cout x cin,
// contraction over the R = B * npoint * nsample rows).
//
// Both operands are built on the fly from what the forward saved (sa_train.cu): dz_l = scale * g + b + c * z_l from the raw
// conv output and the upstream gradient, a_{l-1} = relu(scale' z_{l-1} + shift') (or the grouped input rows for the first
// layer).  The contraction index is the ROW, so the producers store both operands TRANSPOSED into the K-major, 128-byte
// swizzled tiles the MMA reads: a k-block is 32 rows, thread (row r, channel quarter q) builds 32 channels of dz and of a
// for its row and writes element (channel, r) -- for a fixed channel the 32 lanes of a warp write one 128-byte tile row.
// Split precision as everywhere else: D_big += A_hi B_hi, D_small += A_lo B_hi + A_hi B_lo (fp32 accumulation in TMEM).
// Grid = (row splits, cin tiles of 128, cout tiles of 128); each CTA walks its k-blocks with a fixed stride, keeps the
// 128 x 128 accumulators in tensor memory for the whole kernel and writes ONE partial matrix; a fixed-order reduction over
// the splits follows (reduce_partials_kernel) -- bit-reproducible.
#include "../../include/b200_pointnet2.h"
#include "common.cuh"
#include "tc_common.cuh"
#include "sa_train.cuh"

namespace b200 {

constexpr int DW_STAGES = 3;
constexpr int DW_THREADS = 160;  // warps 0-3: producers + epilogue; warp 4: MMA issuer
constexpr uint32_t DW_TILE_BYTES = 128 * 128;  // [128 channels][32 rows] fp32

__global__ void __launch_bounds__(DW_THREADS, 1) dw_tc_kernel(const DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle, by OFFSET from the __shared__ array: a round trip through uintptr_t
  // loses the address space and turns every access below into a generic LD.E / ST.E with a descriptor R2UR pair
  uint8_t *base = smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u);
  // stage s: A_hi | A_lo | B_hi | B_lo, 16 KB each
  __shared__ uint64_t full[DW_STAGES], empty[DW_STAGES], done;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_dz[4][128];   // scale, b, c, shift of the cout tile
  __shared__ float s_act[2][128];  // scale, shift of the cin tile (layer below)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * 128, m0 = blockIdx.z * 128;
  const int nt = min(128, ((p.cin - n0) + 15) & ~15);  // MMA N of this tile (multiple of 16)

  if (warp == 4) tc::tmem_alloc<256>(&tmem_base_s);
  if (tid == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      tc::mbar_init(&full[s], 128);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(&done, 1);
    tc::mbar_fence_init();
  }
  for (int e = tid; e < 128; e += DW_THREADS) {
    const int c = m0 + e, n = n0 + e;
    const bool ci = c < p.cout, ni = n < p.cin;
    s_dz[0][e] = ci ? p.dz.scale[c] : 0.f;
    s_dz[1][e] = ci ? p.coef_b[c] : 0.f;
    s_dz[2][e] = ci ? p.coef_c[c] : 0.f;
    s_dz[3][e] = ci ? p.dz.shift[c] : 0.f;
    s_act[0][e] = (ni && p.act.scale) ? p.act.scale[n] : 1.f;
    s_act[1][e] = (ni && p.act.shift) ? p.act.shift[n] : 0.f;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_s;
  const long long nkb = (p.R + 31) / 32;

  if (warp == 4) {
    // ---- MMA issuer: converged warp, one elected lane issues ----
    const uint32_t idesc = tc::make_idesc_tf32(128, nt);
    const uint32_t sbase = tc::smem_addr(base);
    uint32_t st = 0, ph = 0;
    bool first = true;
    for (long long kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
      mbar_wait_wd(&full[st], ph);
      tc::tc_fence_after_sync();
      const uint32_t a_hi = sbase + st * 4u * DW_TILE_BYTES, a_lo = a_hi + DW_TILE_BYTES, b_hi = a_lo + DW_TILE_BYTES,
                     b_lo = b_hi + DW_TILE_BYTES;
      const uint64_t dah = tc::make_desc_sw128(a_hi), dal = tc::make_desc_sw128(a_lo), dbh = tc::make_desc_sw128(b_hi),
                     dbl = tc::make_desc_sw128(b_lo);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);  // 8 rows = 32 B
          const uint32_t acc = (first && ks == 0) ? 0u : 1u;
          tc::mma_tf32(tmem_d, dah + adv, dbh + adv, idesc, acc);
          tc::mma_tf32(tmem_d + 128u, dal + adv, dbh + adv, idesc, acc);
          tc::mma_tf32(tmem_d + 128u, dah + adv, dbl + adv, idesc, 1u);
        }
        tc::mma_commit(&empty[st]);
      }
      __syncwarp();
      first = false;
      if (++st == DW_STAGES) { st = 0; ph ^= 1u; }
    }
    if (elect_one()) tc::mma_commit(&done);
    __syncwarp();
  } else {
    // ---- producers: thread = (row r of the k-block, channel quarter q) ----
    const int r = tid & 31, q = tid >> 5;
    uint32_t st = 0, ph = 0;
    for (long long kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
      const long long row = kb * 32 + r;
      const bool valid = row < p.R;
      float dzv[32], av[32];
      // dz: channels m0 + q*32 .. +31
      {
        const int cb = m0 + q * 32;
        long long grp = 0;
        int slot = 0;
        if (!p.dz.g) { grp = row / p.dz.ns; slot = (int)(row - grp * p.dz.ns); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = cb + j * 4;
          const bool in = valid && c + 3 < p.cout;
          const float4 z = in ? __ldg(reinterpret_cast<const float4 *>(p.dz.z + row * p.dz.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 g;
          if (p.dz.g) {
            g = in ? __ldg(reinterpret_cast<const float4 *>(p.dz.g + row * p.dz.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            g = in ? __ldg(reinterpret_cast<const float4 *>(p.dz.gout_pm + grp * p.dz.C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int4 a = in ? __ldg(reinterpret_cast<const int4 *>(p.dz.arg_pm + grp * p.dz.C + c)) : make_int4(-1, -1, -1, -1);
            const int e = q * 32 + j * 4;
            g.x = (a.x == slot && fmaf(z.x, s_dz[0][e + 0], s_dz[3][e + 0]) > 0.f) ? g.x : 0.f;
            g.y = (a.y == slot && fmaf(z.y, s_dz[0][e + 1], s_dz[3][e + 1]) > 0.f) ? g.y : 0.f;
            g.z = (a.z == slot && fmaf(z.z, s_dz[0][e + 2], s_dz[3][e + 2]) > 0.f) ? g.z : 0.f;
            g.w = (a.w == slot && fmaf(z.w, s_dz[0][e + 3], s_dz[3][e + 3]) > 0.f) ? g.w : 0.f;
          }
          const int e = q * 32 + j * 4;
          dzv[j * 4 + 0] = in ? fmaf(s_dz[2][e + 0], z.x, fmaf(s_dz[0][e + 0], g.x, s_dz[1][e + 0])) : 0.f;
          dzv[j * 4 + 1] = in ? fmaf(s_dz[2][e + 1], z.y, fmaf(s_dz[0][e + 1], g.y, s_dz[1][e + 1])) : 0.f;
          dzv[j * 4 + 2] = in ? fmaf(s_dz[2][e + 2], z.z, fmaf(s_dz[0][e + 2], g.z, s_dz[1][e + 2])) : 0.f;
          dzv[j * 4 + 3] = in ? fmaf(s_dz[2][e + 3], z.w, fmaf(s_dz[0][e + 3], g.w, s_dz[1][e + 3])) : 0.f;
        }
      }
      // a_{l-1}: channels n0 + q*32 .. +31
      {
        const int nb = n0 + q * 32;
        const float *arow = p.act.rows + row * p.act.ld;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = nb + j * 4;
          float x[4] = {0.f, 0.f, 0.f, 0.f};
          if (valid && n + 3 < p.cin && p.vec_act) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(arow + n));
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
          } else if (valid) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < p.cin) x[e] = arow[n + e];
          }
          if (p.act.scale) {
            const int e0 = q * 32 + j * 4;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              x[e] = (valid && n + e < p.cin) ? fmaxf(fmaf(x[e], s_act[0][e0 + e], s_act[1][e0 + e]), 0.f) : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) av[j * 4 + e] = x[e];
        }
      }
      mbar_wait_wd(&empty[st], ph ^ 1u);
      uint8_t *a_hi = base + (size_t)st * 4 * DW_TILE_BYTES, *a_lo = a_hi + DW_TILE_BYTES, *b_hi = a_lo + DW_TILE_BYTES,
              *b_lo = b_hi + DW_TILE_BYTES;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int ch = q * 32 + j;  // tile row = channel; the 32 rows of the k-block are its 32 floats
        const uint32_t off = tc::sw128_offset(ch, r >> 2) + (uint32_t)((r & 3) << 2);
        float h, l;
        tc::split_tf32(dzv[j], h, l);
        *reinterpret_cast<float *>(a_hi + off) = h;
        *reinterpret_cast<float *>(a_lo + off) = l;
        tc::split_tf32(av[j], h, l);
        *reinterpret_cast<float *>(b_hi + off) = h;
        *reinterpret_cast<float *>(b_lo + off) = l;
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full[st]);
      if (++st == DW_STAGES) { st = 0; ph ^= 1u; }
    }
    // ---- epilogue: TMEM lane = cout row m0 + 32 * warp + lane; D = products + corrections ----
    mbar_wait_wd(&done, 0u);
    tc::tc_fence_after_sync();
    const int c = m0 + warp * 32 + lane;
    const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
    float *dst = p.partial + ((size_t)blockIdx.x * p.cout + c) * p.cin + n0;
    const bool any = blockIdx.x < nkb;  // a CTA without k-blocks never issued an MMA: its partial is zero
    for (int c0 = 0; c0 < nt; c0 += 32) {
      uint32_t a[32], b[32];
      if (any) {
        tc::tmem_ld_32x32(lane_base + (uint32_t)c0, a);
        tc::tmem_ld_32x32(lane_base + 128u + (uint32_t)c0, b);
        tc::tmem_ld_wait();
      }
      if (c < p.cout) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n0 + c0 + i < p.cin) dst[c0 + i] = any ? __uint_as_float(a[i]) + __uint_as_float(b[i]) : 0.f;
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<256>(tmem_d);
}

int dw_tc_launch(DwParams &p, int splits, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)DW_STAGES * 4 * DW_TILE_BYTES;
  static DynSmemOptIn optin;
  B200_CUDA_OK(optin.ensure(dw_tc_kernel, smem));
  dim3 grid(splits, (p.cin + 127) / 128, (p.cout + 127) / 128);
  dw_tc_kernel<<<grid, DW_THREADS, smem, stream>>>(p);
  B200_LAUNCH_OK("dw_tc_kernel");
  return 0;
}

}  // namespace b200
