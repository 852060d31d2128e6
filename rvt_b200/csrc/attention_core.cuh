// Attention core for one (128-row tile, head): S = Q K^T (tcgen05), block-diagonal masked
// softmax over each partition group's P keys (fp32, exp2), O = P V (tcgen05).
// Restates the middle of SelfAttentionCl.forward (models/layers/maxvit/maxvit.py:349-352).
//
// qkv scratch: fp16 [n_tiles*128, 3C], rows in partition order (RowMap), columns per head
// interleaved [q_h | k_h | v_h] exactly like the reference's qkv Linear output (:347).
// out scratch: fp16 [n_tiles*128, C], heads concatenated head-major (:352).
//
// A tile holds 128/rows_per_win partition groups; keys outside a row's own group, and the
// padding rows (p >= P), are masked to probability 0.
#pragma once
#include "umma.cuh"

namespace rvt {

struct AttnArgs {
  const __half* qkv;   // [rows, 3C]
  __half* out;         // [rows, C]
  int C, dh, nh;
  int P, rows_per_win; // keys per group, rows a group occupies (64 or 128)
  int nkeys;           // MMA N for S / K for PV: 128 if two groups per tile else round_up(P,16)
  float scale_log2e;   // dh^-0.5 * log2(e)
  int ab_fmt;
};

constexpr uint32_t kAttnSmemBytes = 1024 + 16384 /*Q*/ + 16384 /*K*/ + 16384 /*Vt*/ + 32768 /*P*/ + 64;

__global__ void __launch_bounds__(128) attention_core_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const uint32_t sQ = base, sK = base + 16384, sVt = base + 32768, sP = base + 49152;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 81920);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int t = threadIdx.x, warp = t >> 5;
  const int mt = blockIdx.x, h = blockIdx.y;
  const int dh = a.dh, dhp = (dh + 15) & ~15;
  const int nchunks = dhp >> 3;

  if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 128);

  // zero V^T (rows d in [dh, dhp) and unused key columns must be exact zeros)
  for (int i = t; i < 16384 / 16; i += 128) st_smem_16B(sVt + i * 16, 0u, 0u, 0u, 0u);
  pdl_trigger();
  __syncthreads();
  pdl_wait();            // qkv is the previous kernel's output

  // ---- stage Q, K (row t) and V^T (column t) ----
  const __half* rowp = a.qkv + (static_cast<size_t>(mt) * 128 + t) * (3 * a.C) + h * 3 * dh;
  for (int c = 0; c < nchunks; ++c) {
    uint4 q = make_uint4(0, 0, 0, 0), k = make_uint4(0, 0, 0, 0);
    if (c * 8 < dh) {
      q = __ldg(reinterpret_cast<const uint4*>(rowp + c * 8));
      k = __ldg(reinterpret_cast<const uint4*>(rowp + dh + c * 8));
    }
    st_smem_16B(sQ + sw128_offset(t, c), q.x, q.y, q.z, q.w);
    st_smem_16B(sK + sw128_offset(t, c), k.x, k.y, k.z, k.w);
  }
  if (t < a.nkeys) {
    const uint32_t atom = t >> 6, kk = t & 63;
    for (int c = 0; c * 8 < dh; ++c) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(rowp + 2 * dh + c * 8));
      const __half* hv = reinterpret_cast<const __half*>(&v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t d = c * 8 + e;
        const uint32_t addr = sVt + atom * (dhp * 128) + sw128_offset(d, kk >> 3) + (kk & 7) * 2;
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(hv[e])) : "memory");
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- S = Q K^T ----
  if (t == 0) {
    const uint32_t idesc = umma_idesc_f16(128, a.nkeys, a.ab_fmt);
    for (int k = 0; k < dhp / 16; ++k)
      umma_f16(tmem, umma_desc_sw128(sQ + k * 32), umma_desc_sw128(sK + k * 32), idesc, k != 0);
    umma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();

  // ---- masked softmax on row t ----
  const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int key_lo = (t / a.rows_per_win) * a.rows_per_win;
  const int key_hi = key_lo + a.P;
  float mx = -INFINITY;
  for (int c0 = 0; c0 < a.nkeys; c0 += 16) {
    float v[16];
    tmem_ld_x16(trow + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int key = c0 + q;
      if (key >= key_lo && key < key_hi) mx = fmaxf(mx, v[q]);
    }
  }
  float sum = 0.f;
  for (int c0 = 0; c0 < a.nkeys; c0 += 16) {
    float v[16];
    tmem_ld_x16(trow + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int key = c0 + q;
      float p = 0.f;
      if (key >= key_lo && key < key_hi) p = ex2_approx((v[q] - mx) * a.scale_log2e);
      // the probabilities are rounded to fp16 for the PV MMA; normalise by the sum of the
      // rounded values so each row of P/sum sums to one exactly as seen by the tensor core
      const float pr = __half2float(__float2half_rn(p));
      sum += pr;
      v[q] = p;
    }
    const uint32_t atom = c0 >> 6, ch = (c0 & 63) >> 3;
    st_smem_16B(sP + atom * 16384 + sw128_offset(t, ch), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    st_smem_16B(sP + atom * 16384 + sw128_offset(t, ch + 1), pack_h2(v[8], v[9]), pack_h2(v[10], v[11]),
                pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
  }
  const float inv = rcp_approx(sum);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- O = P V ----
  if (t == 0) {
    const uint32_t idesc = umma_idesc_f16(128, dhp, a.ab_fmt);
    for (int kk = 0; kk < a.nkeys / 16; ++kk) {
      const uint32_t atom = kk >> 2, ks = kk & 3;
      umma_f16(tmem, umma_desc_sw128(sP + atom * 16384 + ks * 32),
               umma_desc_sw128(sVt + atom * (dhp * 128) + ks * 32), idesc, kk != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  __half* orow = a.out + (static_cast<size_t>(mt) * 128 + t) * a.C + h * dh;
  for (int c0 = 0; c0 < dhp; c0 += 16) {
    float v[16];
    tmem_ld_x16(trow + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (c0 + q * 8 < dh) {
        uint4 o;
        o.x = pack_h2(v[q * 8 + 0] * inv, v[q * 8 + 1] * inv); o.y = pack_h2(v[q * 8 + 2] * inv, v[q * 8 + 3] * inv);
        o.z = pack_h2(v[q * 8 + 4] * inv, v[q * 8 + 5] * inv); o.w = pack_h2(v[q * 8 + 6] * inv, v[q * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + c0 + q * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace rvt
