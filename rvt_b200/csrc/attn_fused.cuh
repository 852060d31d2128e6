// Fused attention half of PartitionAttentionCl (reference maxvit.py:252-268 with
// SelfAttentionCl :343-354 and the window/grid partition + reverse :273-304):
//     x[tok] <- x[tok] + gamma1 * ( Wp * concat_h softmax(q_h k_h^T / sqrt(dh)) v_h + bp )
// for one 128-row tile of partition-ordered tokens, in ONE kernel:
//
//   workers  gather x rows through the partition map, LN1 -> fp16 A operand (smem)
//   loader   per-head [q|k|v] weight tile (bulk copy, single buffer) + the proj weight (once)
//   MMA      QKV_h = A * Wqkv_h^T -> TMEM;   S = Q K^T;   O_h = P V;   out = O * Wp^T
//   workers  QKV_h + bias -> fp16 Q|K (packed in one SW128 tile) and V^T operand tiles
//            masked softmax of S (fp32, exp2; two threads per row, halves exchanged through smem)
//            O_h / rowsum -> fp16 column block h of the proj A operand
//            out + bias, * gamma1, + residual -> scattered back (= partition reverse)
//
// QKV of head h+1 is issued while the softmax of head h runs.  TMEM: [0,96) QKV_h, [96,128) O_h,
// [128,256) S; the proj accumulator reuses [0,C).  C <= 128, dim_head <= 32 (padded to 32; the
// per-head weight rows are zero padded by packing.pack_qkv_weight), P <= 128.
#pragma once
#include "gemm_fused.cuh"
#include "mlp_fused.cuh"

namespace rvt {

struct AttnFusedArgs {
  float* x;                  // [B, H, W, C] in/out
  RowMap map;
  int C, dh, nh;
  const float* ln_w; const float* ln_b; float eps; int do_ln;
  const __half* wqkv;        // pack_qkv_weight: [nh][KC1][96 x 64]
  const float* bqkv;         // [nh][96] padded (zeros where padded / no bias)
  const __half* wproj;       // pack_linear_weight(Wp, bn = C): [1][KC1][C x 64]
  const float* bproj;        // [C] or null
  const float* gamma;        // [C] or null
  float scale_log2e;
};

constexpr int kAfThreads = 320;
constexpr int kAfDhp = 32;                 // padded head dim
constexpr int kAfQkvN = 3 * kAfDhp;        // 96 accumulator columns per head

__host__ __device__ inline size_t attn_fused_smem_bytes(int C) {
  const size_t kc1 = (C + 63) / 64;
  return 1024 + kc1 * kATileBytes /*A*/ + kATileBytes /*Q|K*/ + 2 * kAfDhp * 128 /*Vt*/ + 2 * kATileBytes /*P*/ +
         kc1 * kATileBytes /*O*/ + kc1 * kAfQkvN * 128 /*Wqkv head*/ + kc1 * static_cast<size_t>(C) * 128 /*Wproj*/ +
         2 * 128 * sizeof(float) + 256;
}

template <int KC1>   // K atoms of the C-wide operand: 1 (C <= 64) or 2 (C <= 128); sizes the register-resident x rows
__global__ void __launch_bounds__(kAfThreads, 2) attn_fused_kernel(const __grid_constant__ AttnFusedArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x;
  const int C = a.C, nh = a.nh, dh = a.dh;
  const uint32_t wq_bytes = static_cast<uint32_t>(KC1) * kAfQkvN * 128;
  const uint32_t wp_bytes = static_cast<uint32_t>(KC1) * C * 128;

  const uint32_t sA = base;
  const uint32_t sQK = sA + KC1 * kATileBytes;
  const uint32_t sVt = sQK + kATileBytes;
  const uint32_t sP = sVt + 2 * kAfDhp * 128;
  const uint32_t sO = sP + 2 * kATileBytes;
  const uint32_t sWq = sO + KC1 * kATileBytes;
  const uint32_t sWp = sWq + wq_bytes;
  float* s_red = reinterpret_cast<float*>(sm + (sWp - base) + wp_bytes);       // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 256);
  uint64_t* bar_a = bars + 0;        // 256
  uint64_t* wq_full = bars + 1;      // tx
  uint64_t* wq_empty = bars + 2;     // commit
  uint64_t* wp_full = bars + 3;      // tx
  uint64_t* qkv_full = bars + 4;     // commit
  uint64_t* qkv_smem = bars + 5;     // 256
  uint64_t* s_full = bars + 6;       // commit
  uint64_t* p_full = bars + 7;       // 256
  uint64_t* o_full = bars + 8;       // commit
  uint64_t* so_full = bars + 9;      // 256
  uint64_t* out_full = bars + 10;    // commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  // x-tile gather issued before barrier init / TMEM allocation / CTA sync (setup overlaps the load latency)
  const int j8 = tid & 7, r0 = tid >> 3;
  int tok[4] = {-1, -1, -1, -1};
  float keep[KC1][4][8];
  if (warp < 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) tok[i] = row_to_token(a.map, mt * 128 + r0 + 32 * i);
#pragma unroll
    for (int kc = 0; kc < KC1; ++kc) {
      const int k0 = kc * 64 + j8 * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e) keep[kc][i][e] = 0.f;
        if (tok[i] >= 0 && k0 < C) load8(a.x + static_cast<size_t>(tok[i]) * C + k0, keep[kc][i]);
      }
    }
  }

  if (tid == 0) {
    mbar_init(bar_a, kWorkers); mbar_init(wq_full, 1); mbar_init(wq_empty, 1); mbar_init(wp_full, 1);
    mbar_init(qkv_full, 1); mbar_init(qkv_smem, kWorkers); mbar_init(s_full, 1); mbar_init(p_full, kWorkers);
    mbar_init(o_full, 1); mbar_init(so_full, kWorkers); mbar_init(out_full, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_qkv = tmem, t_o = tmem + kAfQkvN, t_s = tmem + 128, t_out = tmem;

  const int rpw = a.map.rows_per_win, P = a.map.P;
  const int nk_w = (P + 15) & ~15;                       // keys a row attends to, padded to 16
  const int nkeys = rpw == 64 ? 128 : nk_w;              // MMA N of S / K of PV

  if (warp < 8) {
    // zero the P operand once: blocks outside a row's own window stay zero for every head
    for (int i = tid; i < 2 * static_cast<int>(kATileBytes) / 16; i += kWorkers) st_smem_16B(sP + i * 16, 0u, 0u, 0u, 0u);

    // ======================= gather + LN1 -> A operand =======================
    {
      float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f}, s1[4] = {0.f, 0.f, 0.f, 0.f},
            s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) s1[i] += keep[kc][i][e];
      if (a.do_ln) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mean[i] = red8(s1[i]) / C;
#pragma unroll
        for (int kc = 0; kc < KC1; ++kc) {
          if (kc * 64 + j8 * 8 < C) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int e = 0; e < 8; ++e) { const float d = keep[kc][i][e] - mean[i]; s2[i] += d * d; }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) rstd[i] = rsqrtf(red8(s2[i]) / C + a.eps);
      }
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc) {
        const int k0 = kc * 64 + j8 * 8;
        const bool kv = k0 < C;
        float g[8], bb[8];
        if (a.do_ln && kv) { load8(a.ln_w + k0, g); load8(a.ln_b + k0, bb); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (tok[i] >= 0 && kv) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = keep[kc][i][e];
            if (a.do_ln) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean[i]) * rstd[i] * g[e] + bb[e];
            }
          }
          st_smem_16B(sA + kc * kATileBytes + sw128_offset(r0 + 32 * i, j8), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                      pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(bar_a);

    const int q = warp & 3, hsel = warp >> 2;
    const int erow = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int key_lo = (erow / rpw) * rpw;
    const int ksplit = ((nk_w / 16 + 1) / 2) * 16;
    const int kbeg = hsel ? ksplit : 0, kend = hsel ? nk_w : ksplit;     // this thread's keys (relative to key_lo)

    for (int h = 0; h < nh; ++h) {
      const uint32_t par = h & 1;
      // ---- QKV_h accumulators -> Q|K tile and V^T tile ----
      mbar_wait(qkv_full, par);
      tc_fence_after();
      {
        const float* bq = a.bqkv + h * kAfQkvN;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {                       // this thread's 48 of the 96 columns
          const int c0 = hsel * 48 + cc * 16;
          float v[16], bv[16];
          tmem_ld_x16(t_qkv + lane_off + c0, v);
          tmem_ld_wait();
          load16(bq + c0, bv);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += bv[e];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c = c0 + half * 8;
            const int part = c >> 5, d0 = c & 31;              // 0 q, 1 k, 2 v ; first head-dim index
            const float* p8 = v + half * 8;
            if (part < 2) {
              st_smem_16B(sQK + sw128_offset(erow, part * 4 + (d0 >> 3)), pack_h2(p8[0], p8[1]), pack_h2(p8[2], p8[3]),
                          pack_h2(p8[4], p8[5]), pack_h2(p8[6], p8[7]));
            } else {
              const uint32_t atom = erow >> 6, kk = erow & 63;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t addr = sVt + atom * (kAfDhp * 128) + sw128_offset(d0 + e, kk >> 3) + (kk & 7) * 2;
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(__float2half_rn(p8[e]))) : "memory");
              }
            }
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(qkv_smem);

      // ---- masked softmax of row erow over its window's keys ----
      mbar_wait(s_full, par);
      tc_fence_after();
      float mx = -INFINITY;
      for (int k0 = kbeg; k0 < kend; k0 += 16) {
        float v[16];
        tmem_ld_x16(t_s + lane_off + key_lo + k0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (k0 + e < P) mx = fmaxf(mx, v[e]);
      }
      s_red[hsel * 128 + erow] = mx;
      named_bar_sync(1, kWorkers);
      mx = fmaxf(s_red[erow], s_red[128 + erow]);
      named_bar_sync(2, kWorkers);
      float sum = 0.f;
      for (int k0 = kbeg; k0 < kend; k0 += 16) {
        float v[16];
        tmem_ld_x16(t_s + lane_off + key_lo + k0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float p = 0.f;
          if (k0 + e < P) p = ex2_approx((v[e] - mx) * a.scale_log2e);
          sum += p;
          v[e] = p;
        }
        const int key = key_lo + k0;
        const uint32_t atom = key >> 6, ch = (key & 63) >> 3;
        st_smem_16B(sP + atom * kATileBytes + sw128_offset(erow, ch), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                    pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        st_smem_16B(sP + atom * kATileBytes + sw128_offset(erow, ch + 1), pack_h2(v[8], v[9]), pack_h2(v[10], v[11]),
                    pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
      }
      s_red[hsel * 128 + erow] = sum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      named_bar_sync(1, kWorkers);
      const float inv = rcp_approx(s_red[erow] + s_red[128 + erow]);
      named_bar_sync(2, kWorkers);

      // ---- O_h / rowsum -> column block h of the proj A operand ----
      mbar_wait(o_full, par);
      tc_fence_after();
      {
        float v[16];
        tmem_ld_x16(t_o + lane_off + hsel * 16, v);      // this thread's 16 of the 32 (padded) head dims
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int d0 = hsel * 16 + half * 8;
          if (d0 < dh) {
            const int col = h * dh + d0;                  // heads concatenated with stride dh (maxvit.py:352)
            const float* p8 = v + half * 8;
            st_smem_16B(sO + (col >> 6) * kATileBytes + sw128_offset(erow, (col & 63) >> 3), pack_h2(p8[0] * inv, p8[1] * inv),
                        pack_h2(p8[2] * inv, p8[3] * inv), pack_h2(p8[4] * inv, p8[5] * inv), pack_h2(p8[6] * inv, p8[7] * inv));
          }
        }
      }
      tc_fence_before();
    }
    // zero-fill the K padding of the last proj atom (C % 64 != 0) is unnecessary: Wproj is zero
    // padded there, but the smem must hold finite values -> A atoms were fully written above? No:
    // sO columns >= C are never written; they are multiplied by zero weights only if finite.
    if (C & 63) {
      for (int idx = tid; idx < 128 * 8; idx += kWorkers) {
        const int r = idx >> 3, ch = idx & 7;
        if (ch * 8 >= (C & 63)) st_smem_16B(sO + (KC1 - 1) * kATileBytes + sw128_offset(r, ch), 0u, 0u, 0u, 0u);
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(so_full);

    // ======================= proj epilogue: residual + scatter =======================
    // residual rows prefetched before the proj accumulator barrier
    const int etok = row_to_token(a.map, mt * 128 + erow);
    const int ocols = (C + 15) & ~15;
    const int csplit = ((ocols / 16 + 1) / 2) * 16;
    const int cbeg = hsel ? csplit : 0, cend = hsel ? ocols : csplit;
    float res[64];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c0 = cbeg + g * 16;
      if (c0 < cend && etok >= 0) load16(a.x + static_cast<size_t>(etok) * C + c0, res + g * 16);
    }
    mbar_wait(out_full, 0);
    tc_fence_after();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c0 = cbeg + g * 16;
      if (c0 >= cend) break;
      float v[16];
      tmem_ld_x16(t_out + lane_off + c0, v);
      tmem_ld_wait();
      if (etok >= 0) {
        float* xp = a.x + static_cast<size_t>(etok) * C + c0;
        float bv[16];
        if (a.bproj) {
          load16(a.bproj + c0, bv);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += bv[e];
        }
        if (a.gamma) {
          load16(a.gamma + c0, bv);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] *= bv[e];
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          *reinterpret_cast<float4*>(xp + qd * 4) =
              make_float4(res[g * 16 + qd * 4] + v[qd * 4], res[g * 16 + qd * 4 + 1] + v[qd * 4 + 1],
                          res[g * 16 + qd * 4 + 2] + v[qd * 4 + 2], res[g * 16 + qd * 4 + 3] + v[qd * 4 + 3]);
      }
    }
  } else if (warp == 8) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t id_qkv = umma_idesc_f16(128, kAfQkvN, 0);
      const uint32_t id_s = umma_idesc_f16(128, nkeys, 0);
      const uint32_t id_o = umma_idesc_f16(128, kAfDhp, 0);
      const uint32_t id_out = umma_idesc_f16(128, (C + 15) & ~15, 0);
      const int ks1 = (C + 15) / 16;
      auto issue_qkv = [&](int h) {
        mbar_wait(wq_full, h & 1);
        tc_fence_after();
        for (int k = 0; k < ks1; ++k) {
          const uint32_t atom = k >> 2, kk = k & 3;
          umma_f16(t_qkv, umma_desc_sw128(sA + atom * kATileBytes + kk * 32),
                   umma_desc_sw128(sWq + atom * (kAfQkvN * 128) + kk * 32), id_qkv, k != 0);
        }
        umma_commit(qkv_full);
        umma_commit(wq_empty);
      };
      mbar_wait(bar_a, 0);
      tc_fence_after();
      issue_qkv(0);
      for (int h = 0; h < nh; ++h) {
        const uint32_t par = h & 1;
        mbar_wait(qkv_smem, par);
        tc_fence_after();
        for (int k = 0; k < kAfDhp / 16; ++k)                 // S = Q K^T ; Q at chunks 0-3, K at chunks 4-7
          umma_f16(t_s, umma_desc_sw128(sQK + k * 32), umma_desc_sw128(sQK + 64 + k * 32), id_s, k != 0);
        umma_commit(s_full);
        if (h + 1 < nh) issue_qkv(h + 1);
        mbar_wait(p_full, par);
        tc_fence_after();
        for (int kk = 0; kk < nkeys / 16; ++kk) {
          const uint32_t atom = kk >> 2, ks = kk & 3;
          umma_f16(t_o, umma_desc_sw128(sP + atom * kATileBytes + ks * 32),
                   umma_desc_sw128(sVt + atom * (kAfDhp * 128) + ks * 32), id_o, kk != 0);
        }
        umma_commit(o_full);
      }
      mbar_wait(so_full, 0);
      mbar_wait(wp_full, 0);
      tc_fence_after();
      for (int k = 0; k < ks1; ++k) {
        const uint32_t atom = k >> 2, kk = k & 3;
        umma_f16(t_out, umma_desc_sw128(sO + atom * kATileBytes + kk * 32),
                 umma_desc_sw128(sWp + atom * (C * 128) + kk * 32), id_out, k != 0);
      }
      umma_commit(out_full);
    }
    __syncwarp();
  } else {
    // ======================= weight loader =======================
    if (lane == 0) {
      mbar_arrive_expect_tx(wp_full, wp_bytes);
      bulk_g2s(sm + (sWp - base), a.wproj, wp_bytes, wp_full);
      for (int h = 0; h < nh; ++h) {
        mbar_wait(wq_empty, (h & 1) ^ 1);
        mbar_arrive_expect_tx(wq_full, wq_bytes);
        bulk_g2s(sm + (sWq - base), a.wqkv + static_cast<size_t>(h) * KC1 * kAfQkvN * 64, wq_bytes, wq_full);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 256);
}

}  // namespace rvt
