// Fused MLP half of PartitionAttentionCl (reference maxvit.py:241,269 + MLP :85-118):
//     x <- x + gamma2 * ( W2 gelu(W1 LN2(x) + b1) + b2 )            for one 128-token tile
// in ONE kernel.  The 4C-wide hidden activation never leaves the SM:
//
//   workers (8 warps)   LN2(x tile) -> fp16 A operand in smem (all K atoms), once
//   loader  (1 thread)  streams [W1 chunk | W2 chunk] pairs through a bulk-copy ring
//   MMA     (1 thread)  fc1 chunk j:  Hj[128 x 64]  = A * W1_j^T        -> TMEM (double buffered)
//   workers             Hj + b1 -> exact GELU -> fp16 -> smem operand chunk (double buffered)
//   MMA                 fc2:  out[128 x C] += gelu(Hj) * W2_j^T          -> TMEM
//   workers             out + b2, * gamma2, + residual -> x
//
// fc1 of chunk j+1 is issued before the GELU of chunk j is consumed, so tensor pipe, MUFU/FMA
// pipes and the weight stream overlap.  C in {32..256}, C % 16 == 0 (K atoms of 64 zero padded),
// hidden % 64 == 0.
#pragma once
#include "gemm_fused.cuh"

namespace rvt {

struct MlpArgs {
  float* x;                 // [n_tokens, C] in/out
  int n_tokens, C, hidden;
  const float* ln_w; const float* ln_b; float eps;
  const __half* w1p;        // pack_linear_weight(W1[hidden, C], bn = 64): [hidden/64][KC1][64 x 64]
  const float* b1;
  const __half* w2p;        // pack_linear_weight(W2[C, hidden], bn = C):  [1][hidden/64][C x 64]
  const float* b2;
  const float* gamma;       // LayerScale or null
  int stages;               // weight ring depth
  int gelu_f16x2;           // 1: packed-half GELU (gemm_fused.cuh gelu_f16x2, opt-in) instead of the fp32 exact-erf evaluation
};

constexpr int kMlpThreads = 320;      // 8 worker warps + MMA warp + loader warp
constexpr int kMlpHC = 64;            // hidden columns per chunk

__host__ __device__ inline int mlp_kc1(int C) { return (C + 63) / 64; }
__host__ __device__ inline size_t mlp_stage_bytes(int C) {
  return static_cast<size_t>(mlp_kc1(C)) * kMlpHC * 128 + static_cast<size_t>(C) * 128;
}
__host__ __device__ inline size_t mlp_smem_bytes(int C, int stages) {
  return 1024 + static_cast<size_t>(mlp_kc1(C)) * kATileBytes + stages * mlp_stage_bytes(C) + 2 * kATileBytes + 256;
}

// KC1: K atoms of the C-wide operand, 1 (C <= 64) or 2 (C <= 128) — sizes the register-resident x rows.  GELU_H2: the opt-in
// packed-half GELU as a separate instantiation, so the default kernel's code is exactly what was measured.
template <int KC1, bool GELU_H2 = false>
__global__ void __launch_bounds__(kMlpThreads, 2) mlp_fused_kernel(const __grid_constant__ MlpArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x;
  const int C = a.C, S = a.stages;
  const int n_chunks = a.hidden / kMlpHC;
  const uint32_t w1_bytes = static_cast<uint32_t>(KC1) * kMlpHC * 128;
  const uint32_t w2_bytes = static_cast<uint32_t>(C) * 128;
  const uint32_t stage_bytes = w1_bytes + w2_bytes;

  const uint32_t sA = base;
  const uint32_t sRing = sA + KC1 * kATileBytes;
  const uint32_t sH = sRing + S * stage_bytes;
  uint8_t* ring_ptr = sm + KC1 * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (sH - base) + 2 * kATileBytes);
  uint64_t* bar_a = bars;                  // A operand ready (256 arrivals)
  uint64_t* w_full = bars + 1;             // [S]
  uint64_t* w_empty = w_full + kMaxStages; // [S]
  uint64_t* hid_full = w_empty + kMaxStages;   // [2]
  uint64_t* sh_full = hid_full + 2;        // [2] (256 arrivals)
  uint64_t* sh_empty = sh_full + 2;        // [2]
  uint64_t* out_full = sh_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_full + 1);

  const uint32_t out_cols = static_cast<uint32_t>((C + 15) & ~15);
  const uint32_t tmem_cols = tmem_cols_pow2(out_cols + 2 * kMlpHC);

  // The x-tile loads of the LayerNorm prologue are issued before barrier init / TMEM allocation / the CTA
  // sync, so the ~1 us of setup overlaps the ~1 us global-load latency.
  const int j8 = tid & 7, r0 = tid >> 3;
  int tok[4] = {-1, -1, -1, -1};
  float keep[KC1][4][8];         // the tile's x rows stay in registers between the LN passes
  float s1[4] = {0.f, 0.f, 0.f, 0.f};
  if (warp < 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = mt * 128 + r0 + 32 * i;
      tok[i] = row < a.n_tokens ? row : -1;
    }
#pragma unroll
    for (int kc = 0; kc < KC1; ++kc) {             // loads only: no use of the values before the CTA sync
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
    mbar_init(bar_a, kWorkers);
    for (int s = 0; s < S; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&hid_full[b], 1); mbar_init(&sh_full[b], kWorkers); mbar_init(&sh_empty[b], 1); }
    mbar_init(out_full, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_out = tmem, t_hid = tmem + out_cols;

  if (warp < 8) {
    // ======================= LN2 -> A operand =======================
    {
      float mean[4], rstd[4], s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) s1[i] += keep[kc][i][e];
#pragma unroll
      for (int i = 0; i < 4; ++i) mean[i] = red8(s1[i]) / C;
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc) {
        const int k0 = kc * 64 + j8 * 8;
        if (k0 < C) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float d = keep[kc][i][e] - mean[i]; s2[i] += d * d; }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) rstd[i] = rsqrtf(red8(s2[i]) / C + a.eps);
#pragma unroll
      for (int kc = 0; kc < KC1; ++kc) {
        const int k0 = kc * 64 + j8 * 8;
        const bool kv = k0 < C;
        float g[8], bb[8];
        if (kv) { load8(a.ln_w + k0, g); load8(a.ln_b + k0, bb); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (tok[i] >= 0 && kv) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (keep[kc][i][e] - mean[i]) * rstd[i] * g[e] + bb[e];
          }
          st_smem_16B(sA + kc * kATileBytes + sw128_offset(r0 + 32 * i, j8), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                      pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(bar_a);

    // ======================= GELU stage: TMEM chunk -> smem operand =======================
    const int q = warp & 3, hsel = warp >> 2;
    const int erow = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    for (int j = 0; j < n_chunks; ++j) {
      const int b = j & 1;
      mbar_wait(&hid_full[b], (j >> 1) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld_x16(t_hid + lane_off + b * kMlpHC + hsel * 32, v);
      tmem_ld_x16(t_hid + lane_off + b * kMlpHC + hsel * 32 + 16, v + 16);
      tmem_ld_wait();
      const uint32_t dst = sH + b * kATileBytes;
      if (GELU_H2) {
        // opt-in: bias in fp32, then the packed-half GELU on two activations per instruction
        uint32_t hp[16];
        float bv[16];
        load16(a.b1 + j * kMlpHC + hsel * 32, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) hp[e] = gelu_f16x2(pack_h2(v[2 * e] + bv[2 * e], v[2 * e + 1] + bv[2 * e + 1]));
        load16(a.b1 + j * kMlpHC + hsel * 32 + 16, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) hp[8 + e] = gelu_f16x2(pack_h2(v[16 + 2 * e] + bv[2 * e], v[16 + 2 * e + 1] + bv[2 * e + 1]));
        mbar_wait(&sh_empty[b], ((j >> 1) & 1) ^ 1);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_smem_16B(dst + sw128_offset(erow, hsel * 4 + c), hp[4 * c], hp[4 * c + 1], hp[4 * c + 2], hp[4 * c + 3]);
      } else {
        {
          float bv[16];
          load16(a.b1 + j * kMlpHC + hsel * 32, bv);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = gelu_erf(v[e] + bv[e]);
          load16(a.b1 + j * kMlpHC + hsel * 32 + 16, bv);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[16 + e] = gelu_erf(v[16 + e] + bv[e]);
        }
        mbar_wait(&sh_empty[b], ((j >> 1) & 1) ^ 1);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_smem_16B(dst + sw128_offset(erow, hsel * 4 + c), pack_h2(v[8 * c], v[8 * c + 1]), pack_h2(v[8 * c + 2], v[8 * c + 3]),
                      pack_h2(v[8 * c + 4], v[8 * c + 5]), pack_h2(v[8 * c + 6], v[8 * c + 7]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&sh_full[b]);
    }

    // ======================= output epilogue =======================
    // the residual rows are fetched BEFORE waiting on the last fc2, so their latency hides behind the MMAs
    const int row = mt * 128 + erow;
    const int csplit = ((static_cast<int>(out_cols) / 16 + 1) / 2) * 16;
    const int cbeg = hsel ? csplit : 0, cend = hsel ? static_cast<int>(out_cols) : csplit;
    float res[64];                                   // <= 64 columns per thread (C <= 128)
    const bool live = row < a.n_tokens;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c0 = cbeg + g * 16;
      if (c0 < cend && live) load16(a.x + static_cast<size_t>(row) * C + c0, res + g * 16);
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
      if (live) {
        float* xp = a.x + static_cast<size_t>(row) * C + c0;
        float bv[16];
        load16(a.b2 + c0, bv);
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] += bv[e];
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
      const uint32_t idesc1 = umma_idesc_f16(128, kMlpHC, 0);
      const uint32_t idesc2 = umma_idesc_f16(128, out_cols, 0);
      const int ks1 = (C + 15) / 16;           // K steps of fc1 (zero-padded inside the last atom)
      auto issue_fc1 = [&](int j) {
        const int s = j % S;
        mbar_wait(&w_full[s], (j / S) & 1);
        tc_fence_after();
        const uint32_t w1 = sRing + s * stage_bytes;
        for (int k = 0; k < ks1; ++k) {
          const uint32_t atom = k >> 2, kk = k & 3;
          umma_f16(t_hid + (j & 1) * kMlpHC, umma_desc_sw128(sA + atom * kATileBytes + kk * 32),
                   umma_desc_sw128(w1 + atom * (kMlpHC * 128) + kk * 32), idesc1, k != 0);
        }
        umma_commit(&hid_full[j & 1]);
      };
      mbar_wait(bar_a, 0);
      tc_fence_after();
      issue_fc1(0);
      for (int j = 0; j < n_chunks; ++j) {
        if (j + 1 < n_chunks) {
          // TMEM buffer (j+1)&1 is free once the workers have drained chunk j-1 out of it
          if (j >= 1) { mbar_wait(&sh_full[(j + 1) & 1], ((j - 1) >> 1) & 1); tc_fence_after(); }
          issue_fc1(j + 1);
        }
        mbar_wait(&sh_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const int s = j % S;
        const uint32_t w2 = sRing + s * stage_bytes + w1_bytes;
        const uint32_t h = sH + (j & 1) * kATileBytes;
        for (int k = 0; k < kMlpHC / 16; ++k)
          umma_f16(t_out, umma_desc_sw128(h + k * 32), umma_desc_sw128(w2 + k * 32), idesc2, (j | k) != 0);
        umma_commit(&w_empty[s]);
        umma_commit(&sh_empty[j & 1]);
      }
      umma_commit(out_full);
    }
    __syncwarp();
  } else {
    // ======================= weight loader =======================
    if (lane == 0) {
      for (int j = 0; j < n_chunks; ++j) {
        const int s = j % S;
        mbar_wait(&w_empty[s], ((j / S) & 1) ^ 1);
        mbar_arrive_expect_tx(&w_full[s], stage_bytes);
        uint8_t* dst = ring_ptr + static_cast<size_t>(s) * stage_bytes;
        bulk_g2s(dst, a.w1p + static_cast<size_t>(j) * KC1 * kMlpHC * 64, w1_bytes, &w_full[s]);
        bulk_g2s(dst + w1_bytes, a.w2p + static_cast<size_t>(j) * C * 64, w2_bytes, &w_full[s]);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, tmem_cols);
}

}  // namespace rvt
