// Fused MLP half of PartitionAttentionCl, persistent + chunk-parallel version (reference maxvit.py:241,269 + MLP :85-118):
//     x <- x + gamma2 * ( W2 gelu(W1 LN2(x) + b1) + b2 )
//
// One CTA per SM loops over 128-token tiles; both weight matrices stay resident in shared memory (C <= 64, hidden <= 256):
//
//   producer   one thread: W1 / W2 tile images once (bulk copies), then the fp32 x tile of tile i+1 by 2-D TMA into the other
//              half of a double buffer while tile i is being processed
//   MMA        one thread: fc1 = ONE N = hidden MMA chain -> TMEM [0, hidden);  fc2 accumulates chunk by chunk as the
//              GELU'd operand chunks arrive -> TMEM [256, 256 + C)
//   workers    4 warpgroups, thread = tile row = TMEM lane.  Per tile i:
//                GELU(i)   warpgroup j: hidden columns [64j, 64j+64) + b1 -> exact-erf GELU (or the packed-half variant)
//                          -> fp16 operand chunk j of fc2           (all chunks concurrently)
//                LN(i+1)   all: LayerNorm of the NEXT tile (smem fp32 -> fp16 A operand) -- fc1(i+1) then runs while ...
//                EPI(i)    all: out + b2, * gamma2, + residual -> x  ... the epilogue of tile i is still being stored
//
// so the tensor-core latencies and the barrier hand-offs of one tile hide behind the ALU work of its neighbours.
#pragma once
#include "attn_v2.cuh"

namespace rvt {

struct MlpV2Args {
  float* x;                 // [n_tokens, C] in/out
  int n_tokens, C, hidden, n_tiles;
  const float* ln_w; const float* ln_b; float eps;
  const __half* w1p;        // pack_linear_weight(W1[hidden, C], bn = 64): [hidden/64][1][64 x 64]
  const float* b1;
  const __half* w2p;        // pack_linear_weight(W2[C, hidden], bn = C):  [1][hidden/64][C x 64]
  const float* b2;
  const float* gamma;       // LayerScale or null
  long long* trace;         // optional [grid][kTraceTiles][kTracePts] globaltimer stamps (profiling aid)
  int fast_ln;              // 1: C in {32, 64}: x tiles arrive as 32-channel SW128 half tiles, thread-per-row LayerNorm
};

constexpr int kMv2Workers = 512;
constexpr int kMv2Threads = kMv2Workers + 64;          // + MMA warp + producer warp
constexpr uint32_t kMv2XBuf = 128 * 64 * 4;            // one fp32 x tile (C <= 64)
constexpr uint32_t kMv2Smem = 1024 + 2 * kMv2XBuf + kAv2Tile /*A*/ + 4 * kAv2Tile /*H*/ + 4 * 8192 /*W1*/ + 4 * 8192 /*W2*/ +
                              (256 + 4 * 64 + 2 * 256) * 4 + 24 * 8 + 16;

template <bool GELU_H2>
__global__ void __launch_bounds__(kMv2Threads, 1)
mlp_v2_kernel(const __grid_constant__ MlpV2Args a, const __grid_constant__ CUtensorMap tmap_x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, hidden = a.hidden, nch = hidden >> 6;
  const uint32_t sX = base;
  const uint32_t sA = sX + 2 * kMv2XBuf;
  const uint32_t sH = sA + kAv2Tile;
  const uint32_t sW1 = sH + 4 * kAv2Tile;
  const uint32_t sW2 = sW1 + 4 * 8192;
  float* s_b1 = reinterpret_cast<float*>(sm + (sW2 - base) + 4 * 8192);      // [256]
  float* s_b2 = s_b1 + 256;                                                  // [64] each
  float* s_gamma = s_b2 + 64;
  float* s_lnw = s_gamma + 64;
  float* s_lnb = s_lnw + 64;
  float* s_part = s_lnb + 64;                                                // [2][128][2] LayerNorm partial sums (fast_ln)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 256);
  uint64_t* x_full = bars + 0;        // [2] tx
  uint64_t* x_free = bars + 2;        // [2] 512
  uint64_t* a_full = bars + 4;        // 512
  uint64_t* hid_full = bars + 5;      // commit
  uint64_t* out_full = bars + 6;      // commit
  uint64_t* out_free = bars + 7;      // 512
  uint64_t* w_full = bars + 8;        // tx
  uint64_t* sh_full = bars + 9;       // [4] 128
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&x_full[b], 1); mbar_init(&x_free[b], kMv2Workers); }
    mbar_init(a_full, kMv2Workers); mbar_init(hid_full, 1); mbar_init(out_full, 1); mbar_init(out_free, kMv2Workers);
    mbar_init(w_full, 1);
    for (int j = 0; j < 4; ++j) mbar_init(&sh_full[j], 128);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 256; i += kMv2Threads) s_b1[i] = (i < hidden && a.b1) ? a.b1[i] : 0.f;
  for (int i = tid; i < 64; i += kMv2Threads) {
    const bool in = i < C;
    s_b2[i] = (in && a.b2) ? a.b2[i] : 0.f;
    s_gamma[i] = (in && a.gamma) ? a.gamma[i] : 1.f;
    s_lnw[i] = in ? a.ln_w[i] : 1.f;
    s_lnb[i] = in ? a.ln_b[i] : 0.f;
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const uint32_t t_hid = tmem, t_out = tmem + 256;
  const int n_tiles = a.n_tiles;
  const int ks1 = C >> 4;

  if (warp < 16) {
    // =============================================== workers ===============================================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int j8 = tid & 7;

    auto layer_norm = [&](int it) {          // x tile `it` (smem fp32) -> fp16 A operand; 8 lanes per row
      const int b = it & 1;
      mbar_wait(&x_full[b], (it >> 1) & 1);
      const uint32_t xb = sX + b * kMv2XBuf;
      const int k0 = j8 * 8;
      if (a.fast_ln) {
        if (32 * wg < C) {
          if (C == 64) ln_row32_to_operand<2>(xb + wg * kAv2Tile, row, wg, true, true, C, a.eps, s_lnw, s_lnb, s_part, 1, sA);
          else ln_row32_to_operand<1>(xb + wg * kAv2Tile, row, wg, true, true, C, a.eps, s_lnw, s_lnb, s_part, 1, sA);
        }
      } else
      for (int r = tid >> 3; r < 128; r += kMv2Workers / 8) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (k0 < C) {
          const uint32_t src = xb + (static_cast<uint32_t>(r) * C + k0) * 4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(src));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(src + 16));
        }
        float s1 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s1 += v[e];
        const float mean = red8(s1) / C;
        float s2 = 0.f;
        if (k0 < C) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float d = v[e] - mean; s2 += d * d; }
        }
        const float rstd = rsqrtf(red8(s2) / C + a.eps);
        if (k0 < C) {
          float g[8], bb[8];
          lds8(s_lnw + k0, g);
          lds8(s_lnb + k0, bb);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaf((v[e] - mean) * rstd, g[e], bb[e]);
        }
        st_smem_16B(sA + sw128_offset(r, j8), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_full);
      mbar_arrive(&x_free[b]);
    };

    if (static_cast<int>(blockIdx.x) < n_tiles) layer_norm(0);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      RVT_TRACE(a, it, 0);
      // ---------------- GELU(it): hidden chunk wg -> fp16 operand chunk of fc2 ----------------
      mbar_wait(hid_full, par);                // everybody: fc1(it) has finished reading the A operand
      RVT_TRACE(a, it, 1);
      tc_fence_after();
      if (wg < nch) {
        const uint32_t dst = sH + wg * kAv2Tile;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tmem_ld_x32(t_hid + lane_off + wg * 64 + half * 32, v);
          tmem_ld_wait();
          const float* bp = s_b1 + wg * 64 + half * 32;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float bv[8];
            lds8(bp + 8 * c, bv);
            uint32_t o[4];
            if (GELU_H2) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = gelu_f16x2(pack_h2(v[8 * c + 2 * e] + bv[2 * e], v[8 * c + 2 * e + 1] + bv[2 * e + 1]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = pack_h2(gelu_erf(v[8 * c + 2 * e] + bv[2 * e]), gelu_erf(v[8 * c + 2 * e + 1] + bv[2 * e + 1]));
            }
            st_smem_16B(dst + sw128_offset(row, half * 4 + c), o[0], o[1], o[2], o[3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&sh_full[wg]);
      }
      RVT_TRACE(a, it, 2);
      // ---------------- LN(it + 1): the next tile's A operand (fc1(it+1) overlaps the epilogue below) ----------------
      if (tile + static_cast<int>(gridDim.x) < n_tiles) layer_norm(it + 1);
      RVT_TRACE(a, it, 3);
      // ---------------- EPI(it): + b2, * gamma2, + residual -> x ----------------
      if ((C & (C - 1)) == 0 && C >= 32) {
        // coalesced version (see attn_v2.cuh): (acc + b2) * gamma -> swizzled fp32 staging tile in the (now idle) fc2 operand
        // region, then (row, 16-byte chunk) threads read x / write x as whole 128-byte lines
        // residual prefetch in the coalesced (row, chunk) mapping BEFORE the accumulator wait (its latency hides behind fc2)
        const int nch = C >> 2;
        const int ech = tid % nch, er0 = tid / nch, erstep = kMv2Workers / nch;
        constexpr int kRows = 4;                                  // 128 rows * nch chunks / 512 threads: 4 (C = 64), 2 (C = 32)
        const int nrows = 128 / erstep;
        float4 xr[kRows];
        bool ok[kRows];
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
          const int r = er0 + q * erstep;
          ok[q] = q < nrows && tile * 128 + r < a.n_tokens;
          if (ok[q]) xr[q] = *reinterpret_cast<const float4*>(a.x + static_cast<size_t>(tile * 128 + r) * C + ech * 4);
        }
        mbar_wait(out_full, par);
        RVT_TRACE(a, it, 4);
        tc_fence_after();
        const uint32_t srow = sH + static_cast<uint32_t>(row) * C * 4;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c8 = wg + 4 * q;
          if (c8 * 8 >= C) break;
          float v[8];
          tmem_ld_x8(t_out + lane_off + c8 * 8, v);
          tmem_ld_wait();
          float bv[8], gv[8];
          lds8(s_b2 + c8 * 8, bv);
          lds8(s_gamma + c8 * 8, gv);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (v[e] + bv[e]) * gv[e];
          const int ch = c8 * 2;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((ch ^ (row & 7)) << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((ch + 1) ^ (row & 7)) << 4)), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        }
        tc_fence_before();
        mbar_arrive(out_free);                                    // the out accumulator is drained (fc2 of the next tile may start)
        named_bar_sync(2, kMv2Workers);
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
          if (!ok[q]) continue;
          const int r = er0 + q * erstep;
          float4 sv;
          const uint32_t src = sH + static_cast<uint32_t>(r) * C * 4 + ((ech ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w) : "r"(src));
          *reinterpret_cast<float4*>(a.x + static_cast<size_t>(tile * 128 + r) * C + ech * 4) =
              make_float4(xr[q].x + sv.x, xr[q].y + sv.y, xr[q].z + sv.z, xr[q].w + sv.w);
        }
        named_bar_sync(3, kMv2Workers);                           // the staging tile is the next tile's fc2 operand
        RVT_TRACE(a, it, 5);
        continue;
      }
      // ---------------- EPI(it): + b2, * gamma2, + residual -> x ----------------
      const int tok = tile * 128 + row;
      const bool live = tok < a.n_tokens;
      float* xrow = a.x + static_cast<size_t>(live ? tok : 0) * C;
      float res[16];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c8 = wg + 4 * q;
        if (live && c8 * 8 < C) load8(xrow + c8 * 8, res + q * 8);
      }
      mbar_wait(out_full, par);
      RVT_TRACE(a, it, 4);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c8 = wg + 4 * q;
        if (c8 * 8 >= C) break;
        float v[8];
        tmem_ld_x8(t_out + lane_off + c8 * 8, v);
        tmem_ld_wait();
        if (live) {
          float bv[8], gv[8];
          lds8(s_b2 + c8 * 8, bv);
          lds8(s_gamma + c8 * 8, gv);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e] + bv[e], gv[e], res[q * 8 + e]);
          *reinterpret_cast<float4*>(xrow + c8 * 8) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(xrow + c8 * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      tc_fence_before();
      mbar_arrive(out_free);
      RVT_TRACE(a, it, 5);
    }
  } else if (warp == 16) {
    // =============================================== MMA issuer ===============================================
    if (lane == 0) {
      const uint32_t id1 = umma_idesc_f16(128, hidden, 0);
      const uint32_t id2 = umma_idesc_f16(128, C, 0);
      mbar_wait(w_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        mbar_wait(a_full, par);               // A(it) ready; every worker is done reading the hidden accumulators of tile it-1
        tc_fence_after();
        for (int k = 0; k < ks1; ++k)
          umma_f16(t_hid, umma_desc_sw128(sA + k * 32), umma_desc_sw128(sW1 + k * 32), id1, k != 0);
        umma_commit(hid_full);
        for (int j = 0; j < nch; ++j) {
          mbar_wait(&sh_full[j], par);
          if (j == 0 && it > 0) mbar_wait(out_free, (it - 1) & 1);     // epilogue(it-1) has drained the out accumulator
          tc_fence_after();
          for (int k = 0; k < 4; ++k)
            umma_f16(t_out, umma_desc_sw128(sH + j * kAv2Tile + k * 32), umma_desc_sw128(sW2 + j * C * 128 + k * 32), id2, (j | k) != 0);
        }
        umma_commit(out_full);
      }
    }
    __syncwarp();
  } else {
    // =============================================== producer ===============================================
    if (lane == 0 && static_cast<int>(blockIdx.x) < n_tiles) {
      tma_prefetch_desc(&tmap_x);
      const uint32_t w1_bytes = static_cast<uint32_t>(nch) * 8192, w2_bytes = static_cast<uint32_t>(nch) * C * 128;
      mbar_arrive_expect_tx(w_full, w1_bytes + w2_bytes);
      for (int j = 0; j < nch; ++j) {
        bulk_g2s(sm + (sW1 - base) + j * 8192, a.w1p + static_cast<size_t>(j) * 64 * 64, 8192, w_full);
        bulk_g2s(sm + (sW2 - base) + j * C * 128, a.w2p + static_cast<size_t>(j) * C * 64, static_cast<uint32_t>(C) * 128, w_full);
      }
      const uint32_t x_bytes = 128u * C * 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&x_free[b], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&x_full[b], x_bytes);
        if (a.fast_ln) {
          for (int j = 0; 32 * j < C; ++j) tma_load_2d(sX + b * kMv2XBuf + j * kAv2Tile, &tmap_x, 32 * j, tile * 128, &x_full[b]);
        } else {
          tma_load_2d(sX + b * kMv2XBuf, &tmap_x, 0, tile * 128, &x_full[b]);
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}


// ==========================================================================================================================
// C = 128 variant (stage 2 of RVT-B / T; hidden = 512): the same pipeline, with the hidden layer processed in passes of 256
// columns (the TMEM budget: 256 hidden + 128 output accumulator columns) and the weights STREAMED through a bulk-copy ring
// (W1 + W2 = 256 KB do not fit next to the tiles): per pass four [64 x 128] W1 chunks (fc1) then four [128 x 64] W2 chunks
// (fc2).  LayerNorm: 4 warpgroups x 32 channels, thread-per-row on SW128 half tiles.
// ==========================================================================================================================
struct MlpV2xArgs {
  float* x; int n_tokens, C, hidden, n_tiles;
  const float* ln_w; const float* ln_b; float eps;
  const __half* w1p;        // [hidden/64][2][64 x 64]
  const float* b1;
  const __half* w2p;        // [1][hidden/64][C x 64]
  const float* b2; const float* gamma;
  long long* trace;
};

constexpr int kMx2Stages = 3;
constexpr uint32_t kMx2Slot = 16384;                    // one W1 chunk (2 atoms x 8 KB) or one W2 chunk (128 x 64 fp16)
constexpr uint32_t kMx2XBuf = 128 * 128 * 4;            // 64 KB
constexpr uint32_t kMx2Smem = 1024 + kMx2XBuf + 2 * kAv2Tile /*A*/ + 4 * kAv2Tile /*H*/ + kMx2Stages * kMx2Slot +
                              (512 + 4 * 128 + 4 * 256) * 4 + 32 * 8 + 16;

template <bool GELU_H2>
__global__ void __launch_bounds__(kMv2Threads, 1)
mlp_v2x_kernel(const __grid_constant__ MlpV2xArgs a, const __grid_constant__ CUtensorMap tmap_x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, hidden = a.hidden, nch = hidden >> 6, npass = (nch + 3) >> 2;
  const uint32_t sX = base;
  const uint32_t sA = sX + kMx2XBuf;
  const uint32_t sH = sA + 2 * kAv2Tile;
  const uint32_t sW = sH + 4 * kAv2Tile;
  float* s_b1 = reinterpret_cast<float*>(sm + (sW - base) + kMx2Stages * kMx2Slot);     // [512]
  float* s_b2 = s_b1 + 512;                                                             // [128] each
  float* s_gamma = s_b2 + 128;
  float* s_lnw = s_gamma + 128;
  float* s_lnb = s_lnw + 128;
  float* s_part = s_lnb + 128;                                                          // [4][128][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 4 * 256);
  uint64_t* x_full = bars + 0;        // tx
  uint64_t* x_free = bars + 1;        // 512
  uint64_t* a_full = bars + 2;        // 512
  uint64_t* hid_full = bars + 3;      // commit (per pass)
  uint64_t* out_full = bars + 4;      // commit
  uint64_t* out_free = bars + 5;      // 512
  uint64_t* sh_full = bars + 6;       // [4] 128 (per pass)
  uint64_t* w_full = bars + 10;       // [stages] tx
  uint64_t* w_empty = w_full + kMx2Stages;   // [stages] commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kMx2Stages);

  if (tid == 0) {
    mbar_init(x_full, 1); mbar_init(x_free, kMv2Workers); mbar_init(a_full, kMv2Workers); mbar_init(hid_full, 1);
    mbar_init(out_full, 1); mbar_init(out_free, kMv2Workers);
    for (int j = 0; j < 4; ++j) mbar_init(&sh_full[j], 128);
    for (int s = 0; s < kMx2Stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 512; i += kMv2Threads) s_b1[i] = (i < hidden && a.b1) ? a.b1[i] : 0.f;
  for (int i = tid; i < 128; i += kMv2Threads) {
    const bool in = i < C;
    s_b2[i] = (in && a.b2) ? a.b2[i] : 0.f;
    s_gamma[i] = (in && a.gamma) ? a.gamma[i] : 1.f;
    s_lnw[i] = in ? a.ln_w[i] : 1.f;
    s_lnb[i] = in ? a.ln_b[i] : 0.f;
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const uint32_t t_hid = tmem, t_out = tmem + 256;
  const int n_tiles = a.n_tiles;
  const int ks1 = C >> 4;                                  // K steps of fc1 (8 for C = 128)

  if (warp < 16) {
    // =============================================== workers ===============================================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;

    auto layer_norm = [&](int it) {
      mbar_wait(x_full, it & 1);
      ln_row32_to_operand<4>(sX + wg * kAv2Tile, row, wg, true, true, C, a.eps, s_lnw, s_lnb, s_part, 1, sA);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(a_full);
      mbar_arrive(x_free);
    };

    if (static_cast<int>(blockIdx.x) < n_tiles) layer_norm(0);
    int it = 0;
    uint32_t hp = 0;                                       // completions of hid_full / sh_full so far (one per pass)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      for (int p = 0; p < npass; ++p, ++hp) {
        // ---------------- GELU(it, pass p): hidden chunk 4p + wg -> fp16 operand chunk wg of fc2 ----------------
        mbar_wait(hid_full, hp & 1);
        tc_fence_after();
        const int j = 4 * p + wg;
        if (j < nch) {
          const uint32_t dst = sH + wg * kAv2Tile;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld_x32(t_hid + lane_off + wg * 64 + half * 32, v);
            tmem_ld_wait();
            const float* bp = s_b1 + j * 64 + half * 32;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float bv[8];
              lds8(bp + 8 * c, bv);
              uint32_t o[4];
              if (GELU_H2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = gelu_f16x2(pack_h2(v[8 * c + 2 * e] + bv[2 * e], v[8 * c + 2 * e + 1] + bv[2 * e + 1]));
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = pack_h2(gelu_erf(v[8 * c + 2 * e] + bv[2 * e]), gelu_erf(v[8 * c + 2 * e + 1] + bv[2 * e + 1]));
              }
              st_smem_16B(dst + sw128_offset(row, half * 4 + c), o[0], o[1], o[2], o[3]);
            }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&sh_full[wg]);
        }
      }
      // ---------------- LN(it + 1) ----------------
      if (tile + static_cast<int>(gridDim.x) < n_tiles) layer_norm(it + 1);
      // ---------------- EPI(it): coalesced (see mlp_v2_kernel) ----------------
      const int nc4 = C >> 2;
      const int ech = tid % nc4, er0 = tid / nc4, erstep = kMv2Workers / nc4;      // C = 128: 32 chunks, 16 rows per pass, 8 rows
      constexpr int kRows = 8;
      float4 xr[kRows];
      bool ok[kRows];
#pragma unroll
      for (int q = 0; q < kRows; ++q) {
        const int r = er0 + q * erstep;
        ok[q] = r < 128 && tile * 128 + r < a.n_tokens;
        if (ok[q]) xr[q] = *reinterpret_cast<const float4*>(a.x + static_cast<size_t>(tile * 128 + r) * C + ech * 4);
      }
      mbar_wait(out_full, par);
      tc_fence_after();
      const uint32_t srow = sH + static_cast<uint32_t>(row) * C * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c8 = wg + 4 * q;
        if (c8 * 8 >= C) break;
        float v[8];
        tmem_ld_x8(t_out + lane_off + c8 * 8, v);
        tmem_ld_wait();
        float bv[8], gv[8];
        lds8(s_b2 + c8 * 8, bv);
        lds8(s_gamma + c8 * 8, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (v[e] + bv[e]) * gv[e];
        const int ch = c8 * 2;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((ch ^ (row & 7)) << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((ch + 1) ^ (row & 7)) << 4)), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
      }
      tc_fence_before();
      mbar_arrive(out_free);
      named_bar_sync(2, kMv2Workers);
#pragma unroll
      for (int q = 0; q < kRows; ++q) {
        if (!ok[q]) continue;
        const int r = er0 + q * erstep;
        float4 sv;
        const uint32_t src = sH + static_cast<uint32_t>(r) * C * 4 + ((ech ^ (r & 7)) << 4);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w) : "r"(src));
        *reinterpret_cast<float4*>(a.x + static_cast<size_t>(tile * 128 + r) * C + ech * 4) =
            make_float4(xr[q].x + sv.x, xr[q].y + sv.y, xr[q].z + sv.z, xr[q].w + sv.w);
      }
      named_bar_sync(3, kMv2Workers);
    }
  } else if (warp == 16) {
    // =============================================== MMA issuer ===============================================
    if (lane == 0) {
      const uint32_t id1 = umma_idesc_f16(128, 64, 0);
      const uint32_t id2 = umma_idesc_f16(128, C, 0);
      uint32_t wc = 0, hp = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_full, it & 1);
        tc_fence_after();
        for (int p = 0; p < npass; ++p, ++hp) {
          const int nj = nch - 4 * p < 4 ? nch - 4 * p : 4;
          if (p > 0) {                                       // the hidden accumulators of the previous pass have been drained
            for (int j = 0; j < 4 && 4 * (p - 1) + j < nch; ++j) mbar_wait(&sh_full[j], (hp - 1) & 1);
            tc_fence_after();
          }
          for (int j = 0; j < nj; ++j, ++wc) {               // fc1 chunk: [128 x 64] += A [128 x C] W1_chunk^T
            const uint32_t slot = wc % kMx2Stages;
            mbar_wait(&w_full[slot], (wc / kMx2Stages) & 1);
            tc_fence_after();
            const uint32_t w1 = sW + slot * kMx2Slot;
            for (int k = 0; k < ks1; ++k) {
              const uint32_t atom = k >> 2, kk = k & 3;
              umma_f16(t_hid + j * 64, umma_desc_sw128(sA + atom * kAv2Tile + kk * 32), umma_desc_sw128(w1 + atom * 8192 + kk * 32), id1, k != 0);
            }
            umma_commit(&w_empty[slot]);
          }
          umma_commit(hid_full);
          for (int j = 0; j < nj; ++j, ++wc) {               // fc2: out += gelu(H_chunk) W2_chunk^T
            mbar_wait(&sh_full[j], hp & 1);
            if (p == 0 && j == 0 && it > 0) mbar_wait(out_free, (it - 1) & 1);
            const uint32_t slot = wc % kMx2Stages;
            mbar_wait(&w_full[slot], (wc / kMx2Stages) & 1);
            tc_fence_after();
            const uint32_t w2 = sW + slot * kMx2Slot;
            for (int k = 0; k < 4; ++k)
              umma_f16(t_out, umma_desc_sw128(sH + j * kAv2Tile + k * 32), umma_desc_sw128(w2 + k * 32), id2, (p | j | k) != 0);
            umma_commit(&w_empty[slot]);
          }
        }
        umma_commit(out_full);
      }
    }
    __syncwarp();
  } else {
    // =============================================== producer ===============================================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_x);
      const uint32_t x_bytes = 128u * C * 4, w1_bytes = static_cast<uint32_t>(C >> 6) * 8192u, w2_bytes = static_cast<uint32_t>(C) * 128u;
      uint32_t wc = 0;
      int it = 0;
      auto load_x = [&](int tile_, int it_) {
        if (it_ > 0) mbar_wait(x_free, (it_ - 1) & 1);
        mbar_arrive_expect_tx(x_full, x_bytes);
        for (int j = 0; 32 * j < C; ++j) tma_load_2d(sX + j * kAv2Tile, &tmap_x, 32 * j, tile_ * 128, x_full);
      };
      if (static_cast<int>(blockIdx.x) < n_tiles) load_x(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        // the next tile's x first (its buffer frees as soon as LN(it) is done, long before this tile's weights are consumed)
        if (tile + static_cast<int>(gridDim.x) < n_tiles) load_x(tile + gridDim.x, it + 1);
        for (int p = 0; p < npass; ++p) {
          const int nj = nch - 4 * p < 4 ? nch - 4 * p : 4;
          for (int half = 0; half < 2; ++half)
            for (int j = 0; j < nj; ++j, ++wc) {
              const uint32_t slot = wc % kMx2Stages;
              mbar_wait(&w_empty[slot], ((wc / kMx2Stages) & 1) ^ 1);
              const int chunk = 4 * p + j;
              if (half == 0) {
                mbar_arrive_expect_tx(&w_full[slot], w1_bytes);
                bulk_g2s(sm + (sW - base) + slot * kMx2Slot, a.w1p + static_cast<size_t>(chunk) * (C >> 6) * 64 * 64, w1_bytes, &w_full[slot]);
              } else {
                mbar_arrive_expect_tx(&w_full[slot], w2_bytes);
                bulk_g2s(sm + (sW - base) + slot * kMx2Slot, a.w2p + static_cast<size_t>(chunk) * C * 64, w2_bytes, &w_full[slot]);
              }
            }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace rvt
