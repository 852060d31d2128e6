// DWSConvLSTM2d (1x1 path, reference models/layers/rnn.py:36-69), persistent version for the narrow stages (C <= 64):
//     mix = W [x | h_prev] + b;  f, i, o = sigmoid(mix[0:3C]);  g = tanh(mix[3C:4C]);  c_t = f c_prev + i g;  h_t = o tanh(c_t)
//
// One CTA per SM loops over 128-token tiles; the [4C x 2C] gate weight stays resident in shared memory:
//
//   producer   one thread: the fp32 x and h_prev token tiles of tile i+1 by 2-D TMA (one buffer each, refilled as soon as the
//              workers have converted tile i)
//   MMA        one thread: gates(i) = A(i) W^T, N = 4C, into one of TWO TMEM accumulator buffers
//   workers    4 warpgroups, thread = tile row = TMEM lane.  Per iteration i:
//                build A(i+1) = fp16 [x | h_prev]  (smem fp32 -> swizzled operand, double buffered)       -- then MMA(i+1) runs while
//                epilogue(i): 8-channel groups round-robin over the warpgroups: gates + bias, non-linearities (fp32), c_prev
//                from global (prefetched), h_t / c_t fp32 stores (+ optional fp16 copy of h_t for the next stage's conv)
//
// so the MUFU-bound gate math of tile i overlaps the loads, the operand build and the MMA of tile i+1.
#pragma once
#include "attn_v2.cuh"

namespace rvt {

struct LstmV2Args {
  const float* cprev;       // [n_tokens, C] or null (zero state)
  float* hout; float* cout; // [n_tokens, C]
  __half* hout16;           // optional
  int n_tokens, C, n_tiles, has_h;
  const __half* w;          // pack_lstm_weight(cw = C): [1][KC][4C x 64]
  const float* bias;        // [4C] gate-major [f | i | o | g]
  int fast_gates;           // 1: single-MUFU gate non-linearities (gemm_fused.cuh tanh_fast / sigmoid_fast)
};

constexpr int kLv2Workers = 512;
constexpr int kLv2Threads = kLv2Workers + 64;
constexpr uint32_t kLv2XBuf = 128 * 64 * 4;
constexpr uint32_t kLv2Smem = 1024 + 3 * kLv2XBuf /*x, h, c_prev*/ + 2 * 2 * kAv2Tile /*A double buffered, 2 atoms*/ + 2 * 256 * 128 /*W*/ +
                              256 * 4 + 16 * 8 + 16;

__global__ void __launch_bounds__(kLv2Threads, 1)
lstm_v2_kernel(const __grid_constant__ LstmV2Args a, const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_h,
               const __grid_constant__ CUtensorMap tmap_c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, N4 = 4 * C, K2 = 2 * C;
  const int kc_n = (K2 + 63) >> 6;                       // K atoms of [x | h]
  const uint32_t sX = base, sH = sX + kLv2XBuf;
  const uint32_t sC = sH + kLv2XBuf;                     // c_prev tile: C/32 half tiles [128 rows x 32 fp32], 128-byte swizzle
  const uint32_t sA = sC + kLv2XBuf;                     // 2 buffers x 2 atoms
  const uint32_t sW = sA + 4 * kAv2Tile;
  float* s_bias = reinterpret_cast<float*>(sm + (sW - base) + 2 * 256 * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* xh_full = bars + 0;       // tx
  uint64_t* xh_free = bars + 1;       // 512
  uint64_t* a_full = bars + 2;        // [2] 512
  uint64_t* acc_full = bars + 4;      // [2] commit
  uint64_t* acc_free = bars + 6;      // [2] 512
  uint64_t* w_full = bars + 8;        // tx
  uint64_t* c_full = bars + 9;        // tx
  uint64_t* c_free = bars + 10;       // 512
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (tid == 0) {
    mbar_init(xh_full, 1); mbar_init(xh_free, kLv2Workers); mbar_init(w_full, 1);
    mbar_init(c_full, 1); mbar_init(c_free, kLv2Workers);
    for (int b = 0; b < 2; ++b) { mbar_init(&a_full[b], kLv2Workers); mbar_init(&acc_full[b], 1); mbar_init(&acc_free[b], kLv2Workers); }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 256; i += kLv2Threads) s_bias[i] = i < N4 ? a.bias[i] : 0.f;
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const int n_tiles = a.n_tiles;

  if (warp < 16) {
    // =============================================== workers ===============================================
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int chunks = K2 >> 3;                          // 16-byte operand chunks per row

    auto build_a = [&](int it) {                         // tile `it`: smem fp32 x | h -> fp16 A operand buffer it & 1
      mbar_wait(xh_full, it & 1);
      const uint32_t ab = sA + (it & 1) * 2 * kAv2Tile;
      for (int idx = tid; idx < 128 * chunks; idx += kLv2Workers) {
        const int r = idx / chunks, ch = idx - r * chunks;
        const int k0 = ch * 8;
        const bool is_h = k0 >= C;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!is_h || a.has_h) {
          const uint32_t src = (is_h ? sH : sX) + (static_cast<uint32_t>(r) * C + (is_h ? k0 - C : k0)) * 4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(src));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(src + 16));
        }
        st_smem_16B(ab + (k0 >> 6) * kAv2Tile + sw128_offset(r, (k0 & 63) >> 3), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                    pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[it & 1]);
      mbar_arrive(xh_free);
    };

    if (static_cast<int>(blockIdx.x) < n_tiles) build_a(0);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      if (tile + static_cast<int>(gridDim.x) < n_tiles) build_a(it + 1);
      // ---------------- epilogue(it) ----------------
      // Row-per-thread global accesses touch 32 different 128-byte lines per instruction (measured: the dominant cost of the
      // first version), so c_prev comes in through TMA (swizzled half tiles) and c_t / h_t go out through a swizzled fp32
      // staging tile -- the A-operand buffer of THIS tile, idle since its MMA finished -- from which (row, 16-byte chunk)
      // threads write whole lines.
      const bool coalesced = (C & (C - 1)) == 0 && C >= 32;
      const int tok = tile * 128 + row;
      const bool live = tok < a.n_tokens;
      const size_t rbase = static_cast<size_t>(live ? tok : 0) * C;
      float cp[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cp[q][e] = 0.f;
      }
      if (a.cprev) {
        if (coalesced) {
          mbar_wait(c_full, it & 1);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int j0 = (wg + 4 * q) * 8;
            if (j0 < C) {
              const uint32_t half = sC + (j0 >> 5) * kAv2Tile;
              const uint32_t s0 = half + sw128_offset(row, (j0 & 31) >> 2), s1 = half + sw128_offset(row, ((j0 & 31) >> 2) + 1);
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cp[q][0]), "=f"(cp[q][1]), "=f"(cp[q][2]), "=f"(cp[q][3]) : "r"(s0));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cp[q][4]), "=f"(cp[q][5]), "=f"(cp[q][6]), "=f"(cp[q][7]) : "r"(s1));
            }
          }
          mbar_arrive(c_free);
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c8 = wg + 4 * q;
            if (live && c8 * 8 < C) load8(a.cprev + rbase + c8 * 8, cp[q]);
          }
        }
      }
      mbar_wait(&acc_full[b], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem + lane_off + b * 256;
      const uint32_t stage = sA + b * 2 * kAv2Tile;            // fp32 [128 x C] staging tile (C = 64: 32 KB = the A buffer)
      float hn[2][8], cn[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c8 = wg + 4 * q;
        if (c8 * 8 >= C) break;
        const int j0 = c8 * 8;
        float f[8], ig[8], og[8], g[8];
        tmem_ld_x8(tacc + j0, f);
        tmem_ld_x8(tacc + C + j0, ig);
        tmem_ld_x8(tacc + 2 * C + j0, og);
        tmem_ld_x8(tacc + 3 * C + j0, g);
        tmem_ld_wait();
        float bv[8];
        lds8(s_bias + j0, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] += bv[e];
        lds8(s_bias + C + j0, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) ig[e] += bv[e];
        lds8(s_bias + 2 * C + j0, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) og[e] += bv[e];
        lds8(s_bias + 3 * C + j0, bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] += bv[e];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (a.fast_gates) {
            cn[q][e] = sigmoid_fast(f[e]) * cp[q][e] + sigmoid_fast(ig[e]) * tanh_fast(g[e]);
            hn[q][e] = sigmoid_fast(og[e]) * tanh_fast(cn[q][e]);
          } else {
            cn[q][e] = sigmoid_acc(f[e]) * cp[q][e] + sigmoid_acc(ig[e]) * tanh_acc(g[e]);
            hn[q][e] = sigmoid_acc(og[e]) * tanh_acc(cn[q][e]);
          }
        }
        if (!coalesced && live) {
          *reinterpret_cast<float4*>(a.cout + rbase + j0) = make_float4(cn[q][0], cn[q][1], cn[q][2], cn[q][3]);
          *reinterpret_cast<float4*>(a.cout + rbase + j0 + 4) = make_float4(cn[q][4], cn[q][5], cn[q][6], cn[q][7]);
          *reinterpret_cast<float4*>(a.hout + rbase + j0) = make_float4(hn[q][0], hn[q][1], hn[q][2], hn[q][3]);
          *reinterpret_cast<float4*>(a.hout + rbase + j0 + 4) = make_float4(hn[q][4], hn[q][5], hn[q][6], hn[q][7]);
          if (a.hout16)
            *reinterpret_cast<uint4*>(a.hout16 + rbase + j0) = make_uint4(pack_h2(hn[q][0], hn[q][1]), pack_h2(hn[q][2], hn[q][3]),
                                                                           pack_h2(hn[q][4], hn[q][5]), pack_h2(hn[q][6], hn[q][7]));
        }
      }
      if (coalesced) {
        const int nch = C >> 2;
        const int ech = tid % nch, er0 = tid / nch, erstep = kLv2Workers / nch, nrows = 128 / erstep;
        const uint32_t srow = stage + static_cast<uint32_t>(row) * C * 4;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {                  // pass 0: c_t, pass 1: h_t (+ its fp16 copy)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int ch = (wg + 4 * q) * 2;
            if (ch * 4 >= C) break;
            const float* v = pass == 0 ? cn[q] : hn[q];
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((ch ^ (row & 7)) << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((ch + 1) ^ (row & 7)) << 4)), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
          }
          named_bar_sync(1, kLv2Workers);
          for (int q = 0; q < nrows; ++q) {
            const int r = er0 + q * erstep;
            const int t = tile * 128 + r;
            if (t >= a.n_tokens) continue;
            float4 sv;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w)
                         : "r"(stage + static_cast<uint32_t>(r) * C * 4 + ((ech ^ (r & 7)) << 4)));
            const size_t off = static_cast<size_t>(t) * C + ech * 4;
            if (pass == 0) {
              *reinterpret_cast<float4*>(a.cout + off) = sv;
            } else {
              *reinterpret_cast<float4*>(a.hout + off) = sv;
              if (a.hout16) *reinterpret_cast<uint2*>(a.hout16 + off) = make_uint2(pack_h2(sv.x, sv.y), pack_h2(sv.z, sv.w));
            }
          }
          named_bar_sync(2, kLv2Workers);                       // staging reused by the next pass / as the A operand of tile it+2
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free[b]);
    }
  } else if (warp == 16) {
    // =============================================== MMA issuer ===============================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, N4, 0);
      const int ksteps = K2 >> 4;
      mbar_wait(w_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&a_full[b], (it >> 1) & 1);
        if (it >= 2) mbar_wait(&acc_free[b], ((it >> 1) - 1) & 1);      // epilogue(it - 2) has drained this accumulator buffer
        tc_fence_after();
        const uint32_t ab = sA + b * 2 * kAv2Tile;
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t atom = k >> 2, kk = k & 3;
          umma_f16(tmem + b * 256, umma_desc_sw128(ab + atom * kAv2Tile + kk * 32), umma_desc_sw128(sW + atom * (N4 * 128) + kk * 32),
                   idesc, k != 0);
        }
        umma_commit(&acc_full[b]);
      }
    }
    __syncwarp();
  } else {
    // =============================================== producer ===============================================
    if (lane == 0 && static_cast<int>(blockIdx.x) < n_tiles) {
      tma_prefetch_desc(&tmap_x);
      if (a.has_h) tma_prefetch_desc(&tmap_h);
      const uint32_t w_bytes = static_cast<uint32_t>(kc_n) * N4 * 128;
      mbar_arrive_expect_tx(w_full, w_bytes);
      bulk_g2s(sm + (sW - base), a.w, w_bytes, w_full);
      const uint32_t t_bytes = 128u * C * 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        if (it > 0) mbar_wait(xh_free, (it - 1) & 1);
        mbar_arrive_expect_tx(xh_full, a.has_h ? 2 * t_bytes : t_bytes);
        tma_load_2d(sX, &tmap_x, 0, tile * 128, xh_full);
        if (a.has_h) tma_load_2d(sH, &tmap_h, 0, tile * 128, xh_full);
        if (a.cprev && (C & (C - 1)) == 0 && C >= 32) {          // c_prev as 32-channel swizzled half tiles (row-per-thread reads)
          if (it > 0) mbar_wait(c_free, (it - 1) & 1);
          mbar_arrive_expect_tx(c_full, t_bytes);
          for (int j = 0; 32 * j < C; ++j) tma_load_2d(sC + j * kAv2Tile, &tmap_c, 32 * j, tile * 128, c_full);
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
