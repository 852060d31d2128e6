// Backward-pass building blocks of the recurrent backbone (training step, BASELINE configs[2]).
// The reference has no explicit backward code — PyTorch autograd differentiates
// models/layers/maxvit/maxvit.py and models/layers/rnn.py — so every kernel below restates the
// analytic gradient of one forward operator and cites that operator.
//
//   gemm_tn_kernel        G[i,j] += sum_m A1[m,i] * A2[m,j]   (weight gradients: contraction over tokens,
//                         tcgen05, both operands by 2-D TMA, MN-major smem tiles, split over token ranges)
//   ln_rows_any_kernel    LayerNorm rows for any C % 4 == 0 (recompute of the GEMM operand)
//   ln_bwd_kernel         LayerNorm backward (+ residual-stream accumulation, + d weight / d bias)
//   gather_cast_kernel    residual-stream gradient -> fp16 GEMM operand rows in partition order (x LayerScale)
//   attn_core_bwd_kernel  softmax(QK^T)V backward per (partition group, head)
//   lstm_gates_bwd_kernel gate non-linearities + cell update backward (rnn.py:57-67)
//   im2col_kernel / col2im_kernel   downsample-conv operand / input gradient
//   colsum_kernel         bias gradients
#pragma once
#include "gemm_fused.cuh"

namespace rvt {

// ----------------------------------------------------------------------------------------
// G[i*s_i + j*s_j] += sum_{m in split} A1[m, i0+i] * A2[m, j0+j]
// A1: fp16 [M, N1] row-major, A2: fp16 [M, N2] row-major (token-major activations / gradients).
// The contraction index m is the *row* index of both operands, i.e. the operands are "MN-major"
// for the tensor core: a TMA box of 64 tokens x 64 channels (128-byte swizzle) lands as the canonical
// MN-major SW128 atom stack (8 token rows x 128 B per 1024-B atom, SBO = 1024 B between 8-token groups,
// LBO = one box = 8192 B between 64-channel groups); instruction descriptor a_major = b_major = 1.
// `kmajor` selects the alternative operand form: A1t [N1, Mpad], A2t [N2, Mpad] (pre-transposed copies),
// plain K-major tiles exactly like gemm_fused_kernel<LD_TMA>.
// grid (ceil(N1/128), ceil(N2/BN), splits); 192 threads: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue.
// ----------------------------------------------------------------------------------------
struct TnArgs {
  int M, N1, N2, BN;
  int kc_total, kc_per_split;
  int stages, tmem_cols, kmajor;
  int epi_bulk;        // 1: rows of the accumulator tile are staged in smem and added to G with cp.reduce.async.bulk (s_j == 1)
  float* G;
  long long s_i, s_j;
  float* colsum1;      // optional: colsum1[i] += sum_m A1[m, i]  (bias gradient of the layer whose output gradient is A1)
  float* colsum2;      // optional: colsum2[j] += sum_m A2[m, j]
};

constexpr int kTnThreads = 192;
__host__ __device__ inline size_t tn_smem_bytes(int stages, int BN) {
  return 1024 + static_cast<size_t>(stages) * (16384 + static_cast<size_t>(BN) * 128) + (2 * kMaxStages + 1) * 8 + 16;
}


__global__ void __launch_bounds__(kTnThreads) gemm_tn_kernel(const __grid_constant__ TnArgs a,
                                                             const __grid_constant__ CUtensorMap tm1,
                                                             const __grid_constant__ CUtensorMap tm2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = a.stages, BN = a.BN;
  const uint32_t b_bytes = static_cast<uint32_t>(BN) * 128u;
  const uint32_t sA = base, sB = base + stages * 16384u;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + static_cast<size_t>(stages) * (16384 + b_bytes));
  uint64_t* empty = full + kMaxStages;
  uint64_t* accum = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int i0 = blockIdx.x * 128, j0 = blockIdx.y * BN;
  const int kc_lo = blockIdx.z * a.kc_per_split;
  const int kc_hi = min(a.kc_total, kc_lo + a.kc_per_split);
  const int nkc = kc_hi - kc_lo;
  if (nkc <= 0) return;

  // column sums ride along: the idle epilogue warps add up the operand tiles while they sit in shared memory (MN-major form)
  const bool cs1 = a.colsum1 != nullptr && blockIdx.y == 0 && !a.kmajor;
  const bool cs2 = a.colsum2 != nullptr && blockIdx.x == 0 && !a.kmajor;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (cs1 || cs2) ? 5 : 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // boxes that hold at least one valid channel (wholly out-of-range boxes are never loaded; the accumulator
  // rows / columns they would feed are not stored either)
  const int na = a.kmajor ? 1 : min(2, (a.N1 - i0 + 63) / 64);
  const int nb = a.kmajor ? 1 : min(BN / 64 > 0 ? BN / 64 : 1, (a.N2 - j0 + 63) / 64);

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tm1);
      tma_prefetch_desc(&tm2);
      for (int it = 0; it < nkc; ++it) {
        const int s = it % stages;
        mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
        const int m0 = (kc_lo + it) * 64;
        if (a.kmajor) {
          mbar_arrive_expect_tx(&full[s], 16384u + b_bytes);
          tma_load_2d(sA + s * 16384u, &tm1, m0, i0, &full[s]);          // box {64 tokens, 128 channels}
          tma_load_2d(sB + s * b_bytes, &tm2, m0, j0, &full[s]);         // box {64 tokens, BN channels}
        } else {
          mbar_arrive_expect_tx(&full[s], static_cast<uint32_t>(na + nb) * 8192u);
          for (int q = 0; q < na; ++q) tma_load_2d(sA + s * 16384u + q * 8192u, &tm1, i0 + q * 64, m0, &full[s]);
          for (int q = 0; q < nb; ++q) tma_load_2d(sB + s * b_bytes + q * 8192u, &tm2, j0 + q * 64, m0, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t idesc = umma_idesc_f16(128, BN, 0);
      if (!a.kmajor) idesc |= (1u << 15) | (1u << 16);                    // a_major = b_major = MN
      for (int it = 0; it < nkc; ++it) {
        const int s = it % stages;
        mbar_wait(&full[s], (it / stages) & 1);
        tc_fence_after();
        const uint32_t at = sA + s * 16384u, bt = sB + s * b_bytes;
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (it | k) != 0 ? 1u : 0u;
          if (a.kmajor)
            umma_f16(tmem_base, umma_desc_sw128(at + k * 32), umma_desc_sw128(bt + k * 32), idesc, acc);
          else
            umma_f16(tmem_base, umma_desc_sw128_mn(at + k * 2048, 8192), umma_desc_sw128_mn(bt + k * 2048, 8192), idesc, acc);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
    __syncwarp();
  } else {
    if (cs1 || cs2) {
      // thread ch of the 128 epilogue threads owns channel ch of the A tile (and of the B tile): element (token k, channel ch)
      // of an MN-major SW128 box sits at box*8192 + k*128 + (((ch & 63) >> 3) ^ (k & 7))*16 + (ch & 7)*2
      const int ch = (warp - 2) * 32 + lane;
      const uint32_t box_off = static_cast<uint32_t>(ch >> 6) * 8192u, cchunk = (ch & 63) >> 3, cbyte = (ch & 7) * 2;
      float acc1 = 0.f, acc2 = 0.f;
      const bool do1 = cs1 && (ch >> 6) < na, do2 = cs2 && ch < BN && (ch >> 6) < nb;
      for (int it = 0; it < nkc; ++it) {
        const int s = it % stages;
        mbar_wait(&full[s], (it / stages) & 1);
        const uint32_t at = sA + s * 16384u + box_off, bt = sB + s * b_bytes + box_off;
        if (do1) {
#pragma unroll 8
          for (uint32_t k = 0; k < 64; ++k) {
            unsigned short h;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(at + k * 128u + ((cchunk ^ (k & 7u)) << 4) + cbyte));
            acc1 += __half2float(__ushort_as_half(h));
          }
        }
        if (do2) {
#pragma unroll 8
          for (uint32_t k = 0; k < 64; ++k) {
            unsigned short h;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(bt + k * 128u + ((cchunk ^ (k & 7u)) << 4) + cbyte));
            acc2 += __half2float(__ushort_as_half(h));
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (do1 && i0 + ch < a.N1) atomicAdd(a.colsum1 + i0 + ch, acc1);
      if (do2 && j0 + ch < a.N2) atomicAdd(a.colsum2 + j0 + ch, acc2);
      named_bar_sync(1, 128);    // every epilogue warp is done reading the operand stages before any of them is reused as staging
    }
    const int q = warp & 3;                                  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const int i = i0 + r;
    mbar_wait(accum, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (a.epi_bulk) {
      // all MMAs have completed (accum barrier): the operand stages are free -> stage this thread's accumulator row
      // (pitch BN*4 + 16 B: conflict-free 16-byte stores) and add it to G with ONE bulk reduction per row.
      const uint32_t pitch = static_cast<uint32_t>(BN) * 4u + 16u;
      const uint32_t srow = base + static_cast<uint32_t>(r) * pitch;
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (c0 + qd * 4) * 4), "f"(v[qd * 4]), "f"(v[qd * 4 + 1]),
                       "f"(v[qd * 4 + 2]), "f"(v[qd * 4 + 3]) : "memory");
      }
      fence_proxy_async_smem();
      const int ncols = min(BN, a.N2 - j0);
      if (i < a.N1 && ncols > 0) {
        float* dst = a.G + static_cast<long long>(i) * a.s_i + j0;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(srow),
                     "r"(static_cast<uint32_t>(ncols) * 4u) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
        if (i < a.N1) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = j0 + c0 + e;
            if (j < a.N2) atomicAdd(a.G + static_cast<long long>(i) * a.s_i + static_cast<long long>(j) * a.s_j, v[e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// fp16 [M, N] (ld) -> fp16 [N, ldt] transpose (operands of the kmajor form of gemm_tn_kernel); 32x32 smem tiles
__global__ void __launch_bounds__(256) transpose_f16_kernel(const __half* __restrict__ in, int M, int N, int ld,
                                                            __half* __restrict__ out, int ldt) {
  __shared__ __half tile[32][33];
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, n = n0 + tx;
    tile[r][tx] = (m < M && n < N) ? in[static_cast<size_t>(m) * ld + n] : __float2half(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, m = m0 + tx;
    if (n < N && m < ldt) out[static_cast<size_t>(n) * ldt + m] = tile[tx][r];
  }
}

// ----------------------------------------------------------------------------------------
// LayerNorm rows, any C % 4 == 0, C <= 512; one warp per output row.
// OUT_F16: out fp16 [n_rows, ldo] in `map` order (rows without a token -> zeros); else fp32 token order.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Sub-warp row groups: a row of C channels = nv = C/4 float4s is handled by `lpr` lanes (8, 16 or 32), so narrow stages
// (C <= 64) process 2-4 rows per warp instead of idling lanes; reductions shuffle inside the group only.
__device__ __forceinline__ int ln_lanes_per_row(int nv) { return nv > 16 ? 32 : (nv > 8 ? 16 : 8); }
__device__ __forceinline__ float group_sum(float s, int lpr) {
  for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

template <bool OUT_F16, int VPL>   // VPL float4s per lane: 1 for C <= 128, 4 up to C = 512
__global__ void __launch_bounds__(256) ln_rows_any_kernel(const float* __restrict__ x, RowMap map, int n_rows, int C, int do_ln,
                                                          const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                          float eps, void* out, int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 2;
  const int lpr = ln_lanes_per_row(nv), rpw = 32 / lpr;
  const int sub = lane / lpr, sl = lane - sub * lpr;
  const int row = (blockIdx.x * 8 + warp) * rpw + sub;
  const bool in_range = row < n_rows;
  const int tok = in_range ? row_to_token(map, row) : -1;
  float4 v[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int g = sl + lpr * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < nv && tok >= 0) v[i] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(tok) * C + 4 * g);
  }
  if (do_ln) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = group_sum(s, lpr) / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (sl + lpr * i < nv) {
        const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
        ss += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
      }
    const float rstd = rsqrtf(group_sum(ss, lpr) / C + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int g = sl + lpr * i;
      if (g < nv) {
        float4 w = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ln_w != nullptr) { w = __ldg(reinterpret_cast<const float4*>(ln_w) + g); b = __ldg(reinterpret_cast<const float4*>(ln_b) + g); }
        v[i].x = (v[i].x - mean) * rstd * w.x + b.x; v[i].y = (v[i].y - mean) * rstd * w.y + b.y;
        v[i].z = (v[i].z - mean) * rstd * w.z + b.z; v[i].w = (v[i].w - mean) * rstd * w.w + b.w;
      }
    }
  }
  if (!in_range) return;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int g = sl + lpr * i;
    if (g >= nv) continue;
    if (OUT_F16) {
      const float4 t = tok >= 0 ? v[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + static_cast<size_t>(row) * ldo + 4 * g) =
          make_uint2(pack_h2(t.x, t.y), pack_h2(t.z, t.w));
    } else if (tok >= 0) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + static_cast<size_t>(tok) * ldo + 4 * g) = v[i];
    }
  }
}

// ----------------------------------------------------------------------------------------
// LayerNorm backward (F.layer_norm over C, eps; norm1 / norm2 maxvit.py:234,241 and the downsample norm :172).
//   y = xhat * w + b,  xhat = (x - mean) * rstd.   Given dy:
//   dw += sum_rows dy * xhat;  db += sum_rows dy;  g = dy * w;
//   dx = rstd * (g - mean_C(g) - xhat * mean_C(g * xhat))
// DY_F16: dy is fp16 [n_rows, lddy] in `map` order (gradient of a GEMM operand); else fp32 token order [n_tokens, C].
// Outputs: dres (fp32 token order) += dx when dres != null;  dx16 (fp16 [n_rows, lddx], map order) = dx when != null.
// do_ln == 0: the "norm" was Identity (skip_first_norm): dx = dy.
// Persistent CTAs (grid-stride over rows) so the per-channel dw/db partials live in registers and are flushed
// with one atomicAdd per channel per CTA.
// ----------------------------------------------------------------------------------------
template <bool DY_F16, int VPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const void* __restrict__ dy, int lddy, RowMap map,
                                                     int n_rows, int C, int do_ln, const float* __restrict__ ln_w, float eps,
                                                     float* dres, __half* dx16, int lddx, float* dw_acc, float* db_acc) {
  __shared__ float s_acc[4096];                 // [8 warps * rows-per-warp][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = C >> 2;
  const int lpr = ln_lanes_per_row(nv), rpw = 32 / lpr;
  const int sub = lane / lpr, sl = lane - sub * lpr;
  float4 aw[VPL], ab[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) { aw[i] = make_float4(0.f, 0.f, 0.f, 0.f); ab[i] = aw[i]; }
  for (int base_row = (blockIdx.x * 8 + warp) * rpw; base_row < n_rows; base_row += gridDim.x * 8 * rpw) {
    const int row = base_row + sub;
    const bool in_range = row < n_rows;
    const int tok = in_range ? row_to_token(map, row) : -1;
    const bool act = tok >= 0;
    float4 xv[VPL], gv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int g = sl + lpr * i;
      xv[i] = make_float4(0.f, 0.f, 0.f, 0.f); gv[i] = xv[i];
      if (g < nv && act) {
        if (do_ln) xv[i] = *reinterpret_cast<const float4*>(x + static_cast<size_t>(tok) * C + 4 * g);
        if (DY_F16) {
          const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(dy) + static_cast<size_t>(row) * lddy + 4 * g);
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
          gv[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
        } else {
          gv[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + static_cast<size_t>(tok) * lddy + 4 * g);
        }
      }
    }
    if (do_ln) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
      const float mean = group_sum(s, lpr) / C;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (sl + lpr * i < nv) {
          xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
          ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
        }
      const float rstd = rsqrtf(group_sum(ss, lpr) / C + eps);
      float sg = 0.f, sgx = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int g = sl + lpr * i;
        if (g < nv) {
          xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;      // xhat
          if (act) {
            aw[i].x += gv[i].x * xv[i].x; aw[i].y += gv[i].y * xv[i].y; aw[i].z += gv[i].z * xv[i].z; aw[i].w += gv[i].w * xv[i].w;
            ab[i].x += gv[i].x; ab[i].y += gv[i].y; ab[i].z += gv[i].z; ab[i].w += gv[i].w;
          }
          if (ln_w != nullptr) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(ln_w) + g);
            gv[i].x *= w.x; gv[i].y *= w.y; gv[i].z *= w.z; gv[i].w *= w.w;
          }
          sg += gv[i].x + gv[i].y + gv[i].z + gv[i].w;
          sgx += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
        }
      }
      const float mg = group_sum(sg, lpr) / C, mgx = group_sum(sgx, lpr) / C;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        gv[i].x = rstd * (gv[i].x - mg - xv[i].x * mgx); gv[i].y = rstd * (gv[i].y - mg - xv[i].y * mgx);
        gv[i].z = rstd * (gv[i].z - mg - xv[i].z * mgx); gv[i].w = rstd * (gv[i].w - mg - xv[i].w * mgx);
      }
    }
    if (!in_range) continue;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int g = sl + lpr * i;
      if (g >= nv) continue;
      if (dres && act) {
        float4* p = reinterpret_cast<float4*>(dres + static_cast<size_t>(tok) * C + 4 * g);
        float4 r = *p;
        r.x += gv[i].x; r.y += gv[i].y; r.z += gv[i].z; r.w += gv[i].w;
        *p = r;
      }
      if (dx16) {
        const float4 t = act ? gv[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<uint2*>(dx16 + static_cast<size_t>(row) * lddx + 4 * g) = make_uint2(pack_h2(t.x, t.y), pack_h2(t.z, t.w));
      }
    }
  }
  if (!do_ln || (dw_acc == nullptr && db_acc == nullptr)) return;
  // flush dw, then db: (8 warps x rows-per-warp) partial rows -> smem -> one atomicAdd per channel
  const int nparts = 8 * rpw;
  for (int pass = 0; pass < 2; ++pass) {
    float* acc = pass == 0 ? dw_acc : db_acc;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int g = sl + lpr * i;
      if (g < nv) *reinterpret_cast<float4*>(&s_acc[(warp * rpw + sub) * C + 4 * g]) = pass == 0 ? aw[i] : ab[i];
    }
    __syncthreads();
    if (acc != nullptr)
      for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int w = 0; w < nparts; ++w) s += s_acc[w * C + c];
        atomicAdd(acc + c, s);
      }
  }
}

// ----------------------------------------------------------------------------------------
// Residual-stream gradient -> fp16 GEMM operand rows in `map` order:
//   d0[row] = dres[token(row)]  (unscaled: weight-gradient operand),  d1[row] = gamma * dres[token(row)]
//   (LayerScale backward, maxvit.py:45-53; the operand of the data-gradient GEMMs).  Rows without a token: zeros.
// One thread per 8 channels.  d1 may be null (then only d0), gamma null => d1 = d0 values.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_cast_kernel(const float* __restrict__ dres, RowMap map, int n_rows, int C,
                                                          const float* __restrict__ gamma, __half* __restrict__ d0,
                                                          __half* __restrict__ d1) {
  const int per_row = C >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_rows) * per_row) return;
  const int row = static_cast<int>(idx / per_row), c0 = static_cast<int>(idx - static_cast<long long>(row) * per_row) * 8;
  const int tok = row_to_token(map, row);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (tok >= 0) load8(dres + static_cast<size_t>(tok) * C + c0, v);
  if (d0)
    *reinterpret_cast<uint4*>(d0 + static_cast<size_t>(row) * C + c0) =
        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  if (d1) {
    if (gamma) {
      float g[8];
      load8(gamma + c0, g);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= g[e];
    }
    *reinterpret_cast<uint4*>(d1 + static_cast<size_t>(row) * C + c0) =
        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  }
}

// ----------------------------------------------------------------------------------------
// Attention core backward for one (partition group, head)  (SelfAttentionCl.forward maxvit.py:349-352):
//   S = scale * Q K^T, A = softmax(S), O = A V.   Given dO:
//   dV = A^T dO;  dA = dO V^T;  dS = A * (dA - rowsum(dA * A));  dQ = scale * dS K;  dK = scale * dS^T Q.
// qkv / dqkv: fp16 [rows, 3C], per-head interleaved [q_h | k_h | v_h]; dO: fp16 [rows, C] head-major.
// Only the P real tokens of the group take part (padding rows of dqkv are written as zeros).
// fp32 SIMT in shared memory: the tiles are tiny (P <= 128, dh <= 32).
// ----------------------------------------------------------------------------------------
struct AttnBwdArgs {
  const __half* qkv; const __half* o; const __half* dout; __half* dqkv;
  int C, dh, nh, P, rows_per_win, n_groups;
  float scale;
};
constexpr int kAbPitch = 33;   // dh <= 32
__host__ __device__ inline size_t attn_bwd_smem_bytes(int P) {
  return (static_cast<size_t>(4) * P * kAbPitch + static_cast<size_t>(P) * (P + 1) + P) * sizeof(float);
}

// C(i, j) = sum_k A(i, k) * B(j, k) on TM x TN register tiles (operands in shared memory through accessors)
template <int TM, int TN, class FA, class FB, class FC>
__device__ __forceinline__ void tile_mm(int M, int N, int K, FA fa, FB fb, FC fc, int tid, int nthreads) {
  const int tm = (M + TM - 1) / TM, tn = (N + TN - 1) / TN;
  for (int t = tid; t < tm * tn; t += nthreads) {
    const int i0 = (t / tn) * TM, j0 = (t - (t / tn) * tn) * TN;
    int ir[TM], jr[TN];
#pragma unroll
    for (int r = 0; r < TM; ++r) ir[r] = min(i0 + r, M - 1);
#pragma unroll
    for (int q = 0; q < TN; ++q) jr[q] = min(j0 + q, N - 1);
    float acc[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int q = 0; q < TN; ++q) acc[r][q] = 0.f;
    for (int k = 0; k < K; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int r = 0; r < TM; ++r) av[r] = fa(ir[r], k);
#pragma unroll
      for (int q = 0; q < TN; ++q) bv[q] = fb(jr[q], k);
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int q = 0; q < TN; ++q) acc[r][q] = fmaf(av[r], bv[q], acc[r][q]);
    }
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int q = 0; q < TN; ++q)
        if (i0 + r < M && j0 + q < N) fc(i0 + r, j0 + q, acc[r][q]);
  }
}

__global__ void __launch_bounds__(256) attn_core_bwd_kernel(const __grid_constant__ AttnBwdArgs a) {
  extern __shared__ float sf[];
  const int P = a.P, dh = a.dh, SP = P + 1;
  float* sQ = sf;
  float* sK = sQ + P * kAbPitch;
  float* sV = sK + P * kAbPitch;
  float* sD = sV + P * kAbPitch;      // dO
  float* sS = sD + P * kAbPitch;      // S -> A -> dS
  float* sDelta = sS + P * SP;        // rowsum(dO * O) = rowsum(dA * A)
  const int g = blockIdx.x, hd = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t row0 = static_cast<size_t>(g) * a.rows_per_win;
  const int C3 = 3 * a.C;
  const float scale = a.scale;

  for (int idx = tid; idx < P * dh; idx += 256) {
    const int i = idx / dh, d = idx - i * dh;
    const __half* rp = a.qkv + (row0 + i) * C3 + hd * 3 * dh + d;
    sQ[i * kAbPitch + d] = __half2float(rp[0]);
    sK[i * kAbPitch + d] = __half2float(rp[dh]);
    sV[i * kAbPitch + d] = __half2float(rp[2 * dh]);
    sD[i * kAbPitch + d] = __half2float(a.dout[(row0 + i) * a.C + hd * dh + d]);
  }
  __syncthreads();
  for (int i = warp; i < P; i += 8) {
    float s = 0.f;
    if (lane < dh) s = sD[i * kAbPitch + lane] * __half2float(a.o[(row0 + i) * a.C + hd * dh + lane]);
    s = warp_sum(s);
    if (lane == 0) sDelta[i] = s;
  }
  // S = scale * Q K^T
  tile_mm<4, 4>(P, P, dh, [&](int i, int k) { return sQ[i * kAbPitch + k]; }, [&](int j, int k) { return sK[j * kAbPitch + k]; },
                [&](int i, int j, float v) { sS[i * SP + j] = v * scale; }, tid, 256);
  __syncthreads();
  // row softmax
  for (int i = warp; i < P; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < P; j += 32) mx = fmaxf(mx, sS[i * SP + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < P; j += 32) { const float e = __expf(sS[i * SP + j] - mx); sS[i * SP + j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < P; j += 32) sS[i * SP + j] *= inv;
  }
  __syncthreads();
  // dV[j][d] = sum_i A[i][j] dO[i][d]
  tile_mm<2, 4>(P, dh, P, [&](int j, int i) { return sS[i * SP + j]; }, [&](int d, int i) { return sD[i * kAbPitch + d]; },
                [&](int j, int d, float v) { a.dqkv[(row0 + j) * C3 + hd * 3 * dh + 2 * dh + d] = __float2half_rn(v); }, tid, 256);
  __syncthreads();
  // dS[i][j] = A[i][j] * (dA[i][j] - delta_i) * scale,  dA[i][j] = sum_d dO[i][d] V[j][d]   (in place over A)
  tile_mm<4, 4>(P, P, dh, [&](int i, int k) { return sD[i * kAbPitch + k]; }, [&](int j, int k) { return sV[j * kAbPitch + k]; },
                [&](int i, int j, float v) { sS[i * SP + j] = sS[i * SP + j] * (v - sDelta[i]) * scale; }, tid, 256);
  __syncthreads();
  // dQ[i][d] = sum_j dS[i][j] K[j][d];  dK[j][d] = sum_i dS[i][j] Q[i][d]
  tile_mm<2, 4>(P, dh, P, [&](int i, int j) { return sS[i * SP + j]; }, [&](int d, int j) { return sK[j * kAbPitch + d]; },
                [&](int i, int d, float v) { a.dqkv[(row0 + i) * C3 + hd * 3 * dh + d] = __float2half_rn(v); }, tid, 256);
  tile_mm<2, 4>(P, dh, P, [&](int j, int i) { return sS[i * SP + j]; }, [&](int d, int i) { return sQ[i * kAbPitch + d]; },
                [&](int j, int d, float v) { a.dqkv[(row0 + j) * C3 + hd * 3 * dh + dh + d] = __float2half_rn(v); }, tid, 256);
  // padding rows of this head: zeros
  const int pad = a.rows_per_win - P;
  for (int idx = tid; idx < pad * 3 * dh; idx += 256) {
    const int i = P + idx / (3 * dh), d = idx % (3 * dh);
    a.dqkv[(row0 + i) * C3 + hd * 3 * dh + d] = __float2half_rn(0.f);
  }
}

// ----------------------------------------------------------------------------------------
// Attention core backward on the tensor cores: one CTA per (128-row tile, head), thread t = tile row t.
//   S = Q K^T and dA = dO V^T            (K-major operands: rows = tokens, 64 B of head dim, two tensors per 128-B row)
//   A = softmax(S) over the row's own partition group, dS = A * (dA - delta) * scale, delta = rowsum(dO * O)
//   dV = A^T dO, dK = dS^T Q             (contraction over QUERY rows: A / dS and dO / Q are consumed in place as MN-major
//                                         operands — the row index is the K dimension — so nothing is transposed)
//   dQ = dS K                            (dS K-major, K as MN-major B)
// TMEM: S [0,128), dA [128,256); afterwards dV / dK / dQ accumulator blocks at [0,64), [64,128), [128,192).  Block-diagonal structure (two 64-row
// groups per tile) and padding rows / keys are handled by zeros in the A / dS tiles.
// pair_tiles = 1: Q|K share one SW128 tile (chunks 0-3 / 4-7 of a row) and V|dO another; 0: four separate tiles.
// ----------------------------------------------------------------------------------------
struct AttnBwdTcArgs {
  const __half* qkv; const __half* o; const __half* dout; __half* dqkv;
  int C, dh, nh, P, rows_per_win, n_groups, nkeys, pair_tiles;
  float scale, scale_log2e;
};
__host__ __device__ inline size_t attn_bwd_tc_smem_bytes(int pair_tiles) {
  return 1024 + (pair_tiles ? 2 : 4) * 16384 + 2 * 32768 + 64;
}

__global__ void __launch_bounds__(128) attn_core_bwd_tc_kernel(const __grid_constant__ AttnBwdTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const int nt = a.pair_tiles ? 2 : 4;
  // operand homes: tile base + byte offset inside the 128-byte row
  const uint32_t tQ = base, oQ = 0;
  const uint32_t tK = a.pair_tiles ? base : base + 16384, oK = a.pair_tiles ? 64 : 0;
  const uint32_t tV = a.pair_tiles ? base + 16384 : base + 32768, oV = 0;
  const uint32_t tD = a.pair_tiles ? base + 16384 : base + 49152, oD = a.pair_tiles ? 64 : 0;
  const uint32_t sP = base + nt * 16384, sDS = sP + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + nt * 16384 + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int t = threadIdx.x, warp = t >> 5;
  const int mt = blockIdx.x, hd = blockIdx.y;
  const int dh = a.dh, P = a.P, rpw = a.rows_per_win, nkeys = a.nkeys;
  const int C3 = 3 * a.C;

  if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  // zero the A / dS tiles (off-diagonal blocks, padding rows and keys stay zero) and the operand tiles' unused chunks
  for (int i = t; i < (nt * 16384 + 65536) / 16; i += 128) st_smem_16B(base + i * 16, 0u, 0u, 0u, 0u);
  __syncthreads();

  // ---- stage row t of Q, K, V, dO; delta_t = sum_d dO[t][d] * O[t][d] ----
  const size_t grow = static_cast<size_t>(mt) * 128 + t;
  const __half* qrow = a.qkv + grow * C3 + hd * 3 * dh;
  const __half* drow = a.dout + grow * a.C + hd * dh;
  const __half* orow = a.o + grow * a.C + hd * dh;
  float delta = 0.f;
  for (int c = 0; c * 8 < dh; ++c) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(qrow + c * 8));
    const uint4 k = __ldg(reinterpret_cast<const uint4*>(qrow + dh + c * 8));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(qrow + 2 * dh + c * 8));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(drow + c * 8));
    const uint4 ov = __ldg(reinterpret_cast<const uint4*>(orow + c * 8));
    st_smem_16B(tQ + sw128_offset(t, (oQ >> 4) + c), q.x, q.y, q.z, q.w);
    st_smem_16B(tK + sw128_offset(t, (oK >> 4) + c), k.x, k.y, k.z, k.w);
    st_smem_16B(tV + sw128_offset(t, (oV >> 4) + c), v.x, v.y, v.z, v.w);
    st_smem_16B(tD + sw128_offset(t, (oD >> 4) + c), d.x, d.y, d.z, d.w);
    const __half2* dh2 = reinterpret_cast<const __half2*>(&d);
    const __half2* oh2 = reinterpret_cast<const __half2*>(&ov);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f0 = __half22float2(dh2[e]), f1 = __half22float2(oh2[e]);
      delta = fmaf(f0.x, f1.x, fmaf(f0.y, f1.y, delta));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_s = tmem, t_da = tmem + 128;

  if (t == 0) {
    const uint32_t idesc = umma_idesc_f16(128, nkeys, 0);
    for (int k = 0; k < 2; ++k) {                               // head dim padded to 32 = 2 K steps
      umma_f16(t_s, umma_desc_sw128(tQ + oQ + k * 32), umma_desc_sw128(tK + oK + k * 32), idesc, k != 0);
      umma_f16(t_da, umma_desc_sw128(tD + oD + k * 32), umma_desc_sw128(tV + oV + k * 32), idesc, k != 0);
    }
    umma_commit(&bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();

  // ---- row t: softmax over its own group's keys, dS ----
  const uint32_t trow = static_cast<uint32_t>(warp * 32) << 16;
  const int key_lo = (t / rpw) * rpw;
  const bool row_valid = (t - key_lo) < P && static_cast<int>(grow / rpw) < a.n_groups;
  // NOTE tcgen05.ld is warp-collective (.sync.aligned): every lane executes every load, only the arithmetic / the stores
  // are predicated (a tile's padding rows share warps with real rows).
  {
    float mx = -INFINITY;
    for (int c0 = 0; c0 < P; c0 += 16) {
      float v[16];
      tmem_ld_x16(t_s + trow + key_lo + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (c0 + e < P) mx = fmaxf(mx, v[e]);
    }
    float sum = 0.f;
    for (int c0 = 0; c0 < P; c0 += 16) {
      float v[16];
      tmem_ld_x16(t_s + trow + key_lo + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (c0 + e < P) sum += ex2_approx((v[e] - mx) * a.scale_log2e);
    }
    const float inv = 1.0f / sum;
    for (int c0 = 0; c0 < P; c0 += 16) {
      float v[16], g[16];
      tmem_ld_x16(t_s + trow + key_lo + c0, v);
      tmem_ld_x16(t_da + trow + key_lo + c0, g);
      tmem_ld_wait();
      if (row_valid) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float p = 0.f, ds = 0.f;
          if (c0 + e < P) {
            p = ex2_approx((v[e] - mx) * a.scale_log2e) * inv;
            ds = p * (g[e] - delta) * a.scale;
          }
          v[e] = p; g[e] = ds;
        }
        const int key = key_lo + c0;
        const uint32_t off = static_cast<uint32_t>(key >> 6) * 16384u, ch = (key & 63) >> 3;
        st_smem_16B(sP + off + sw128_offset(t, ch), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        st_smem_16B(sP + off + sw128_offset(t, ch + 1), pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
        st_smem_16B(sDS + off + sw128_offset(t, ch), pack_h2(g[0], g[1]), pack_h2(g[2], g[3]), pack_h2(g[4], g[5]), pack_h2(g[6], g[7]));
        st_smem_16B(sDS + off + sw128_offset(t, ch + 1), pack_h2(g[8], g[9]), pack_h2(g[10], g[11]), pack_h2(g[12], g[13]), pack_h2(g[14], g[15]));
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // MN-major B operands: with paired tiles the B operand is the whole 128-byte row (N = 64 = one full MN-major atom, the
  // layout gemm_tn_kernel uses): the wanted 32 head-dim columns are the half of the accumulator block that belongs to the
  // tensor we need, the other half (the row's other tensor) is ignored.  Separate tiles use N = 32 at offset 0.
  const int nB = a.pair_tiles ? 64 : 32;
  const uint32_t c_dv = 0 + (a.pair_tiles ? 32 : 0);        // dO is the second half of [V | dO]
  const uint32_t c_dk = 64 + 0;                             // Q is the first half of [Q | K]
  const uint32_t c_dq = 128 + (a.pair_tiles ? 32 : 0);      // K is the second half of [Q | K]
  if (t == 0) {
    const uint32_t id_mn = umma_idesc_f16(128, nB, 0) | (1u << 15) | (1u << 16);     // A, B both MN-major
    const uint32_t id_kmn = umma_idesc_f16(128, nB, 0) | (1u << 16);                 // A K-major, B MN-major
    for (int k = 0; k < 8; ++k) {                               // contraction over the 128 query rows, 16 per step
      umma_f16(tmem + 0, umma_desc_sw128_mn(sP + k * 2048, 16384), umma_desc_sw128_mn(tD + k * 2048, 16384), id_mn, k != 0);
      umma_f16(tmem + 64, umma_desc_sw128_mn(sDS + k * 2048, 16384), umma_desc_sw128_mn(tQ + k * 2048, 16384), id_mn, k != 0);
    }
    for (int kk = 0; kk < nkeys / 16; ++kk) {                   // contraction over the keys
      const uint32_t atom = kk >> 2, ks = kk & 3;
      umma_f16(tmem + 128, umma_desc_sw128(sDS + atom * 16384 + ks * 32), umma_desc_sw128_mn(tK + kk * 2048, 16384), id_kmn, kk != 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  {
    float dv[32], dk[32], dq[32];
    tmem_ld_x32(tmem + trow + c_dv, dv);
    tmem_ld_x32(tmem + trow + c_dk, dk);
    tmem_ld_x32(tmem + trow + c_dq, dq);
    tmem_ld_wait();
    __half* orow2 = a.dqkv + grow * C3 + hd * 3 * dh;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c * 8 >= dh) break;
      const float* p;
      p = dq + c * 8;
      *reinterpret_cast<uint4*>(orow2 + c * 8) = make_uint4(pack_h2(p[0], p[1]), pack_h2(p[2], p[3]), pack_h2(p[4], p[5]), pack_h2(p[6], p[7]));
      p = dk + c * 8;
      *reinterpret_cast<uint4*>(orow2 + dh + c * 8) = make_uint4(pack_h2(p[0], p[1]), pack_h2(p[2], p[3]), pack_h2(p[4], p[5]), pack_h2(p[6], p[7]));
      p = dv + c * 8;
      *reinterpret_cast<uint4*>(orow2 + 2 * dh + c * 8) = make_uint4(pack_h2(p[0], p[1]), pack_h2(p[2], p[3]), pack_h2(p[4], p[5]), pack_h2(p[6], p[7]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ----------------------------------------------------------------------------------------
// Conv-LSTM gate backward (rnn.py:57-67): with activated gates f, i, o, g, c_t = f c_{t-1} + i g, h_t = o tanh(c_t):
//   dc = dc_t + dh_t * o * (1 - tanh(c_t)^2);  dc_{t-1} = dc * f;
//   dpre = [dc * c_{t-1} * f(1-f) | dc * g * i(1-i) | dh_t * tanh(c_t) * o(1-o) | dc * i * (1 - g^2)]
// gates: fp16 [N, 4C] ([f|i|o|g]); dpre: fp16 [N, 4C] (same order = rows of conv1x1.weight). 4 channels per thread.
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_gates_bwd_kernel(const __half* __restrict__ gates, const float* __restrict__ c_prev,
                                                             const float* __restrict__ c_new, const float* __restrict__ dh,
                                                             const float* __restrict__ dc_in, long long n_tokens, int C,
                                                             __half* __restrict__ dpre, float* __restrict__ dc_prev) {
  const int per_tok = C >> 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_tokens * per_tok) return;
  const long long tok = idx / per_tok;
  const int c0 = static_cast<int>(idx - tok * per_tok) * 4;
  const size_t off = static_cast<size_t>(tok) * C + c0;
  const __half* gp = gates + static_cast<size_t>(tok) * 4 * C + c0;
  float f[4], ig[4], og[4], gg[4];
  auto ld4 = [](const __half* p, float* o) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  };
  ld4(gp, f); ld4(gp + C, ig); ld4(gp + 2 * C, og); ld4(gp + 3 * C, gg);
  const float4 cn = *reinterpret_cast<const float4*>(c_new + off);
  float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), dci = cp, dhv = cp;
  if (c_prev) cp = *reinterpret_cast<const float4*>(c_prev + off);
  if (dc_in) dci = *reinterpret_cast<const float4*>(dc_in + off);
  if (dh) dhv = *reinterpret_cast<const float4*>(dh + off);
  const float cnv[4] = {cn.x, cn.y, cn.z, cn.w}, cpv[4] = {cp.x, cp.y, cp.z, cp.w};
  const float dcv[4] = {dci.x, dci.y, dci.z, dci.w}, dhh[4] = {dhv.x, dhv.y, dhv.z, dhv.w};
  float pf[4], pi[4], po[4], pg[4], dcp[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float tc = tanhf(cnv[e]);
    const float dc = dcv[e] + dhh[e] * og[e] * (1.0f - tc * tc);
    dcp[e] = dc * f[e];
    pf[e] = dc * cpv[e] * f[e] * (1.0f - f[e]);
    pi[e] = dc * gg[e] * ig[e] * (1.0f - ig[e]);
    po[e] = dhh[e] * tc * og[e] * (1.0f - og[e]);
    pg[e] = dc * ig[e] * (1.0f - gg[e] * gg[e]);
  }
  __half* dp = dpre + static_cast<size_t>(tok) * 4 * C + c0;
  *reinterpret_cast<uint2*>(dp) = make_uint2(pack_h2(pf[0], pf[1]), pack_h2(pf[2], pf[3]));
  *reinterpret_cast<uint2*>(dp + C) = make_uint2(pack_h2(pi[0], pi[1]), pack_h2(pi[2], pi[3]));
  *reinterpret_cast<uint2*>(dp + 2 * C) = make_uint2(pack_h2(po[0], po[1]), pack_h2(po[2], po[3]));
  *reinterpret_cast<uint2*>(dp + 3 * C) = make_uint2(pack_h2(pg[0], pg[1]), pack_h2(pg[2], pg[3]));
  if (dc_prev) *reinterpret_cast<float4*>(dc_prev + off) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
}

// ----------------------------------------------------------------------------------------
// Downsample conv (maxvit.py:166-175) operand / input-gradient helpers.  K order: (ky, kx, ci) for channels-last inputs,
// (ci, ky, kx) — the conv weight's own order — for NCHW inputs (the stem), so consecutive threads read consecutive
// pixels either way; col rows = output tokens, ldc = round_up(K, 8) (pad columns are zeros).
// im2col: in NCHW (u8 / f32 / f16) or NHWC (f32 / f16); rows / cols of the virtual input beyond (Hin, Win) are zero.
// col2im: d_in[b, iy, ix, ci] = sum over the taps that read this pixel of dcol[(b, oy, ox), (ky, kx, ci)]  (gather form,
// deterministic); d_in fp32 NHWC.
// ----------------------------------------------------------------------------------------
struct ConvGeom { int B, Cin, Hin, Win, KS, stride, pad, Hout, Wout, K, ldc, in_dtype, in_nchw; };

__global__ void __launch_bounds__(256) im2col_kernel(const void* __restrict__ in, ConvGeom g, __half* __restrict__ col) {
  const int half_ld = g.ldc >> 1;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_tok = static_cast<long long>(g.B) * g.Hout * g.Wout;
  if (idx >= n_tok * half_ld) return;
  const long long tok = idx / half_ld;
  const int k0 = static_cast<int>(idx - tok * half_ld) * 2;
  const int hw = g.Hout * g.Wout;
  const int b = static_cast<int>(tok / hw), rem = static_cast<int>(tok - static_cast<long long>(b) * hw);
  const int oy = rem / g.Wout, ox = rem - oy * g.Wout;
  float v[2] = {0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = k0 + e;
    if (k >= g.K) continue;
    int ci, ky, kx;
    if (g.in_nchw) { const int kk = g.KS * g.KS; ci = k / kk; const int rem2 = k - ci * kk; ky = rem2 / g.KS; kx = rem2 - ky * g.KS; }
    else { const int tap = k / g.Cin; ci = k - tap * g.Cin; ky = tap / g.KS; kx = tap - ky * g.KS; }
    const int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
    if (iy < 0 || iy >= g.Hin || ix < 0 || ix >= g.Win) continue;
    const size_t off = g.in_nchw ? ((static_cast<size_t>(b) * g.Cin + ci) * g.Hin + iy) * g.Win + ix
                                 : ((static_cast<size_t>(b) * g.Hin + iy) * g.Win + ix) * g.Cin + ci;
    if (g.in_dtype == 1) v[e] = static_cast<float>(__ldg(reinterpret_cast<const uint8_t*>(in) + off));
    else if (g.in_dtype == 2) v[e] = __half2float(__ldg(reinterpret_cast<const __half*>(in) + off));
    else v[e] = __ldg(reinterpret_cast<const float*>(in) + off);
  }
  *reinterpret_cast<uint32_t*>(col + static_cast<size_t>(tok) * g.ldc + k0) = pack_h2(v[0], v[1]);
}

// Channels-last fast path (Cin % 8 == 0, fp16 or fp32 input): one thread per (token, tap, 8-channel group) = one 16-byte
// store; K = KS*KS*Cin = ldc.
__global__ void __launch_bounds__(256) im2col_nhwc8_kernel(const void* __restrict__ in, ConvGeom g, __half* __restrict__ col) {
  const int cg = g.Cin >> 3, per_tok = g.KS * g.KS * cg;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_tok = static_cast<long long>(g.B) * g.Hout * g.Wout;
  if (idx >= n_tok * per_tok) return;
  const long long tok = idx / per_tok;
  const int rem0 = static_cast<int>(idx - tok * per_tok);
  const int tap = rem0 / cg, c0 = (rem0 - tap * cg) * 8;
  const int ky = tap / g.KS, kx = tap - ky * g.KS;
  const int hw = g.Hout * g.Wout;
  const int b = static_cast<int>(tok / hw), rem = static_cast<int>(tok - static_cast<long long>(b) * hw);
  const int oy = rem / g.Wout, ox = rem - oy * g.Wout;
  const int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (iy >= 0 && iy < g.Hin && ix >= 0 && ix < g.Win) {
    const size_t off = ((static_cast<size_t>(b) * g.Hin + iy) * g.Win + ix) * g.Cin + c0;
    if (g.in_dtype == 2) {
      o = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(in) + off));
    } else {
      float v[8];
      load8(reinterpret_cast<const float*>(in) + off, v);
      o = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
  }
  *reinterpret_cast<uint4*>(col + static_cast<size_t>(tok) * g.ldc + tap * g.Cin + c0) = o;
}

// NCHW (u8 / f32 / f16) -> channels-last fp16 [B, H, W, Cp], channels [C, Cp) zero (Cp % 8 == 0): lets the stem use the
// vectorised im2col above.  One thread per pixel: reads are coalesced along W per channel, the Cp-channel row is written
// as 16-byte vectors.
__global__ void __launch_bounds__(256) nchw_to_nhwc_f16_kernel(const void* __restrict__ in, int in_dtype, int B, int C, int H, int W,
                                                               int Cp, __half* __restrict__ out) {
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long hw = static_cast<long long>(H) * W;
  if (pix >= B * hw) return;
  const long long b = pix / hw, r = pix - b * hw;
  __half* op = out + pix * Cp;
  for (int c0 = 0; c0 < Cp; c0 += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      v[e] = 0.f;
      if (c < C) {
        const size_t off = (static_cast<size_t>(b) * C + c) * hw + r;
        if (in_dtype == 1) v[e] = static_cast<float>(__ldg(reinterpret_cast<const uint8_t*>(in) + off));
        else if (in_dtype == 2) v[e] = __half2float(__ldg(reinterpret_cast<const __half*>(in) + off));
        else v[e] = __ldg(reinterpret_cast<const float*>(in) + off);
      }
    }
    *reinterpret_cast<uint4*>(op + c0) = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
  }
}

// NCHW input (the stem): one thread per (token, ci, ky) writes the KS consecutive columns (ci, ky, 0..KS-1) from KS
// consecutive pixels of one input row; the last thread slot of a token zero-fills the pad columns [K, ldc).
__global__ void __launch_bounds__(256) im2col_nchw_kernel(const void* __restrict__ in, ConvGeom g, __half* __restrict__ col) {
  const int per_tok = g.Cin * g.KS + 1;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_tok = static_cast<long long>(g.B) * g.Hout * g.Wout;
  if (idx >= n_tok * per_tok) return;
  const long long tok = idx / per_tok;
  const int slot = static_cast<int>(idx - tok * per_tok);
  __half* crow = col + static_cast<size_t>(tok) * g.ldc;
  if (slot == per_tok - 1) {
    for (int k = g.K; k < g.ldc; ++k) crow[k] = __float2half(0.f);
    return;
  }
  const int ci = slot / g.KS, ky = slot - ci * g.KS;
  const int hw = g.Hout * g.Wout;
  const int b = static_cast<int>(tok / hw), rem = static_cast<int>(tok - static_cast<long long>(b) * hw);
  const int oy = rem / g.Wout, ox = rem - oy * g.Wout;
  const int iy = oy * g.stride - g.pad + ky, ix0 = ox * g.stride - g.pad;
  __half* dst = crow + (ci * g.KS + ky) * g.KS;
  const bool row_ok = iy >= 0 && iy < g.Hin;
  const size_t base = ((static_cast<size_t>(b) * g.Cin + ci) * g.Hin + (row_ok ? iy : 0)) * g.Win;
  for (int kx = 0; kx < g.KS; ++kx) {
    const int ix = ix0 + kx;
    float v = 0.f;
    if (row_ok && ix >= 0 && ix < g.Win) {
      if (g.in_dtype == 1) v = static_cast<float>(__ldg(reinterpret_cast<const uint8_t*>(in) + base + ix));
      else if (g.in_dtype == 2) v = __half2float(__ldg(reinterpret_cast<const __half*>(in) + base + ix));
      else v = __ldg(reinterpret_cast<const float*>(in) + base + ix);
    }
    dst[kx] = __float2half_rn(v);
  }
}

__global__ void __launch_bounds__(256) col2im_kernel(const __half* __restrict__ dcol, ConvGeom g, float* __restrict__ d_in) {
  const int half_c = g.Cin >> 1;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_pix = static_cast<long long>(g.B) * g.Hin * g.Win;
  if (idx >= n_pix * half_c) return;
  const long long pix = idx / half_c;
  const int ci = static_cast<int>(idx - pix * half_c) * 2;
  const int hw = g.Hin * g.Win;
  const int b = static_cast<int>(pix / hw), rem = static_cast<int>(pix - static_cast<long long>(b) * hw);
  const int iy = rem / g.Win, ix = rem - iy * g.Win;
  float s0 = 0.f, s1 = 0.f;
  for (int ky = 0; ky < g.KS; ++ky) {
    const int ty = iy + g.pad - ky;
    if (ty < 0 || ty % g.stride != 0) continue;
    const int oy = ty / g.stride;
    if (oy >= g.Hout) continue;
    for (int kx = 0; kx < g.KS; ++kx) {
      const int tx = ix + g.pad - kx;
      if (tx < 0 || tx % g.stride != 0) continue;
      const int ox = tx / g.stride;
      if (ox >= g.Wout) continue;
      const size_t tok = (static_cast<size_t>(b) * g.Hout + oy) * g.Wout + ox;
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(dcol + tok * g.ldc + (ky * g.KS + kx) * g.Cin + ci));
      s0 += f.x; s1 += f.y;
    }
  }
  *reinterpret_cast<float2*>(d_in + static_cast<size_t>(pix) * g.Cin + ci) = make_float2(s0, s1);
}

// acc[n] += sum_m a[m, n]   (bias gradients).  fp16 [M, N] (ld), N % 8 == 0, N <= 2048.  16-byte loads: thread = 8 columns
// of one row slot; a CTA sweeps rows [m_lo, m_hi), reduces its row slots through shared memory, one atomicAdd per column.
__global__ void __launch_bounds__(256) colsum_kernel(const __half* __restrict__ a, long long M, int N, int ld, float* acc,
                                                     int rows_per_block) {
  __shared__ float s[2048];
  const int tpr = N >> 3;                    // threads per row
  const int slots = 256 / tpr;               // rows in flight per sweep
  const int cg = threadIdx.x % tpr, slot = threadIdx.x / tpr;
  const long long m_lo = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long m_hi = min(M, m_lo + rows_per_block);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (slot < slots) {
    long long m = m_lo + slot;
    for (; m + 3LL * slots < m_hi; m += 4LL * slots) {          // four independent 16-byte loads in flight
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = __ldg(reinterpret_cast<const uint4*>(a + static_cast<size_t>(m + q * slots) * ld + cg * 8));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __half2* h = reinterpret_cast<const __half2*>(&u[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
      }
    }
    for (; m < m_hi; m += slots) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(a + static_cast<size_t>(m) * ld + cg * 8));
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
    }
  }
  if (slot < slots) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s[slot * N + cg * 8 + e] = v[e];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += 256) {
    float t = 0.f;
    for (int w = 0; w < slots; ++w) t += s[w * N + n];
    atomicAdd(acc + n, t);
  }
}

}  // namespace rvt
