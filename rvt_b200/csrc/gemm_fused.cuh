// One warp-specialised tcgen05 GEMM mainloop with pluggable A-operand loaders (prologues)
// and accumulator epilogues.  Every dense contraction of the RVT backbone stage
// (reference: models/layers/maxvit/maxvit.py, models/layers/rnn.py) is an instance:
//
//   D[128 x BN] (fp32, TMEM) = A[128 x K] (fp16, built in smem by the loader) * W[BN x K]^T
//
//   loader LD_F16   : A rows are fp16 rows of a scratch matrix
//          LD_LN    : A = LayerNorm(x[token(row)]) (or x itself), tokens gathered through the
//                     window / grid partition map  (maxvit.py:234,241,252-265,273-304)
//          LD_CONV  : A = im2col of the strided downsample conv input (maxvit.py:166-175)
//          LD_XH    : A = cat(x, [dwconv3x3](h_prev)) for the Conv-LSTM 1x1 (rnn.py:50-55)
//   epilogue EP_F16 : +bias, optional exact-erf GELU, fp16 store
//            EP_RES : x[token] = res[token] + gamma * (acc + bias)   (LayerScale + residual,
//                     scattered back through the partition map = window/grid reverse)
//            EP_LN  : LayerNorm over the C output channels of the conv (+ mask token)
//            EP_LSTM: gates -> (h_t, c_t)  (rnn.py:57-67)
//
// Roles: warps 0-3 build A tiles (registers -> swizzled smem) and later run the epilogue
// (thread t owns accumulator row t = TMEM lane t); thread 0 also streams the pre-packed
// weight tiles with 1-D bulk async copies; warp 4 lane 0 issues tcgen05.mma and commits.
#pragma once
#include "umma.cuh"

namespace rvt {

enum { LD_F16 = 0, LD_LN = 1, LD_CONV = 2, LD_XH = 3 };
enum { EP_F16 = 0, EP_RES = 1, EP_LN = 2, EP_LSTM = 3 };
enum { MAP_IDENTITY = 0, MAP_WINDOW = 1, MAP_GRID = 2 };

// tile row -> token of a [B, H, W, C] channels-last tensor
struct RowMap {
  int mode;
  int H, W;
  int ph, pw;        // partition size
  int ny, nx;        // groups along y / x  (H/ph, W/pw)
  int P;             // ph*pw
  int rows_per_win;  // rows a partition group occupies in a tile (64 or 128)
  int n_groups;      // B*ny*nx
  int n_tokens;      // B*H*W
};

__device__ __forceinline__ int row_to_token(const RowMap& m, int row) {
  if (m.mode == MAP_IDENTITY) return row < m.n_tokens ? row : -1;
  const int g = row / m.rows_per_win, p = row - g * m.rows_per_win;
  if (p >= m.P || g >= m.n_groups) return -1;
  const int per_img = m.ny * m.nx;
  const int b = g / per_img, gi = g - b * per_img;
  const int gy = gi / m.nx, gx = gi - gy * m.nx;
  const int py = p / m.pw, px = p - py * m.pw;
  int y, x;
  if (m.mode == MAP_WINDOW) { y = gy * m.ph + py; x = gx * m.pw + px; }
  else                      { y = py * m.ny + gy; x = px * m.nx + gx; }
  return (b * m.H + y) * m.W + x;
}

struct GemmArgs {
  // tiling
  int K, KC, BN, stages, tmem_cols, ab_fmt;
  const __half* Wp;    // packed weights [n_ntiles][KC][BN x 64 SW128 image]
  const float* bias;   // [n_ntiles*BN] in tile column order, or null
  RowMap map;
  // LD_F16
  const __half* a16; int lda; int a_rows;
  // LD_LN / LD_XH
  const float* x; int C;
  const float* ln_w; const float* ln_b; float eps; int do_ln;
  // LD_XH
  const float* hprev; const float* dw_w; const float* dw_b; int dws_mode; int dws_ks;
  // LD_CONV
  const void* cin; int in_dtype; int in_nchw; int Cin, Hin, Win, KS, cstride, cpad, Hout, Wout;
  // EP_F16
  __half* o16; int ldo; int act;
  // EP_RES
  const float* res; float* xout; const float* gamma;
  // EP_LN
  float* yout; const float* eln_w; const float* eln_b; float eeps;
  const uint8_t* token_mask; const float* mask_token;
  // EP_LSTM
  const float* cprev; float* hout; float* cout; int cw;
};

constexpr int kMaxStages = 6;
constexpr uint32_t kATileBytes = 128 * 128;   // 128 rows x 64 fp16

__host__ __device__ inline size_t gemm_smem_bytes(int stages, int BN) {
  return 1024 /*align slack*/ + static_cast<size_t>(stages) * (kATileBytes + static_cast<size_t>(BN) * 128) +
         2 * 128 * sizeof(float) + (2 * kMaxStages + 1) * sizeof(uint64_t) + 16;
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float sigmoid_acc(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------
// depthwise KSxKS conv of 8 consecutive channels at one token (zero padded), fp32
// (rnn.py:24-28,50-54: nn.Conv2d(groups=dim, padding=k//2) with bias)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void dwconv8(const float* __restrict__ src, int Csrc, int ch_src, const float* __restrict__ w,
                                        const float* __restrict__ bias, int D, int ch_w, int ks, int b, int y, int x,
                                        int H, int W, float* out) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + ch_w));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + ch_w + 4));
  out[0] = b0.x; out[1] = b0.y; out[2] = b0.z; out[3] = b0.w;
  out[4] = b1.x; out[5] = b1.y; out[6] = b1.z; out[7] = b1.w;
  if (src == nullptr) return;  // zero state: conv(0) = bias
  const int r = ks >> 1;
  for (int dy = 0; dy < ks; ++dy) {
    const int yy = y + dy - r;
    if (yy < 0 || yy >= H) continue;
    for (int dx = 0; dx < ks; ++dx) {
      const int xx = x + dx - r;
      if (xx < 0 || xx >= W) continue;
      const float* sp = src + (static_cast<size_t>(b * H + yy) * W + xx) * Csrc + ch_src;
      const float* wp = w + static_cast<size_t>(dy * ks + dx) * D + ch_w;
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp + 4));
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      out[0] = fmaf(s0.x, w0.x, out[0]); out[1] = fmaf(s0.y, w0.y, out[1]);
      out[2] = fmaf(s0.z, w0.z, out[2]); out[3] = fmaf(s0.w, w0.w, out[3]);
      out[4] = fmaf(s1.x, w1.x, out[4]); out[5] = fmaf(s1.y, w1.y, out[5]);
      out[6] = fmaf(s1.z, w1.z, out[6]); out[7] = fmaf(s1.w, w1.w, out[7]);
    }
  }
}

// ----------------------------------------------------------------------------------------
// The kernel
// ----------------------------------------------------------------------------------------
template <int LOADER, int EPI>
__global__ void __launch_bounds__(160) gemm_fused_kernel(const __grid_constant__ GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base_addr - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int stages = a.stages, KC = a.KC, BN = a.BN;
  const uint32_t b_bytes = static_cast<uint32_t>(BN) * 128u;

  const uint32_t sA_addr = base_addr;
  const uint32_t sB_addr = base_addr + stages * kATileBytes;
  uint8_t* sB = sm + stages * kATileBytes;
  float* s_mean = reinterpret_cast<float*>(sB + static_cast<size_t>(stages) * b_bytes);
  float* s_rstd = s_mean + 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_rstd + 128);
  uint64_t* empty = full + kMaxStages;
  uint64_t* accum = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== A-tile producers ===========================
    const int j = tid & 7;          // 16-byte chunk (8 fp16) inside the 64-wide K chunk
    const int r0 = tid >> 3;        // rows r0 + 16*i, i = 0..7
    int tok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = mt * 128 + r0 + 16 * i;
      if (LOADER == LD_F16) tok[i] = row < a.a_rows ? row : -1;
      else tok[i] = row_to_token(a.map, row);
    }

    if (LOADER == LD_LN && a.do_ln) {
      // per-row LayerNorm statistics (two-pass in registers; C <= 512 -> <= 16 values / lane)
      const int C = a.C;
      for (int rr = 0; rr < 32; ++rr) {
        const int r = warp * 32 + rr;
        const int t = row_to_token(a.map, mt * 128 + r);
        float v[16];
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int c = lane + 32 * q;
          v[q] = (t >= 0 && c < C) ? __ldg(a.x + static_cast<size_t>(t) * C + c) : 0.f;
          s += v[q];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / C;
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int c = lane + 32 * q;
          const float d = (c < C) ? v[q] - mean : 0.f;
          ss += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) { s_mean[r] = mean; s_rstd[r] = rsqrtf(ss / C + a.eps); }
      }
      named_bar_sync(1, 128);
    }

    // conv: per-row input window origin
    int cb[8], ciy[8], cix[8];
    if (LOADER == LD_CONV || LOADER == LD_XH) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int hw = (LOADER == LD_CONV) ? a.Hout * a.Wout : a.map.H * a.map.W;
        const int wd = (LOADER == LD_CONV) ? a.Wout : a.map.W;
        const int t = tok[i] < 0 ? 0 : tok[i];
        const int b = t / hw, rem = t - b * hw;
        const int oy = rem / wd, ox = rem - oy * wd;
        cb[i] = b;
        ciy[i] = (LOADER == LD_CONV) ? oy * a.cstride - a.cpad : oy;
        cix[i] = (LOADER == LD_CONV) ? ox * a.cstride - a.cpad : ox;
      }
    }

    for (int kc = 0; kc < KC; ++kc) {
      const int s = kc % stages;
      const uint32_t ph = (kc / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (tid == 0) {
        mbar_expect_tx(&full[s], b_bytes);
        bulk_g2s(sB + static_cast<size_t>(s) * b_bytes,
                 a.Wp + (static_cast<size_t>(nt) * KC + kc) * static_cast<size_t>(BN) * 64, b_bytes, &full[s]);
      }
      const uint32_t tile = sA_addr + s * kATileBytes;
      const int k0 = kc * 64 + j * 8;

      if (LOADER == LD_F16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (tok[i] >= 0 && k0 < a.K)
            v = __ldg(reinterpret_cast<const uint4*>(a.a16 + static_cast<size_t>(tok[i]) * a.lda + k0));
          st_smem_16B(tile + sw128_offset(r0 + 16 * i, j), v.x, v.y, v.z, v.w);
        }
      } else if (LOADER == LD_LN) {
        const bool kv = k0 < a.C;
        float g[8], bb[8];
        if (a.do_ln && kv) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.ln_w + k0));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.ln_w + k0 + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.ln_b + k0));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.ln_b + k0 + 4));
          g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
          bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (tok[i] >= 0 && kv) {
            const float* p = a.x + static_cast<size_t>(tok[i]) * a.C + k0;
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
            v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
            if (a.do_ln) {
              const float mu = s_mean[r0 + 16 * i], rs = s_rstd[r0 + 16 * i];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = (v[e] - mu) * rs * g[e] + bb[e];
            }
          }
          st_smem_16B(tile + sw128_offset(r0 + 16 * i, j), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                      pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      } else if (LOADER == LD_XH) {
        const int C = a.C;
        const bool is_h = k0 >= C;
        const int ch = is_h ? k0 - C : k0;
        const bool kv = k0 < 2 * C;
        const bool conv_this = kv && ((a.dws_mode == 1 && is_h) || a.dws_mode == 2);
        const float* src = is_h ? a.hprev : a.x;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (tok[i] >= 0 && kv) {
            if (conv_this) {
              const int D = a.dws_mode == 2 ? 2 * C : C;
              dwconv8(src, C, ch, a.dw_w, a.dw_b, D, a.dws_mode == 2 ? k0 : ch, a.dws_ks, cb[i], ciy[i], cix[i],
                      a.map.H, a.map.W, v);
            } else if (src != nullptr) {
              const float* p = src + static_cast<size_t>(tok[i]) * C + ch;
              const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
              const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
              v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
            }
          }
          st_smem_16B(tile + sw128_offset(r0 + 16 * i, j), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]),
                      pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      } else {  // LD_CONV
        __half hv[8][8];
        const int KS = a.KS, Cin = a.Cin, Hin = a.Hin, Win = a.Win;
        if (a.in_nchw) {
          // k = ci*KS*KS + ky*KS + kx  (natural [Cout, Cin, KS, KS] weight order); scalar gathers
          const int kk = KS * KS;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = k0 + e;
            const int ci = k / kk, rem = k - ci * kk;
            const int ky = rem / KS, kx = rem - ky * KS;
            const bool kv = k < a.K;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int iy = ciy[i] + ky, ix = cix[i] + kx;
              float f = 0.f;
              if (kv && tok[i] >= 0 && iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
                const size_t off = (static_cast<size_t>(cb[i]) * Cin + ci) * Hin * Win + static_cast<size_t>(iy) * Win + ix;
                if (a.in_dtype == 1) f = static_cast<float>(__ldg(reinterpret_cast<const uint8_t*>(a.cin) + off));
                else if (a.in_dtype == 2) f = __half2float(__ldg(reinterpret_cast<const __half*>(a.cin) + off));
                else f = __ldg(reinterpret_cast<const float*>(a.cin) + off);
              }
              hv[i][e] = __float2half_rn(f);
            }
          }
        } else {
          // channels-last input, k = (ky*KS + kx)*Cin + ci, Cin % 8 == 0
          const int tap = k0 / Cin, ci = k0 - tap * Cin;
          const int ky = tap / KS, kx = tap - ky * KS;
          const bool kv = k0 < a.K;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int iy = ciy[i] + ky, ix = cix[i] + kx;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (kv && tok[i] >= 0 && iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
              const float* p = reinterpret_cast<const float*>(a.cin) +
                               (static_cast<size_t>(cb[i] * Hin + iy) * Win + ix) * Cin + ci;
              const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
              const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
              v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) hv[i][e] = __float2half_rn(v[e]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t* u = reinterpret_cast<const uint32_t*>(&hv[i][0]);
          st_smem_16B(tile + sw128_offset(r0 + 16 * i, j), u[0], u[1], u[2], u[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
    }

    // =========================== epilogue ===========================
    mbar_wait(accum, 0);
    tc_fence_after();
    const int row = mt * 128 + tid;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int etok = (EPI == EP_F16) ? row : row_to_token(a.map, row);

    if (EPI == EP_F16) {
      __half* dst = a.o16 + static_cast<size_t>(row) * a.ldo + nt * BN;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int w = min(32, BN - c0);
        float v[32];
        if (w == 32) tmem_ld_x32(trow + c0, v); else tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          if (q < w) {
            float t = v[q] + (a.bias ? __ldg(a.bias + nt * BN + c0 + q) : 0.f);
            if (a.act == 1) t = gelu_erf(t);
            v[q] = t;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 8 < w) {
            uint4 o;
            o.x = pack_h2(v[q * 8 + 0], v[q * 8 + 1]); o.y = pack_h2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_h2(v[q * 8 + 4], v[q * 8 + 5]); o.w = pack_h2(v[q * 8 + 6], v[q * 8 + 7]);
            *reinterpret_cast<uint4*>(dst + c0 + q * 8) = o;
          }
        }
      }
    } else if (EPI == EP_RES) {
      const int C = a.C;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int w = min(32, BN - c0);
        float v[32];
        if (w == 32) tmem_ld_x32(trow + c0, v); else tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
        if (etok >= 0) {
          const int col = nt * BN + c0;
          const float* rp = a.res + static_cast<size_t>(etok) * C + col;
          float* op = a.xout + static_cast<size_t>(etok) * C + col;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q * 4 < w) {
              const float4 r = *reinterpret_cast<const float4*>(rp + q * 4);  // plain load: res may alias xout
              float4 o;
              float acc[4] = {v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]};
              float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float t = acc[e] + (a.bias ? __ldg(a.bias + col + q * 4 + e) : 0.f);
                if (a.gamma) t *= __ldg(a.gamma + col + q * 4 + e);
                rr[e] += t;
              }
              o.x = rr[0]; o.y = rr[1]; o.z = rr[2]; o.w = rr[3];
              *reinterpret_cast<float4*>(op + q * 4) = o;
            }
          }
        }
      }
    } else if (EPI == EP_LN) {
      // LayerNorm over the BN == C conv output channels held in this thread's TMEM lane
      float s = 0.f;
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) s += v[q];
      }
      const float mean = s / BN;
      float ss = 0.f;
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) { const float d = v[q] - mean; ss += d * d; }
      }
      const float rstd = rsqrtf(ss / BN + a.eeps);
      const bool masked = (etok >= 0) && a.token_mask && a.token_mask[etok];
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
        if (etok >= 0) {
          float* op = a.yout + static_cast<size_t>(etok) * BN + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c0 + q * 4 + e;
              float t = (v[q * 4 + e] - mean) * rstd;
              if (a.eln_w) t = t * __ldg(a.eln_w + c) + __ldg(a.eln_b + c);
              if (masked) t = __ldg(a.mask_token + c);
              o[e] = t;
            }
            *reinterpret_cast<float4*>(op + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    } else {  // EP_LSTM: tile columns = [f | i | o | g] x cw channels  (rnn.py:57-67)
      const int cw = a.cw, C = a.C;
      for (int j0 = 0; j0 < cw; j0 += 16) {
        float f[16], ig[16], og[16], g[16];
        tmem_ld_x16(trow + j0, f);
        tmem_ld_x16(trow + cw + j0, ig);
        tmem_ld_x16(trow + 2 * cw + j0, og);
        tmem_ld_x16(trow + 3 * cw + j0, g);
        tmem_ld_wait();
        if (etok >= 0) {
          const int ch0 = nt * cw + j0;
          const float* bt = a.bias + nt * BN;
          const size_t off = static_cast<size_t>(etok) * C + ch0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 cp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.cprev) cp = __ldg(reinterpret_cast<const float4*>(a.cprev + off + q * 4));
            const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
            float hn[4], cn[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = j0 + q * 4 + e;
              const float fg = sigmoid_acc(f[q * 4 + e] + __ldg(bt + jj));
              const float i_ = sigmoid_acc(ig[q * 4 + e] + __ldg(bt + cw + jj));
              const float o_ = sigmoid_acc(og[q * 4 + e] + __ldg(bt + 2 * cw + jj));
              const float g_ = tanhf(g[q * 4 + e] + __ldg(bt + 3 * cw + jj));
              cn[e] = fg * cpv[e] + i_ * g_;
              hn[e] = o_ * tanhf(cn[e]);
            }
            *reinterpret_cast<float4*>(a.cout + off + q * 4) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(a.hout + off + q * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
          }
        }
      }
    }
  } else {
    // =========================== MMA issuer (warp 4, one lane) ===========================
    if (lane == 0) {
      const int n0 = BN > 256 ? 256 : BN;
      const int n1 = BN - n0;
      const uint32_t idesc0 = umma_idesc_f16(128, n0, a.ab_fmt);
      const uint32_t idesc1 = n1 > 0 ? umma_idesc_f16(128, n1, a.ab_fmt) : 0u;
      for (int kc = 0; kc < KC; ++kc) {
        const int s = kc % stages;
        const uint32_t ph = (kc / stages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t at = sA_addr + s * kATileBytes;
        const uint32_t bt = sB_addr + s * b_bytes;
        const int krem = a.K - kc * 64;
        const int ksteps = krem >= 64 ? 4 : (krem + 15) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t acc = (kc | k) != 0 ? 1u : 0u;
          const uint64_t ad = umma_desc_sw128(at + k * 32);
          umma_f16(tmem_base, ad, umma_desc_sw128(bt + k * 32), idesc0, acc);
          if (n1 > 0) umma_f16(tmem_base + 256, ad, umma_desc_sw128(bt + 256 * 128 + k * 32), idesc1, acc);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(accum);
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, a.tmem_cols);
}

}  // namespace rvt
