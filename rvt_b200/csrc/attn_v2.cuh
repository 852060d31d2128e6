// Fused attention half of PartitionAttentionCl, persistent + head-parallel version (reference maxvit.py:252-268 with
// SelfAttentionCl :343-354 and the window / grid partition + reverse :273-304):
//     x[tok] <- x[tok] + gamma1 * ( Wp * concat_h softmax(q_h k_h^T / sqrt(dh)) v_h + bp )
//
// One CTA is resident per SM slot and loops over 128-row tiles (two partition groups of P <= 64 tokens, padded to 64 rows):
//
//   x-producer  one thread: the two groups' [P x C] fp32 token boxes of the NEXT tile by cp.async.bulk.tensor (5-D tensor
//               map: window = box (C, pw, 1, ph, 1) of (C, pw, nx, ph, B*ny); grid = box (C, 1, pw, 1, ph) of
//               (C, nx, pw, ny, B*ph) -- the partition never exists in memory), issued as soon as the PV MMAs of the current
//               tile have drained the region it aliases
//   w-producer  one thread: [q|k|v] weight tile per head + proj weight K-atoms through a bulk-copy ring (resident when they fit)
//   MMA         one thread: QKV_h = A Wqkv_h^T (all heads back to back), S_h = Q_h K_h^T, O_h = P_h V_h, out = O Wp^T
//   workers     NH warpgroups, warpgroup h owns head h, thread = tile row = TMEM lane:
//                 all : LayerNorm of the x tile (smem fp32 -> fp16 A operand), 8 lanes per row
//                 h   : QKV_h + bias -> fp16 Q_h | K_h | V_h operand tiles      (all heads concurrently)
//                 h   : masked softmax of S_h over the row's own group, P_h tile (compact [128 x 64 keys]); 1/rowsum in a register
//                 h   : O_h / rowsum -> fp16 column block h of the proj A operand
//                 all : out + bias, * gamma1, + residual -> x (scatter through the partition map = window / grid reverse),
//                       warpgroup h stores column block h
//
// V is consumed in place as an MN-major B operand (no transpose); the block-diagonal P V product is two M = 128 MMAs per head
// (one per partition group) into separate TMEM column blocks, each row reading the block of its own group.
//
// Shared memory (regions are reused along the tile's life):  R1: x tile fp32 -> Q | K operand atoms -> P_h tiles;
// R2: A operand -> O operand;  R3: V;  weight ring;  staged bias / LayerNorm / LayerScale vectors.
// TMEM (128*NH columns): QKV_h [96h, 96h+96) -> S_h [128h, 128h+128) -> O_h group blocks [128h + 64g, +64) -> out [0, C).
//
// Limits: P <= 64, dim_head <= 32 (padded to 32 by packing.pack_qkv_weight), nh in {1, 2, 4}, C <= 128, C % 16 == 0.
#pragma once
#include "gemm_fused.cuh"

namespace rvt {

struct AttnV2Args {
  float* x;                  // [B, H, W, C] in/out
  RowMap map;                // rows_per_win == 64
  int C, dh, nh, n_tiles;
  const float* ln_w; const float* ln_b; float eps; int do_ln;
  const __half* wqkv;        // pack_qkv_weight: [nh][KC1][96 x 64]
  const float* bqkv;         // [nh][96] padded
  const __half* wproj;       // pack_linear_weight(Wp, bn = C): [1][KC1][C x 64]
  const float* bproj;        // [C] or null
  const float* gamma;        // [C] or null
  float scale_log2e;
  long long* trace;          // optional [grid][kTraceTiles][kTracePts] globaltimer stamps of warp 0 lane 0 (profiling aid)
  int fast_ln;               // 1: C == 32 * nh and the x tile arrives as nh 32-channel SW128 half tiles (thread-per-row LayerNorm)
};

constexpr int kTraceTiles = 8, kTracePts = 12;
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define RVT_TRACE(args, it, pt) do { if ((args).trace && tid == 0 && (it) < kTraceTiles) (args).trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + (it)) * kTracePts + (pt)] = gtime(); } while (0)

constexpr int kAv2Stages = 3;
constexpr uint32_t kAv2Tile = 16384;     // one [128 x 64] fp16 operand atom

template <int NH, int KC1>
struct AttnV2Cfg {
  static constexpr int NW = NH * 128;                       // worker threads
  static constexpr int NA = NH >= 2 ? NH / 2 : 1;           // 64-column atoms holding all heads' padded head dims
  static constexpr int THREADS = NW + 128;                  // + one auxiliary warpgroup: MMA warp, x-producer warp, w-producer warp, spare
  // NH = 2 runs two CTAs per SM: 384 threads x 80 registers at launch; the auxiliary warpgroup then hands registers to the
  // workers (setmaxnreg; the pool is per CTA): workers 96, auxiliary 48 -> 2 x 128 x 96 + 128 x 48 = the same 30720 per CTA
  static constexpr bool REBALANCE_REGS = NH == 2;
  static constexpr uint32_t R1 = (32768u * KC1 > 2u * NA * kAv2Tile) ? 32768u * KC1 : 2u * NA * kAv2Tile;
  static constexpr uint32_t R2 = KC1 * kAv2Tile;
  static constexpr uint32_t R3 = NA * kAv2Tile;
  static constexpr uint32_t SLOT = KC1 * 96 * 128;          // >= proj K-atom (C * 128 <= 64 * KC1 * 128)
  static constexpr int CHUNKS = NH + KC1;                   // weight chunks per tile
  static constexpr bool RESIDENT = CHUNKS <= kAv2Stages;
  static constexpr uint32_t PAR_FLOATS = NH * 96 + 4 * 64 * KC1 + NH * 256;     // + per-row LayerNorm partial sums
  static constexpr uint32_t SMEM = 1024 + R1 + R2 + R3 + kAv2Stages * SLOT + PAR_FLOATS * 4 + 64 * 4 + 4 * 4 + 128 * 4 + 32 * 8 + 16;
  static constexpr int TMEM_COLS = 128 * NH;
  static constexpr int CTAS_PER_SM = NH <= 2 ? 2 : 1;
};

// 8 consecutive floats from a 32-byte aligned generic pointer into shared memory (two LDS.128; broadcast when warp-uniform)
__device__ __forceinline__ void lds8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// Thread-per-row LayerNorm on a [128 rows x 32 fp32] half tile that TMA wrote with SWIZZLE_128B (row r's 16-byte chunk c
// sits at r*128 + ((c ^ (r & 7)) << 4): the 8 lanes of a quarter warp hit 8 different bank groups, conflict free).  `part`
// exchanges the per-row partial sums between the NWG warpgroups that share a row; one named barrier (id `bar_id`, NWG*128
// threads).  The result goes to 16-byte chunks [4*wg, 4*wg+4) of the row's 128-byte operand row, i.e. K columns [32*wg, +32).
template <int NWG>
__device__ __forceinline__ void ln_row32_to_operand(uint32_t half_tile, int row, int wg, bool valid, bool do_ln, int C, float eps,
                                                    const float* s_lnw, const float* s_lnb, float* part, int bar_id,
                                                    uint32_t a_tile_base) {
  float v[32];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t src = half_tile + sw128_offset(row, c);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3]) : "r"(src));
  }
  if (!valid) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = 0.f;
  }
  if (do_ln) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) { s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
    if (NWG > 1) {
      part[(wg * 128 + row) * 2] = s1;
      part[(wg * 128 + row) * 2 + 1] = s2;
      named_bar_sync(bar_id, NWG * 128);
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int w = 0; w < NWG; ++w) { s1 += part[(w * 128 + row) * 2]; s2 += part[(w * 128 + row) * 2 + 1]; }
    }
    const float mean = s1 / C;
    const float rstd = rsqrtf(fmaxf(s2 / C - mean * mean, 0.f) + eps);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float g[8], bb[8];
      lds8(s_lnw + 32 * wg + 8 * c, g);
      lds8(s_lnb + 32 * wg + 8 * c, bb);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[8 * c + e] = valid ? fmaf((v[8 * c + e] - mean) * rstd, g[e], bb[e]) : 0.f;
    }
  }
  const uint32_t dst = a_tile_base + ((32 * wg) >> 6) * kAv2Tile;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    st_smem_16B(dst + sw128_offset(row, (wg & 1) * 4 + c), pack_h2(v[8 * c], v[8 * c + 1]), pack_h2(v[8 * c + 2], v[8 * c + 3]),
                pack_h2(v[8 * c + 4], v[8 * c + 5]), pack_h2(v[8 * c + 6], v[8 * c + 7]));
}

template <int NH, int KC1>
__global__ void __launch_bounds__(AttnV2Cfg<NH, KC1>::THREADS, AttnV2Cfg<NH, KC1>::CTAS_PER_SM)
attn_v2_kernel(const __grid_constant__ AttnV2Args a, const __grid_constant__ CUtensorMap tmap_x) {
  using Cfg = AttnV2Cfg<NH, KC1>;
  constexpr int NW = Cfg::NW, NA = Cfg::NA;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, dh = a.dh, P = a.map.P;
  const int nkg = (P + 15) & ~15;                       // keys of one group, padded to the MMA K step

  const uint32_t sR1 = base;
  const uint32_t sQ = sR1, sK = sR1 + NA * kAv2Tile;
  const uint32_t sA = sR1 + Cfg::R1;                    // A operand, later the O operand
  const uint32_t sV = sA + Cfg::R2;
  const uint32_t sW = sV + Cfg::R3;
  float* s_par = reinterpret_cast<float*>(sm + (sW - base) + kAv2Stages * Cfg::SLOT);
  float* s_bqkv = s_par;                                // [NH * 96]
  float* s_bproj = s_par + NH * 96;                     // [64 * KC1] each
  float* s_gamma = s_bproj + 64 * KC1;
  float* s_lnw = s_gamma + 64 * KC1;
  float* s_lnb = s_lnw + 64 * KC1;
  float* s_part = s_lnb + 64 * KC1;                          // [NH][128][2] LayerNorm partial sums (fast_ln)
  int* s_lut = reinterpret_cast<int*>(s_part + NH * 256);    // [64] token offset of position p inside its group
  int* s_tbase = s_lut + 64;                                 // [2 parities][2 groups] first token of the group, -1 = no group
  int* s_tok = s_tbase + 4;                                  // [128] token of every tile row (-1 = padding), for the coalesced epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_tok + 128) + 7) & ~static_cast<uintptr_t>(7));
  uint64_t* x_full = bars + 0;        // tx
  uint64_t* x_free = bars + 1;        // commit (PV done: R1 may be overwritten)
  uint64_t* a_full = bars + 2;        // NW
  uint64_t* qk_ready = bars + 3;      // NW
  uint64_t* s_full = bars + 4;        // commit
  uint64_t* so_full = bars + 5;       // NW
  uint64_t* out_full = bars + 6;      // commit
  uint64_t* qkv_full = bars + 7;      // [NH] commit
  uint64_t* p_full = qkv_full + 4;    // [NH] 128
  uint64_t* o_full = p_full + 4;      // [NH] commit
  uint64_t* w_full = o_full + 4;      // [stages] tx
  uint64_t* w_empty = w_full + kAv2Stages;   // [stages] commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kAv2Stages);

  if (tid == 0) {
    mbar_init(x_full, 1); mbar_init(x_free, 1); mbar_init(a_full, NW); mbar_init(qk_ready, NW); mbar_init(s_full, 1);
    mbar_init(so_full, NW); mbar_init(out_full, 1);
    for (int h = 0; h < NH; ++h) { mbar_init(&qkv_full[h], 1); mbar_init(&p_full[h], 128); mbar_init(&o_full[h], 1); }
    for (int s = 0; s < kAv2Stages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    fence_mbar_init();
  }
  if (warp == NH * 4) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  // per-CTA constants: bias / LayerNorm / LayerScale vectors and the in-group token offsets
  for (int i = tid; i < NH * 96; i += Cfg::THREADS) s_bqkv[i] = a.bqkv[i];
  for (int i = tid; i < 64 * KC1; i += Cfg::THREADS) {
    const bool in = i < C;
    s_bproj[i] = (in && a.bproj) ? a.bproj[i] : 0.f;
    s_gamma[i] = (in && a.gamma) ? a.gamma[i] : 1.f;
    s_lnw[i] = (in && a.do_ln) ? a.ln_w[i] : 1.f;
    s_lnb[i] = (in && a.do_ln) ? a.ln_b[i] : 0.f;
  }
  for (int p = tid; p < 64; p += Cfg::THREADS) {
    const int py = p / a.map.pw, px = p - py * a.map.pw;
    s_lut[p] = a.map.mode == MAP_WINDOW ? py * a.map.W + px : py * a.map.ny * a.map.W + px * a.map.nx;
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();            // x (TMA loads, residual reads) is the previous kernel's output; weights / bias vectors above are constants

  const int n_tiles = a.n_tiles;
  const int ks1 = C >> 4;                                // K steps of the C-wide contractions

  if (warp < NH * 4) {
    // =============================================== workers ===============================================
    if (Cfg::REBALANCE_REGS) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory");
    const int h = warp >> 2;                             // this warpgroup's head
    const int row = (warp & 3) * 32 + lane;              // tile row == TMEM lane
    const int grp = row >> 6, pos = row & 63;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int j8 = tid & 7;
    const int cw = C / NH;                               // output columns this warpgroup stores (multiple of 8)
    const uint32_t sP = sR1 + h * kAv2Tile;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      // ---------------- LayerNorm of the x tile -> A operand (8 lanes per row) ----------------
      RVT_TRACE(a, it, 0);
      mbar_wait(x_full, par);
      RVT_TRACE(a, it, 1);
      const int* tb = s_tbase + 2 * par;
      if (a.fast_ln) {
        ln_row32_to_operand<NH>(sR1 + h * kAv2Tile, row, h, pos < P && tb[grp] >= 0, a.do_ln != 0, C, a.eps, s_lnw, s_lnb, s_part, 1, sA);
      } else
      for (int r = tid >> 3; r < 128; r += NW / 8) {
        const bool valid = (r & 63) < P && tb[r >> 6] >= 0;
        float v[KC1][8];
#pragma unroll
        for (int kc = 0; kc < KC1; ++kc) {
          const int k0 = kc * 64 + j8 * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[kc][e] = 0.f;
          if (valid && k0 < C) {
            const uint32_t src = sR1 + (static_cast<uint32_t>(r) * C + k0) * 4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[kc][0]), "=f"(v[kc][1]), "=f"(v[kc][2]), "=f"(v[kc][3]) : "r"(src));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[kc][4]), "=f"(v[kc][5]), "=f"(v[kc][6]), "=f"(v[kc][7]) : "r"(src + 16));
          }
        }
        if (a.do_ln) {
          float s1 = 0.f;
#pragma unroll
          for (int kc = 0; kc < KC1; ++kc)
#pragma unroll
            for (int e = 0; e < 8; ++e) s1 += v[kc][e];
          const float mean = red8(s1) / C;
          float s2 = 0.f;
#pragma unroll
          for (int kc = 0; kc < KC1; ++kc)
            if (kc * 64 + j8 * 8 < C) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { const float d = v[kc][e] - mean; s2 += d * d; }
            }
          const float rstd = rsqrtf(red8(s2) / C + a.eps);
#pragma unroll
          for (int kc = 0; kc < KC1; ++kc) {
            const int k0 = kc * 64 + j8 * 8;
            if (valid && k0 < C) {
              float g[8], bb[8];
              lds8(s_lnw + k0, g);
              lds8(s_lnb + k0, bb);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[kc][e] = fmaf((v[kc][e] - mean) * rstd, g[e], bb[e]);
            }
          }
        }
#pragma unroll
        for (int kc = 0; kc < KC1; ++kc)
          st_smem_16B(sA + kc * kAv2Tile + sw128_offset(r, j8), pack_h2(v[kc][0], v[kc][1]), pack_h2(v[kc][2], v[kc][3]),
                      pack_h2(v[kc][4], v[kc][5]), pack_h2(v[kc][6], v[kc][7]));
      }
      fence_proxy_async_smem();
      tc_fence_before();            // orders this thread's TMEM reads of the previous tile before the MMAs that follow a_full
      mbar_arrive(a_full);
      RVT_TRACE(a, it, 2);

      // residual row segment of this thread, fetched early (the latency hides behind the whole attention chain)
      const int gbase = tb[grp];
      const bool live = pos < P && gbase >= 0;
      const int tok = live ? gbase + s_lut[pos] : -1;
      if (h == 0) s_tok[row] = tok;                      // read by the coalesced epilogue (all workers) much later

      // ---------------- QKV_h accumulators + bias -> Q_h | K_h | V_h operand tiles ----------------
      mbar_wait(&qkv_full[h], par);
      RVT_TRACE(a, it, 3);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        float v[32];
        tmem_ld_x32(tmem + lane_off + 96 * h + 32 * part, v);
        tmem_ld_wait();
        const float* bq = s_bqkv + h * 96 + part * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float bv[8];
          lds8(bq + 8 * c, bv);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[8 * c + e] += bv[e];
        }
        const uint32_t dst = (part == 0 ? sQ : (part == 1 ? sK : sV)) + (h >> 1) * kAv2Tile;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_smem_16B(dst + sw128_offset(row, (h & 1) * 4 + c), pack_h2(v[8 * c], v[8 * c + 1]), pack_h2(v[8 * c + 2], v[8 * c + 3]),
                      pack_h2(v[8 * c + 4], v[8 * c + 5]), pack_h2(v[8 * c + 6], v[8 * c + 7]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(qk_ready);
      RVT_TRACE(a, it, 4);

      // ---------------- masked softmax of S_h over the row's own group ----------------
      mbar_wait(s_full, par);
      RVT_TRACE(a, it, 5);
      tc_fence_after();
      const uint32_t ts = tmem + lane_off + 128 * h + 64 * grp;
      float mx = -INFINITY;
      for (int k0 = 0; k0 < nkg; k0 += 16) {
        float v[16];
        tmem_ld_x16(ts + k0, v);
        tmem_ld_wait();
        if (k0 + 16 <= P) {                  // full chunk: no masking (P = 60: three of the four chunks)
#pragma unroll
          for (int e = 0; e < 16; ++e) mx = fmaxf(mx, v[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (k0 + e < P) mx = fmaxf(mx, v[e]);
        }
      }
      const float mxs = mx * a.scale_log2e;
      float sum = 0.f;
      for (int k0 = 0; k0 < nkg; k0 += 16) {
        float v[16];
        tmem_ld_x16(ts + k0, v);
        tmem_ld_wait();
        if (k0 + 16 <= P) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { v[e] = ex2_approx(fmaf(v[e], a.scale_log2e, -mxs)); sum += v[e]; }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p = (k0 + e < P) ? ex2_approx(fmaf(v[e], a.scale_log2e, -mxs)) : 0.f;
            sum += p;
            v[e] = p;
          }
        }
        st_smem_16B(sP + sw128_offset(row, k0 >> 3), pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        st_smem_16B(sP + sw128_offset(row, (k0 >> 3) + 1), pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]),
                    pack_h2(v[14], v[15]));
      }
      const float inv = rcp_approx(sum);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[h]);
      RVT_TRACE(a, it, 6);

      // ---------------- O_h / rowsum -> column block h of the proj A operand ----------------
      mbar_wait(&o_full[h], par);
      RVT_TRACE(a, it, 7);
      tc_fence_after();
      {
        float v[32];
        tmem_ld_x32(tmem + lane_off + 128 * h + 64 * grp + 32 * (h & 1), v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 8 >= dh) break;
          const int col0 = h * dh + 8 * c;               // heads concatenated with stride dh (maxvit.py:352)
          st_smem_16B(sA + (col0 >> 6) * kAv2Tile + sw128_offset(row, (col0 & 63) >> 3), pack_h2(v[8 * c] * inv, v[8 * c + 1] * inv),
                      pack_h2(v[8 * c + 2] * inv, v[8 * c + 3] * inv), pack_h2(v[8 * c + 4] * inv, v[8 * c + 5] * inv),
                      pack_h2(v[8 * c + 6] * inv, v[8 * c + 7] * inv));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(so_full);
      RVT_TRACE(a, it, 8);

      // ---------------- proj epilogue: + bias, * gamma, + residual, scatter (= partition reverse) ----------------
      if ((C & (C - 1)) == 0 && C >= 32) {
        // Coalesced version.  A thread owns a ROW of the accumulator (TMEM lane), but row-per-thread global accesses touch 32
        // different 128-byte lines per instruction (the first profile's top cost: 3.5 us per tile).  So: (1) acc + bias, * gamma
        // -> fp32 staging tile in shared memory (the A/O and V regions, dead by now; 16-byte chunks XOR-swizzled with row & 7),
        // (2) the workers re-map to (row, 16-byte column chunk) so that the lanes of a warp cover whole 128-byte lines of x for
        // the residual read and the store.
        // residual prefetch in the coalesced (row, chunk) mapping, issued BEFORE the accumulator wait: the L2 latency of these
        // loads hides behind the proj MMA (s_tok was published by the row threads at the top of the tile; the qk_ready / s_full
        // hand-offs in between order those writes before these reads)
        const int nch = C >> 2;                                   // 16-byte chunks per row (power of two)
        const int ech = tid % nch, er0 = tid / nch, erstep = NW / nch;          // 128 rows * nch chunks / NW threads = 8 rows each
        constexpr int kRows = 8;
        float4 xr[kRows];
        int tk[kRows];
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
          tk[q] = s_tok[er0 + q * erstep];
          if (tk[q] >= 0) xr[q] = *reinterpret_cast<const float4*>(a.x + static_cast<size_t>(tk[q]) * C + ech * 4);
        }
        mbar_wait(out_full, par);
        RVT_TRACE(a, it, 9);
        tc_fence_after();
        const uint32_t srow = sA + static_cast<uint32_t>(row) * C * 4;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          if (c8 * 8 >= cw) break;
          float v[8];
          tmem_ld_x8(tmem + lane_off + h * cw + c8 * 8, v);
          tmem_ld_wait();
          const int col = h * cw + c8 * 8;
          float bv[8], gv[8];
          lds8(s_bproj + col, bv);
          lds8(s_gamma + col, gv);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (v[e] + bv[e]) * gv[e];
          const int ch = col >> 2;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((ch) ^ (row & 7)) << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((ch + 1) ^ (row & 7)) << 4)), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        }
        tc_fence_before();
        named_bar_sync(2, NW);
#pragma unroll
        for (int q = 0; q < kRows; ++q) {
          if (tk[q] < 0) continue;
          const int r = er0 + q * erstep;
          float4 sv;
          const uint32_t src = sA + static_cast<uint32_t>(r) * C * 4 + ((ech ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w) : "r"(src));
          *reinterpret_cast<float4*>(a.x + static_cast<size_t>(tk[q]) * C + ech * 4) =
              make_float4(xr[q].x + sv.x, xr[q].y + sv.y, xr[q].z + sv.z, xr[q].w + sv.w);
        }
        named_bar_sync(3, NW);                                    // the staging tile is the next tile's A operand / V
      } else {
      float res[32];
      float* xrow = a.x + static_cast<size_t>(tok < 0 ? 0 : tok) * C + h * cw;
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)
        if (live && c8 * 8 < cw) load8(xrow + c8 * 8, res + c8 * 8);
      mbar_wait(out_full, par);
      RVT_TRACE(a, it, 9);
      tc_fence_after();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        if (c8 * 8 >= cw) break;
        float v[8];
        tmem_ld_x8(tmem + lane_off + h * cw + c8 * 8, v);
        tmem_ld_wait();
        if (live) {
          const int col = h * cw + c8 * 8;
          float bv[8], gv[8];
          lds8(s_bproj + col, bv);
          lds8(s_gamma + col, gv);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e] + bv[e], gv[e], res[c8 * 8 + e]);
          *reinterpret_cast<float4*>(xrow + c8 * 8) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(xrow + c8 * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      }
      RVT_TRACE(a, it, 10);
    }
  } else {
   if (Cfg::REBALANCE_REGS) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
   if (warp == NH * 4) {
    // =============================================== MMA issuer ===============================================
    if (lane == 0) {
      const uint32_t id_qkv = umma_idesc_f16(128, 96, 0);
      const uint32_t id_s = umma_idesc_f16(128, 128, 0);
      const uint32_t id_pv = umma_idesc_f16(128, 64, 0) | (1u << 16);        // B (= V) MN-major
      const uint32_t id_out = umma_idesc_f16(128, C, 0);
      uint32_t wc = 0;                                   // weight chunks consumed (ring position)
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        mbar_wait(a_full, par);
        tc_fence_after();
        for (int h = 0; h < NH; ++h) {
          const uint32_t slot = Cfg::RESIDENT ? h : wc % kAv2Stages;
          mbar_wait(&w_full[slot], Cfg::RESIDENT ? 0u : ((wc / kAv2Stages) & 1u));
          tc_fence_after();
          const uint32_t wq = sW + slot * Cfg::SLOT;
          for (int k = 0; k < ks1; ++k) {
            const uint32_t atom = k >> 2, kk = k & 3;
            umma_f16(tmem + 96 * h, umma_desc_sw128(sA + atom * kAv2Tile + kk * 32), umma_desc_sw128(wq + atom * (96 * 128) + kk * 32),
                     id_qkv, k != 0);
          }
          umma_commit(&qkv_full[h]);
          if (!Cfg::RESIDENT) umma_commit(&w_empty[slot]);
          ++wc;
        }
        mbar_wait(qk_ready, par);
        tc_fence_after();
        for (int h = 0; h < NH; ++h) {
          const uint32_t q = sQ + (h >> 1) * kAv2Tile + (h & 1) * 64, k_ = sK + (h >> 1) * kAv2Tile + (h & 1) * 64;
          for (int k = 0; k < 2; ++k)                    // head dim padded to 32 = 2 K steps
            umma_f16(tmem + 128 * h, umma_desc_sw128(q + k * 32), umma_desc_sw128(k_ + k * 32), id_s, k != 0);
        }
        umma_commit(s_full);                             // ONE barrier: P tiles overwrite the Q / K atoms of other heads
        for (int h = 0; h < NH; ++h) {
          mbar_wait(&p_full[h], par);
          tc_fence_after();
          const uint32_t p = sR1 + h * kAv2Tile, v = sV + (h >> 1) * kAv2Tile;
          for (int g = 0; g < 2; ++g)
            for (int kk = 0; kk < nkg / 16; ++kk)
              umma_f16(tmem + 128 * h + 64 * g, umma_desc_sw128(p + kk * 32), umma_desc_sw128_mn(v + g * 8192 + kk * 2048, 16384),
                       id_pv, kk != 0);
          umma_commit(&o_full[h]);
        }
        umma_commit(x_free);
        mbar_wait(so_full, par);
        tc_fence_after();
        for (int kc = 0; kc < KC1; ++kc) {
          const uint32_t slot = Cfg::RESIDENT ? NH + kc : wc % kAv2Stages;
          mbar_wait(&w_full[slot], Cfg::RESIDENT ? 0u : ((wc / kAv2Stages) & 1u));
          tc_fence_after();
          const uint32_t wp = sW + slot * Cfg::SLOT;
          const int ksteps = ks1 - 4 * kc < 4 ? ks1 - 4 * kc : 4;
          for (int kk = 0; kk < ksteps; ++kk)
            umma_f16(tmem, umma_desc_sw128(sA + kc * kAv2Tile + kk * 32), umma_desc_sw128(wp + kk * 32), id_out, (kc | kk) != 0);
          if (!Cfg::RESIDENT) umma_commit(&w_empty[slot]);
          ++wc;
        }
        umma_commit(out_full);
      }
    }
    __syncwarp();
  } else if (warp == NH * 4 + 1) {
    // =============================================== x producer (TMA) ===============================================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_x);
      const int per_img = a.map.ny * a.map.nx;
      const uint32_t grp_bytes = static_cast<uint32_t>(P) * C * 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        if (it > 0) mbar_wait(x_free, (it - 1) & 1);
        int* tb = s_tbase + 2 * (it & 1);
        int c1[2], c2[2], c3[2], c4[2], nv = 0;
        for (int g = 0; g < 2; ++g) {
          const int gi_ = tile * 2 + g;
          if (gi_ < a.map.n_groups) {
            const int b = gi_ / per_img, gi = gi_ - b * per_img;
            const int gy = gi / a.map.nx, gx = gi - gy * a.map.nx;
            if (a.map.mode == MAP_WINDOW) {
              tb[g] = (b * a.map.H + gy * a.map.ph) * a.map.W + gx * a.map.pw;
              c1[g] = 0; c2[g] = gx; c3[g] = 0; c4[g] = b * a.map.ny + gy;
            } else {
              tb[g] = (b * a.map.H + gy) * a.map.W + gx;
              c1[g] = gx; c2[g] = 0; c3[g] = gy; c4[g] = b * a.map.ph;
            }
            ++nv;
          } else {
            tb[g] = -1;
          }
        }
        mbar_arrive_expect_tx(x_full, nv * grp_bytes);
        for (int g = 0; g < 2; ++g) {
          if (tb[g] < 0) continue;
          if (a.fast_ln) {           // NH half tiles of 32 channels, each [128 rows x 128 B] with the 128-byte swizzle
            for (int j = 0; j < NH; ++j)
              tma_load_5d(sR1 + j * kAv2Tile + g * 64 * 128, &tmap_x, 32 * j, c1[g], c2[g], c3[g], c4[g], x_full);
          } else {
            tma_load_5d(sR1 + g * 64 * C * 4, &tmap_x, 0, c1[g], c2[g], c3[g], c4[g], x_full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == NH * 4 + 2) {
    // =============================================== weight producer ===============================================
    if (lane == 0) {
      const uint32_t wq_bytes = static_cast<uint32_t>(KC1) * 96 * 128, wp_bytes = static_cast<uint32_t>(C) * 128;
      auto load_chunk = [&](int ci, uint32_t slot) {
        uint8_t* dst = sm + (sW - base) + slot * Cfg::SLOT;
        if (ci < NH) {
          mbar_arrive_expect_tx(&w_full[slot], wq_bytes);
          bulk_g2s(dst, a.wqkv + static_cast<size_t>(ci) * KC1 * 96 * 64, wq_bytes, &w_full[slot]);
        } else {
          mbar_arrive_expect_tx(&w_full[slot], wp_bytes);
          bulk_g2s(dst, a.wproj + static_cast<size_t>(ci - NH) * C * 64, wp_bytes, &w_full[slot]);
        }
      };
      if (Cfg::RESIDENT) {
        if (blockIdx.x < static_cast<unsigned>(n_tiles))
          for (int ci = 0; ci < Cfg::CHUNKS; ++ci) load_chunk(ci, ci);
      } else {
        uint32_t wc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
          for (int ci = 0; ci < Cfg::CHUNKS; ++ci, ++wc) {
            const uint32_t slot = wc % kAv2Stages;
            mbar_wait(&w_empty[slot], ((wc / kAv2Stages) & 1u) ^ 1u);
            load_chunk(ci, slot);
          }
      }
    }
    __syncwarp();
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NH * 4) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

}  // namespace rvt
