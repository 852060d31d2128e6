// Stem of stage 1, persistent + pipelined: ConvDownsampling_Cf2Cl with the 7x7 / stride-4 / pad-3 conv on uint8 NCHW event
// histograms, LayerNorm over the C output channels and the optional mask token (reference maxvit.py:161-178,
// maxvit_rnn.py:174-176).  Same arithmetic as gemm_fused<LD_STEM, EP_LN> (fp16 operands that are exact for uint8 counts, fp32
// accumulate, two-pass LayerNorm); the one-tile-per-CTA version serialises patch load -> 18 operand builds -> epilogue inside a
// CTA, here one CTA per SM keeps the roles busy at once.  A tile = 8 x 16 output tokens of one sample = 128 accumulator rows;
// K = (kyi, ci, kx8) = 7 * Cin * 8 in chunks of 64, kernel rows in the order ky = stem_ky(kyi) = 0, 4, 1, 5, 2, 6, 3.
//
// stem_v2_kernel<ATMEM = true> (default; DESIGN.md 4.1b has the measurements that led here):
//   producer   one thread: ALL packed weight chunks once (resident, 144 KB at C = 64); per tile the uint8 input patch as FOUR
//              row-phase planes (input rows iy0 + p + 4 k: one 3-D TMA box each with row stride 4; out-of-image rows / columns
//              and the rows of a zero-padded model resolution are the TMA unit's zero fill).  Plane p is dead once the kernel
//              rows that read it are done (2/7, 4/7, 6/7, 7/7 of the K loop), so its reload for the next tile overlaps this one
//   builders   8 warps, thread = tile row = TMEM lane: patch bytes -> fp16 (exact byte-permute trick) in registers -> tcgen05.st
//              into a 4-slot OPERAND RING IN TENSOR MEMORY, one slot = a step of three K chunks (96 packed columns); patch-row
//              offsets of the (kyi, ci) pairs come from a small shared-memory table
//   MMA        one thread: TS-form tcgen05.mma (A from tensor memory, B = resident weight chunk), 12 MMAs per hand-off, TWO
//              accumulators (TMEM columns [64 b, 64 b + C)); ring at columns [128, 512)
//   epilogue   4 warps, thread = token = TMEM lane: row statistics from one TMEM round trip, normalised rows into
//              [4 token rows x 16 tokens x 32 channels] boxes in the 128-byte swizzle, out through 3-D TMA tensor stores
//              (per-thread coalesced stores when C is not a multiple of 32); runs on tile i while the builders are on tile i+1
// stem_v2_kernel<false> (RVT_STEM_V2=1, and shapes the tensor-memory variant does not take): operand chunks built smem -> smem
// into a 3-deep ring, weight chunks streamed through a 6-deep ring, two full patches; same epilogue.
#pragma once
#include "attn_v2.cuh"

namespace rvt {

struct StemV2Args {
  const __half* wp;         // pack_stem_weight_u8: [KC][C x 64] SW128 K-major tiles
  float* y;                 // [B, Hout, Wout, C]
  int Cin, Hout, Wout, C, KC, n_tiles, ny, nx;
  const float* ln_w; const float* ln_b; float eps;      // null: no affine / no LayerNorm is not supported (ln_w may be null)
  const uint8_t* token_mask; const float* mask_token;
  long long* trace;         // optional [grid][kTraceTiles][kTracePts] globaltimer stamps of the role leaders (profiling aid)
  int tma_store;            // 1: C % 32 == 0 and tmap_out is valid: the normalised tile leaves through TMA tensor stores
};

#define SV2_TRACE(leader, it, pt) do { if (a.trace && (leader) && (it) < kTraceTiles) a.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + (it)) * kTracePts + (pt)] = gtime(); } while (0)
constexpr int kSv2Builders = 256;
constexpr int kSv2Epi = 128;
constexpr int kSv2Threads = kSv2Epi + kSv2Builders + 64;      // + producer warp + MMA warp
constexpr int kSv2Stages = 3;                                 // A-chunk ring in shared memory (ATMEM = false)
constexpr int kSv2TStages = 4;                                // ATMEM: operand ring in tensor memory, one slot = one STEP of kSv2KS K chunks
constexpr int kSv2KS = 3;                                     // K chunks (of 64) per step: 3 x 32 packed columns per slot
constexpr uint32_t kSv2TAcol = 128;                           // TMEM: accumulators [0, 128), operand ring [128, 128 + 4 * 96) = 512 columns
constexpr int kSv2WStages = 6;                                // weight-chunk ring: deep, an L2 -> smem bulk copy takes ~1000 cycles
constexpr uint32_t kSv2WBytes = 64 * 128;                     // one weight chunk (C <= 64 rows x 64 halves)
constexpr int kSv2PlaneRows = kStemTileH + 1;                 // ATMEM: rows of one row-phase plane of the input patch (iy = iy0 + p + 4 k)
__host__ __device__ inline uint32_t stem_v2_plane_bytes(int cin) {
  return (static_cast<uint32_t>(cin) * kSv2PlaneRows * kStemPatchPitch + 127u) & ~127u;
}
__host__ __device__ inline uint32_t stem_v2_patch_bytes(int cin) {
  return (static_cast<uint32_t>(cin) * kStemPatchRows * kStemPatchPitch + 1023u) & ~1023u;
}
// ATMEM variant: the patch as four row-phase planes, ALL weight chunks resident, no operand ring in shared memory
__host__ __device__ inline uint32_t stem_v2t_smem_bytes(int cin, int c, int kc) {
  return 1024 + ((4 * stem_v2_plane_bytes(cin) + 1023u) & ~1023u) + static_cast<uint32_t>(kc) * kSv2WBytes + 64 * (static_cast<uint32_t>(c) * 4 + 16) + 3 * 64 * 4 +
         48 * 8 + 16 + 256 * 4 /*pair -> patch offset LUT*/;
}
__host__ __device__ inline uint32_t stem_v2_smem_bytes(int cin, int c) {
  return 1024 + 2 * stem_v2_patch_bytes(cin) + kSv2Stages * kATileBytes + kSv2WStages * kSv2WBytes +
         64 * (static_cast<uint32_t>(c) * 4 + 16) + 3 * 64 * 4 + 48 * 8 + 16;
}

template <bool ATMEM>
__global__ void __launch_bounds__(kSv2Threads, 1)
stem_v2_kernel(const __grid_constant__ StemV2Args a, const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = a.C, Cin = a.Cin, KC = a.KC, n_tiles = a.n_tiles;
  const uint32_t patch_bytes = stem_v2_patch_bytes(Cin);
  const uint32_t sP = base;                                    // [2] patches ([1] with ATMEM)
  const uint32_t plane_bytes = stem_v2_plane_bytes(Cin);
  const uint32_t sS = sP + (ATMEM ? ((4 * plane_bytes + 1023u) & ~1023u) : 2 * patch_bytes);      // [3] A chunks (none with ATMEM)
  const uint32_t sW = sS + (ATMEM ? 0 : kSv2Stages) * kATileBytes;      // [6] weight chunk ring / ATMEM: all KC chunks, resident
  const uint32_t sO = sW + (ATMEM ? static_cast<uint32_t>(KC) : kSv2WStages) * kSv2WBytes;     // fp32 staging of HALF a tile (64 rows), row pitch C * 4 + 16 bytes
  const uint32_t o_pitch = static_cast<uint32_t>(C) * 4 + 16;
  float* s_lnw = reinterpret_cast<float*>(sm + (sO - base) + 64 * o_pitch);
  float* s_lnb = s_lnw + 64;
  float* s_mask = s_lnb + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mask + 64);
  constexpr int NST = ATMEM ? kSv2TStages : kSv2Stages;
  uint64_t* patch_full = bars + 0;     // [2] tx
  uint64_t* patch_free = bars + 2;     // [2] 8 (one per builder warp)
  uint64_t* full = bars + 4;           // [NST] one arrival per builder warp of the chunk (8 smem / 4 TMEM)
  uint64_t* empty = bars + 12;         // [NST] commit
  uint64_t* acc_full = bars + 20;      // [2] commit
  uint64_t* acc_free = bars + 22;      // [2] 128
  uint64_t* w_full = bars + 24;        // [6] tx
  uint64_t* w_empty = bars + 30;       // [6] commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);
  uint64_t* plane_full = bars + 26;    // ATMEM [4] tx      (slots of the weight ring the ATMEM variant does not use)
  uint64_t* plane_free = bars + 30;    // ATMEM [4] 8 builder warps
  uint32_t* s_lut = tmem_slot + 4;      // ATMEM: [KC * 8] byte offset of the patch row of K pair q = (ky, ci), ~0u = zero padding of K

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&patch_full[b], 1); mbar_init(&patch_free[b], kSv2Builders / 32);
      mbar_init(&acc_full[b], 1); mbar_init(&acc_free[b], kSv2Epi);
    }
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], ATMEM ? 4 : kSv2Builders / 32); mbar_init(&empty[s], 1); }
    if (ATMEM) {
      mbar_init(&w_full[0], 1);
      for (int pl = 0; pl < 4; ++pl) { mbar_init(&plane_full[pl], 1); mbar_init(&plane_free[pl], kSv2Builders / 32); }
    } else {
      for (int s = 0; s < kSv2WStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    }
    fence_mbar_init();
  }
  if (warp == 13) tmem_alloc(tmem_slot, ATMEM ? 512 : 128);
  if (ATMEM) {
    for (int q = tid; q < KC * 8; q += kSv2Threads) {
      const int kyi = q / Cin, ci = q - kyi * Cin;
      const int ky = stem_ky(kyi < 7 ? kyi : 0);                // plane ky & 3 holds input rows iy0 + (ky & 3) + 4 k; this row is k = oy + (ky >> 2)
      s_lut[q] = q < 7 * Cin ? (ky & 3) * plane_bytes + static_cast<uint32_t>((ci * kSv2PlaneRows + (ky >> 2)) * kStemPatchPitch) : 0xffffffffu;
    }
  }
  for (int i = tid; i < 64; i += kSv2Threads) {
    const bool in = i < C;
    s_lnw[i] = (in && a.ln_w) ? a.ln_w[i] : 1.f;
    s_lnb[i] = (in && a.ln_b) ? a.ln_b[i] : 0.f;
    s_mask[i] = (in && a.mask_token) ? a.mask_token[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int per_img = a.ny * a.nx;

  if (warp < 4) {
    // =============================================== epilogue ===============================================
    const int row = tid;                                        // accumulator row == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const int nch = C >> 2;                                     // 16-byte chunks per token row
    const float inv_c = 1.f / static_cast<float>(C);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const int tb = tile / per_img, tt = tile - tb * per_img;
      const int ty = tt / a.nx, tx = tt - ty * a.nx;
      const int tok0 = (tb * a.Hout + ty * kStemTileH) * a.Wout + tx * kStemTileW;      // token of row 0; row r: + (r>>4) * Wout + (r&15)
      const int my_tok = tok0 + (row >> 4) * a.Wout + (row & 15);
      const bool masked = a.token_mask && a.token_mask[my_tok];
      SV2_TRACE(tid == 0, it, 3);
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      SV2_TRACE(tid == 0, it, 4);
      tc_fence_after();
      // statistics: the whole row in registers ONCE (one TMEM round trip), exact two-pass mean / variance
      const uint32_t trow = tmem + lane_off + ab * 64;
      float mean, rstd;
      {
        float v[64];
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16)
          if (c0 < C) tmem_ld_x16(trow + c0, v + c0);
        tmem_ld_wait();
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) if (c < C) s += v[c];
        mean = s * inv_c;
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) if (c < C) { const float d = v[c] - mean; ss += d * d; }
        rstd = rsqrtf(ss * inv_c + a.eps);
      }
      SV2_TRACE(tid == 0, it, 5);
      // two passes of 64 rows through the half-tile staging buffer.  tma_store: the buffer is 4 token rows x C/32 boxes of
      // [16 tokens x 32 channels] in the 128-byte swizzle (conflict-free for row-per-thread writes) and leaves through TMA tensor
      // stores issued by one thread; otherwise padded rows and (token, chunk) threads that write whole lines.
      const int rl = row & 63;
      long long ecyc_free = 0, ecyc_norm = 0;                   // profiling: SM cycles waiting for the staging buffer / load + normalise + barrier
      const uint32_t srow = sO + static_cast<uint32_t>(rl) * o_pitch;
      const uint32_t sbox = sO + static_cast<uint32_t>(rl) * 128u;      // [channel half][4 token rows x 16 tokens][32 channels], 8 KB per half
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const long long e0 = a.trace ? clock64() : 0;
        if (a.tma_store) {
          if ((tid & 31) == 0 && (tid >> 5) < (C >> 5)) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // my previous store has read the buffer
          named_bar_sync(2, kSv2Epi);
        }
        const long long e1 = a.trace ? clock64() : 0;
        if ((row >> 6) == half) {
#pragma unroll 1
          for (int c0 = 0; c0 < C; c0 += 32) {
            float v[32];
            tmem_ld_x32(trow + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              if (c0 + c4 * 4 >= C) break;
              float4 o;
              if (masked) {
                o = *reinterpret_cast<const float4*>(s_mask + c0 + c4 * 4);
              } else {
                const float4 g = *reinterpret_cast<const float4*>(s_lnw + c0 + c4 * 4), bb = *reinterpret_cast<const float4*>(s_lnb + c0 + c4 * 4);
                o.x = fmaf((v[c4 * 4 + 0] - mean) * rstd, g.x, bb.x);
                o.y = fmaf((v[c4 * 4 + 1] - mean) * rstd, g.y, bb.y);
                o.z = fmaf((v[c4 * 4 + 2] - mean) * rstd, g.z, bb.z);
                o.w = fmaf((v[c4 * 4 + 3] - mean) * rstd, g.w, bb.w);
              }
              const uint32_t dst = a.tma_store ? sbox + static_cast<uint32_t>(c0 >> 5) * 8192u + ((static_cast<uint32_t>(c4) ^ (rl & 7)) << 4)
                                               : srow + (c0 + c4 * 4) * 4;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
            }
          }
          tc_fence_before();
          mbar_arrive(&acc_free[ab]);                           // this row's accumulator is drained (128 arrivals: tile it + 2 may start)
        }
        if (a.tma_store) {
          fence_proxy_async_smem();
          named_bar_sync(1, kSv2Epi);
          if (a.trace) { ecyc_free += e1 - e0; ecyc_norm += clock64() - e1; }
          if ((tid & 31) == 0 && (tid >> 5) < (C >> 5)) {      // lane 0 of warp h: ONE 3-D box [4 token rows x 16 tokens x 32 channels]
            const int h = tid >> 5;
            tma_store_3d(&tmap_out, h * 32, tx * kStemTileW, tb * a.Hout + ty * kStemTileH + half * 4, sO + static_cast<uint32_t>(h) * 8192u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          continue;
        }
        named_bar_sync(1, kSv2Epi);
        for (int idx = tid; idx < 64 * nch; idx += kSv2Epi) {
          const int rr = idx / nch, ch = idx - rr * nch;
          float4 sv;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sv.x), "=f"(sv.y), "=f"(sv.z), "=f"(sv.w)
                       : "r"(sO + static_cast<uint32_t>(rr) * o_pitch + ch * 16));
          const int r = half * 64 + rr;
          const int tok = tok0 + (r >> 4) * a.Wout + (r & 15);
          *reinterpret_cast<float4*>(a.y + static_cast<size_t>(tok) * C + ch * 4) = sv;
        }
        named_bar_sync(2, kSv2Epi);                             // the staging buffer is free again
      }
      SV2_TRACE(tid == 0, it, 6);
      if (ATMEM && a.trace && tid == 0 && it < kTraceTiles) {     // (overrides the builder / MMA wait counters of slots 10, 11)
        a.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + it) * kTracePts + 10] = ecyc_free;
        a.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + it) * kTracePts + 11] = ecyc_norm;
      }
    }
    if (a.tma_store && (tid & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all tensor stores complete before the CTA retires
  } else if (warp < 12) {
    // =============================================== builders ===============================================
    const int bt = tid - kSv2Epi;
    if (ATMEM) {
      // Operand in TENSOR memory: thread = tile row = TMEM lane builds the halves of ITS row in registers and writes them with
      // tcgen05.st (32 packed columns per K chunk) -- no shared-memory stores, no proxy fence, and the MMA reads the operand without
      // touching shared memory.  One hand-off = a STEP of 3 K chunks (the per-chunk version was bound by the hand-off latencies, not by
      // any throughput).  Warps 4..7 / 8..11 own the four lane quarters; the two warps of a quarter take alternate steps.
      const int qt = (warp - 4) & 3, hpar = (warp - 4) >> 2;
      const int r = qt * 32 + lane;
      const uint32_t src_row = (r >> 4) * kStemPatchPitch + (r & 15) * 4 + 12;      // inside a plane consecutive rows are 4 input rows apart
      const int npairs = 7 * Cin;
      const uint32_t t_a = tmem + kSv2TAcol + (static_cast<uint32_t>(qt * 32) << 16);
      const int n_steps = (KC + kSv2KS - 1) / kSv2KS;
      const __half2 k1024 = __half2half2(__ushort_as_half(static_cast<unsigned short>(0x6400)));
      int it = 0;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        SV2_TRACE(bt == 0, it, 0);
        const uint32_t patch = sP + src_row;
        long long cyc_wait = 0;                                 // profiling: SM cycles the leader waited for a free operand slot
        for (int st = 0; st < n_steps; ++st, ++g) {
          // planes this step reads: pairs [24 st, 24 st + 24) are (kyi, ci) with kyi = q / Cin; the plane index rises with q
          const int q0 = st * kSv2KS * 8;
          const int q1 = min(q0 + kSv2KS * 8, npairs) - 1;
          const int p_lo = stem_ky(min(q0 / Cin, 6)) & 3, p_hi = q1 >= q0 ? (stem_ky(q1 / Cin) & 3) : -1;
          // a plane is released by every builder warp once the warp is past the last step that reads it (planes 0..2 end with
          // kyi = 1, 3, 5; plane 3 with the last pair), so its reload for the NEXT tile overlaps the rest of this tile
          auto release = [&]() {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
              for (int pl = 0; pl < 4; ++pl) {
                const int last = (pl < 3 ? (2 * pl + 2) * Cin : npairs) - 1;
                if (last >= q0 && last < q0 + kSv2KS * 8) mbar_arrive(&plane_free[pl]);
              }
            }
          };
          if (static_cast<int>(g & 1) != hpar) { release(); continue; }
          for (int pl = p_lo; pl <= p_hi; ++pl) mbar_wait(&plane_full[pl], it & 1);
          if (st == 0) SV2_TRACE(bt == 0, it, 1);
          const uint32_t s = g % kSv2TStages, ph = (g / kSv2TStages) & 1;
          const long long c0 = a.trace ? clock64() : 0;
          mbar_wait(&empty[s], ph ^ 1);
          if (a.trace) cyc_wait += clock64() - c0;
          tc_fence_after();
          const uint32_t* lut = s_lut + st * kSv2KS * 8;
#pragma unroll 1
          for (int atom = 0; atom < kSv2KS; ++atom) {
            if (st * kSv2KS + atom >= KC) break;
            uint32_t o[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t off = lut[atom * 8 + j];             // warp-uniform: one broadcast load
              uint32_t w0 = 0, w1 = 0;
              if (off != 0xffffffffu) {
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(patch + off));
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(patch + off + 4));
              }
              // u8 -> fp16 exactly: bytes (b, 0x64) form the half 1024 + b; subtract 1024 (zero bytes give exact zeros)
              uint32_t p0 = __byte_perm(w0, 0x64646464u, 0x4140), p1 = __byte_perm(w0, 0x64646464u, 0x4342);
              uint32_t p2 = __byte_perm(w1, 0x64646464u, 0x4140), p3 = __byte_perm(w1, 0x64646464u, 0x4342);
              const __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&p0), k1024);
              const __half2 h1 = __hsub2(*reinterpret_cast<__half2*>(&p1), k1024);
              const __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&p2), k1024);
              const __half2 h3 = __hsub2(*reinterpret_cast<__half2*>(&p3), k1024);
              o[4 * j + 0] = *reinterpret_cast<const uint32_t*>(&h0); o[4 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
              o[4 * j + 2] = *reinterpret_cast<const uint32_t*>(&h2); o[4 * j + 3] = *reinterpret_cast<const uint32_t*>(&h3);
            }
            tmem_st_x32(t_a + s * (kSv2KS * 32) + atom * 32, o);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[s]);
          release();
        }
        SV2_TRACE(bt == 0, it, 2);
        (void)cyc_wait;
      }
    } else {
    const int j = bt & 7;                                       // 16-byte chunk of the 128-byte operand row = one (ky, ci) pair
    const int r0 = bt >> 3;                                     // rows r0 + 32 i
    const int npairs = 7 * Cin;
    uint32_t src_off[4], dst_off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + 32 * i;
      src_off[i] = (r >> 4) * (4 * kStemPatchPitch) + (r & 15) * 4 + 12;
      dst_off[i] = sw128_offset(r, j);
    }
    const __half2 k1024 = __half2half2(__ushort_as_half(static_cast<unsigned short>(0x6400)));
    int it = 0;
    uint32_t g = 0;                                             // running K-chunk counter (ring position)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      SV2_TRACE(bt == 0, it, 0);
      mbar_wait(&patch_full[b], (it >> 1) & 1);
      SV2_TRACE(bt == 0, it, 1);
      const uint32_t patch = sP + b * patch_bytes;
      int q = j, ky = j / Cin, ci = j - ky * Cin;               // (ky, ci) pair of this thread's chunk in K chunk 0
      long long cyc_wait = 0, cyc_fence = 0;                    // profiling: SM cycles the leader spent waiting for a free slot / publishing
      for (int kc = 0; kc < KC; ++kc, ++g) {
        const uint32_t s = g % kSv2Stages, ph = (g / kSv2Stages) & 1;
        const long long c0 = a.trace ? clock64() : 0;
        mbar_wait(&empty[s], ph ^ 1);
        if (a.trace) cyc_wait += clock64() - c0;
        const uint32_t tile_a = sS + s * kATileBytes;
        const bool qv = q < npairs;
        const uint32_t prow = patch + (ci * kStemPatchRows + stem_ky(ky < 7 ? ky : 0)) * kStemPatchPitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
          if (qv) {
            uint32_t w0, w1;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(prow + src_off[i]));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(prow + src_off[i] + 4));
            // u8 -> fp16 exactly: bytes (b, 0x64) form the half 1024 + b; subtract 1024
            uint32_t p0 = __byte_perm(w0, 0x64646464u, 0x4140), p1 = __byte_perm(w0, 0x64646464u, 0x4342);
            uint32_t p2 = __byte_perm(w1, 0x64646464u, 0x4140), p3 = __byte_perm(w1, 0x64646464u, 0x4342);
            const __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&p0), k1024);
            const __half2 h1 = __hsub2(*reinterpret_cast<__half2*>(&p1), k1024);
            const __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&p2), k1024);
            const __half2 h3 = __hsub2(*reinterpret_cast<__half2*>(&p3), k1024);
            o0 = *reinterpret_cast<const uint32_t*>(&h0); o1 = *reinterpret_cast<const uint32_t*>(&h1);
            o2 = *reinterpret_cast<const uint32_t*>(&h2); o3 = *reinterpret_cast<const uint32_t*>(&h3);
          }
          st_smem_16B(tile_a + dst_off[i], o0, o1, o2, o3);
        }
        const long long c1 = a.trace ? clock64() : 0;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);                   // one arrival per warp (same-address arrivals serialise)
        if (a.trace) cyc_fence += clock64() - c1;
        q += 8; ci += 8;
        while (ci >= Cin) { ci -= Cin; ++ky; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&patch_free[b]);
      SV2_TRACE(bt == 0, it, 2);
      if (a.trace && bt == 0 && it < kTraceTiles) {
        a.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + it) * kTracePts + 10] = cyc_wait;
        a.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + it) * kTracePts + 11] = cyc_fence;
      }
    }
    }  // !ATMEM
  } else if (warp == 12) {
    // =============================================== producer ===============================================
    if (ATMEM && lane == 0 && static_cast<int>(blockIdx.x) < n_tiles) {
      tma_prefetch_desc(&tmap_in);
      const uint32_t plane_box_bytes = static_cast<uint32_t>(Cin) * kSv2PlaneRows * kStemPatchPitch;
      const uint32_t w_bytes = static_cast<uint32_t>(C) * 128;
      mbar_arrive_expect_tx(&w_full[0], w_bytes * KC);           // all weight chunks once, resident for the CTA's lifetime
      for (int kc = 0; kc < KC; ++kc)
        bulk_g2s(sm + (sW - base) + kc * kSv2WBytes, a.wp + static_cast<size_t>(kc) * C * 64, w_bytes, &w_full[0]);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int tb = tile / per_img, tt = tile - tb * per_img;
        const int ty = tt / a.nx, tx = tt - ty * a.nx;
        // four row-phase planes (TMA row stride 4): plane pl is free again as soon as the builders are past the kernel rows that
        // read it, long before the tile ends, so these loads overlap the previous tile
        for (int pl = 0; pl < 4; ++pl) {
          mbar_wait(&plane_free[pl], (it & 1) ^ 1);
          mbar_arrive_expect_tx(&plane_full[pl], plane_box_bytes);
          tma_load_3d(sP + pl * plane_bytes, &tmap_in, tx * kStemTileW * 4 - 16, ty * kStemTileH * 4 - 3 + pl, tb * Cin, &plane_full[pl]);
        }
      }
    }
    if (!ATMEM && lane == 0 && static_cast<int>(blockIdx.x) < n_tiles) {
      tma_prefetch_desc(&tmap_in);
      const uint32_t box_bytes = static_cast<uint32_t>(Cin) * kStemPatchRows * kStemPatchPitch;
      const uint32_t w_bytes = static_cast<uint32_t>(C) * 128;
      auto load_patch = [&](int it2, int tile2) {
        const int b = it2 & 1;
        const int tb = tile2 / per_img, tt = tile2 - tb * per_img;
        const int ty = tt / a.nx, tx = tt - ty * a.nx;
        mbar_wait(&patch_free[b], ((it2 >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&patch_full[b], box_bytes);
        tma_load_3d(sP + b * patch_bytes, &tmap_in, tx * kStemTileW * 4 - 16, ty * kStemTileH * 4 - 3, tb * Cin, &patch_full[b]);
      };
      load_patch(0, blockIdx.x);
      int it = 0;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        if (tile + static_cast<int>(gridDim.x) < n_tiles) load_patch(it + 1, tile + gridDim.x);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const uint32_t s = g % kSv2WStages, ph = (g / kSv2WStages) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&w_full[s], w_bytes);
          bulk_g2s(sm + (sW - base) + s * kSv2WBytes, a.wp + static_cast<size_t>(kc) * C * 64, w_bytes, &w_full[s]);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================================== MMA issuer ===============================================
    if (ATMEM && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, C, 0);
      const int n_steps = (KC + kSv2KS - 1) / kSv2KS;
      if (static_cast<int>(blockIdx.x) < n_tiles) mbar_wait(&w_full[0], 0);
      int it = 0;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        SV2_TRACE(true, it, 7);
        mbar_wait(&acc_free[ab], ((it >> 1) & 1) ^ 1);          // the epilogue of tile it - 2 has drained this accumulator
        SV2_TRACE(true, it, 8);
        tc_fence_after();
        const uint32_t t_acc = tmem + ab * 64;
        long long cyc_full = 0;                                 // profiling: SM cycles waiting for the builders
        for (int st = 0; st < n_steps; ++st, ++g) {
          const uint32_t s = g % kSv2TStages, ph = (g / kSv2TStages) & 1;
          const long long c0 = a.trace ? clock64() : 0;
          mbar_wait(&full[s], ph);
          if (a.trace) cyc_full += clock64() - c0;
          tc_fence_after();
          for (int atom = 0; atom < kSv2KS; ++atom) {
            const int kc = st * kSv2KS + atom;
            if (kc >= KC) break;
            const uint32_t ta = tmem + kSv2TAcol + s * (kSv2KS * 32) + atom * 32, tw = sW + kc * kSv2WBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(t_acc, ta + k * 8, umma_desc_sw128(tw + k * 32), idesc, (kc | k) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[ab]);
        SV2_TRACE(true, it, 9);
        (void)cyc_full;
      }
    }
    if (!ATMEM && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, C, 0);
      int it = 0;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        SV2_TRACE(true, it, 7);
        mbar_wait(&acc_free[ab], ((it >> 1) & 1) ^ 1);          // the epilogue of tile it - 2 has drained this accumulator
        SV2_TRACE(true, it, 8);
        tc_fence_after();
        const uint32_t t_acc = tmem + ab * 64;
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const uint32_t s = g % NST, ph = (g / NST) & 1;
          const uint32_t ws = g % kSv2WStages, wph = (g / kSv2WStages) & 1;
          mbar_wait(&w_full[ws], wph);
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t ta = sS + s * kATileBytes, tw = sW + ws * kSv2WBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(t_acc, umma_desc_sw128(ta + k * 32), umma_desc_sw128(tw + k * 32), idesc, (kc | k) != 0);
          umma_commit(&empty[s]);
          umma_commit(&w_empty[ws]);
        }
        umma_commit(&acc_full[ab]);
        SV2_TRACE(true, it, 9);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, ATMEM ? 512 : 128);
}

}  // namespace rvt
