// Callers / data formats either side of the hot path.  All of it is integer / byte / copy work bound by
// HBM bandwidth: coalesced 16-byte accesses, grid-stride loops sized to the SM count, no tensor cores.
//
//  f3  harness glue inside the L-step sequence (modules/utils/detection.py):
//        state_reset_kernel   RNNStates.reset -> recursive_reset: state[mask] = 0 in place          (:96-113, modules/detection.py:117,217)
//        gather_rows_kernel   BackboneFeatureSelector: cat_t( feat_t[selected_indices_t] )            (:24-46)
//  f4  preprocessing neighbours of the voxelizer (scripts/genx/preprocess_dataset.py, data/utils/representations.py):
//        downsample2_nearest_kernel   downsample_ev_repr(scale 0.5, 'nearest-exact')                   (preprocess_dataset.py:467-477,525-528)
//        cummax_*_kernel              H5Reader._correct_time: t[i] = max(t[0..i]) (monotone fix-up)     (:163-172)
//        searchsorted_kernel          window boundaries np.searchsorted(ev_ts, ts, side)               (:511-516)
//        mixed_density_*_kernel       MixedDensityEventStack.construct                                  (representations.py:130-218)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvt {

// ---------------------------------------------------------------------------------------------------------------------------
// f3
// ---------------------------------------------------------------------------------------------------------------------------
// h, c: [batch, per_sample] fp32 (per_sample % 4 == 0, 16-byte aligned); mask: u8 [batch]; rows with mask != 0 are zeroed.
__global__ void __launch_bounds__(256) state_reset_kernel(float* __restrict__ h, float* __restrict__ c,
                                                          const uint8_t* __restrict__ mask, int batch, int64_t per4) {
  const int64_t total = static_cast<int64_t>(batch) * per4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int b = static_cast<int>(i / per4);
    if (mask[b]) {
      reinterpret_cast<float4*>(h)[i] = z;
      if (c) reinterpret_cast<float4*>(c)[i] = z;
    }
  }
}

// dst[j, :] = src[idx[j], :] for j < n_idx (idx[j] < 0 or >= n_src_rows -> zeros).  row4 = row length in float4.
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n_idx,
                                                          int64_t n_src_rows, int64_t row4, float* __restrict__ dst) {
  const int64_t total = static_cast<int64_t>(n_idx) * row4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t j = i / row4, k = i - j * row4;
    const int64_t s = idx[j];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0 && s < n_src_rows) v = __ldcs(reinterpret_cast<const float4*>(src) + s * row4 + k);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// f4
// ---------------------------------------------------------------------------------------------------------------------------
// torch.nn.functional.interpolate(x, scale_factor=0.5, mode='nearest-exact') on [C, H, W] bytes (uint8 or int8 alike):
// out[c, y, x] = in[c, min(2y+1, H-1), min(2x+1, W-1)], Ho = H / 2, Wo = W / 2.
__global__ void __launch_bounds__(256) downsample2_nearest_kernel(const uint8_t* __restrict__ in, int C, int H, int W,
                                                                  uint8_t* __restrict__ out, int Ho, int Wo) {
  const int64_t total = static_cast<int64_t>(C) * Ho * Wo;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x = static_cast<int>(i % Wo);
    const int64_t r = i / Wo;
    const int y = static_cast<int>(r % Ho), c = static_cast<int>(r / Ho);
    const int sy = min(2 * y + 1, H - 1), sx = min(2 * x + 1, W - 1);
    out[i] = __ldg(in + (static_cast<int64_t>(c) * H + sy) * W + sx);
  }
}

// ---- running maximum of an int64 array (three passes: chunk maxima, scan of the maxima, in-chunk scan with carry) ----
constexpr int kScanChunk = 4096;            // elements per CTA (256 threads x 16)

__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int64_t u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
  return v;
}

__global__ void __launch_bounds__(256) cummax_chunk_max_kernel(const int64_t* __restrict__ t, int64_t n, int64_t* __restrict__ part) {
  __shared__ int64_t s[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanChunk;
  int64_t m = INT64_MIN;
  for (int k = threadIdx.x; k < kScanChunk; k += 256) {
    const int64_t i = base + k;
    if (i < n) { const int64_t v = t[i]; m = v > m ? v : m; }
  }
  m = warp_max_i64(m);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = s[w] > m ? s[w] : m;
    part[blockIdx.x] = m;
  }
}

// exclusive running maximum of the chunk maxima (single CTA; `floor` = the reference's initial time_last = 0)
__global__ void __launch_bounds__(1024) cummax_scan_parts_kernel(int64_t* __restrict__ part, int n_parts, int64_t floor_v) {
  __shared__ int64_t s_w[32];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = floor_v;
  __syncthreads();
  for (int base = 0; base < n_parts; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n_parts ? part[i] : INT64_MIN;
    int64_t incl = v;                                      // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl = u > incl ? u : incl;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    int64_t pre = s_carry;                                 // carry + maxima of the previous warps
    for (int w = 0; w < (threadIdx.x >> 5); ++w) pre = s_w[w] > pre ? s_w[w] : pre;
    int64_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if ((threadIdx.x & 31) == 0) excl = INT64_MIN;
    excl = excl > pre ? excl : pre;
    if (i < n_parts) part[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) { const int64_t tot = incl > pre ? incl : pre; s_carry = tot; }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) cummax_apply_kernel(int64_t* __restrict__ t, int64_t n, const int64_t* __restrict__ part) {
  __shared__ int64_t s_w[8];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanChunk + static_cast<int64_t>(threadIdx.x) * 16;
  int64_t v[16];
  int64_t m = INT64_MIN;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    v[e] = (base + e < n) ? t[base + e] : INT64_MIN;
    m = v[e] > m ? v[e] : m;
    v[e] = m;                                               // inclusive running max of this thread's 16 elements
  }
  int64_t incl = m;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl = u > incl ? u : incl;
  }
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
  __syncthreads();
  int64_t pre = part[blockIdx.x];
  for (int w = 0; w < (threadIdx.x >> 5); ++w) pre = s_w[w] > pre ? s_w[w] : pre;
  int64_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if ((threadIdx.x & 31) == 0) excl = INT64_MIN;
  pre = excl > pre ? excl : pre;
#pragma unroll
  for (int e = 0; e < 16; ++e)
    if (base + e < n) t[base + e] = v[e] > pre ? v[e] : pre;
}

// out[q] = first index i in the sorted array a[0..n) with a[i] >= v (left) / a[i] > v (right)   (np.searchsorted)
__global__ void __launch_bounds__(128) searchsorted_kernel(const int64_t* __restrict__ a, int64_t n, const int64_t* __restrict__ q,
                                                           int64_t nq, int right, int64_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int64_t v = q[i];
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    const int64_t am = __ldg(a + mid);
    const bool go_right = right ? (am <= v) : (am < v);
    if (go_right) lo = mid + 1; else hi = mid;
  }
  out[i] = lo;
}

// ---- MixedDensityEventStack (representations.py:130-218) ----
//   t_norm = clamp(f32(t - t0) / f32(max(t1 - t0, 1)), lo, hi);  t_idx = floor(clamp(bins - log(t_norm)/log(1/2), 0))
//   rep[t_idx, y, x] += 2*pol - 1  (int8, wraps);  rep[i] = int8(sum_{j<=i} rep[j]);  clamp(+-cutoff)
// The fp32 log is not reproduced on the device: t_idx = #{k : t_norm >= thr[k]} with the bins-1 fp32 thresholds found on the host
// by bisection over the reference's own torch expression (rvt_b200/representations.py), so the binning is bit-identical.
__global__ void __launch_bounds__(256) mixed_density_accumulate_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ y,
                                                                       const int64_t* __restrict__ pol, const int64_t* __restrict__ t,
                                                                       int64_t n, int bins, int H, int W, float lo, float hi,
                                                                       const float* __restrict__ thr, int32_t* __restrict__ counts,
                                                                       int* __restrict__ err) {
  const int64_t t0 = __ldg(t), t1 = __ldg(t + n - 1);
  const int64_t dt = t1 - t0;
  const float denom = __ll2float_rn(dt > 1 ? dt : 1);
  const int64_t hw = static_cast<int64_t>(H) * W;
  if (blockIdx.x == 0 && threadIdx.x == 0 && dt < 0) atomicOr(err, 1);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t xi = __ldcs(x + i), yi = __ldcs(y + i), pi = __ldcs(pol + i), ti = __ldcs(t + i);
    if (pi < 0 || pi > 1) { atomicOr(err, 2); continue; }
    if (xi < 0 || xi >= W || yi < 0 || yi >= H) { atomicOr(err, 4); continue; }
    float q = __fdiv_rn(__ll2float_rn(ti - t0), denom);
    q = fminf(fmaxf(q, lo), hi);
    int b = 0;
    for (int k = 0; k < bins - 1; ++k) b += (q >= __ldg(thr + k)) ? 1 : 0;
    atomicAdd(counts + (b * hw + yi * W + xi), pi ? 1 : -1);
  }
}

// per pixel: running sum over the bins, int8 wrap, clamp; the scratch is re-zeroed
__global__ void __launch_bounds__(256) mixed_density_finalize_kernel(int32_t* __restrict__ counts, int8_t* __restrict__ out, int bins,
                                                                     int64_t hw, int cutoff) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  int run = 0;
  for (int b = 0; b < bins; ++b) {
    run += counts[b * hw + p];
    counts[b * hw + p] = 0;
    int v = static_cast<int>(static_cast<int8_t>(run & 0xFF));
    if (cutoff >= 0) v = v > cutoff ? cutoff : (v < -cutoff ? -cutoff : v);
    out[b * hw + p] = static_cast<int8_t>(v);
  }
}

}  // namespace rvt
