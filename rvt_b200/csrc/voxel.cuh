// Event stream -> StackedHistogram (reference: data/utils/representations.py:76-121).
// Integer / byte work, HBM + L2-atomic bound: 32 B/event in (the reference's int64 x,y,p,t
// interface), 1 B/bin out.  Bit-exact contract:
//   t_idx = min(floor(f32(t - t0) / f32(max(t1 - t0, 1)) * f32(bins)), bins-1)   (IEEE fp32, RN)
//   idx   = x + W*y + H*W*t_idx + bins*H*W*pol
//   fastmode: out = min(count mod 256, cutoff);  else out = clamp(int16(count mod 65536), 0, cutoff)
// Counts accumulate in a u32 scratch image (CUDA has no byte atomics; the mod is applied in
// the finalize pass, which also re-zeroes the scratch so it is clean for the next window).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvt {

// One event -> bin index (0xFFFFFFFF if rejected).  n_out < 2^32 is checked on the host.
__device__ __forceinline__ uint32_t voxel_bin(int64_t xi, int64_t yi, int64_t pi, int64_t ti, int64_t t0, float denom, float fb,
                                              int bins, int H, int W, int64_t hw, int* err) {
  const float q = __fdiv_rn(__ll2float_rn(ti - t0), denom);
  float f = floorf(__fmul_rn(q, fb));
  f = fminf(f, fb - 1.0f);
  if (!(f >= 0.0f)) { atomicOr(err, 1); return 0xFFFFFFFFu; }   // t < t[0]: time not sorted
  if (pi < 0 || pi > 1) { atomicOr(err, 2); return 0xFFFFFFFFu; }
  if (xi < 0 || xi >= W || yi < 0 || yi >= H) { atomicOr(err, 4); return 0xFFFFFFFFu; }
  return static_cast<uint32_t>(xi + W * yi + hw * static_cast<int64_t>(f) + bins * hw * pi);
}

// Adaptive warp aggregation: events are time sorted, so a hot pixel shows up as equal bin indices in
// neighbouring lanes.  One shuffle + vote decides per warp whether to pay for match_any (one atomic
// per distinct bin) or to issue plain reductions (the spatially-random common case, where match_any's
// MIO cost dominated: profiles/ncu_r01.md).
__device__ __forceinline__ void voxel_commit(uint32_t idx, uint32_t* __restrict__ counts) {
  const uint32_t nb = __shfl_down_sync(0xffffffffu, idx, 1);
  const bool dup = idx != 0xFFFFFFFFu && idx == nb && (threadIdx.x & 31) != 31;
  if (__any_sync(0xffffffffu, dup)) {
    const unsigned peers = __match_any_sync(0xffffffffu, idx);
    if (idx != 0xFFFFFFFFu && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + idx, static_cast<uint32_t>(__popc(peers)));
  } else if (idx != 0xFFFFFFFFu) {
    atomicAdd(counts + idx, 1u);
  }
}

__global__ void __launch_bounds__(256) voxel_accumulate_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ y,
                                                               const int64_t* __restrict__ pol,
                                                               const int64_t* __restrict__ t, int64_t n, int bins,
                                                               int H, int W, uint32_t* __restrict__ counts,
                                                               int* __restrict__ err) {
  const int64_t t0 = __ldg(t), t1 = __ldg(t + n - 1);
  const int64_t dt = t1 - t0;
  const float denom = __ll2float_rn(dt > 1 ? dt : 1);
  const float fb = static_cast<float>(bins);
  const int64_t hw = static_cast<int64_t>(H) * W;
  if (blockIdx.x == 0 && threadIdx.x == 0 && dt < 0) atomicOr(err, 1);  // time not sorted

  // two consecutive events per thread per iteration: 16-byte loads from each of the four arrays
  const int64_t npairs = n >> 1;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t pairs_round = (npairs + 31) & ~static_cast<int64_t>(31);     // keep warps converged
  const longlong2* x2 = reinterpret_cast<const longlong2*>(x);
  const longlong2* y2 = reinterpret_cast<const longlong2*>(y);
  const longlong2* p2 = reinterpret_cast<const longlong2*>(pol);
  const longlong2* t2 = reinterpret_cast<const longlong2*>(t);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs_round; i += stride) {
    uint32_t ia = 0xFFFFFFFFu, ib = 0xFFFFFFFFu;
    if (i < npairs) {
      const longlong2 xv = __ldcs(x2 + i), yv = __ldcs(y2 + i), pv = __ldcs(p2 + i), tv = __ldcs(t2 + i);
      ia = voxel_bin(xv.x, yv.x, pv.x, tv.x, t0, denom, fb, bins, H, W, hw, err);
      ib = voxel_bin(xv.y, yv.y, pv.y, tv.y, t0, denom, fb, bins, H, W, hw, err);
    }
    voxel_commit(ia, counts);
    voxel_commit(ib, counts);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd tail event
    const uint32_t il = voxel_bin(x[n - 1], y[n - 1], pol[n - 1], t[n - 1], t0, denom, fb, bins, H, W, hw, err);
    if (il != 0xFFFFFFFFu) atomicAdd(counts + il, 1u);
  }
}

// out[j] = clamp(wrap(counts[j])); counts[j] = 0.  16 bins per thread (4 x uint4 in, 1 x uint4 out).
__global__ void __launch_bounds__(256) voxel_finalize_kernel(uint32_t* __restrict__ counts, uint8_t* __restrict__ out,
                                                             int64_t n_out, int cutoff, int fastmode) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
  if (i >= n_out) return;
  auto fin = [&](uint32_t c) -> uint32_t {
    int v;
    if (fastmode) v = static_cast<int>(c & 0xFFu);
    else { v = static_cast<int>(static_cast<int16_t>(c & 0xFFFFu)); v = v < 0 ? 0 : v; }
    return static_cast<uint32_t>(v > cutoff ? cutoff : v);
  };
  if (i + 16 <= n_out) {
    uint4* cp = reinterpret_cast<uint4*>(counts + i);
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 c = cp[q];
      o[q] = fin(c.x) | (fin(c.y) << 8) | (fin(c.z) << 16) | (fin(c.w) << 24);
      cp[q] = make_uint4(0, 0, 0, 0);
    }
    *reinterpret_cast<uint4*>(out + i) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
    for (int64_t j = i; j < n_out; ++j) { out[j] = static_cast<uint8_t>(fin(counts[j])); counts[j] = 0; }
  }
}

}  // namespace rvt
