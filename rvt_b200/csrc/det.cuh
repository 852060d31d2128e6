// Detection head neighbours of the backbone: element-wise / small kernels around the conv GEMMs of
// YOLOPAFPN + YOLOXHead (models/detection/yolox_extension/models/yolo_pafpn.py:109-139, yolox/models/yolo_head.py:165-290)
// and the post-processing (yolox/utils/boxes.py:32-76).  The convolutions themselves are gemm_fused_kernel<LD_CONV, EP_F16>
// with BatchNorm folded into weight + bias and SiLU in the epilogue.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvt {

// fp32 [B, H, W, C] with arbitrary (element) strides -> fp16 channel slice of a wider channels-last buffer:
// dst[(b*H + y)*W + x][0..C) with pixel pitch `dpitch` (th.cat along channels = writing slices of one buffer)
__global__ void __launch_bounds__(256) cast_slice_f16_kernel(const float* __restrict__ src, int64_t sb, int64_t sy, int64_t sx,
                                                             int64_t sc, int B, int H, int W, int C, __half* __restrict__ dst,
                                                             int dpitch) {
  const int64_t total = static_cast<int64_t>(B) * H * W * C;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % C);
    int64_t r = i / C;
    const int x = static_cast<int>(r % W); r /= W;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    dst[((static_cast<int64_t>(b) * H + y) * W + x) * dpitch + c] = __float2half_rn(src[b * sb + y * sy + x * sx + c * sc]);
  }
}

// nearest-exact x2 upsample (yolo_pafpn.py:47) of an fp16 channels-last slice into another slice: dst[b, y, x] = src[b, y/2, x/2]
__global__ void __launch_bounds__(256) upsample2_slice_f16_kernel(const __half* __restrict__ src, int spitch, int B, int H, int W, int C,
                                                                  __half* __restrict__ dst, int dpitch) {
  const int c8n = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * (2 * H) * (2 * W) * c8n;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c8 = static_cast<int>(i % c8n);
    int64_t r = i / c8n;
    const int x = static_cast<int>(r % (2 * W)); r /= (2 * W);
    const int y = static_cast<int>(r % (2 * H));
    const int b = static_cast<int>(r / (2 * H));
    const uint4 v = *reinterpret_cast<const uint4*>(src + ((static_cast<int64_t>(b) * H + (y >> 1)) * W + (x >> 1)) * spitch + c8 * 8);
    *reinterpret_cast<uint4*>(dst + ((static_cast<int64_t>(b) * 2 * H + y) * 2 * W + x) * dpitch + c8 * 8) = v;
  }
}

// YOLOXHead inference output of one level + decode_outputs (yolo_head.py:228-232,241-290):
//   out[b, a0 + y*W + x] = [(reg_xy + (x, y)) * stride, exp(reg_wh) * stride, sigmoid(obj), sigmoid(cls_0..)]
// regobj: f16 [rows, rp] columns [reg(4) | obj(1)], cls: f16 [rows, cp] columns [cls(nc)], rows = (b*H + y)*W + x.
__global__ void __launch_bounds__(256) yolox_decode_kernel(const __half* __restrict__ regobj, int rp, const __half* __restrict__ cls, int cp,
                                                           int B, int H, int W, int nc, float stride_px, int a0, int a_total,
                                                           float* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(B) * H * W;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  const int y = static_cast<int>((i / W) % H);
  const int b = static_cast<int>(i / (static_cast<int64_t>(W) * H));
  const __half* r = regobj + i * rp;
  float* o = out + (static_cast<int64_t>(b) * a_total + a0 + y * W + x) * (5 + nc);
  o[0] = (__half2float(r[0]) + x) * stride_px;
  o[1] = (__half2float(r[1]) + y) * stride_px;
  o[2] = expf(__half2float(r[2])) * stride_px;
  o[3] = expf(__half2float(r[3])) * stride_px;
  o[4] = 1.0f / (1.0f + expf(-__half2float(r[4])));
  const __half* c = cls + i * cp;
  for (int k = 0; k < nc; ++k) o[5 + k] = 1.0f / (1.0f + expf(-__half2float(c[k])));
}

// postprocess (boxes.py:32-76) of one image per CTA: corner boxes, class_conf / class_pred = max / argmax over classes,
// keep obj * class_conf >= conf_thre, sort by that score (descending), greedy per-class NMS (torchvision batched_nms),
// detections (x1, y1, x2, y2, obj_conf, class_conf, class_pred) in score order.  A (anchors per image) <= 8192.
constexpr int kNmsMax = 8192;
__global__ void __launch_bounds__(1024) yolox_postprocess_kernel(const float* __restrict__ pred, int A, int nc, float conf_thre,
                                                                 float nms_thre, float* __restrict__ det, int* __restrict__ count) {
  extern __shared__ uint8_t nms_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(nms_smem);      // [npow2] (score bits << 32) | (0xFFFFFFFF - index)
  uint8_t* supp = reinterpret_cast<uint8_t*>(keys + kNmsMax);                       // [kNmsMax]
  __shared__ int s_m;
  __shared__ float s_box[6];
  const int img = blockIdx.x, tid = threadIdx.x;
  const float* p = pred + static_cast<int64_t>(img) * A * (5 + nc);
  int npow2 = 1;
  while (npow2 < A) npow2 <<= 1;
  if (tid == 0) s_m = 0;
  __syncthreads();
  for (int i = tid; i < npow2; i += 1024) {
    unsigned long long key = 0ull;
    if (i < A) {
      const float* q = p + static_cast<int64_t>(i) * (5 + nc);
      float cmax = q[5];
      for (int k = 1; k < nc; ++k) cmax = fmaxf(cmax, q[5 + k]);
      const float score = q[4] * cmax;
      if (score >= conf_thre) {
        key = (static_cast<unsigned long long>(__float_as_uint(score)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - i);
        atomicAdd(&s_m, 1);
      }
    }
    keys[i] = key;
    if (i < kNmsMax) supp[i] = 0;
  }
  __syncthreads();
  // bitonic sort, descending (scores are positive floats: their bit patterns order like the values; ties: lower index first)
  for (int k = 2; k <= npow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool desc = (i & k) == 0;
          if (desc ? a < b : a > b) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  const int M = s_m;
  auto load_box = [&](int idx, float* bx) {
    const float* q = p + static_cast<int64_t>(idx) * (5 + nc);
    const float cx = q[0], cy = q[1], w = q[2], h = q[3];
    bx[0] = cx - w / 2; bx[1] = cy - h / 2; bx[2] = cx + w / 2; bx[3] = cy + h / 2;
    float cmax = q[5]; int cls = 0;
    for (int k = 1; k < nc; ++k) if (q[5 + k] > cmax) { cmax = q[5 + k]; cls = k; }
    bx[4] = cmax; bx[5] = static_cast<float>(cls);
  };
  int kept = 0;
  for (int i = 0; i < M; ++i) {
    if (supp[i]) continue;                      // uniform: every thread reads the same shared byte after the barrier below
    const int idx_i = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[i] & 0xFFFFFFFFull));
    if (tid == 0) {
      float bx[6];
      load_box(idx_i, bx);
      const float* q = p + static_cast<int64_t>(idx_i) * (5 + nc);
      float* d = det + (static_cast<int64_t>(img) * A + kept) * 7;
      d[0] = bx[0]; d[1] = bx[1]; d[2] = bx[2]; d[3] = bx[3]; d[4] = q[4]; d[5] = bx[4]; d[6] = bx[5];
#pragma unroll
      for (int e = 0; e < 6; ++e) s_box[e] = bx[e];
    }
    ++kept;
    __syncthreads();
    const float ax1 = s_box[0], ay1 = s_box[1], ax2 = s_box[2], ay2 = s_box[3], acls = s_box[5];
    const float aarea = (ax2 - ax1) * (ay2 - ay1);
    for (int j = i + 1 + tid; j < M; j += 1024) {
      if (supp[j]) continue;
      float bx[6];
      load_box(static_cast<int>(0xFFFFFFFFu - static_cast<unsigned>(keys[j] & 0xFFFFFFFFull)), bx);
      if (bx[5] != acls) continue;
      const float iw = fmaxf(fminf(ax2, bx[2]) - fmaxf(ax1, bx[0]), 0.f), ih = fmaxf(fminf(ay2, bx[3]) - fmaxf(ay1, bx[1]), 0.f);
      const float inter = iw * ih;
      const float iou = inter / (aarea + (bx[2] - bx[0]) * (bx[3] - bx[1]) - inter);
      if (iou > nms_thre) supp[j] = 1;
    }
    __syncthreads();
  }
  if (tid == 0) count[img] = kept;
}

}  // namespace rvt
