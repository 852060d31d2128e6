/* rvt_b200 — C-ABI of the B200-native (sm_100a) RVT hot path.
 *
 * The reference (uzh-rpg/RVT) is pure Python/PyTorch and has no FFI of its own; every entry
 * point below replaces the body of one reference function on the hot path
 * and is what a ctypes/cffi binding on the reference side would bind.
 * Plain pointers and sizes only: all pointers are DEVICE pointers unless noted, buffers are
 * caller-allocated (no allocation inside), `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success or a cudaError_t / negative argument-error code;
 * rvt_error_string() renders it.  Launches are asynchronous on `stream`.
 *
 * Tensors are channels-last ("NHWC", token-major): x[b][y][x][c].  fp32 residual stream and
 * LSTM states; fp16 tensor-core operands with fp32 accumulation.
 */
#ifndef RVT_B200_H_
#define RVT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVT_B200_ABI_VERSION 1

int rvt_abi_version(void);
const char* rvt_error_string(int code);

/* ---- tiling contract shared with the host-side weight packer (rvt_b200/packing.py) ---- */
/* N-tile (columns per CTA) used for a Linear with n_total output and k input features. */
int rvt_tile_n(int n_total, int k);
/* 1 when rvt_partition_attention runs as one fused kernel (dim <= 128, dim_head <= 32): then
 * wqkv_packed / bqkv must come from packing.pack_qkv_weight() (per-head [q|k|v] tiles, head dim
 * zero-padded to 32), wproj_packed from pack_linear_weight(bn = dim), and no scratch is touched. */
int rvt_attention_is_fused(int dim, int dim_head);
/* N-tiles the MLP weights must be packed with (fc1 [hidden, dim], fc2 [dim, hidden]); returns 1 when
 * rvt_mlp_block runs as one fused kernel (dim <= 128), 0 for the two-GEMM path. */
int rvt_mlp_tiles(int dim, int hidden, int* bn_fc1, int* bn_fc2);
/* N-tile of the downsample conv with cout output channels (cout itself = fused LayerNorm). */
int rvt_conv_tile_n(int cout);
/* K slices the wide-stage (Cout >= 256) downsample conv is split into when given a workspace (1 = none). */
int rvt_conv_split_k(int64_t n_tokens, int cout, int k);
/* Channels per CTA for the Conv-LSTM gate GEMM (tile = [f|i|o|g] x cw columns). */
int rvt_lstm_cw(int dim);
/* Rows one partition group occupies in a 128-row tile (64 or 128; <0 if P > 128). */
int rvt_rows_per_group(int partition_tokens);
/* Rows of the attention scratch matrices for B x H x W tokens and partition (ph, pw). */
int64_t rvt_attention_scratch_rows(int batch, int height, int width, int ph, int pw);

/* ---- a10: StackedHistogram.construct  (data/utils/representations.py:76-121) ----------
 * x, y, pol, t: int64[n] device arrays (t sorted).  counts: u32[2*bins*H*W] scratch that must
 * be zero on entry and is left zero on exit.  out: u8[2*bins*H*W] laid out [2*bins][H][W],
 * channel = pol*bins + t_idx.  err_flag: device int, OR-ed with 1 (time not sorted),
 * 2 (pol outside {0,1}), 4 (coordinate outside the frame); never cleared here. n == 0 -> zeros. */
int rvt_stacked_histogram(const int64_t* x, const int64_t* y, const int64_t* pol, const int64_t* t,
                          int64_t n, int bins, int height, int width, int count_cutoff, int fastmode,
                          uint32_t* counts, uint8_t* out, int* err_flag, void* stream);

/* ---- a3: ConvDownsampling_Cf2Cl.forward  (models/layers/maxvit/maxvit.py:143-178) -------
 * Strided conv (no bias) + LayerNorm(C_out) [+ mask token, maxvit_rnn.py:174-176].
 * in: in_nchw ? [B,Cin,Hin,Win] : [B,Hin,Win,Cin]; in_dtype 0=f32 1=u8 2=f16 (u8 only with in_nchw).  Rows/cols of the virtual input beyond (Hin,Win) read as zero, which folds
 * the harness' zero padding (utils/padding.py:29-44) into the conv.  out: f32 [B,Hout,Wout,Cout].
 * w_packed: rvt_b200.packing.pack_conv_weight().  ln_w/ln_b may be NULL (norm_affine=False).
 * token_mask: u8 [B,Hout,Wout] or NULL; mask_token: f32 [Cout].
 * s2d_scratch (stem fast path, in_nchw only): f16 [B, Hin, Wout, stride*Cin] workspace; when given
 * the input is first re-laid out space-to-depth so every conv tap is a 16-byte vector load, and
 * w_packed must come from packing.pack_stem_weight_s2d().  NULL selects the generic gather path.
 * stem_mode: 0/1 = paths above; 2 = uint8 NCHW 7x7/s4 stem with the input patch staged in shared memory
 * (needs rvt_stem_u8_ok(); w_packed from packing.pack_stem_weight_u8(): K order (kyi, ci, kx8) with kernel rows
 * ky = 0, 4, 1, 5, 2, 6, 3 and kx8 = kx + 1 (kx8 = 0 carries zero weights); no scratch).
 * Channels-last inputs (stages 2-4) with Cout >= 256: `s2d_scratch` is instead an optional split-K workspace, f32
 * [rvt_conv_split_k(n_tokens, Cout, K)][n_tokens, Cout]; NULL = one K loop per CTA. */
int rvt_downsample_cf2cl(const void* in, int in_dtype, int in_nchw, int batch, int cin, int hin, int win,
                         int ksize, int stride, int pad, int hout, int wout, int cout,
                         const void* w_packed, const float* ln_w, const float* ln_b, float eps,
                         const uint8_t* token_mask, const float* mask_token, float* out,
                         void* s2d_scratch, int stem_mode, void* stream);
/* 1 if the uint8 smem-patch stem (stem_mode 2) supports this geometry. */
int rvt_stem_u8_ok(int cin, int ksize, int stride, int pad, int win, int hout, int wout, int cout);

/* ---- a4-a7: attention half of PartitionAttentionCl.forward  (maxvit.py:252-268, 273-354) -
 * x <- x + gamma1 * proj(attn(partition(norm1(x))))   in place, x: f32 [B,H,W,C].
 * grid = 0: window partition, 1: grid partition.  n1_w/n1_b NULL => norm1 = Identity.
 * gamma1 NULL => LayerScale = Identity.  scratch_qkv: f16 [rows, 3C], scratch_o and scratch_xn:
 * f16 [rows, C], rows = rvt_attention_scratch_rows() (scratch_xn is only touched when C >= 256). */
int rvt_partition_attention(float* x, int batch, int height, int width, int dim, int ph, int pw, int grid,
                            int dim_head, const float* n1_w, const float* n1_b, float eps,
                            const void* wqkv_packed, const float* bqkv, const void* wproj_packed,
                            const float* bproj, const float* gamma1, void* scratch_qkv, void* scratch_o,
                            void* scratch_xn, void* stream);

/* ---- a8: MLP half of PartitionAttentionCl.forward  (maxvit.py:85-118, 269) --------------
 * x <- x + gamma2 * fc2(gelu(fc1(norm2(x))))   in place, x: f32 [n_tokens, C].
 * w1_packed / w2_packed: pack_linear_weight() with the N-tiles from rvt_mlp_tiles().
 * scratch_hidden: f16 [round_up(n_tokens,128), hidden]; scratch_xn: f16 [round_up(n_tokens,128), C]
 * (only touched when C >= 256). */
int rvt_mlp_block(float* x, int64_t n_tokens, int dim, int hidden, const float* n2_w, const float* n2_b,
                  float eps, const void* w1_packed, const float* b1, const void* w2_packed, const float* b2,
                  const float* gamma2, void* scratch_hidden, void* scratch_xn, void* stream);

/* ---- a9: DWSConvLSTM2d.forward  (models/layers/rnn.py:36-69) ---------------------------
 * x, h_prev, c_prev, h_out, c_out: f32 [B,H,W,C]; h_prev/c_prev NULL => zero state.
 * dws_mode 0: no depthwise conv; 1: depthwise ks x ks (+bias) on h_prev only; 2: on cat(x,h).
 * dw_w: f32 [ks*ks][D] (tap-major), dw_b: f32 [D], D = C (mode 1) or 2C (mode 2).
 * w_packed / bias_tiled: rvt_b200.packing.pack_lstm_weight() (gate-interleaved tiles).
 * scratch_xh: optional f16 [round_up(B*H*W,128), 2C] workspace (used when dim >= 256 and dws_mode == 0).
 * h_out_f16: optional f16 [B,H,W,C] second copy of h_t (feeds the next stage's rvt_downsample_cf2cl, in_dtype 2). */
int rvt_dws_conv_lstm(const float* x, const float* h_prev, const float* c_prev, int batch, int height,
                      int width, int dim, const void* w_packed, const float* bias_tiled, const float* dw_w,
                      const float* dw_b, int dws_mode, int dws_ks, float* h_out, float* c_out, void* scratch_xh,
                      void* h_out_f16, void* stream);

/* ---- building block exposed for tests: D = A W^T + b, f16 in / f16 out ------------------
 * a: f16 [m, k] row-major (k % 8 == 0), w_packed: pack_linear_weight(W[n,k]), out: f16
 * [round_up(m,128), n]. act: 0 none, 1 exact-erf GELU. */
int rvt_linear_f16(const void* a, int64_t m, int k, int n, const void* w_packed, const float* bias, int act,
                   void* out, void* stream);

/* ======================================================================================
 * Training step (BASELINE configs[2]).  The reference has no backward code of its own: PyTorch
 * autograd differentiates maxvit.py / rnn.py inside modules/detection.py:150-199 (training_step).
 * The entries below are (1) the training-mode forward of each operator, which additionally saves
 * the intermediates its gradient needs, and (2) the building blocks of the analytic backward,
 * composed per operator in rvt_b200/train.py.  Gradient matrices are fp32 accumulators the
 * kernels ADD into (zero them once per backward pass; contributions of all unrolled timesteps
 * accumulate in place).  Gradient signals inside a branch are fp16 (as under the reference's
 * precision-16 AMP), the residual-stream / state gradients are fp32.
 * ====================================================================================== */

/* a3 forward, training: also stores the conv output before LayerNorm (raw_out f32 [B,Hout,Wout,Cout]);
 * s2d_scratch / stem_mode as in rvt_downsample_cf2cl. */
int rvt_downsample_cf2cl_train(const void* in, int in_dtype, int in_nchw, int batch, int cin, int hin, int win,
                               int ksize, int stride, int pad, int hout, int wout, int cout, const void* w_packed,
                               const float* ln_w, const float* ln_b, float eps, const uint8_t* token_mask,
                               const float* mask_token, float* out, float* raw_out, void* s2d_scratch, int stem_mode,
                               void* stream);
/* a4-a7 forward, training: out of place (x_out = x_in + ...), always the row-LN + three-kernel path; weights packed with
 * pack_linear_weight(bn = rvt_tile_n(...)); xn_save f16 [rows,C] (= norm1(x) rows in partition order), qkv_save f16
 * [rows,3C] and o_save f16 [rows,C] are kept for the backward. */
int rvt_partition_attention_train(const float* x_in, float* x_out, int batch, int height, int width, int dim, int ph,
                                  int pw, int grid, int dim_head, const float* n1_w, const float* n1_b, float eps,
                                  const void* wqkv_packed, const float* bqkv, const void* wproj_packed,
                                  const float* bproj, const float* gamma1, void* qkv_save, void* o_save,
                                  void* xn_save, void* stream);
/* a8 forward, training: out of place, row-LN + two-GEMM path; xn_save f16 [round_up(n,128), C] = norm2(x);
 * pre_save / act_save f16 [round_up(n,128), hidden] = fc1 output before / after GELU. */
int rvt_mlp_block_train(const float* x_in, float* x_out, int64_t n_tokens, int dim, int hidden, const float* n2_w,
                        const float* n2_b, float eps, const void* w1_packed, const float* b1, const void* w2_packed,
                        const float* b2, const float* gamma2, void* pre_save, void* act_save, void* xn_save,
                        void* stream);
/* a9 forward, training (dws_mode 0 only): xh_save f16 [round_up(n,128), 2C] = [x | h_prev], gates_save f16 [n, 4C] =
 * activated gates [f|i|o|g]. */
int rvt_dws_conv_lstm_train(const float* x, const float* h_prev, const float* c_prev, int batch, int height, int width,
                            int dim, const void* w_packed, const float* bias_tiled, float* h_out, float* c_out,
                            void* xh_save, void* gates_save, void* stream);

/* D = A W^T (+bias) with A f16 [m,k], W packed by pack_linear_weight(W[n,k], rvt_tile_n(n,k)).
 * out_f32 = 0: f16 out [round_up(m,128), n]; act 0 none, 1 GELU, 2 multiply by gelu'(aux[m,n] f16) (MLP backward).
 * out_f32 = 1: f32 out [m, n] (no bias / act).  Data gradients dX = dY W are this with W := W^T packed. */
int rvt_linear_ex(const void* a, int64_t m, int k, int n, const void* w_packed, const float* bias, int act,
                  const void* aux, void* out, int out_f32, void* stream);
/* Weight gradient: g[i*s_i + j*s_j] += sum_m a1[m,i] * a2[m,j]; a1 f16 [m,n1] (ld1), a2 f16 [m,n2] (ld2).
 * mode 0: operands consumed in place as MN-major tcgen05 tiles; mode 1: transposed copies in scratch_t
 * (f16, rvt_gemm_tn_scratch_elems() elements) and K-major tiles.  colsum1 / colsum2 (optional, f32 [n1] / [n2]):
 * += the column sums of a1 / a2 (bias gradients), accumulated from the operand tiles while they sit in shared memory. */
int rvt_gemm_tn(const void* a1, int ld1, int n1, const void* a2, int ld2, int n2, int64_t m, float* g, int64_t s_i,
                int64_t s_j, int mode, void* scratch_t, float* colsum1, float* colsum2, void* stream);
int64_t rvt_gemm_tn_scratch_elems(int64_t m, int n1, int n2);
/* Row maps: map_mode 0 identity (rows = tokens), 1 window, 2 grid partition order (rvt_attention_scratch_rows rows). */
/* out16[row] = LayerNorm(x[token(row)]) (x itself if !do_ln), f16 [rows, dim]; rows without a token are zero. */
int rvt_ln_rows_f16(const float* x, int map_mode, int batch, int height, int width, int dim, int ph, int pw,
                    const float* ln_w, const float* ln_b, int do_ln, float eps, void* out16, void* stream);
/* LayerNorm backward.  dy: f16 [rows, dim] in map order (dy_is_f16) or f32 [tokens, dim].  dres (f32 tokens) += dx when
 * given; dx16 (f16 [rows, dim]) = dx when given; dw_acc / db_acc (f32 [dim]) += parameter gradients. do_ln = 0: dx = dy. */
int rvt_ln_bwd(const float* x, const void* dy, int dy_is_f16, int map_mode, int batch, int height, int width, int dim,
               int ph, int pw, const float* ln_w, int do_ln, float eps, float* dres, void* dx16, float* dw_acc,
               float* db_acc, void* stream);
/* d0[row] = f16(dres[token(row)]), d1[row] = f16(gamma * dres[token(row)])  (LayerScale backward, maxvit.py:45-53). */
int rvt_gather_cast(const float* dres, int map_mode, int batch, int height, int width, int dim, int ph, int pw,
                    const float* gamma, void* d0, void* d1, void* stream);
/* softmax(QK^T)V backward per (partition group, head) (maxvit.py:349-352): qkv, dqkv f16 [rows,3C]; o (the forward
 * output of the core) and dout f16 [rows,C]. */
int rvt_attn_core_bwd(const void* qkv, const void* o, const void* dout, void* dqkv, int batch, int height, int width,
                      int dim, int ph, int pw, int dim_head, void* stream);
/* Conv-LSTM gates backward (rnn.py:57-67): dpre f16 [n,4C] ([f|i|o|g] = rows of conv1x1.weight), dc_prev f32 [n,C]. */
int rvt_lstm_gates_bwd(const void* gates, const float* c_prev, const float* c_new, const float* dh, const float* dc,
                       int64_t n_tokens, int dim, void* dpre, float* dc_prev, void* stream);
/* Downsample conv operand: col f16 [B*Hout*Wout, round_up(k*k*cin, 8)]; K order (ky, kx, ci) for channels-last inputs,
 * (ci, ky, kx) for NCHW inputs. */
int rvt_im2col(const void* in, int in_dtype, int in_nchw, int batch, int cin, int hin, int win, int ksize, int stride,
               int pad, int hout, int wout, void* col, void* stream);
/* NCHW (in_dtype 0 f32 / 1 u8 / 2 f16) -> channels-last f16 [B,H,W,channels_padded] (zero padded channels). */
int rvt_nchw_to_nhwc_f16(const void* in, int in_dtype, int batch, int channels, int height, int width,
                         int channels_padded, void* out, void* stream);
/* d_in f32 [B,Hin,Win,Cin] = col2im(dcol f16 [B*Hout*Wout, round_up(k*k*cin, 8)]). */
int rvt_col2im(const void* dcol, int batch, int cin, int hin, int win, int ksize, int stride, int pad, int hout,
               int wout, float* d_in, void* stream);
/* acc[n] += sum_m a[m,n]  (a f16, leading dimension ld): bias gradients. */
int rvt_colsum(const void* a, int64_t m, int n, int ld, float* acc, void* stream);

/* ======================================================================================
 * SURVEY.md §8(f) "next" rows, built to the same bar: the callers / data formats either side of the path.
 * ====================================================================================== */

/* ---- f3: harness glue inside an L-step sequence (modules/utils/detection.py) -------------
 * RNNStates.reset -> recursive_reset (:96-113; modules/detection.py:117,217): state[mask] = 0 in place.
 * h, c: f32 [batch, per_sample] (c may be NULL); mask: u8 [batch] (is_first_sample). */
int rvt_state_reset(float* h, float* c, const uint8_t* mask, int batch, int64_t per_sample, void* stream);
/* BackboneFeatureSelector (:24-46): dst[j, :] = src[idx[j], :], j < n_idx; src f32 [n_src_rows, row_elems] holds the L*B
 * feature maps of a sequence, idx = t*B + b of the labelled (step, sample) pairs; idx[j] < 0 => zero row. */
int rvt_gather_rows(const float* src, const int32_t* idx, int n_idx, int64_t n_src_rows, int64_t row_elems, float* dst,
                    void* stream);

/* ---- f4: preprocessing neighbours of the voxelizer ---------------------------------------
 * downsample_ev_repr(x, 0.5) = F.interpolate(mode='nearest-exact') (scripts/genx/preprocess_dataset.py:467-477,525-528):
 * out[c,y,x] = in[c, min(2y+1,H-1), min(2x+1,W-1)], out: [channels, height/2, width/2] bytes (uint8 or int8 alike). */
int rvt_downsample2_nearest(const uint8_t* in, int channels, int height, int width, uint8_t* out, void* stream);
/* H5Reader._correct_time (preprocess_dataset.py:163-172): t[i] = max(floor_value, t[0..i]) in place (floor_value = 0 there).
 * scratch: int64 [rvt_cummax_scratch_elems(n)]. */
int64_t rvt_cummax_scratch_elems(int64_t n);
int rvt_cummax_i64(int64_t* t, int64_t n, int64_t floor_value, int64_t* scratch, void* stream);
/* np.searchsorted(sorted, queries, side) (preprocess_dataset.py:511-516): right = 0 'left', 1 'right'. */
int rvt_searchsorted_i64(const int64_t* sorted, int64_t n, const int64_t* queries, int64_t n_queries, int right, int64_t* out,
                         void* stream);
/* MixedDensityEventStack.construct (data/utils/representations.py:130-218) -> int8 [bins, H, W].  thresholds: f32 [bins-1],
 * t_idx = #{k: t_norm >= thresholds[k]} (found on the host from the reference's own fp32 expression, so binning is bit-exact);
 * t_lo / t_hi: the fp32 clamp bounds of t_norm; count_cutoff < 0 = none; counts: i32 [bins*H*W] scratch, zero at rest;
 * err_flag as rvt_stacked_histogram. */
int rvt_mixed_density_stack(const int64_t* x, const int64_t* y, const int64_t* pol, const int64_t* t, int64_t n, int bins,
                            int height, int width, int count_cutoff, float t_lo, float t_hi, const float* thresholds,
                            int32_t* counts, int8_t* out, int* err_flag, void* stream);

/* ---- f2: YOLOPAFPN + YOLOXHead inference + postprocess (the step right after the backbone) -------------
 * (models/detection/yolox_extension/models/yolo_pafpn.py:109-139, yolox/models/yolo_head.py:165-290, yolox/utils/boxes.py:32-76)
 * BaseConv = Conv2d(bias=False) + BatchNorm2d + SiLU (network_blocks.py:29-51) as ONE implicit-GEMM launch: BatchNorm folded
 * into w_packed (packing.pack_conv_weight of the scaled kernel, K order (ky, kx, ci)) + bias.  Tensors are channels-last
 * f16 with a pixel pitch, so th.cat along channels is writing channel slices of one buffer.  in: [B,Hin,Win] pixels of `cin`
 * channels, pitch in_pitch; out: [round_up(B*Hout*Wout,128)] pixels of cout_padded (multiple of 16) channels, pitch out_pitch.
 * act: 0 none, 4 SiLU. */
int rvt_conv2d_nhwc_f16(const void* in, int in_pitch, int batch, int cin, int hin, int win, int ksize, int stride, int pad, int hout,
                        int wout, int cout_padded, const void* w_packed, const float* bias, int act, void* out, int out_pitch,
                        void* stream);
/* f32 [B,H,W,C] given by element strides (any layout) -> f16 channel slice dst (pixel pitch dst_pitch). */
int rvt_cast_slice_f16(const float* src, int64_t sb, int64_t sy, int64_t sx, int64_t sc, int batch, int height, int width, int channels,
                       void* dst, int dst_pitch, void* stream);
/* F.interpolate(scale_factor=2, mode='nearest-exact') (yolo_pafpn.py:47) between f16 channel slices; src is [B,height,width]. */
int rvt_upsample2_slice_f16(const void* src, int src_pitch, int batch, int height, int width, int channels, void* dst, int dst_pitch,
                            void* stream);
/* One head level: [reg(4)|obj(1)] and [cls(nc)] f16 rows -> decoded f32 out[b, anchor_offset + y*W + x, 5+nc] (yolo_head.py:228-232,271-290). */
int rvt_yolox_decode(const void* regobj, int regobj_pitch, const void* cls, int cls_pitch, int batch, int height, int width,
                     int num_classes, float stride_px, int anchor_offset, int anchors_total, float* out, void* stream);
/* postprocess (boxes.py:32-76): prediction f32 [B, anchors, 5+nc] (cx, cy, w, h, obj, cls..) -> detections f32 [B, anchors, 7]
 * (x1, y1, x2, y2, obj_conf, class_conf, class_pred), the first counts[b] rows of image b valid, in score order. anchors <= 8192. */
int rvt_yolox_postprocess(const float* prediction, int batch, int anchors, int num_classes, float conf_thre, float nms_thre,
                          float* detections, int* counts, void* stream);

/* Profiling aid, not part of the reference boundary: device buffer int64 [grid][8][12] the persistent kernels fill with
 * %globaltimer stamps of their phase boundaries (profiles/trace_v2.py); NULL disables. */
int rvt_debug_set_trace(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* RVT_B200_H_ */
