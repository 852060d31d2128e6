"""Per-operator timing at the RVT-B 1Mpx bs=8 shapes (CUDA events, L2 flushed between timed
launches), with algorithmic FLOPs / bytes per launch -> achieved TF/s and GB/s.
Also the driver for `ncu` captures:  python profiles/op_bench.py --only mlp --stage 0 --iters 1
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

B, H0, W0 = 8, 384, 640


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None, help='conv|attn|mlp|lstm')
    ap.add_argument('--stage', type=int, default=None)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    import rvt_b200
    from rvt_b200 import ops
    from oracle import backbone_oracle as bo
    from tests.test_host_cpu import make_cfg
    dev = torch.device('cuda:0')
    spec = bo.BackboneSpec(embed_dim=64, dim_head=32, partition_size=(6, 10))
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(bo.synth_params(spec, 0), strict=True)
    m = m.to(dev).eval()
    packed = m._ensure_packed(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    x8 = bo.synth_events_tensor(0, B, 20, H0, W0).to(dev)
    rows = []

    def timeit(name, s, fn, flops, bytes_):
        if args.only and args.only != name:
            return
        if args.stage is not None and args.stage != s:
            return
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        rows.append(dict(op=name, stage=s + 1, us=us, gflop=flops / 1e9, mbytes=bytes_ / 1e6,
                         tflops=flops / us / 1e6, gbs=bytes_ / us / 1e3))
        print(f'{name:5s} S{s + 1}: {us:8.1f} us  {flops / 1e9:7.2f} GFLOP {flops / us / 1e6:7.1f} TF/s   '
              f'{bytes_ / 1e6:7.1f} MB {bytes_ / us / 1e3:7.0f} GB/s')

    def split_ws(s, n, c, cin, d):     # the K-split workspace RNNDetector hands to the wide-stage convs
        if s == 0 or c < 256:
            return None
        from rvt_b200 import _lib
        k = _lib.lib().rvt_conv_split_k(n, c, cin * d.kernel_size ** 2)
        return torch.empty(k * n * c, dtype=torch.float32, device=dev) if k > 1 else None

    with torch.inference_mode():
        h, w = H0, W0
        cin = 20
        for s, (st, pk) in enumerate(zip(m.stages, packed)):
            d = st.downsample_cf2cl
            c = st.dim
            f = d.factor
            ho, wo = h // f, w // f
            n = B * ho * wo
            if s == 0:
                # the product path for uint8 events: smem-staged patch loader (stem_mode 2), no scratch tensor
                src, nchw, cw, s2d = x8, True, pk['conv_w_u8'], None
                in_bytes = x8.numel()
            else:
                # stages 2-4 read the fp16 copy of the previous stage's h_t (RNNDetector._stage_step)
                src, nchw, cw, s2d = torch.randn(B, h, w, cin, device=dev).half(), False, pk['conv_w'], None
                in_bytes = src.numel() * 2
            timeit('conv', s, lambda: ops.downsample_cf2cl(src, nchw, cw, c, d.kernel_size, f, d.padding, pk['ds_ln_w'],
                                                           pk['ds_ln_b'], s2d_scratch=s2d, stem_mode=2 if s == 0 else 0, split_ws=split_ws(s, n, c, cin, d)),
                   2 * n * c * cin * d.kernel_size ** 2, in_bytes + n * c * 4)
            xs = torch.randn(B, ho, wo, c, device=dev)
            blk = pk['blocks'][1]
            rows_s = ops.attention_scratch_rows(B, ho, wo, blk['part'])
            sq = torch.empty(rows_s * 3 * c, dtype=torch.float16, device=dev)
            so = torch.empty(rows_s * c, dtype=torch.float16, device=dev)
            sxn = torch.empty(max(rows_s, ((n + 127) // 128) * 128) * c, dtype=torch.float16, device=dev)
            sh = torch.empty(((n + 127) // 128) * 128 * 4 * c, dtype=torch.float16, device=dev)
            P = blk['part'][0] * blk['part'][1]
            timeit('attn', s, lambda: ops.partition_attention_(xs, blk, sq, so, sxn), 8 * n * c * c + 4 * n * P * c, 2 * n * c * 4)
            timeit('mlp', s, lambda: ops.mlp_block_(xs, blk, sh, sxn), 16 * n * c * c, 2 * n * c * 4)
            hp, cp = torch.randn_like(xs), torch.randn_like(xs)
            sxh = torch.empty(((n + 127) // 128) * 128 * 2 * c, dtype=torch.float16, device=dev)
            timeit('lstm', s, lambda: ops.dws_conv_lstm(xs, hp, cp, pk, 3, sxh), 16 * n * c * c, 5 * n * c * 4)
            h, w, cin = ho, wo, c
    if rows:
        tot = sum(r['us'] for r in rows)
        print(f'sum {tot:.1f} us (a timestep runs attn and mlp twice per stage)')
    if args.json:
        json.dump(rows, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
