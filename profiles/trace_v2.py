"""Phase timeline of the persistent v2 kernels from in-kernel %globaltimer stamps (rvt_debug_set_trace):
   python profiles/trace_v2.py attn|mlp [--stage 0]
Prints, per tile slot of a CTA, the mean duration (us) between consecutive stamps over all CTAs."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ATTN_PTS = ['loop top', 'x_full ok', 'LN done (a_full arrive)', 'qkv_full ok', 'QKV epi done (qk_ready)', 's_full ok',
            'softmax done (p_full)', 'o_full ok', 'O epi done (so_full)', 'out_full ok', 'tile end']
MLP_PTS = ['loop top', 'hid_full ok', 'GELU done', 'LN(next) done', 'out_full ok', 'epilogue done']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('which', choices=['attn', 'mlp'])
    ap.add_argument('--stage', type=int, default=0)
    args = ap.parse_args()
    import rvt_b200
    from rvt_b200 import _lib, ops
    from oracle import backbone_oracle as bo
    from tests.test_host_cpu import make_cfg
    dev = torch.device('cuda:0')
    spec = bo.BackboneSpec(embed_dim=64, dim_head=32, partition_size=(6, 10))
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(bo.synth_params(spec, 0), strict=True)
    m = m.to(dev).eval()
    pk = m._ensure_packed(dev)[args.stage]
    blk = pk['blocks'][0]
    s = args.stage
    c = 64 << s
    xs = torch.randn(8, 96 >> s, 160 >> s, c, device=dev)
    dummy = torch.empty(1, dtype=torch.float16, device=dev)
    sh = torch.empty(xs.numel() * 4, dtype=torch.float16, device=dev)
    fn = (lambda: ops.partition_attention_(xs, blk, dummy, dummy, dummy)) if args.which == 'attn' else (lambda: ops.mlp_block_(xs, blk, sh, dummy))
    pts = ATTN_PTS if args.which == 'attn' else MLP_PTS
    grid, T, P = 4096, 8, 12
    trace = torch.zeros(grid * T * P, dtype=torch.int64, device=dev)
    with torch.inference_mode():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        _lib.lib().rvt_debug_set_trace(trace.data_ptr())
        fn()
        torch.cuda.synchronize()
        _lib.lib().rvt_debug_set_trace(None)
    tr = trace.view(grid, T, P).cpu()
    used = tr[:, 0, 0] > 0
    tr = tr[used]
    t0 = tr[:, 0, 0].min()
    print(f'{args.which} stage {s + 1}: {int(used.sum())} CTAs traced; kernel span {(tr.max() - t0) / 1e3:.1f} us')
    n = len(pts)
    for it in range(T):
        ok = tr[:, it, 0] > 0
        if not bool(ok.any()):
            break
        seg = tr[ok, it, :n].double()
        d = (seg[:, 1:] - seg[:, :-1]) / 1e3
        start = (seg[:, 0] - float(t0)).mean() / 1e3
        tot = (seg[:, n - 1] - seg[:, 0]).mean() / 1e3
        print(f' tile slot {it}: {int(ok.sum())} CTAs, starts at {start:7.2f} us, lasts {tot:6.2f} us')
        for k in range(n - 1):
            print(f'    {pts[k]:>28s} -> {pts[k + 1]:<28s} {d[:, k].mean():6.2f} us  (max {d[:, k].max():6.2f})')


if __name__ == '__main__':
    main()
