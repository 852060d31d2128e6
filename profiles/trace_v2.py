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
    ap.add_argument('which', choices=['attn', 'mlp', 'stem'])
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
    if args.which == 'stem':
        return trace_stem(m, pk, dev)
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


def trace_stem(m, pk, dev):
    """stem_v2 (RVT_STEM_V2=1): stamps of the role leaders.  builders 0 tile start, 1 patch ready, 2 K loop done; epilogue 3 waiting,
    4 accumulator ready, 5 statistics done, 6 tile stored; MMA thread 7 waiting for a free accumulator, 8 got it, 9 last commit issued."""
    from rvt_b200 import _lib, ops
    from oracle import backbone_oracle as bo
    st = m.stages[0]
    d = st.downsample_cf2cl
    x8 = bo.synth_events_tensor(0, 8, 20, 384, 640).to(dev)
    fn = lambda: ops.downsample_cf2cl(x8, True, pk['conv_w_u8'], st.dim, d.kernel_size, d.factor, d.padding, pk['ds_ln_w'], pk['ds_ln_b'],
                                      stem_mode=2)
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
    tr = trace.view(grid, T, P).cpu().double()
    used = tr[:, 0, 0] > 0
    tr = tr[used]
    t0 = tr[:, 0, 0].min()
    print(f'stem_v2: {int(used.sum())} CTAs traced; span {(tr.max() - t0) / 1e3:.1f} us')
    segs = [('builders: wait for the patch', 0, 1), ('builders: K loop (18 chunks)', 1, 2), ('epilogue: wait for the accumulator', 3, 4),
            ('epilogue: LayerNorm statistics', 4, 5), ('epilogue: normalise + store', 5, 6), ('MMA: wait for a free accumulator', 7, 8),
            ('MMA: K loop issue', 8, 9)]
    for it in range(T):
        ok = tr[:, it, 0] > 0
        if not bool(ok.any()):
            break
        print(f' tile slot {it}: {int(ok.sum())} CTAs, builders start at {((tr[ok, it, 0] - t0).mean()) / 1e3:7.2f} us')
        for name, i0, i1 in segs:
            dd = (tr[ok, it, i1] - tr[ok, it, i0]) / 1e3
            print(f'    {name:<40s} {dd.mean():6.2f} us  (max {dd.max():6.2f})')
        if os.environ.get('RVT_STEM_V2', '2') == '2':
            print(f'    SM cycles per tile: epilogue waiting for the staging buffer (previous TMA stores) {tr[ok, it, 10].mean():8.0f}   TMEM load + normalise + staging + barrier {tr[ok, it, 11].mean():8.0f}')
        else:
            print(f'    builder leader, SM cycles per tile: waiting for a free A slot {tr[ok, it, 10].mean():8.0f}   fence + arrive {tr[ok, it, 11].mean():8.0f}')


if __name__ == '__main__':
    main()
