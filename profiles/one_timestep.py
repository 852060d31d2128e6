"""Driver for ncu: RVT-B 1Mpx bs 8, `--steps` eager timesteps through RNNDetector.forward (one stream, no graph) so that
`ncu -s <launches of the warm-up steps> -c <launches per step>` captures exactly one timestep's kernels.
  python profiles/one_timestep.py --steps 3 [--count]      (--count prints the number of kernel launches per timestep)"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--count', action='store_true')
    ap.add_argument('--voxel', action='store_true', help='also run one 50 M-event StackedHistogram.construct at the end')
    args = ap.parse_args()
    import bench
    dev = torch.device('cuda:0')
    model = bench.build_model(0).to(dev).eval()
    model.pad_to_hw = (bench.PAD_H, bench.PAD_W)
    seq = bench.make_uint8_sequence(1, args.steps, bench.B_PER_GPU).to(dev)
    st = None
    with torch.inference_mode():
        if args.count:
            _, st = model(seq[0], st)
            n, names = bench.count_kernel_launches(lambda: model(seq[1], st), dev)
            print(n)
            for k, v in names.items():
                print(f'  {v:3d} {k[:100]}')
            return
        for t in range(args.steps):
            _, st = model(seq[t], st)
        torch.cuda.synchronize()
        if args.voxel:
            import rvt_b200
            n = 50_000_000
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randint(0, 1280, (n,), generator=g, device=dev, dtype=torch.int64)
            y = torch.randint(0, 720, (n,), generator=g, device=dev, dtype=torch.int64)
            p = torch.randint(0, 2, (n,), generator=g, device=dev, dtype=torch.int64)
            t = torch.sort(torch.randint(0, 50000, (n,), generator=g, device=dev, dtype=torch.int64)).values
            sh = rvt_b200.StackedHistogram(10, 720, 1280, 10, validate=False)
            sh.construct(x, y, p, t)
            torch.cuda.synchronize()


if __name__ == '__main__':
    main()
