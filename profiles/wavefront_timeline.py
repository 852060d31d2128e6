"""Where the wavefront's time goes: CUDA-event start / end of every (stage, timestep) cell of one eager multi-stream
RNNDetector.forward_sequence on the bench workload (RVT-B 1Mpx bs 8, L = 21).  Prints, per stage, the mean busy time of a
cell, the mean gap between consecutive cells of the stage (idle = waiting for its input or for SMs) and the steady-state period.
   python profiles/wavefront_timeline.py [--no-wavefront]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--no-wavefront', action='store_true')
    args = ap.parse_args()
    import bench
    dev = torch.device('cuda:0')
    model = bench.build_model(0).to(dev).eval()
    model.pad_to_hw = (bench.PAD_H, bench.PAD_W)
    L = bench.SEQ_LEN
    seq = bench.make_uint8_sequence(1, L, bench.B_PER_GPU).to(dev)
    xs = [seq[t] for t in range(L)]
    with torch.inference_mode():
        st = None
        for _ in range(2):
            _, st = model.forward_sequence(xs, st, wavefront=not args.no_wavefront)
        torch.cuda.synchronize()
        model.debug_timeline = []
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
        _, st = model.forward_sequence(xs, st, wavefront=not args.no_wavefront)
        torch.cuda.synchronize()
        tl = model.debug_timeline
        model.debug_timeline = None
    cells = {}
    for s, t, e0, e1 in tl:
        cells[(s, t)] = (t0.elapsed_time(e0) * 1e3, t0.elapsed_time(e1) * 1e3)
    n = 4
    total = max(v[1] for v in cells.values())
    print(f'sequence of {L} steps: {total:.0f} us  ({total / L:.0f} us / step)')
    for s in range(n):
        busy = [cells[(s, t)][1] - cells[(s, t)][0] for t in range(L)]
        gaps = [cells[(s, t)][0] - cells[(s, t - 1)][1] for t in range(1, L)]
        per = [cells[(s, t)][1] - cells[(s, t - 1)][1] for t in range(4, L - 3)]
        lag = [cells[(s, t)][0] - cells[(s - 1, t)][1] for t in range(L)] if s else [0.0]
        print(f'stage {s + 1}: cell {sum(busy) / len(busy):7.1f} us (min {min(busy):6.1f} max {max(busy):6.1f})   '
              f'gap to previous cell {sum(gaps) / len(gaps):7.1f} us   wait after stage {s} output {sum(lag) / len(lag):7.1f} us   '
              f'period {sum(per) / len(per):7.1f} us')
    print('cells (start, end) in us:')
    for t in range(L):
        print(f' t={t:2d} ' + '  '.join(f'S{s + 1} {cells[(s, t)][0]:7.0f}-{cells[(s, t)][1]:7.0f}' for s in range(n)))


if __name__ == '__main__':
    main()
