"""Profiling driver: one small training step (RVT-Base 1Mpx, bs 3, L timesteps) — warm-up iteration, then one
iteration between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0')
m = bench.build_model(0).to(dev).train()
m.pad_to_hw = (bench.PAD_H, bench.PAD_W)
seq = bench.make_uint8_sequence(1, L, 3).to(dev)


def step():
    st = None
    for t in range(L):
        out, st = m(seq[t], st)
    loss = sum((out[s].float() ** 2).mean() for s in (1, 2, 3, 4)) * 65536.0
    m.zero_grad(set_to_none=True)
    loss.backward()


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
