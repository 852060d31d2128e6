"""Secondary BASELINE configs: (5) raw-event voxelization 50M events -> 2x10x720x1280, (4) RVT-S Gen1
bs=64 eval.  CUDA-event timings; prints JSON.  Used by bench.py's `aux` block and for ncu captures."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bench_voxel(n_events=50_000_000, iters=5, hot=0.0, cpu_sample=5_000_000):
    import rvt_b200
    from oracle import voxel_oracle as vo
    dev = torch.device('cuda:0')
    h, w, bins = 720, 1280, 10
    x, y, p, t = vo.synth_events(0, n_events, h, w, hot_fraction=hot)
    tx, ty, tp, tt = (torch.from_numpy(a).to(dev) for a in (x, y, p, t))
    sh = rvt_b200.StackedHistogram(bins, h, w, 10, True, validate=False)
    out = torch.empty(sh.get_shape(), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sh.construct(tx, ty, tp, tt, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sh.construct(tx, ty, tp, tt, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    sec = ts[len(ts) // 2]
    algo_bytes = 32 * n_events + out.numel()
    # CPU baseline (single-thread C port of the reference algorithm) on a bounded sample
    import ctypes
    import subprocess
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libvoxel_oracle.so'))
    m = min(cpu_sample, n_events)
    ref = np.zeros(2 * bins * h * w, np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    t0 = time.perf_counter()
    lib.rvt_oracle_stacked_histogram(P(x[:m]), P(y[:m]), P(p[:m]), P(t[:m]), ctypes.c_int64(m), bins, h, w, 10, 1, P(ref))
    cpu_s = time.perf_counter() - t0
    return {'events': n_events, 'hot_fraction': hot, 'ms': sec * 1e3, 'events_per_s': n_events / sec,
            'algorithmic_bytes': algo_bytes, 'achieved_gbs': algo_bytes / sec / 1e9,
            'cpu_port_events_per_s': m / cpu_s, 'cpu_sample_events': m}


def bench_rvt_s_gen1(batch=64, L=5, iters=3):
    import rvt_b200
    from oracle import backbone_oracle as bo
    from tests.test_host_cpu import make_cfg
    dev = torch.device('cuda:0')
    spec = bo.BackboneSpec(embed_dim=48, dim_head=24, partition_size=(8, 10))
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(bo.synth_params(spec, 0), strict=True)
    m = m.to(dev).eval()
    m.pad_to_hw = (256, 320)
    xs = torch.stack([bo.synth_events_tensor(i, batch, 20, 240, 304) for i in range(L)]).to(dev)
    g = rvt_b200.capture_sequence(m, xs)
    g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / iters
    return {'config': 'RVT-S Gen1 240x304 (model 256x320, P=80, dim_head 24) bs=64', 'timesteps': L,
            'ms_per_timestep': sec / L * 1e3, 'frames_per_s': batch * L / sec, 'gflop_per_frame': 4.12,
            'tflops': batch * L / sec * 4.12e-3}


if __name__ == '__main__':
    res = {}
    if 'voxel' in sys.argv or len(sys.argv) == 1:
        res['voxel_uniform'] = bench_voxel()
        res['voxel_hot1pct'] = bench_voxel(hot=0.01, iters=3)
    if 'gen1' in sys.argv or len(sys.argv) == 1:
        res['rvt_s_gen1_bs64'] = bench_rvt_s_gen1()
    print(json.dumps(res, indent=1))
