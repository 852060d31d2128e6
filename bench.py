#!/usr/bin/env python
"""bench.py — RVT hot-path throughput on B200 (contract: task statement §④ + base contract).

Workload (BASELINE.json configs[1]): RVT-Base 1Mpx, event tensor 8 x 20 x 360 x 640 (uint8, zero
padding to the model's 384x640 folded into the stem), seq_len 21, states carried, inference.
One *step* = one 21-timestep sequence for the local batch (8 samples/GPU) = 168 frames/GPU.
Batch-sharded over N GPUs with no data-path collective (weak scaling).

  value : frames/s, whole job, inputs resident in HBM (774 MB of uint8 sequences > L2)
  e2e   : frames/s through rvt_b200.RNNDetector.forward with HOST (pinned) uint8 inputs copied
          H2D every timestep and the stage-4 feature map read back D2H every timestep
  --impl reference : the reference's CPU path (oracle port, torch fp32, all host threads) on a
          bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'RVT-B 1Mpx seq_len=21 backbone frames/sec'
B_PER_GPU, SEQ_LEN, IN_C, IN_H, IN_W, PAD_H, PAD_W = 8, 21, 20, 360, 640, 384, 640
# per timestep: S1 s2d+conv, 4 fused attn/mlp, lstm; S2 conv, 4 fused, lstm; S3/S4 conv+ln, 2 x (ln,qkv,core,proj,ln,fc1,fc2), lstm
LAUNCHES_PER_TIMESTEP = (2 + 4 + 1) + (1 + 4 + 1) + 2 * (2 + 2 * 7 + 1)
TRAIN_METRIC = 'RVT-B 1Mpx seq_len=21 backbone training frames/sec'
LOSS_SCALE = 65536.0      # static stand-in for the harness' GradScaler (precision 16, config/general.yaml:6)
GFLOP_PER_FRAME = 20.62                                 # BASELINE.md §3 (algorithmic, MAC = 2 FLOP)


_REAL_STDOUT = None


def protect_stdout():
    """stdout must carry exactly ONE JSON line, but NCCL prints its version banner to fd 1 (at NCCL_DEBUG=VERSION, the
    image default, and at WARN too) and other libraries may chat as well: keep a private copy of the real stdout for the
    result line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')


def emit(obj):
    data = (json.dumps(obj) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def rvt_b_spec():
    """oracle-side description of RVT-Base 1Mpx (reference arms only)."""
    from oracle import backbone_oracle as bo
    return bo.BackboneSpec(embed_dim=64, dim_head=32, partition_size=(6, 10))


def rvt_b_cfg():
    """RVT-Base 1Mpx `mdl_config` (config/model/maxvit_yolox/default.yaml + base.yaml, partition size as
    config/modifier.py:36-41 derives it for 384x640) as a plain dict — what the reference hands to the backbone."""
    return dict(
        name='MaxViTRNN', compile=dict(enable=False, args=dict(mode='reduce-overhead')),
        input_channels=IN_C, enable_masking=False, partition_split_32=2, embed_dim=64, dim_multiplier=[1, 2, 4, 8],
        num_blocks=[1, 1, 1, 1], T_max_chrono_init=[4, 8, 16, 32], stem=dict(patch_size=4),
        stage=dict(downsample=dict(type='patch', overlap=True, norm_affine=True),
                   attention=dict(use_torch_mha=False, partition_size=(6, 10), dim_head=32, attention_bias=True,
                                  mlp_activation='gelu', mlp_gated=False, mlp_bias=True, mlp_ratio=4, drop_mlp=0, drop_path=0,
                                  ls_init_value=1e-5, norm_eps=1e-5),
                   lstm=dict(dws_conv=False, dws_conv_only_hidden=True, dws_conv_kernel_size=3, drop_cell_update=0)))


def build_model(seed=0):
    """rvt_b200.RNNDetector with random-init weights of the RVT-Base architecture: N(0, 1/fan_in) weights, N(0, 0.1) biases,
    LayerNorm weights and LayerScale gammas ~ U(0.5, 1.5) (so no branch is numerically dead).  Nothing of
    oracle/ is touched on the product arm."""
    import rvt_b200
    model = rvt_b200.RNNDetector(rvt_b_cfg())
    g = torch.Generator(device='cpu').manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('gamma') or ('norm' in name and name.endswith('weight')):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif name.endswith('bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith('mask_token'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / fan_in ** 0.5)
    return model


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [v.strip() for v in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_reference_fps(n_timesteps, batch, threads):
    """The reference's CPU implementation of the path (oracle port: torch fp32 eager ops, the same
    ATen kernels the reference dispatches to), states carried; returns frames/s."""
    from oracle import backbone_oracle as bo
    spec = rvt_b_spec()
    torch.set_num_threads(threads)
    params = bo.synth_params(spec, 0)
    xs = [torch.nn.functional.pad(bo.synth_events_tensor(i, batch, IN_C, IN_H, IN_W).float(),
                                  (0, PAD_W - IN_W, 0, PAD_H - IN_H)) for i in range(n_timesteps + 1)]
    st = None
    with torch.inference_mode():
        _, st = bo.backbone_forward(xs[0], st, params, spec)      # warm-up timestep
        t0 = time.perf_counter()
        for i in range(n_timesteps):
            _, st = bo.backbone_forward(xs[1 + i], st, params, spec)
        dt = time.perf_counter() - t0
    return batch * n_timesteps / dt


def best_cpu_threads():
    """torch's intra-op pool at os.cpu_count() threads is far slower than a moderate pool on big hosts
    (measured 0.56 frames/s at 128 threads vs ~10 at 8-32): calibrate on one bs=1 timestep."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    best, best_fps = cands[0], 0.0
    for c in cands:
        fps = cpu_reference_fps(1, 1, c)
        if fps > best_fps:
            best, best_fps = c, fps
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = best_cpu_threads()
    n_ts = 2                     # bounded sample: 2 timesteps x 8 samples per "step"
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_reference_fps(n_ts, B_PER_GPU, cores))
    v = sum(vals) / len(vals)
    sample = (f'{n_ts} timesteps x batch {B_PER_GPU} of the 21-timestep sequence per step (states carried), fp32; '
              f'{cores} torch threads (best of a sweep up to os.cpu_count()={os.cpu_count()})')
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * B_PER_GPU * SEQ_LEN / v,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'RVT-Base 1Mpx 360x640 (padded 384x640) T=10 seq_len=21 bs=8 inference, CPU'},
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


def make_uint8_sequence(seed, length, batch):
    g = torch.Generator(device='cpu').manual_seed(seed)
    shape = (length, batch, IN_C, IN_H, IN_W)
    seq = torch.randint(1, 11, shape, generator=g, dtype=torch.uint8)            # counts 1..10 ...
    seq.mul_(torch.randint(0, 10, shape, generator=g, dtype=torch.uint8) == 0)   # ... on ~10 % of the bins
    return seq


def run_reference_gpu(args, rank, world, local_rank):
    """Informational arm (the denominator of north_star's '>= 10x PyTorch-eager on one B200'): the reference's
    op-by-op PyTorch path (oracle port = the same ATen/cuDNN/cuBLAS calls) on the GPU under fp16 autocast and
    inference_mode, same workload, inputs resident.  Not the driver's reference arm (that is --impl reference, CPU)."""
    if rank != 0:
        return
    from oracle import backbone_oracle as bo
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    spec = rvt_b_spec()
    params = {k: v.to(dev) for k, v in bo.synth_params(spec, 0).items()}
    if args.mode == 'train':
        # the same training step as run_train (fwd 21 timesteps + loss + bwd + unscale + fused Adam), PyTorch eager + autograd
        B = args.train_batch
        plist = [v.requires_grad_(True) for v in params.values()]
        opt = torch.optim.Adam(plist, lr=2e-4, fused=True)
        seq = make_uint8_sequence(4321, SEQ_LEN, B).to(dev)

        def tstep():
            opt.zero_grad(set_to_none=True)
            st, out = None, None
            with torch.autocast('cuda', dtype=torch.float16):
                for t in range(SEQ_LEN):
                    x = torch.nn.functional.pad(seq[t].float(), (0, PAD_W - IN_W, 0, PAD_H - IN_H))
                    out, st = bo.backbone_forward(x, st, params, spec)
            loss = sum((out[s].float() ** 2).mean() for s in (1, 2, 3, 4)) * LOSS_SCALE
            loss.backward()
            torch._foreach_mul_([p.grad for p in plist], 1.0 / LOSS_SCALE)
            opt.step()

        for _ in range(max(args.warmup, 3)):
            tstep()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tstep()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        emit({
            'impl': 'reference-gpu', 'metric': TRAIN_METRIC, 'value': B * SEQ_LEN * args.steps / (ms * 1e-3), 'unit': 'frames/s',
            'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'dtype': 'f16 autocast', 'data': 'synthetic',
            'config': {'workload': f'RVT-Base 1Mpx bs={B} TBPTT seq_len=21 training step, PyTorch eager + autograd on the GPU '
                                   '(oracle port of the reference op sequence), fp16 autocast'}})
        return
    seq = make_uint8_sequence(1234, SEQ_LEN, B_PER_GPU).to(dev)

    def step():
        st = None
        for t in range(SEQ_LEN):
            x = torch.nn.functional.pad(seq[t].float(), (0, PAD_W - IN_W, 0, PAD_H - IN_H))   # modules/detection.py:133-134
            _, st = bo.backbone_forward(x, st, params, spec)
        return st

    with torch.inference_mode(), torch.autocast('cuda', dtype=torch.float16):
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    v = B_PER_GPU * SEQ_LEN * args.steps / (ms * 1e-3)
    emit({
        'impl': 'reference-gpu', 'metric': METRIC, 'value': v, 'unit': 'frames/s', 'n_gpus': 1, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f16 autocast', 'data': 'synthetic',
        'config': {'workload': 'RVT-Base 1Mpx 360x640 (padded 384x640) seq_len=21 bs=8 inference, PyTorch eager on the GPU '
                               '(oracle port of the reference op sequence), fp16 autocast, inputs resident'}})




def run_train(args, rank, world, local_rank):
    """BASELINE configs[2]: RVT-Base 1Mpx training step, TBPTT over seq_len 21, 3 samples per GPU, batch-sharded,
    ONE NCCL all-reduce over the flat gradient buffer, fused Adam on the backbone parameters.  A step = forward of
    21 timesteps (states carried) + synthetic loss on the last step's four feature maps + backward through all 21
    timesteps + gradient all-reduce + unscale + optimizer step."""
    import torch.distributed as dist
    import rvt_b200  # noqa: F401
    from rvt_b200 import sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    B = args.train_batch
    model = build_model(0).to(dev).train()
    model.pad_to_hw = (PAD_H, PAD_W)
    model.train_wavefront = args.train_wavefront      # stage-per-stream schedule (measured slower in training: off by default)
    lo, hi = sharding.batch_slice(B * world, rank, world)
    n_seq = 4                                                        # rotate sequences: 4 x 290 MB of uint8 inputs > L2
    seqs = [make_uint8_sequence(4321 + lo * 10 + i, SEQ_LEN, B).to(dev) for i in range(n_seq)]
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=2e-4, fused=True)
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for k in ('fwd', 'bwd', 'ar', 'opt')}
    acc_ms = {k: 0.0 for k in ev}
    counter = [0]

    cpu_ms = {'fwd': 0.0, 'bwd': 0.0}
    graphed = {}

    def fwd_bwd(seq, record=False):
        if record:
            ev['fwd'][0].record()
        t0 = time.perf_counter()
        st, out = None, None
        for t in range(SEQ_LEN):
            out, st = model(seq[t], st)
        loss = sum((out[s].float() ** 2).mean() for s in (1, 2, 3, 4)) * LOSS_SCALE
        t1 = time.perf_counter()
        if record:
            ev['fwd'][1].record(); ev['bwd'][0].record()
        loss.backward()
        t2 = time.perf_counter()
        if record:
            ev['bwd'][1].record()
            cpu_ms['fwd'] += (t1 - t0) * 500; cpu_ms['bwd'] += (t2 - t1) * 500    # mean of the 2 recorded steps, ms
        return loss

    def capture():
        # ONE CUDA graph of forward (21 timesteps) + loss + backward (rvt_b200.graph.capture_training_step): the weight
        # re-packing, every kernel and the gradient hand-over are replayed without Python; the all-reduce and the optimizer
        # stay outside.  p.grad become static tensors that each replay overwrites.
        from rvt_b200.graph import capture_training_step
        static_seq = torch.empty_like(seqs[0])
        graphed['seq'] = static_seq
        graphed['run'] = capture_training_step(model, lambda: fwd_bwd(static_seq), params)

    def step(record=False):
        seq = seqs[counter[0] % n_seq]
        counter[0] += 1
        if graphed:
            graphed['seq'].copy_(seq, non_blocking=True)
            if record:
                ev['fwd'][0].record(); ev['fwd'][1].record(); ev['bwd'][0].record()
            loss = graphed['run']()
            if record:
                ev['bwd'][1].record()
        else:
            opt.zero_grad(set_to_none=True)
            loss = fwd_bwd(seq, record)
        if record:
            ev['ar'][0].record()
        n_coll = sharding.allreduce_gradients(params)
        if record:
            ev['ar'][1].record(); ev['opt'][0].record()
        torch._foreach_mul_([p.grad for p in params if p.grad is not None], 1.0 / LOSS_SCALE)
        opt.step()
        if record:
            ev['opt'][1].record()
        return loss, n_coll

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if not args.train_eager:
        capture()            # (no eager step before this: AccumulateGrad nodes must first be created on the capture stream)
    for _ in range(max(args.warmup, 3)):
        loss, n_coll = step()
    assert torch.isfinite(loss.detach()).item(), 'non-finite training loss'
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, n_coll = step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):                                               # phase breakdown (separate, event-instrumented steps)
        step(record=True)
        torch.cuda.synchronize(dev)
        for k in ev:
            acc_ms[k] += ev[k][0].elapsed_time(ev[k][1]) / 2
    grad_ok = all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in params)
    frames = B * SEQ_LEN * args.steps * world
    if rank == 0:
        n_par = sum(p.numel() for p in params)
        emit({
            'metric': TRAIN_METRIC, 'value': frames / (ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': f'RVT-Base 1Mpx {B}x20x360x640 uint8 per GPU (model res 384x640) TBPTT seq_len=21 training step: '
                                   'fwd + bwd through 21 timesteps + grad all-reduce + fused Adam (backbone only)',
                       'global_batch': B * world, 'frames_per_step': B * SEQ_LEN,
                       'l2_policy': f'{n_seq} rotating input sequences ({n_seq * seqs[0].numel() >> 20} MB) > L2',
                       'parallelism': f'batch-sharded x{world}; {n_coll} NCCL all-reduce of {n_par * 4 >> 20} MB fp32 gradients per step',
                       'loss_scale': LOSS_SCALE},
            'schedule': ('eager launches' if args.train_eager else 'fwd+bwd replayed as one CUDA graph; all-reduce + Adam eager') +
                        ('; stage-per-stream wavefront (4 streams)' if args.train_wavefront else ''),
            'clocks': clocks, 'phases_ms': acc_ms, 'cpu_issue_ms': cpu_ms if args.train_eager else None, 'final_loss': float(loss.detach()) / LOSS_SCALE, 'grads_finite': grad_ok,
        })
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference-gpu'])
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'],
                    help="infer: BASELINE configs[1] (the headline metric); train: configs[2], the batch-sharded training step")
    ap.add_argument('--train-batch', type=int, default=3, help='samples per GPU in --mode train (BASELINE configs[2]: 3)')
    ap.add_argument('--train-wavefront', action='store_true', help='--mode train: stage-per-stream schedule (fwd and bwd)')
    ap.add_argument('--train-eager', action='store_true',
                    help='--mode train: launch forward+backward eagerly instead of replaying one CUDA graph of them')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch eagerly instead of replaying CUDA graphs')
    ap.add_argument('--no-wavefront', action='store_true', help='run the four stages strictly one after the other')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    protect_stdout()

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.impl == 'reference-gpu':
        run_reference_gpu(args, rank, world, local_rank)
        return
    if args.mode == 'train':
        run_train(args, rank, world, local_rank)
        return

    import torch.distributed as dist
    import rvt_b200
    from rvt_b200 import sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner to STDOUT at NCCL_DEBUG=VERSION (the image default); stdout must carry
        # exactly one JSON line, so keep NCCL at WARN unless the user asked for more.
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)

    model = build_model(0).to(dev).eval()
    model.pad_to_hw = (PAD_H, PAD_W)

    # weak scaling: the global batch is 8*N samples; this rank owns [lo, hi) and its states (no exchange)
    lo, hi = sharding.batch_slice(B_PER_GPU * world, rank, world)
    assert hi - lo == B_PER_GPU
    # synthetic uint8 event tensors: SEQ_LEN timesteps, resident on the device (37 MB each)
    g = torch.Generator(device='cpu').manual_seed(1234 + lo)
    shape = (SEQ_LEN, B_PER_GPU, IN_C, IN_H, IN_W)
    seq_host = torch.randint(1, 11, shape, generator=g, dtype=torch.uint8)         # counts 1..10 ...
    seq_host.mul_(torch.randint(0, 10, shape, generator=g, dtype=torch.uint8) == 0)  # ... on ~10 % of the bins
    seq_host = seq_host.pin_memory()
    seq_dev = seq_host.to(dev)

    wavefront = not args.no_wavefront

    def run_sequence_resident():
        _, st = model.forward_sequence(seq_dev, None, wavefront=wavefront)
        return st

    feat_host = torch.empty((B_PER_GPU, 512, PAD_H // 32, PAD_W // 32), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(dev)

    seq_stage = torch.empty_like(seq_dev)        # device landing buffers of the per-timestep H2D copies
    d2h_stream = torch.cuda.Stream(dev)

    def run_sequence_e2e():
        """Public API with HOST buffers: every timestep's uint8 event tensor is copied H2D from pinned
        memory (copy stream, overlapping compute), the sequence runs through
        RNNDetector.forward_sequence, and every timestep's stage-4 feature map is read back D2H."""
        main_s = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main_s)
        ready = []
        with torch.cuda.stream(copy_stream):
            for tstep in range(SEQ_LEN):
                seq_stage[tstep].copy_(seq_host[tstep], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append(ev)
        outs, st = model.forward_sequence(seq_stage, None, wavefront=wavefront, input_ready=ready)
        with torch.cuda.stream(d2h_stream):
            for tstep in range(SEQ_LEN):
                d2h_stream.wait_event(model.last_step_events[tstep])
                feat_host.copy_(outs[tstep][4], non_blocking=True)
        main_s.wait_stream(d2h_stream)
        main_s.wait_stream(copy_stream)
        return st

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)     # device time, max over ranks
        barrier()
        return ms

    with torch.inference_mode():
        if args.no_graph:
            step_resident, step_e2e = run_sequence_resident, run_sequence_e2e
        else:
            # whole-sequence CUDA graphs (rvt_b200.graph): kernels + H2D/D2H copies of all 4+2 streams
            step_resident = rvt_b200.GraphedCallable(run_sequence_resident, warmup=2)
            step_e2e = rvt_b200.GraphedCallable(run_sequence_e2e, warmup=2)
        for _ in range(max(args.warmup, 3)):
            step_resident()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ms = timed(step_resident, args.steps)
        clocks = sampler.stop() if rank == 0 else None
        for _ in range(3):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel, timed live with CUDA events on its launch stream ----
    # attn_fused_kernel, stage-1 instance (largest share of the step in profiles/launches_r01.txt): one
    # launch = the attention half of one PartitionAttentionCl block over 8 x 96 x 160 tokens, C = 64.
    roof = None
    if rank == 0:
        from rvt_b200 import ops
        with torch.inference_mode():
            pk = model._ensure_packed(dev)[0]
            blk = pk['blocks'][0]
            n_tok, c, P = B_PER_GPU * 96 * 160, 64, 60
            bufs = [torch.randn(B_PER_GPU, 96, 160, c, device=dev) for _ in range(10)]   # 315 MB > L2: cold x every launch
            dummy = torch.empty(1, dtype=torch.float16, device=dev)
            for xb in bufs[:3]:
                ops.partition_attention_(xb, blk, dummy, dummy, dummy)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                for xb in bufs:
                    ops.partition_attention_(xb, blk, dummy, dummy, dummy)
            e1.record()
            torch.cuda.synchronize(dev)
            us = e0.elapsed_time(e1) * 1e3 / (reps * len(bufs))
            del bufs
        algo_bytes = 2 * n_tok * c * 4                               # x read + x written, fp32
        algo_flops = 8 * n_tok * c * c + 4 * n_tok * P * c           # qkv + proj + QK^T + PV
        roof = {'kernel': 'attn_fused_kernel (stage 1, C=64, 122880 tokens)', 'us_per_launch': us,
                'algorithmic_bytes': algo_bytes, 'algorithmic_flops': algo_flops}

    frames = B_PER_GPU * SEQ_LEN * args.steps * world
    value = frames / (ms * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_gbs = peaks.get('hbm_gbs', 6650.0)
        peak_src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)'
        ach_tf = value / world * GFLOP_PER_FRAME / 1e3
        ach_gbs = roof['algorithmic_bytes'] / roof['us_per_launch'] / 1e3
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic_r01.json'))).get('attn_fused_s1_dram_bytes')
        except Exception:
            pass
        line = {
            'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': 'RVT-Base 1Mpx 8x20x360x640 uint8 (model res 384x640) seq_len=21 bs=8/GPU '
                                   'inference, states carried', 'frames_per_step': B_PER_GPU * SEQ_LEN,
                       'l2_policy': 'inputs larger than L2 (774 MB of uint8 sequences per GPU)',
                       'parallelism': f'batch-sharded x{world}, no collective',
                       'schedule': ('wavefront over 4 streams' if wavefront else 'sequential') +
                                   (', eager launches' if args.no_graph else ', CUDA-graph replay')},
            'e2e': {'value': e2e, 'unit': 'frames/s',
                    'h2d_bytes_per_step': SEQ_LEN * B_PER_GPU * IN_C * IN_H * IN_W,
                    'd2h_bytes_per_step': SEQ_LEN * feat_host.numel() * 4},
            'gpu_launches': LAUNCHES_PER_TIMESTEP * SEQ_LEN * args.steps,
            'clocks': clocks,
            # C = 64: 94 FLOP/B algorithmic intensity, below the 263 FLOP/B ridge -> HBM is the bounding roof
            'roofline': {'bound': 'hbm', 'achieved': ach_gbs, 'peak': peak_gbs, 'unit': 'GB/s',
                         'frac': ach_gbs / peak_gbs, 'traffic': traffic, 'kernel': roof['kernel'],
                         'us_per_launch': roof['us_per_launch'], 'peak_source': peak_src,
                         'tensor_tflops': roof['algorithmic_flops'] / roof['us_per_launch'] / 1e6,
                         'note': 'algorithmic bytes = x read + written (fp32) = 62.9 MB per launch; see DESIGN.md §6'},
            'whole_step': {'algorithmic_tflops': ach_tf, 'frac_of_sustained_bf16_peak': ach_tf / peak_tf,
                           'gflop_per_frame': GFLOP_PER_FRAME},
        }
        if not args.no_cpu_baseline:
            cores = best_cpu_threads()
            v = cpu_reference_fps(2, B_PER_GPU, cores)
            line['cpu_baseline'] = {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                    'sample': '2 timesteps x batch 8 (after 1 warm-up timestep), fp32 torch CPU ops; '
                                              f'{cores} threads = best of a sweep up to os.cpu_count()={os.cpu_count()}'}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
