"""CUDA-graph capture of whole recurrent sequences.

The per-timestep work is ~50 kernel launches spread over four streams (RNNDetector.forward_sequence's
wavefront schedule); replaying it from a captured graph removes the Python / launch overhead and lets
the GPU's scheduler see all the inter-stage parallelism at once — the B200-native replacement of the
reference's optional ``torch.compile(mode='reduce-overhead')`` (maxvit_rnn.py:43-52)."""
from typing import Callable

import torch


class GraphedCallable:
    """Capture ``fn()`` (which must only enqueue work on CUDA streams forked from the current stream and
    use static input/output buffers) after ``warmup`` eager runs; ``__call__`` replays the graph and
    returns fn's captured return value (static tensors)."""

    def __init__(self, fn: Callable[[], object], warmup: int = 2):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result


def capture_sequence(model, xs, prev_states=None, wavefront: bool = True, warmup: int = 2) -> GraphedCallable:
    """Graph of ``model.forward_sequence(xs, prev_states)``.  ``xs`` (and ``prev_states``) are the static
    buffers: refill them in place (``xs.copy_(...)``) before each replay.  Returns a callable giving
    (per-step feature dicts, final states) as static tensors."""
    def run():
        with torch.no_grad():
            return model.forward_sequence(xs, prev_states, wavefront=wavefront)
    return GraphedCallable(run, warmup)


def capture_training_step(model, fwd_bwd: Callable[[], torch.Tensor], params, warmup: int = 2):
    """Capture ``loss = fwd_bwd()`` — forward over the unrolled sequence, loss and ``loss.backward()`` through
    rvt_b200.train — into ONE CUDA graph.  ``fwd_bwd`` must read its inputs from static tensors.  After the capture every
    ``p.grad`` is a static tensor (a view of the flat gradient buffer) that each replay OVERWRITES; run the all-reduce and
    the optimizer eagerly after the replay and do not call ``zero_grad(set_to_none=True)`` any more.  The re-packing of
    the (optimizer-updated) weights is part of the graph.  Call this BEFORE any eager training step that is still
    referenced (a live loss keeps AccumulateGrad nodes bound to the default stream, which cannot join a capture).
    Returns a callable giving the static loss tensor."""
    eng = model._train_engine()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            for p in params:
                p.grad = None
            fwd_bwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for p in params:
        p.grad = None
    import gc
    gc.collect()                          # drop dead autograd graphs: their AccumulateGrad nodes pin the stream they were made on
    eng.invalidate()                      # the packed-weight refresh must be recorded in the graph
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        loss = fwd_bwd()

    def replay():
        g.replay()
        return loss
    replay.graph = g
    return replay
