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
