"""CUDA-graph capture of a whole training / inference step.

A CSA training step is ~120 kernel launches, half of them 2-3 us pieces of the compatibility glue, the
classifier head and the loss; launched one by one they leave the GPU idle between kernels.  GraphedStep
captures the step once (static input tensors, outputs and gradients live in the graph's private memory
pool) and replays it with a single launch.  Everything the hot path does is capture-safe: the C-ABI calls
only enqueue kernels on the current stream, work tables are cached on the device after the first call and
the loss scale is computed on the device (no host synchronisation inside a step).
"""
from __future__ import annotations

import torch

from . import _lib as L


class GraphedStep:
    """fn(*static_inputs) -> tensor(s), captured after `warmup` eager calls on a side stream.

    replay() re-runs the captured kernels on the current contents of the static inputs and returns the
    (static) outputs; parameter .grad tensors written by fn are static as well."""

    def __init__(self, fn, *static_inputs, warmup: int = 2):
        self.fn, self.inputs = fn, static_inputs
        # warm-up and capture run on side streams while the parameters' AccumulateGrad nodes may predate them
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*static_inputs)
        self.launches = L.launch_count() - n0   # C-ABI kernel launches recorded in the graph
        # the graph holds raw device pointers to the work tables of its step: keep them allocated while it lives
        from . import engine as _E
        self._tables = _E.pin_tables()

    def replay(self, epoch: int | None = None):
        """epoch: training-mode steps (dropout on) pass a different value per replay -- e.g. the step counter --
        so that every replay draws fresh masks: the seeds were frozen into the graph when it was captured, the
        device-resident epoch word (`csn_set_drop_epoch`) offsets them.  None leaves the word as it is."""
        if epoch is not None:
            set_drop_epoch(epoch)
        self.graph.replay()
        return self.outputs


def set_drop_epoch(epoch: int) -> None:
    """Offset of every dropout seed on the current device (0 = plain seeds), set on the current stream."""
    L.check(L.lib().csn_set_drop_epoch(int(epoch) & 0xFFFFFFFF, L.stream_ptr()), "csn_set_drop_epoch")
