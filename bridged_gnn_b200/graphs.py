"""CUDA-graph replay of a fixed step (B200-first: streams and graphs instead of a tracing compiler).

A KT-GNN training step on a bridged graph launches ~60 library kernels and ~100 small torch ops; issuing them from Python
takes ~10 ms of host time.  On one GPU at 10^6 nodes that hides under 11 ms of device time, but it bounds the
destination-partitioned step (each of N ranks has 1/N of the device work and the SAME host work) and every office-scale
step.  The reference trains hundreds of epochs on one graph with fixed shapes, so the whole step -- forward, loss,
backward, the NCCL halo exchanges and the gradient all-reduce included -- can be captured once and replayed.
"""
import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    """``fn()`` captured as ONE CUDA graph.  ``fn`` must be free of host synchronisation (no ``.item()``, no boolean-mask
    compaction) after its warm-up calls have filled every cache (CSR, masks, communicators) and must read its inputs from
    tensors that stay alive (the closure's).  Returns on every call the SAME output tensors, refreshed by the replay.
    Gradients accumulated by ``fn`` (``loss.backward()`` inside it) land in static ``.grad`` tensors, ready for an
    optimiser step after the replay."""

    def __init__(self, fn, warmup=3):
        self.fn = fn
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other threads (NCCL's watchdog) may touch the CUDA API while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
