"""CUDA-graph replay of a whole step for the launch-bound shapes of the path (single process, no collectives inside).

At small batches (SPARC at 64 samples: ~0.3 ms of device work behind ~150 kernel launches and a dozen allocations) the step
time is host time.  The C ABI never synchronises and takes every buffer from the caller, and the tensor maps are encoded
on the host with fixed addresses, so a full forward + backward is capturable: `GraphedStep` warms the step up on a side
stream, captures it once with `torch.cuda.graph` and replays it.  Inputs are STATIC tensors: copy new data into them
(`.copy_()`) before `replay()`; gradients land in the `.grad` tensors of the captured run (also static).  Steps with a
host round trip (the ragged hard-negative gather of `OpenClipLoss` reads the per-rank counts) cannot be captured, and
steps with NCCL collectives inside are not supported by this helper (a first attempt hung during capture).
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, warmup=3, capture_error_mode="global"):
        """`step_fn()` runs one forward + backward on static tensors and returns the loss tensor.
        `capture_error_mode="thread_local"` tolerates CUDA calls of other threads during capture (the NCCL watchdog
        of `torch.distributed` polls events) -- needed when the step contains collectives."""
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.loss = step_fn()

    def replay(self):
        self.graph.replay()
        return self.loss
