"""CUDA-graph capture of the loss step ("with CUDA-graph capture of the loss").

``Loss.forward`` + ``loss.backward()`` is four launches of ours plus autograd's bookkeeping; replayed from a graph the
step costs its kernels and nothing else (bench.py measures exactly this).  ``GraphedLossStep`` captures it once on a set
of STATIC tensors -- the caller refreshes them in place (``tensor.copy_(...)``, or lets the producing nets write into them)
and calls ``replay()``; the loss scalars and the ``.grad`` of the leaf tensors are refreshed by the replay.
"""
from __future__ import annotations

import torch


class GraphedLossStep:
    """Captures ``losses = loss(inputs, frame_ids, flows, mobiles, instances, scales, cams)[1]; losses["loss"].backward()``.

    flows / mobiles / cams whose tensors require grad must be LEAF tensors; after ``replay()`` their ``.grad`` holds the
    gradient of the step (the same graph-owned buffers every time).  ``instances`` (DS / DC) are prepared inside the
    capture, so their mask tensors must be static too."""

    def __init__(self, loss, inputs, frame_ids, flows, mobiles, instances, scales, cams, warmup=3):
        self.loss_module = loss
        self.args = (inputs, list(frame_ids), flows, mobiles, instances, list(scales), cams)
        self.leaves = [t for d in (flows, mobiles, cams) for t in d.values() if t.requires_grad]
        for t in self.leaves:
            if not t.is_leaf:
                raise ValueError("tensors that require grad must be leaves of the captured step")
        dev = next(iter(flows.values())).device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):      # warm-up outside the capture: lazy kernel attributes, allocator pools
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = self._step()
        self.loss = self.losses["loss"]

    def _step(self):
        for t in self.leaves:
            t.grad = None
        _, losses = self.loss_module(*self.args)
        losses["loss"].backward()
        return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in losses.items()}

    def replay(self):
        """Re-runs the captured step on the current contents of the static tensors; returns the (static) losses dict."""
        self.graph.replay()
        return self.losses
