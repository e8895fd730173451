"""Whole training step (forward, fused loss, backward, optimizer) captured in ONE CUDA graph.

Every kernel of libb200unet.so launches on the stream it is given, never synchronises and never allocates, so a
training step is capturable as is.  For small / narrow configurations (e.g. the repo's feature net, 3x192x640) the
eager step is bound by Python launch overhead; one `cudaGraphLaunch` removes it.

    step = GraphedTrainStep(model, torch.optim.Adam(model.parameters(), capturable=True, fused=True), x, y)
    loss = step(x_new, y_new)        # copies into the static inputs, replays, returns the (static) loss tensor

Data parallel (`model` is a b200unet.ddp.DataParallel with more than one rank): the step becomes TWO graphs with one
eager collective between them — graph 1 = forward + loss + backward writing every gradient into the flat arena, then a
single `all_reduce(arena, AVG)`, then graph 2 = the optimizer.  The narrow networks this is meant for (the repo's
feature net: ~10^5 parameters) have nothing to overlap the reduction with anyway.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, x: torch.Tensor, y: torch.Tensor, warmup: int = 3):
        self.model, self.opt = model, optimizer
        self.x, self.y = x.clone(), y.clone()
        # warm-up on a side stream: autograd nodes born on the legacy default stream would invalidate the capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib
        n0 = _lib.load().b200unet_launch_count()
        self._bucketer = getattr(model, "bucketer", None)
        self._ddp = self._bucketer is not None and self._bucketer.world > 1
        if not self._ddp:
            with torch.cuda.graph(self.graph):
                self.loss = self._step()
        else:
            self._bucketer.defer = True
            with torch.cuda.graph(self.graph):
                loss = self.model.loss(self.x, self.y)
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                self.loss = loss.detach()
            self._arena = self._bucketer.deferred_arena
            if self._arena is None:
                raise RuntimeError("GraphedTrainStep: the data-parallel backward did not hand over its gradient arena")
            self._bucketer.reduce_all(self._arena)
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self.opt.step()
        self.launches_per_replay = int(_lib.load().b200unet_launch_count() - n0)  # kernels of this library in the graph(s)

    def _step(self) -> torch.Tensor:
        loss = self.model.loss(self.x, self.y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        if self._ddp:
            self._bucketer.reduce_all(self._arena)
            self.graph_opt.replay()
        # the replay updated the weights without running any Python of the module: bf16 operand copies an eval-mode
        # forward made earlier are stale now (UNet._packed reuses them only within one value of this counter)
        m = getattr(self.model, "module", self.model)
        if hasattr(m, "_train_forwards"):
            m._train_forwards += 1
        return self.loss
