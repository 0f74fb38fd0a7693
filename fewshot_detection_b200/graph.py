"""CUDA-graph capture of the whole meta-training step.

The eager step issues ~3,000 kernel launches through ctypes (~37 ms of host time
per step at config 2, about as long as the GPU work).  `GraphedTrainStep`
captures forward + RegionLossV2 + backward (+ SGD) once into a CUDA graph and
replays it: the host cost drops to a few launches and the GPU runs back to back.

Semantics kept from the eager loop (train_meta.py:201-226):
  * inputs are copied into static device buffers before each replay (the query
    batch, the support images/masks and the float64 target tensor);
  * the learning-rate schedule keeps working: lr / momentum / weight decay are read
    by the fused SGD kernel from device memory (`FusedSGD.sync_hyper`);
  * `region_loss.seen` only selects the warm-up branch of build_targets
    (seen < 12800); the graph is re-captured when that regime changes;
  * multi-GPU: the gradient all-reduce runs between two graphs (backward | SGD).
Requirements: `cfg.neg_ratio == 'full'` (the numeric neg_filter draws host random
numbers every step), fixed shapes, at least one eager step done before capture
(so that lazy initialisation and the momentum buffers exist).
"""
import torch

from .cfg import cfg
from .distributed import GradAllReducer


class GraphedTrainStep(object):
    def __init__(self, model, region_loss, optimizer, reducer=None):
        self.model, self.loss_mod, self.opt = model, region_loss, optimizer
        self.reducer = reducer if reducer is not None else GradAllReducer(model)
        self.graph_fb = None
        self.graph_opt = None
        self.static = None
        self.regime = None
        self.loss = None

    def _eager(self, x, metax, mask, target):
        self.reducer.begin_step()
        out = self.model(x, metax, mask)
        loss = self.loss_mod(out, target)
        loss.backward()
        self.reducer.finish()
        self.opt.step()
        return loss

    def _capture(self, x, metax, mask, target):
        if cfg.neg_ratio != 'full':
            raise RuntimeError("GraphedTrainStep needs cfg.neg_ratio == 'full' (neg_filter draws host random numbers)")
        dev = x.device
        self.static = [t.clone() for t in (x, metax, mask)] + [target.to(dev).clone()]
        self.regime = self.loss_mod.seen < 12800
        self.loss_mod.verbose = False
        self.opt.capturable = True
        self.opt.sync_hyper()
        self.opt.prepare()
        self.reducer.overlap = False
        torch.cuda.synchronize()
        multi = self.reducer.world > 1
        self.graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fb):
            self.reducer.begin_step()
            out = self.model(*self.static[:3])
            self.loss = self.loss_mod(out, self.static[3])
            self.loss.backward()
            if not multi:
                self.opt.step()
        if multi:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph_fb.pool()):
                self.opt.step()

    def __call__(self, x, metax, mask, target):
        """One training step. Returns the (static, device) loss tensor."""
        if not any('momentum_buffer' in self.opt.state[p] for g in self.opt.param_groups for p in g['params']):
            return self._eager(x, metax, mask, target)      # the very first step runs eagerly
        if self.graph_fb is None or (self.loss_mod.seen < 12800) != self.regime:
            self._capture(x, metax, mask, target)
        for s, t in zip(self.static, (x, metax, mask, target)):
            s.copy_(t, non_blocking=True)
        self.opt.sync_hyper()
        self.graph_fb.replay()
        if self.graph_opt is not None:
            self.reducer.finish()          # one NCCL all-reduce over the flat gradient buffer
            self.graph_opt.replay()
        return self.loss
