"""One training step replayed from a CUDA graph.

The step is a fixed sequence of ~180 launches on two streams (engine.py); issued one by one, the few microseconds
between dependent launches add up to ~2 % of a 23 ms step and the host spends ~6 ms per step in Python.  After two
eager steps `GraphedTrainStep` captures `zero_grad -> forward -> loss -> backward -> FusedAdam.step` once and replays it;
every call is still exactly one optimizer step on the batch it is given (utils/trainer.py:177-195).  The only
step-dependent scalars (Adam bias corrections, learning rate, gradient scale) live in device memory and are refreshed
before each replay (FusedAdam.refresh_dynamic_scalars), so LR schedulers keep working.

Restrictions: FusedAdam, one process (the bucketed NCCL all-reduce of parallel.GradSync is issued eagerly), fixed
batch shape per graph (a new shape gets its own graph; more than `max_graphs` shapes fall back to eager steps)."""
import torch

from . import ops
from .optim import FusedAdam


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, eager_steps: int = 2, max_graphs: int = 2,
                 capture_collectives: bool = False):
        if not isinstance(optimizer, FusedAdam):
            raise TypeError("GraphedTrainStep needs the FusedAdam optimizer (its step reads device-side scalars)")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.eager_steps, self.max_graphs = eager_steps, max_graphs
        self.capture_collectives = capture_collectives   # record the NCCL bucket all-reduces too (data-parallel ranks)
        self._seen = {}      # shape key -> eager steps taken
        self._graphs = {}    # shape key -> (graph, static_x, static_y, static_loss, launches)
        self.disabled = None  # reason, once capture has failed or is not applicable
        self.replays = 0

    def _eager(self, x, y):
        self.optimizer.zero_grad()
        loss = self.criterion(self.model(x), y)
        loss.backward()
        self.optimizer.step()
        # the step is complete: hand out the value without its (already consumed) autograd graph.  A caller that kept
        # the previous eager step's loss alive WITH its graph made the capture of the next step fail
        # (cudaErrorStreamCaptureInvalidated at capture_end; tools/capture_probe.py variants D / K / L)
        return loss.detach()

    def _capture(self, key, x, y):
        static_x, static_y = x.clone(), y.clone()
        self.optimizer.refresh_dynamic_scalars(advance=False)
        self.model.engine._pack_key = None   # the weight-operand packing kernels must be part of the recorded step
        graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count
        # recorded without programmatic dependent launch: PDL buys 0.8 ms per step on eager launches, but a graph of
        # programmatic edges replays 0.2 ms SLOWER than one of plain edges (tools/ab_pdl.py, both orders)
        pdl_was = ops.set_pdl(False)
        try:
            with torch.cuda.graph(graph):
                static_loss = self._eager(static_x, static_y)
        finally:
            ops.set_pdl(pdl_was)
        self._graphs[key] = (graph, static_x, static_y, static_loss, ops.launch_count - l0)

    def __call__(self, x, y):
        eng = self.model.engine
        if self.disabled is None and ((eng.grad_sync is not None and not self.capture_collectives) or
                                      not self.model.training):
            self.disabled = "data-parallel gradient sync installed" if eng.grad_sync is not None else "model in eval mode"
        if self.disabled is not None or not x.is_cuda:
            return self._eager(x, y)
        key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            n = self._seen.get(key, 0)
            if n < self.eager_steps or len(self._graphs) >= self.max_graphs:
                self._seen[key] = n + 1
                return self._eager(x, y)
            try:
                self._capture(key, x, y)
            except Exception as e:  # capture is an optimisation: say why it is off and carry on eagerly
                self.disabled = f"capture failed: {e!r}"
                torch.cuda.synchronize()
                return self._eager(x, y)
            entry = self._graphs[key]
        graph, static_x, static_y, static_loss, launches = entry
        static_x.copy_(x, non_blocking=True)
        static_y.copy_(y, non_blocking=True)
        self.optimizer.refresh_dynamic_scalars(advance=True)
        graph.replay()
        eng.external_epoch += 1          # parameter memory changed behind torch's back, as in FusedAdam.step
        eng._shadow_key = eng.current_key()
        eng._pack_key = None             # an eager forward after replays must repack its weight operands
        ops.launch_count += launches
        self.replays += 1
        return static_loss.clone()   # a fresh tensor per step, as the eager path returns (4-byte copy on this stream)
