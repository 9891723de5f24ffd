"""Fused Adam over the engine's flat parameter buffer (torch.optim.Adam semantics, utils/trainer.py:113-117):
coupled L2 weight decay, bias correction, eps outside the sqrt, no amsgrad.  One kernel launch per step."""
import torch

from . import ops
from .unet3d import UNet3D


class FusedAdam(torch.optim.Optimizer):
    """``FusedAdam(model)`` or ``FusedAdam(model.parameters(), model=model)``; lr is re-read from ``param_groups``
    every step so torch LR schedulers (ReduceLROnPlateau) keep working."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, model=None):
        if isinstance(params, UNet3D):
            model, params = params, params.parameters()
        if model is None:
            raise ValueError("FusedAdam needs the UNet3D whose flat parameter buffer it updates (model=...)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single param group (the whole model)")
        given = self.param_groups[0]["params"]
        if {id(p) for p in given} != {id(p) for p in model.parameters()} or not all(p.requires_grad for p in given):
            # one launch updates the whole flat buffer (weight decay included): a subset or frozen layers would be
            # modified behind the caller's back
            raise ValueError("FusedAdam updates every parameter of the UNet3D in one launch: pass exactly "
                             "model.parameters(), all with requires_grad=True (use torch.optim.Adam for subsets)")
        self.model = model
        self._step = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self.grad_scale = 1.0     # multiply gradients on the fly (1/world for DP sums, 1/loss_scale, clip coefficient)
        self.found_inf = None     # optional device flag: skip the update when non-zero
        self._dyn = None          # device float[3] read by a captured launch (see graph.GraphedTrainStep)
        self._dyn_host = None

    def _ensure_state(self):
        eng = self.model.engine
        dev = next(self.model.parameters()).device
        eng.prepare(dev)
        if self.exp_avg is None or self.exp_avg.numel() != eng.flat_param.numel() or self.exp_avg.device != dev:
            self.exp_avg = torch.zeros_like(eng.flat_param)
            self.exp_avg_sq = torch.zeros_like(eng.flat_param)
        return eng

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng = self._ensure_state()
        g = self.param_groups[0]
        if torch.cuda.is_current_stream_capturing():
            # the launch being recorded reads its step-dependent scalars from device memory; the step counter is
            # advanced by whoever replays the graph (refresh_dynamic_scalars)
            if self._dyn is None:
                raise RuntimeError("FusedAdam.step() under CUDA-graph capture needs refresh_dynamic_scalars() first")
            ops.adam_step(eng.flat_param, eng.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0],
                          g["betas"][1], g["eps"], g["weight_decay"], max(self._step, 1), self.grad_scale,
                          self.found_inf, bf16_shadow=eng.flat_bf16, dyn_scalars=self._dyn)
            return loss
        self._step += 1
        ops.adam_step(eng.flat_param, eng.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0],
                      g["betas"][1], g["eps"], g["weight_decay"], self._step, self.grad_scale, self.found_inf,
                      bf16_shadow=eng.flat_bf16)
        eng.external_epoch += 1
        eng._shadow_key = eng.current_key()  # the bf16 operand shadow was rewritten in the same pass
        return loss

    def refresh_dynamic_scalars(self, advance: bool = True):
        """For replays of a captured step: advance the step count and upload (lr / (1 - b1^t), sqrt(1 - b2^t),
        grad_scale) — computed in double exactly as the eager launch does — to the device scalars the captured
        launch reads.  The upload is an asynchronous copy from pinned memory on the current stream."""
        self._ensure_state()
        g = self.param_groups[0]
        if advance:
            self._step += 1
        t = max(self._step, 1)
        if self._dyn is None:
            self._dyn = torch.zeros(3, device=self.exp_avg.device, dtype=torch.float32)
            self._dyn_host = torch.zeros(3, dtype=torch.float32).pin_memory()
        bc1 = 1.0 - g["betas"][0] ** t
        bc2 = 1.0 - g["betas"][1] ** t
        self._dyn_host[0] = g["lr"] / bc1
        self._dyn_host[1] = bc2 ** 0.5
        self._dyn_host[2] = self.grad_scale
        self._dyn.copy_(self._dyn_host, non_blocking=True)

    def zero_grad(self, set_to_none: bool = True):
        # gradients live in the flat buffer; dropping the views lets the next backward zero it with one memset
        super().zero_grad(set_to_none=set_to_none)

    # ---- checkpoint format: exactly torch.optim.Adam's (utils/trainer.py:262 stores optimizer.state_dict()), so that a
    # checkpoint written here resumes under torch.optim.Adam and a reference checkpoint resumes here
    def _param_views(self, flat):
        from .engine import _slot_view
        eng = self.model.engine
        return [_slot_view(flat, eng._slots[id(p)][0], p) for p in self.param_groups[0]["params"]]

    def state_dict(self):
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        n = len(self.param_groups[0]["params"])
        group["params"] = list(range(n))
        state = {}
        if self.exp_avg is not None and self._step > 0:
            m, v = self._param_views(self.exp_avg), self._param_views(self.exp_avg_sq)
            for i in range(n):
                state[i] = {"step": torch.tensor(float(self._step)), "exp_avg": m[i].clone().contiguous(),
                            "exp_avg_sq": v[i].clone().contiguous()}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.param_groups[0]["params"]):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        state = sd.get("state", {})
        legacy = sd.get("b200")   # flat-buffer format of earlier builds of this package
        self._ensure_state()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self._step = 0
        if legacy is not None:
            self._step = int(legacy["step"])
            if legacy["exp_avg"] is not None:
                self.exp_avg.copy_(legacy["exp_avg"])
                self.exp_avg_sq.copy_(legacy["exp_avg_sq"])
            return
        if state:
            m, v = self._param_views(self.exp_avg), self._param_views(self.exp_avg_sq)
            steps = set()
            for i, st in state.items():
                i = int(i)
                m[i].copy_(st["exp_avg"].to(m[i].device))
                v[i].copy_(st["exp_avg_sq"].to(v[i].device))
                steps.add(int(float(st["step"])))
            if len(steps) != 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): FusedAdam keeps one step count")
            self._step = steps.pop()
