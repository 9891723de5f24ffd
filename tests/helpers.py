"""Shared helpers for the GPU parity tests."""
import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    den = b.norm().item()
    if den == 0.0:
        return a.norm().item()
    return ((a - b).norm() / den).item()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def to_act(ops, x_ncdhw: torch.Tensor, ld: int = None, c_off: int = 0):
    """fp32 (N,C,D,H,W) -> ActView over a fresh NDHWC bf16 buffer with pitch ld (other channels NaN-poisoned)."""
    n, c, d, h, w = x_ncdhw.shape
    ld = ld or c
    buf = torch.full((n, d, h, w, ld), float("nan"), device=x_ncdhw.device, dtype=torch.bfloat16)
    buf[..., c_off:c_off + c] = x_ncdhw.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return ops.ActView(buf, c_off, c)


def from_act(v) -> torch.Tensor:
    """ActView -> fp32 (N,C,D,H,W) through torch indexing (independent of the unpack kernel)."""
    return v.as_torch().permute(0, 4, 1, 2, 3).to(torch.float32).contiguous()


def empty_act(ops, n, c, d, h, w, device, ld=None, c_off=0, poison=True):
    ld = ld or c
    buf = torch.full((n, d, h, w, ld), float("nan") if poison else 0.0, device=device, dtype=torch.bfloat16)
    return ops.ActView(buf, c_off, c)
