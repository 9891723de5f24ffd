"""Needs >= 2 GPUs (run with `-m multigpu` on a multi-GPU box; `-m gpu` is the one-GPU suite): one process driving
models on two devices — every op must run on the device that owns its tensors."""
import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.multigpu


def test_model_on_second_device_while_first_is_current():
    """every op runs on the device that owns its tensors (stream, kernel attributes), whatever the current device is"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pkg = load_pkg()
    cuda_dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    dev1 = torch.device("cuda:1")
    torch.manual_seed(0)
    m1 = pkg.UNet3D(5, 1, init_features=16).to(dev1).train()
    m0 = pkg.UNet3D(5, 1, init_features=16).to(cuda_dev).train()
    m0.load_state_dict({k: v.to(cuda_dev) for k, v in m1.state_dict().items()})
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 5, 32, 32, 32, generator=g)
    y = (torch.rand(2, 1, 32, 32, 32, generator=g) > 0.8).float()
    outs = []
    for m, d in ((m1, dev1), (m0, cuda_dev)):
        opt = pkg.FusedAdam(m, lr=1e-3)
        opt.zero_grad()
        loss = pkg.BCEDiceLoss()(m(x.to(d)), y.to(d))
        loss.backward()
        grad = m.engine.flat_grad.detach().cpu().clone()
        opt.step()
        torch.cuda.synchronize(d)
        assert torch.isfinite(m.engine.flat_param).all()
        outs.append((loss.item(), grad))
        assert torch.cuda.current_device() == 0
    # the forward is deterministic; weight gradients are accumulated with fp32 atomics whose order differs between two
    # physical GPUs, so they agree to summation order (and Adam's first step, lr * sign(g), would amplify that noise)
    assert outs[0][0] == outs[1][0]
    assert ((outs[0][1] - outs[1][1]).norm() / outs[1][1].norm()).item() < 1e-5
