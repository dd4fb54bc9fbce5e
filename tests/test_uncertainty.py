"""Uncertainty-weighting epilogue (reference mgnet/modeling/mg_net.py:360-372)."""
import pytest
import torch

KEYS = ["loss_sem_seg", "loss_center", "loss_offset", "loss_photometric", "loss_smoothness"]   # dict order in MGNet.forward
TOL = 1e-6   # fp32 relative: expf on the device vs torch.exp, everything else rounds identically


def test_cpu_tensors_fail_loudly():
    from mgnet_b200.uncertainty import apply_uncertainty
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        apply_uncertainty({"loss_photometric": torch.tensor(0.1)}, torch.zeros(5))
    with pytest.raises(IndexError):
        apply_uncertainty({k: torch.tensor(0.1) for k in KEYS}, torch.zeros(3))


@pytest.mark.gpu
@pytest.mark.parametrize("keys", [KEYS, KEYS[3:], KEYS[:1]])
def test_matches_reference_expression(keys):
    from mgnet_b200.uncertainty import apply_uncertainty
    from oracle.torch_port import reference_uncertainty
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    vals = (torch.rand(len(keys), generator=g) * 2 + 0.01)
    lv0 = torch.randn(5, generator=g) * 0.7
    up = torch.rand(len(keys), generator=g) + 0.5

    def run(fn, device):
        leaves = [v.clone().to(device).requires_grad_(True) for v in vals]
        lv = lv0.clone().to(device).requires_grad_(True)
        log = {}
        if fn is apply_uncertainty:
            out = fn(dict(zip(keys, leaves)), lv, log)
        else:
            out, log = fn(dict(zip(keys, leaves)), lv)
        total = sum(u * out[k] for u, k in zip(up.to(device), keys))
        total.backward()
        return ([out[k].detach().cpu() for k in keys], [x.grad.cpu() for x in leaves], lv.grad.cpu(),
                {k: float(v) for k, v in log.items()})

    ours = run(apply_uncertainty, dev)
    ref = run(reference_uncertainty, "cpu")
    for a, b in zip(ours[0], ref[0]):
        assert abs(float(a) - float(b)) <= TOL * abs(float(b))
    for a, b in zip(ours[1], ref[1]):
        assert abs(float(a) - float(b)) <= TOL * abs(float(b))
    assert torch.allclose(ours[2], ref[2], rtol=TOL, atol=1e-7)
    assert float(ours[2][len(keys):].abs().sum()) == 0.0          # unused log_vars get no gradient
    assert set(ours[3]) == set(ref[3])
    for k in ref[3]:
        assert abs(ours[3][k] - ref[3][k]) <= TOL * abs(ref[3][k])


@pytest.mark.gpu
def test_with_the_fused_loss_end_to_end():
    """The two depth losses straight out of the fused loss module, weighted, backward through both."""
    from mgnet_b200 import MultiViewPhotometricLoss
    from mgnet_b200.synthetic import make_inputs
    from mgnet_b200.uncertainty import apply_uncertainty
    dev = torch.device("cuda:0")
    pred, tgt = make_inputs(2, 64, 128, 3, seed=9, noise=0.0, shift_sources=True)
    mod = MultiViewPhotometricLoss(0.85, 1.0, 1e-3, True, "min", "zeros")
    t = {k: v.to(dev) for k, v in tgt.items()}
    grads = []
    for weighted in (False, True):
        p = {"depth": [d.to(dev).requires_grad_(True) for d in pred["depth"]], "poses": pred["poses"].to(dev).requires_grad_(True)}
        lv = torch.tensor([0.3, -0.2], device=dev, requires_grad=True)
        out = mod(p, t)
        if weighted:
            w = apply_uncertainty(out, lv)
            (w["loss_photometric"] + w["loss_smoothness"]).backward()
        else:
            (0.5 * torch.exp(-lv[0].detach()) * out["loss_photometric"] + 0.5 * torch.exp(-lv[1].detach()) * out["loss_smoothness"]).backward()
        grads.append([d.grad.clone() for d in p["depth"]] + [p["poses"].grad.clone()])
    for a, b in zip(*grads):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-12)
