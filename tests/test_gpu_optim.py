"""optim.FusedClipAdam (csrc/optimizer.cu) vs torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (reference train.py:55,94-97)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from transformerscandobayesianinference_b200 import optim


def _params(dev, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(512, 1536), (100, 1024), (7,), (1024,), (513, 129), (1,), (128, 128)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


@pytest.mark.parametrize("max_norm,wd", [(1.0, 0.0), (None, 0.0), (1.0, 0.01)])
def test_fused_clip_adam_matches_torch(cuda_device, max_norm, wd):
    ours, ref = _params(cuda_device, 0), _params(cuda_device, 0)
    opt = optim.FusedClipAdam(ours, lr=3e-3, weight_decay=wd, max_grad_norm=max_norm)
    topt = torch.optim.Adam(ref, lr=3e-3, weight_decay=wd)
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        scale = 10.0 if step % 2 == 0 else 1e-3          # alternately clipped / not clipped
        for a, b in zip(ours, ref):
            gr = (torch.randn(a.shape, generator=g) * scale).to(cuda_device)
            a.grad, b.grad = gr.clone(), gr.clone()
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in ref))
        if max_norm:
            torch.nn.utils.clip_grad_norm_(ref, max_norm)
        topt.step()
        opt.step()
        if max_norm:
            assert abs(opt.last_grad_norm_sq.sqrt().item() - total.item()) <= 1e-4 * total.item()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=1e-4, atol=2e-6), (step, a.shape, (a - b).abs().max().item())   # lr 3e-3: a 1e-3 relative slip of one update
    for a, b in zip(ours, ref):
        sa, sb = opt.state[a], topt.state[b]
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"], rtol=1e-4, atol=1e-5)        # gradients of scale 10: fma-vs-mul rounding near zero crossings
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"], rtol=1e-4, atol=1e-6)
        assert float(sa["step"]) == float(sb["step"]) == 6
    # state_dict round trip into a torch.optim.Adam of the same layout
    topt2 = torch.optim.Adam(_params(cuda_device, 0), lr=3e-3, weight_decay=wd)
    topt2.load_state_dict(opt.state_dict())


def test_bf16_shadow_follows_the_parameter(cuda_device):
    ps = _params(cuda_device, 3)
    opt = optim.FusedClipAdam(ps, lr=1e-2, max_grad_norm=1.0)
    w = ps[0]
    assert getattr(w, "_pfn_shadow", None) is None
    c0 = optim.cast_weight(w, torch.bfloat16)
    assert torch.equal(c0, w.detach().to(torch.bfloat16))
    for p in ps:
        p.grad = torch.randn_like(p)
    opt.step()
    sh = optim.cast_weight(w, torch.bfloat16)
    assert sh.data_ptr() == w._pfn_shadow[0].data_ptr()                   # the shadow is used ...
    assert torch.equal(sh, w.detach().to(torch.bfloat16))                 # ... and mirrors the updated weight
    assert getattr(ps[2], "_pfn_shadow", None) is None                    # 1-D parameters get none
    with torch.no_grad():
        w.mul_(2.0)                                                       # an in-place edit invalidates it
    fresh = optim.cast_weight(w, torch.bfloat16)
    assert fresh.data_ptr() != w._pfn_shadow[0].data_ptr() and torch.equal(fresh, w.detach().to(torch.bfloat16))
    for p in ps:
        p.grad = torch.randn_like(p)
    opt.step()                                                            # the next step refreshes it
    assert optim.cast_weight(w, torch.bfloat16).data_ptr() == w._pfn_shadow[0].data_ptr()
    assert torch.equal(w._pfn_shadow[0], w.detach().to(torch.bfloat16))
