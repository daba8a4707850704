"""torchrun worker for tests/test_gpu_multi.py: one training step through train.Trainer with the overlapped bucketed
all-reduce, compared with (a) the plain end-of-step all-reduce and (b) a single-process run over the concatenated batches."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformerscandobayesianinference_b200 import bar_distribution, encoders, parallel, priors, train as T_  # noqa: E402


def build(k):
    torch.manual_seed(7)
    crit = bar_distribution.FullSupportBarDistribution(torch.linspace(-4, 4, 21))
    tr = T_.build_trainer(priors.fast_gp.DataLoader, crit, encoders.Linear, emsize=256, nhid=256, nlayers=2, nhead=2, dropout=0.0,
                          epochs=1, steps_per_epoch=2, batch_size=8, bptt=40, lr=0.0, warmup_epochs=0,
                          y_encoder_generator=encoders.Linear, extra_prior_kwargs_dict=dict(num_features=1),
                          single_eval_pos_gen=20, aggregate_k_gradients=k)
    tr.model.precision = "fp32"
    with torch.no_grad():
        g = torch.Generator().manual_seed(3)
        for l in tr.model.transformer_encoder.layers:
            l.linear2.weight.copy_(torch.randn(l.linear2.weight.shape, generator=g) * 0.05)
            l.self_attn.out_proj.weight.copy_(torch.randn(l.self_attn.out_proj.weight.shape, generator=g) * 0.05)
    return tr


def main():
    rank, world, dev = parallel.init_from_env("cuda")
    tr = build(1)
    assert tr.reducer is not None and str(dev) == tr.device
    torch.manual_seed(100 + rank)
    x = torch.rand(40, 4, 1, device=dev); y = torch.randn(40, 4, device=dev).clamp(-3, 3)
    # (1) overlapped path, but stop before the optimizer touches anything: lr = 0, grads stay in p.grad until zero_grad
    tr.optimizer.zero_grad()
    tr.aggregate_k_gradients = 10 ** 9                      # never step inside .step(); we finish by hand
    from transformerscandobayesianinference_b200 import engine
    out = tr.model((x, y), single_eval_pos=20)
    loss = tr.criterion.to(dev)(out.reshape(-1, 20), y[20:].flatten()).mean()
    tr.reducer.install(engine)
    loss.backward()
    tr.reducer.uninstall(engine)
    tr.reducer.finish(tr.params)
    got = [p.grad.clone() for p in tr.params]
    # (2) plain path
    tr.optimizer.zero_grad()
    out = tr.model((x, y), single_eval_pos=20)
    loss2 = tr.criterion(out.reshape(-1, 20), y[20:].flatten()).mean()
    loss2.backward()
    parallel.allreduce_gradients(tr.params)
    plain = [p.grad.clone() for p in tr.params]
    for a, b, p in zip(got, plain, tr.params):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), f"overlapped != plain for a tensor of shape {tuple(p.shape)}"
    # (3) single-device gradient of the global batch (gather the shards on every rank)
    xs = [torch.zeros_like(x) for _ in range(world)]; ys = [torch.zeros_like(y) for _ in range(world)]
    dist.all_gather(xs, x); dist.all_gather(ys, y)
    X, Y = torch.cat(xs, 1), torch.cat(ys, 1)
    tr.optimizer.zero_grad()
    out = tr.model((X, Y), single_eval_pos=20)
    tr.criterion(out.reshape(-1, 20), Y[20:].flatten()).mean().backward()
    for a, p in zip(got, tr.params):
        assert torch.allclose(a, p.grad, rtol=2e-4, atol=1e-6), f"DP gradient != global-batch gradient, shape {tuple(p.shape)}"
    # the per-rank prior draws differ (seeding in Trainer)
    (bx, by), _ = tr.dl.gbm(**tr.dl.get_batch_kwargs, fuse_x_y=False)
    bxs = [torch.zeros_like(bx) for _ in range(world)]
    dist.all_gather(bxs, bx.contiguous())
    assert not torch.equal(bxs[0], bxs[1])
    if rank == 0:
        print("DP_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
