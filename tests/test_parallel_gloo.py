"""Data-parallel plumbing on CPU: world_size 2, gloo backend."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from transformerscandobayesianinference_b200 import parallel
    r, w, dev = parallel.init_from_env(device_type="cpu")
    assert (r, w) == (rank, world) and parallel.world_size() == world
    torch.manual_seed(100 + rank)                          # ranks seed differently (distinct prior draws) ...
    model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    parallel.broadcast_parameters(model)                   # ... so weights are synchronised explicitly
    flat0 = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, flat0)
    assert all(torch.equal(gathered[0], g) for g in gathered)
    # global batch of 8 rows, sharded 4 + 4: averaged shard gradients == gradient of the global-batch mean
    g = torch.Generator().manual_seed(7)
    X, Y = torch.randn(8, 4, generator=g), torch.randn(8, 1, generator=g)
    xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
    ((model(xs) - ys) ** 2).mean().backward()
    params = list(model.parameters())
    parallel.allreduce_gradients(params)
    ref = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    ref.load_state_dict(model.state_dict())
    ((ref(X) - Y) ** 2).mean().backward()
    for p, q in zip(params, ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-6)
    # train.build_trainer under WORLD_SIZE=2: identical weights, DIFFERENT prior draws per rank (the ranks arrive with the
    # same generator state), the local batch is global / world, the python `random` stream (single_eval_pos) stays shared
    import random
    import numpy as np
    from transformerscandobayesianinference_b200 import encoders, priors, train as T_
    torch.manual_seed(5); np.random.seed(5); random.seed(5)
    tr = T_.build_trainer(priors.ridge.DataLoader, torch.nn.MSELoss(reduction='none'), encoders.Linear, emsize=16, nhid=32,
                          nlayers=1, nhead=2, dropout=0.0, epochs=1, steps_per_epoch=2, batch_size=6, bptt=5, lr=1e-3,
                          warmup_epochs=0, y_encoder_generator=encoders.Linear,
                          extra_prior_kwargs_dict=dict(num_features=3), single_eval_pos_gen=2, gpu_device='cpu')
    assert tr.dl.get_batch_kwargs['batch_size'] == 3
    w = torch.cat([p.detach().flatten() for p in tr.model.parameters()])
    ws = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    assert torch.equal(ws[0], ws[1])
    (x, y), _ = tr.dl.gbm(**tr.dl.get_batch_kwargs, fuse_x_y=False)
    xs_ = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(xs_, x.contiguous())
    assert not torch.equal(xs_[0], xs_[1]), "ranks drew the same prior batch"
    draws = [torch.zeros(2) for _ in range(world)]
    dist.all_gather(draws, torch.tensor([random.random(), float(np.random.rand() )]))
    assert draws[0][0] == draws[1][0] and draws[0][1] != draws[1][1]
    seps = parallel.broadcast_object([3, 1, 4] if rank == 0 else None)
    assert seps == [3, 1, 4]
    m = parallel.allreduce_mean_scalar(torch.tensor(float(rank)))
    assert m.item() == pytest.approx(0.5)
    ret[rank] = True
    dist.destroy_process_group()


def test_gloo_world2_dp_semantics():
    world = 2
    port = 29600 + os.getpid() % 300
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
