"""Data parallelism for the PFN step: one process per GPU (torchrun), independent prior draws per rank, ONE NCCL
all-reduce of the flat gradient buffer per optimizer step.  The reference has no distributed
code at all; the semantics implemented here are the ones that reproduce its single-device gradient of the
global-batch mean: grads are summed over ranks, divided by world size, THEN clipped (reference train.py:95-96),
and `single_eval_pos` is identical on every rank for a given step (reference train.py:69 draws one per step).
"""
import os

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def env_world_size():
    return int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(device_type="cuda"):
    """Initialise torch.distributed from torchrun's environment (no-op for a single process).  Returns
    (rank, world_size, device)."""
    world = env_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if device_type == "cuda" and torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    else:
        device = torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if device.type == "cuda" else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            dist.init_process_group(backend=backend, device_id=device)
        else:
            dist.init_process_group(backend=backend)
    rank = dist.get_rank() if dist.is_initialized() else 0
    return rank, world, device


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s weights (ranks seed torch differently for distinct prior draws)."""
    if world_size() == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers() if b.is_floating_point()]
    if not tensors:
        return
    flat = _flatten_dense_tensors(tensors)
    dist.broadcast(flat, src)
    for t, synced in zip(tensors, _unflatten_dense_tensors(flat, tensors)):
        t.copy_(synced)


def allreduce_gradients(params):
    """Average the gradients of `params` over all ranks with a single all-reduce of one flat buffer."""
    w = world_size()
    if w == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = _flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(w)
    for g, synced in zip(grads, _unflatten_dense_tensors(flat, grads)):
        g.copy_(synced)


def broadcast_object(obj, src=0):
    """Broadcast a small picklable python object (e.g. the epoch's single_eval_pos schedule)."""
    if world_size() == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def allreduce_mean_scalar(t):
    """Mean over ranks of a scalar / small tensor (loss bookkeeping once per epoch)."""
    if world_size() == 1:
        return t
    t = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t / world_size()
