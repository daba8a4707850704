"""Host-side helpers with the reference's names and semantics (reference utils.py): LR schedules, the
`single_eval_pos` samplers, SeqBN, the OpenAI LR rule and the argparse KEY=VAL action."""
import argparse
import math
import random

import torch
from torch import nn
from torch.optim.lr_scheduler import LambdaLR


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, num_cycles=0.5, last_epoch=-1):
    """Linear warm-up to 1 over `num_warmup_steps`, then cosine decay (reference utils.py:10-22).  `train()` steps
    it once per EPOCH, so with warmup_epochs > 0 the whole first epoch runs at lr 0 (reference train.py:56,134)."""
    def factor(step):
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        progress = (step - num_warmup_steps) / max(1, num_training_steps - num_warmup_steps)
        return max(0.0, 0.5 * (1.0 + math.cos(2.0 * math.pi * num_cycles * progress)))
    return LambdaLR(optimizer, factor, last_epoch)


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    """Linear warm-up then linear decay to zero (reference utils.py:25-51)."""
    def factor(step):
        if step < num_warmup_steps:
            return step / max(1, num_warmup_steps)
        return max(0.0, (num_training_steps - step) / max(1, num_training_steps - num_warmup_steps))
    return LambdaLR(optimizer, factor, last_epoch)


def get_openai_lr(transformer_model):
    """lr = 0.003239 - 0.0001395 ln(n_params)  (reference utils.py:54-56)."""
    n = sum(p.numel() for p in transformer_model.parameters())
    return 0.003239 - 0.0001395 * math.log(n)


def get_weighted_single_eval_pos_sampler(max_len):
    """P(sep = i) proportional to 1 / (max_len - i), i in [0, max_len)  (reference utils.py:59-65).  Uses the
    python `random` stream, so seeding `random` identically on every rank keeps sep identical across ranks."""
    positions = range(max_len)
    weights = [1.0 / (max_len - i) for i in positions]
    return lambda: random.choices(positions, weights)[0]


def get_uniform_single_eval_pos_sampler(max_len):
    """Uniform over [0, max_len)  (reference utils.py:68-73)."""
    positions = range(max_len)
    return lambda: random.choices(positions)[0]


class SeqBN(nn.Module):
    """BatchNorm1d over all T*B token rows (reference utils.py:76-86); stays in PyTorch (off by default)."""

    def __init__(self, d_model):
        super().__init__()
        self.bn = nn.BatchNorm1d(d_model)
        self.d_model = d_model

    def forward(self, x):
        assert self.d_model == x.shape[-1]
        return self.bn(x.reshape(-1, self.d_model)).reshape(x.shape)


def set_locals_in_self(locals):
    obj = locals['self']
    for name, value in locals.items():
        if name != 'self':
            setattr(obj, name, value)


default_device = 'cuda:0' if torch.cuda.is_available() else 'cpu:0'


class StoreDictKeyPair(argparse.Action):
    """`--flag K1=V1 K2=V2` -> dict, values eval'd when they are python literals/expressions (reference utils.py:99-113)."""

    def __init__(self, option_strings, dest, nargs=None, **kwargs):
        self._nargs = nargs
        super().__init__(option_strings, dest, nargs=nargs, **kwargs)

    def __call__(self, parser, namespace, values, option_string=None):
        parsed = {}
        for item in values:
            key, raw = item.split("=")
            try:
                parsed[key] = eval(raw)
            except NameError:
                parsed[key] = raw
        setattr(namespace, self.dest, parsed)
        print("dict values: {}".format(parsed))
