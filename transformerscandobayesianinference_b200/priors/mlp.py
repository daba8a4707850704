"""`priors.mlp` — the BNN tabular prior (reference priors/mlp.py:62-203).

Each of the `batch_size // batch_size_per_gp_sample` random MLPs draws its scalar hyperparameters on the host
(dropout_prob, noise_std, init_std, num_features_used, categorical spec, num_layers, hidden_dim; reference :92-104),
gets every weight and bias ~ N(0, (init_std / (1 - p))^2) * Bernoulli(1 - p) (p = 0 for the first tensor; :126-130)
and is evaluated on fresh N(0,1) / U(0,1) causes for each of its datasets (:133-197).

Device formulation: the reference calls every model `batch_size_per_gp_sample` times on a [T, 1, c] input (B tiny
launches-chains per batch).  Here all datasets of a model ride together as one [T, g, c] tensor through the same
layers — the layers act on the last dimension, the Gaussian noise is drawn per element, normalisation / median /
ordering reduce over T only — so each model is one chain of batched device ops and the distribution is unchanged.
"""
import random

import numpy as np
import torch
from torch import nn

from ..utils import default_device
from .utils import get_batch_to_dataloader
from .utils import order_by_y, normalize_data, normalize_by_used_features_f, Binarize  # noqa: F401
from .utils import (trunc_norm_sampler_f, beta_sampler_f, gamma_sampler_f, uniform_sampler_f, zipf_sampler_f,  # noqa: F401
                    scaled_beta_sampler_f, uniform_int_sampler_f)

DEFAULT_NUM_LAYERS = 2
DEFAULT_HIDDEN_DIM = 100
DEFAULT_ACTIVATION_MODULE = torch.nn.ReLU
DEFAULT_INIT_STD = .1
DEFAULT_HIDDEN_NOISE_STD = .1
DEFAULT_FIXED_DROPOUT = 0.
DEFAULT_IS_BINARY_CLASSIFICATION = False


def canonical_pre_processing(x, canonical_args):
    assert x.shape[2] == len(canonical_args)
    for dim, num_classes in enumerate(canonical_args):
        if num_classes is not None:
            rang = torch.arange(num_classes).float()
            x[:, :, dim] = (x[:, :, dim] - rang.mean()) / rang.std()
    return x


class GaussianNoise(nn.Module):
    def __init__(self, std):
        super().__init__()
        self.std = std

    def forward(self, x):
        return x + torch.normal(torch.zeros_like(x), self.std)


def causes_sampler_f(num_causes_sampler):
    num_causes = num_causes_sampler()
    means = np.random.normal(0, 1, (num_causes))
    std = np.abs(np.random.normal(0, 1, (num_causes)) * means)
    return means, std


def categorical_features_sampler(max_features):
    """Random categorical-feature spec: per feature an array of class thresholds and an is-ordinal flag (:45-59)."""
    features, ordinal = [], []
    n_cat = scaled_beta_sampler_f(0.5, .8, max_features, 0)
    classes_nominal = scaled_beta_sampler_f(0.1, 2.0, 10, 1)
    classes_ordinal = scaled_beta_sampler_f(0.1, 2.0, 200, 1)
    for _ in range(0, n_cat()):
        is_ord = random.choice([True, False])
        ordinal.append(is_ord)
        features.append(np.random.rand(classes_ordinal() if is_ord else classes_nominal()))
    return features, ordinal


def _order_groups_by_y(x, y):
    """Per-dataset version of utils.order_by_y for x [T,g,F], y [T,g,1]: each dataset gets its own random sign."""
    T, g = y.shape[0], y.shape[1]
    sign = torch.tensor([1.0 if random.randint(0, 1) else -1.0 for _ in range(g)], device=y.device)
    order = torch.argsort(y[:, :, 0] * sign, dim=0)                       # [T,g]
    order = order.reshape(2, -1, g).transpose(0, 1).reshape(T, g)         # interleave the two halves
    xo = torch.gather(x, 0, order.unsqueeze(-1).expand(-1, -1, x.shape[-1]))
    yo = torch.gather(y, 0, order.unsqueeze(-1))
    return xo, yo


def _draw_model_specs(num_models, hyperparameters):
    """Host-side scalar hyperparameters of every random MLP, drawn with the reference's samplers in the reference's
    call order (priors/mlp.py:92-104: dropout_prob, noise_std, init_std, num_features_used, categorical spec, num_layers,
    hidden_dim) for model 0, then model 1, ... (`models = [get_model() for _ in range(num_models)]`, :195)."""
    (num_layers_sampler, hidden_dim_sampler, _act, init_std_sampler, noise_std_sampler, dropout_prob_sampler, _bin,
     num_features_used_sampler, _causes, _is_causal, _psc, _psw, _yie, _oy, _nbuf, categorical_features_sampler_, _nan) = hyperparameters
    specs = []
    for _ in range(num_models):
        sp = {"dropout_prob": float(dropout_prob_sampler()), "noise_std": float(noise_std_sampler()),
              "init_std": float(init_std_sampler()), "num_features_used": int(num_features_used_sampler())}
        sp["cat_features"], sp["cat_is_ordinal"] = categorical_features_sampler_(sp["num_features_used"])
        sp["num_layers"] = int(num_layers_sampler())
        sp["hidden_dim"] = int(hidden_dim_sampler())
        assert sp["num_layers"] > 2
        specs.append(sp)
    return specs


def _replay_hyperparameters(hyperparameters, specs):
    """The same 17-tuple with the scalar samplers replaced by iterators over already-drawn `specs` (used when the
    vectorised path bails out after the host draws were consumed, so that the per-model path sees the same models)."""
    it = {k: iter([sp[k] for sp in specs]) for k in ("dropout_prob", "noise_std", "init_std", "num_features_used", "num_layers", "hidden_dim")}
    cats = iter([(sp["cat_features"], sp["cat_is_ordinal"]) for sp in specs])
    hp = list(hyperparameters)
    hp[0] = lambda: next(it["num_layers"])
    hp[1] = lambda: next(it["hidden_dim"])
    hp[3] = lambda: next(it["init_std"])
    hp[4] = lambda: next(it["noise_std"])
    hp[5] = lambda: next(it["dropout_prob"])
    hp[7] = lambda: next(it["num_features_used"])
    hp[15] = lambda n: next(cats)
    return tuple(hp)


def _get_batch_vectorized(specs, g, seq_len, num_features, device, hyperparameters, sampling, num_outputs):
    """All `len(specs)` random MLPs x `g` datasets each as one chain of batched device ops (non-causal prior without
    categorical features, reference priors/mlp.py:113-189 per dataset).  Models differ in input width, hidden width,
    init / noise scale and sparsity: tensors are padded to the batch maxima and masked, models with different depth form
    separate groups.  Distribution per dataset is the reference's: weights and biases ~ N(0, (init_std/(1-p))^2) *
    Bernoulli(1-p) (p = 0 for the first weight matrix, :126-130), causes ~ N(0,1) / U(0,1) (:133-141), Gaussian noise of
    std noise_std after every layer but the first (:113-121), x = causes, y = last output (:157-158), normalisation over the
    sequence (:177), median binarisation per dataset (:180), optional /(used/num_features) (:182-183), optional order_by_y
    with one `random.randint` sign per dataset (:185-186), zero padding to num_features (:189)."""
    activation_module, is_binary_classification, order_y = hyperparameters[2], hyperparameters[6], hyperparameters[13]
    normalize_by_used = hyperparameters[14]
    dev = torch.device(device)
    M, T = len(specs), seq_len
    out_x = torch.zeros(T, M * g, num_features, device=dev)
    out_y = torch.empty(T, M * g, device=dev)
    act = activation_module()
    signs = None
    if is_binary_classification and order_y:     # one host draw per dataset, dataset-major like the reference's sample loop
        signs = torch.tensor([1.0 if random.randint(0, 1) else -1.0 for _ in range(M * g)], dtype=torch.float32)
        signs = (signs.pin_memory() if dev.type == 'cuda' else signs).to(dev, non_blocking=True)
    by_depth = {}
    for mi, sp in enumerate(specs):
        by_depth.setdefault(sp["num_layers"], []).append(mi)
    for depth, idxs in by_depth.items():
        m = len(idxs)
        used = torch.tensor([specs[i]["num_features_used"] for i in idxs])
        hid = torch.tensor([specs[i]["hidden_dim"] for i in idxs])
        C, Hd = int(used.max()), int(hid.max())
        host = torch.stack([used.float(), hid.float(),
                            torch.tensor([specs[i]["init_std"] for i in idxs]),
                            torch.tensor([specs[i]["noise_std"] for i in idxs]),
                            torch.tensor([specs[i]["dropout_prob"] for i in idxs])])
        host = (host.pin_memory() if dev.type == 'cuda' else host).to(dev, non_blocking=True)
        used_d, hid_d, init_std, noise_std, p_drop = host[0], host[1], host[2], host[3], host[4]
        cmask = (torch.arange(C, device=dev).view(1, C) < used_d.view(m, 1)).float()        # [m, C]
        hmask = (torch.arange(Hd, device=dev).view(1, Hd) < hid_d.view(m, 1)).float()       # [m, Hd]
        scale = (init_std / (1. - p_drop)).view(m, 1, 1)
        keep = (1. - p_drop).view(m, 1, 1)

        def param(shape_tail, first, mask):
            w = torch.randn((m,) + shape_tail, device=dev)
            if first:
                w = w * init_std.view(m, 1, 1)                                                # p = 0 for the first tensor (:127)
            else:
                w = w * scale * torch.bernoulli(keep.expand((m,) + shape_tail))
            return w * mask
        # layer 0: Linear(c -> h); layers 1 .. depth-2: Linear(h -> h); last: Linear(h -> num_outputs)
        W = [param((Hd, C), True, hmask.view(m, Hd, 1) * cmask.view(m, 1, C))]
        b = [param((1, Hd), False, hmask.view(m, 1, Hd))]
        for li in range(1, depth):
            last = li == depth - 1
            od = num_outputs if last else Hd
            omask = torch.ones(m, od, device=dev) if last else hmask
            W.append(param((od, Hd), False, omask.view(m, od, 1) * hmask.view(m, 1, Hd)))
            b.append(param((1, od), False, omask.view(m, 1, od)))
        if sampling == 'normal':
            c = torch.randn(m, T * g, C, device=dev)
        elif sampling == 'uniform':
            c = torch.rand(m, T * g, C, device=dev)
        else:
            raise ValueError(f'Sampling is set to invalid setting: {sampling}.')
        c = c * cmask.view(m, 1, C)
        h = torch.baddbmm(b[0], c, W[0].transpose(1, 2))                                      # [m, T*g, Hd]
        for li in range(1, depth):
            h = torch.baddbmm(b[li], act(h), W[li].transpose(1, 2))
            h = h + torch.randn_like(h) * noise_std.view(m, 1, 1)
            if li < depth - 1:
                h = h * hmask.view(m, 1, Hd)
        # rows of the (T*g) axis are (t, dataset-in-group) pairs: -> [T, m, g, .]
        y = h.view(m, T, g, num_outputs)[..., 0].permute(1, 0, 2)                             # [T, m, g]
        x = c.view(m, T, g, C).permute(1, 0, 2, 3)                                            # [T, m, g, C]
        x = (x - x.mean(0)) / (x.std(0) + .000001)                                            # padded columns stay 0
        y = (y - y.mean(0)) / (y.std(0) + .000001)
        if is_binary_classification:
            med = torch.median(y, dim=0, keepdim=True)[0]                                     # lower median per dataset (Binarize)
            y = (y > med).float()
        if normalize_by_used:
            x = x / (used_d / num_features).view(1, m, 1, 1)
        cols = torch.tensor([i * g + k for i in idxs for k in range(g)], device=dev)
        x = x.reshape(T, m * g, C)
        y = y.reshape(T, m * g)
        if signs is not None:
            order = torch.argsort(y * signs[cols].view(1, -1), dim=0)                         # [T, m*g]
            order = order.reshape(2, -1, m * g).transpose(0, 1).reshape(T, m * g)             # interleave the two halves
            x = torch.gather(x, 0, order.unsqueeze(-1).expand(-1, -1, C))
            y = torch.gather(y, 0, order)
        out_x[:, cols, :C] = x
        out_y[:, cols] = y
    return out_x, out_y, out_y


def get_batch(batch_size, seq_len, num_features, device=default_device,
              hyperparameters=(DEFAULT_NUM_LAYERS, DEFAULT_HIDDEN_DIM, DEFAULT_ACTIVATION_MODULE, DEFAULT_INIT_STD,
                               DEFAULT_HIDDEN_NOISE_STD, DEFAULT_FIXED_DROPOUT, DEFAULT_IS_BINARY_CLASSIFICATION),
              batch_size_per_gp_sample=None, num_outputs=1, canonical_args=None, sampling='normal'):
    """-> x [T,B,num_features], y [T,B], target_y [T,B].  `hyperparameters` is the 17-tuple built by
    tabular.get_mlp_prior_hyperparameters (reference tabular.py:47-70); like the reference, shorter tuples do not
    unpack (:65)."""
    assert num_outputs == 1
    (num_layers_sampler, hidden_dim_sampler, activation_module, init_std_sampler, noise_std_sampler,
     dropout_prob_sampler, is_binary_classification, num_features_used_sampler, causes_sampler, is_causal,
     pre_sample_causes, pre_sample_weights, y_is_effect, order_y, normalize_by_used_features,
     categorical_features_sampler_, nan_prob) = hyperparameters

    sample_batch_size = batch_size
    batch_size_per_gp_sample = batch_size_per_gp_sample or sample_batch_size // 8
    assert sample_batch_size % batch_size_per_gp_sample == 0, \
        'Please choose a batch_size divisible by batch_size_per_gp_sample.'
    num_models = sample_batch_size // batch_size_per_gp_sample
    g = batch_size_per_gp_sample

    if not is_causal and torch.device(device).type == 'cuda':
        # every model of the batch in ONE chain of batched device ops (the shipped BNN-prior configuration)
        spec = _draw_model_specs(num_models, hyperparameters)
        if not any(len(sp["cat_features"]) for sp in spec):
            return _get_batch_vectorized(spec, g, seq_len, num_features, device, hyperparameters, sampling, num_outputs)
        hyperparameters = _replay_hyperparameters(hyperparameters, spec)   # the host draws were consumed: replay them
        (num_layers_sampler, hidden_dim_sampler, activation_module, init_std_sampler, noise_std_sampler,
         dropout_prob_sampler, is_binary_classification, num_features_used_sampler, causes_sampler, is_causal,
         pre_sample_causes, pre_sample_weights, y_is_effect, order_y, normalize_by_used_features,
         categorical_features_sampler_, nan_prob) = hyperparameters

    def sample_model():
        dropout_prob = dropout_prob_sampler()
        noise_std = noise_std_sampler()
        init_std = init_std_sampler()
        num_features_used = num_features_used_sampler()
        cat_features, cat_is_ordinal = categorical_features_sampler_(num_features_used)
        causes = None
        if is_causal:
            means, stds = causes_sampler()
            causes = (torch.tensor(means, device=device).view(1, 1, -1).tile((seq_len, 1, 1)),
                      torch.tensor(stds, device=device).view(1, 1, -1).tile((seq_len, 1, 1)))
            num_causes = causes[0].shape[2]
        else:
            num_causes = num_features_used
        num_layers = num_layers_sampler()
        hidden_dim = hidden_dim_sampler()
        if is_causal:
            hidden_dim = max(hidden_dim, 2 * num_features_used + 1)
        assert num_layers > 2

        dims = [(num_causes, hidden_dim)]
        for layer_idx in range(num_layers - 1):
            dims.append((hidden_dim, num_outputs if layer_idx == num_layers - 2 else hidden_dim))
        weights, noise_stds = [], []
        pidx = 0
        for li, (fin, fout) in enumerate(dims):
            tensors = []
            for shape in ((fout, fin), (fout,)):
                p_drop = dropout_prob if pidx > 0 else 0.0
                t = torch.randn(shape, device=device) * (init_std / (1. - p_drop))
                t = t * torch.bernoulli(torch.full(shape, 1. - p_drop, device=device))
                tensors.append(t)
                pidx += 1
            weights.append(tensors)
            if li > 0:
                if pre_sample_weights:
                    noise_stds.append(torch.abs(torch.normal(torch.zeros(fout, device=device), float(noise_std))))
                else:
                    noise_stds.append(noise_std)

        act = activation_module()

        def forward():
            if sampling == 'normal':
                if is_causal and pre_sample_causes:
                    c = torch.normal(causes[0].expand(-1, g, -1), causes[1].abs().expand(-1, g, -1)).float()
                else:
                    c = torch.randn(seq_len, g, num_causes, device=device)
            elif sampling == 'uniform':
                c = torch.rand(seq_len, g, num_causes, device=device)
            else:
                raise ValueError(f'Sampling is set to invalid setting: {sampling}.')
            outputs = [c]
            h = c @ weights[0][0].t() + weights[0][1]
            outputs.append(h)
            for li in range(1, len(dims)):
                h = act(h) @ weights[li][0].t() + weights[li][1]
                h = h + torch.randn_like(h) * noise_stds[li - 1]
                outputs.append(h)
            outputs = outputs[2:]

            if is_causal:
                flat = torch.cat(outputs, -1)                              # [T, g, D]
                D = flat.shape[-1]
                xs, ys = [], []
                for d in range(g):                                         # an independent permutation per dataset
                    perm = torch.randperm(D - 1, device=device)
                    iy = torch.tensor([D - 1], device=device) if y_is_effect else perm[0:num_outputs]
                    ys.append(flat[:, d:d + 1, :][:, :, iy])
                    xs.append(flat[:, d:d + 1, :][:, :, perm[num_outputs:num_outputs + num_features_used]])
                x, y = torch.cat(xs, 1), torch.cat(ys, 1)
            else:
                y = outputs[-1]
                x = c

            if len(cat_features) > 0:
                x = x.clone()
                for d in range(g):
                    perm = torch.randperm(x.shape[-1], device=device)
                    for i, (cf, is_ord) in enumerate(zip(cat_features, cat_is_ordinal)):
                        idx = perm[i]
                        temp = normalize_data(x[:, d:d + 1, idx])
                        thr = torch.tensor(cf, device=device, dtype=torch.float32).view(-1, 1, 1) - 0.5
                        q = (temp > thr).sum(axis=0)
                        x[:, d:d + 1, idx] = q if is_ord else q * (127 * len(cf) + 1) % len(cf)

            x, y = normalize_data(x), normalize_data(y)
            if is_binary_classification:
                med = torch.median(y, dim=0, keepdim=True)[0]              # per-dataset lower median (Binarize)
                y = (y > med).float()
            if normalize_by_used_features:
                x = normalize_by_used_features_f(x, num_features_used, num_features)
            if is_binary_classification and order_y:
                x, y = _order_groups_by_y(x, y)
            pad = torch.zeros((x.shape[0], x.shape[1], num_features - num_features_used), device=device)
            return torch.cat([x, pad], -1), y

        return forward

    parts = [sample_model()() for _ in range(num_models)]
    xs, ys = zip(*parts)
    y = torch.cat(ys, 1).squeeze(-1).detach()
    x = torch.cat(xs, 1).detach()
    return x, y, y


DataLoader = get_batch_to_dataloader(get_batch)
DataLoader.num_outputs = 1
