"""CPU oracle for the PFN hot path — TEST INFRASTRUCTURE ONLY.

A plain-torch (CPU, fp32 or fp64) restatement of the reference algorithm, written from the formulas and not
sharing any code with the CUDA engine.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product package never does.

Pinning: the transformer / bar-distribution / mask parts are checked against the *actual* reference modules
(imported from /root/reference by `oracle/make_golden.py`, outputs committed under tests/golden/) in
tests/test_oracle_golden.py.  The GP sampler parts restate gpytorch 1.5.0 / botorch 0.6.0 maths
(requirements.txt:2,14 — neither package is installed here and the reference has no test pinning their
outputs): **parity unpinned** for `gp_sample_ref`; it is validated distributionally and against LAPACK.

Each function cites the reference file:line it follows.
"""
import math

import torch

ICDF_HALFNORMAL_HALF = 0.6744897501960817  # HalfNormal(1).icdf(0.5)  (bar_distribution.py:85-87)


# --------------------------------------------------------------------------------------------------
# mask  (reference transformer.py:35-41, generate_D_q_matrix)
# --------------------------------------------------------------------------------------------------
def d_q_mask(sz, query_size, dtype=torch.float32):
    """Additive mask M[i, j] = 0 if (j < sz - query_size) or (i == j) else -inf."""
    train = sz - query_size
    if train < 0:                      # python slice semantics of `mask[:, train_size:]` (transformer.py:38)
        train = max(sz + train, 0)
    i = torch.arange(sz).unsqueeze(1)
    j = torch.arange(sz).unsqueeze(0)
    allowed = (j < train) | (i == j)
    return torch.zeros(sz, sz, dtype=dtype).masked_fill(~allowed, float("-inf"))


# --------------------------------------------------------------------------------------------------
# transformer forward (reference transformer.py:55-91; torch nn/modules/transformer.py:951-982;
# torch nn/functional.py:6478-6690)
# --------------------------------------------------------------------------------------------------
def gelu_erf(u):
    return 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))


def layernorm_ref(z, gamma, beta, eps=1e-5):
    mean = z.mean(-1, keepdim=True)
    var = ((z - mean) ** 2).mean(-1, keepdim=True)  # biased
    return (z - mean) / torch.sqrt(var + eps) * gamma + beta


def attention_ref(qkv, T, B, H, dh, sep):
    """qkv [T*B, 3E] (token = t*B + b) -> out [T*B, E], lse [B*H, T].  Dense softmax under d_q_mask."""
    E = H * dh
    q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]

    def heads(t):  # [T*B, E] -> [B, H, T, dh]
        return t.reshape(T, B, H, dh).permute(1, 2, 0, 3)

    qh, kh, vh = heads(q), heads(k), heads(v)
    scores = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    scores = scores + d_q_mask(T, T - sep, dtype=qkv.dtype)
    lse = torch.logsumexp(scores, -1)
    out = torch.softmax(scores, -1) @ vh  # [B,H,T,dh]
    out = out.permute(2, 0, 1, 3).reshape(T * B, E)
    return out, lse.reshape(B * H, T)


def encoder_layer_ref(h, lp, T, B, nhead, sep):
    """One post-norm layer.  lp: dict with in_w,in_b,out_w,out_b,w1,b1,w2,b2,g1,be1,g2,be2."""
    E = h.shape[1]
    qkv = h @ lp["in_w"].T + lp["in_b"]
    a, _ = attention_ref(qkv, T, B, nhead, E // nhead, sep)
    a = a @ lp["out_w"].T + lp["out_b"]
    h1 = layernorm_ref(h + a, lp["g1"], lp["be1"])
    m = gelu_erf(h1 @ lp["w1"].T + lp["b1"]) @ lp["w2"].T + lp["b2"]
    return layernorm_ref(h1 + m, lp["g2"], lp["be2"])


def embed_ref(x, y, enc_w, enc_b, yenc_w, yenc_b, sep):
    """x [T,B,F], y [T,B] -> [T*B, E]   (transformer.py:68-74)."""
    T, B, F = x.shape
    xs = x.reshape(T * B, F) @ enc_w.T + enc_b
    ys = y.reshape(T * B, 1) @ yenc_w.T + yenc_b
    train_rows = sep * B
    src = xs.clone()
    src[:train_rows] += ys[:train_rows]
    return src


def transformer_forward_ref(P, x, y, sep, nhead):
    """Full model forward.  P: dict(enc_w, enc_b, yenc_w, yenc_b, layers=[...], dec_w0, dec_b0, dec_w2, dec_b2).
    Returns logits [T - sep, B, n_out]  (decoder applied to query rows; transformer.py:85,91)."""
    T, B, _ = x.shape
    sep = sep % T if sep < 0 else sep
    h = embed_ref(x, y, P["enc_w"], P["enc_b"], P["yenc_w"], P["yenc_b"], sep)
    for lp in P["layers"]:
        h = encoder_layer_ref(h, lp, T, B, nhead, sep)
    hq = h[sep * B:]
    out = gelu_erf(hq @ P["dec_w0"].T + P["dec_b0"]) @ P["dec_w2"].T + P["dec_b2"]
    return out.reshape(T - sep, B, -1)


def params_from_state_dict(sd, nlayers, dtype=torch.float64):
    """Map a reference TransformerModel state_dict (keys as in results/*.ckpt) to the oracle's dict."""
    g = lambda k: sd[k].detach().to("cpu", dtype)
    P = {
        "enc_w": g("encoder.weight"), "enc_b": g("encoder.bias"),
        "yenc_w": g("y_encoder.weight"), "yenc_b": g("y_encoder.bias"),
        "dec_w0": g("decoder.0.weight"), "dec_b0": g("decoder.0.bias"),
        "dec_w2": g("decoder.2.weight"), "dec_b2": g("decoder.2.bias"),
        "layers": [],
    }
    for i in range(nlayers):
        p = f"transformer_encoder.layers.{i}."
        P["layers"].append({
            "in_w": g(p + "self_attn.in_proj_weight"), "in_b": g(p + "self_attn.in_proj_bias"),
            "out_w": g(p + "self_attn.out_proj.weight"), "out_b": g(p + "self_attn.out_proj.bias"),
            "w1": g(p + "linear1.weight"), "b1": g(p + "linear1.bias"),
            "w2": g(p + "linear2.weight"), "b2": g(p + "linear2.bias"),
            "g1": g(p + "norm1.weight"), "be1": g(p + "norm1.bias"),
            "g2": g(p + "norm2.weight"), "be2": g(p + "norm2.bias"),
        })
    return P


# --------------------------------------------------------------------------------------------------
# bar distribution (reference bar_distribution.py:19-33, 83-117)
# --------------------------------------------------------------------------------------------------
def bucket_idx_ref(y, borders):
    """searchsorted-left minus one, with the two edge fix-ups (bar_distribution.py:19-23).  Pure-python
    bisect so that it does not share torch.searchsorted with the reference."""
    b = borders.tolist()
    n_bars = len(b) - 1
    out = []
    for v in y.flatten().tolist():
        lo, hi = 0, len(b)
        while lo < hi:
            mid = (lo + hi) // 2
            if b[mid] < v:
                lo = mid + 1
            else:
                hi = mid
        idx = lo - 1
        if v == b[0]:
            idx = 0
        if v == b[-1]:
            idx = n_bars - 1
        out.append(idx)
    return torch.tensor(out, dtype=torch.int64).reshape(y.shape)


def halfnormal_logpdf(v, s):
    return 0.5 * math.log(2.0 / math.pi) - torch.log(s) - v ** 2 / (2.0 * s ** 2)


def bar_nll_ref(logits, y, borders, full_support=False):
    """logits [N, n_bars], y [N] -> nll [N]  (bar_distribution.py:25-33 / :89-108)."""
    n_bars = borders.numel() - 1
    widths = borders[1:] - borders[:-1]
    idx = bucket_idx_ref(y, borders)
    if full_support:
        idx = idx.clamp(0, n_bars - 1)
    logp = torch.log_softmax(logits, -1) - torch.log(widths)
    lp = logp.gather(-1, idx.unsqueeze(-1)).squeeze(-1)
    if full_support:
        s0 = widths[0] / ICDF_HALFNORMAL_HALF
        s1 = widths[-1] / ICDF_HALFNORMAL_HALF
        first = idx == 0
        last = idx == n_bars - 1
        lp = lp.clone()
        lp[first] += halfnormal_logpdf((borders[1] - y[first]).clamp(min=1e-8), s0) + torch.log(widths[0])
        lp[last] += halfnormal_logpdf(y[last] - borders[-2], s1) + torch.log(widths[-1])
    return -lp


def bar_mean_ref(logits, borders, full_support=False):
    widths = borders[1:] - borders[:-1]
    means = borders[:-1] + widths / 2
    if full_support:
        means = means.clone()
        means[0] = borders[1] - (widths[0] / ICDF_HALFNORMAL_HALF) * math.sqrt(2.0 / math.pi)
        means[-1] = borders[-2] + (widths[-1] / ICDF_HALFNORMAL_HALF) * math.sqrt(2.0 / math.pi)
    return torch.softmax(logits, -1) @ means


# --------------------------------------------------------------------------------------------------
# GP prior sample (reference priors/fast_gp.py:13-32,48-56; priors/fast_gp_mix.py:24-55,88-99)
# gpytorch maths restated; PARITY UNPINNED (gpytorch/botorch not installed, no reference test vectors).
# --------------------------------------------------------------------------------------------------
def gp_kernel_ref(x, ls, os_, noise, kernel="rbf"):
    """x [B,T,F], ls [B,F], os_ [B], noise [B] -> K [B,T,T] (float64 recommended)."""
    xs = x / ls.unsqueeze(1)
    d2 = ((xs.unsqueeze(2) - xs.unsqueeze(1)) ** 2).sum(-1)
    if kernel == "rbf":
        k = torch.exp(-0.5 * d2)
    else:
        r = torch.sqrt(d2)
        nu = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[kernel]
        e = torch.exp(-math.sqrt(2 * nu) * r)
        if nu == 0.5:
            k = e
        elif nu == 1.5:
            k = (1 + math.sqrt(3) * r) * e
        else:
            k = (1 + math.sqrt(5) * r + 5.0 / 3.0 * d2) * e
    T = x.shape[1]
    return os_.view(-1, 1, 1) * k + noise.view(-1, 1, 1) * torch.eye(T, dtype=x.dtype)


def gp_sample_ref(x, z, ls, os_, noise, kernel="rbf", jitter=0.0):
    """y = chol(K + jitter I) z, batched."""
    K = gp_kernel_ref(x, ls, os_, noise, kernel)
    if jitter:
        K = K + jitter * torch.eye(K.shape[-1], dtype=K.dtype)
    L = torch.linalg.cholesky(K)
    return (L @ z.unsqueeze(-1)).squeeze(-1), L


# --------------------------------------------------------------------------------------------------
# exact-GP predictive baseline (reference priors/fast_gp.py:88-120): for every t, condition on rows < t and score row t
# --------------------------------------------------------------------------------------------------
def gp_exact_predictive_ref(x, y, lengthscale, outputscale, noise, use_mse=False):
    """x [T,B,F], y [T,B] (float64 recommended) -> losses [T-1, B]: row k scores position t = k + 1 given rows < t, with
    the Gaussian predictive of the noisy observation (`likelihood(model(x_t))`, constant zero mean, RBF kernel).
    Straight per-t restatement (a fresh t x t solve for every t), sharing nothing with the one-factor device path."""
    T, B, F = x.shape
    xb, yb = x.transpose(0, 1), y.transpose(0, 1)
    d2 = ((xb.unsqueeze(2) - xb.unsqueeze(1)) / lengthscale).pow(2).sum(-1)
    K = outputscale * torch.exp(-0.5 * d2)
    out = []
    for t in range(1, T):
        Ktt = K[:, :t, :t] + noise * torch.eye(t, dtype=K.dtype)
        kst = K[:, :t, t]
        sol = torch.linalg.solve(Ktt, torch.stack([yb[:, :t], kst], -1))      # [B,t,2]
        mean = (kst * sol[..., 0]).sum(-1)
        var = K[:, t, t] + noise - (kst * sol[..., 1]).sum(-1)
        if use_mse:
            out.append((mean - yb[:, t]) ** 2)
        else:
            out.append(0.5 * (math.log(2 * math.pi) + torch.log(var) + (yb[:, t] - mean) ** 2 / var))
    return torch.stack(out)


# --------------------------------------------------------------------------------------------------
# encoder layer with dropout at the reference's four sites (torch nn/modules/transformer.py:961-982;
# torch nn/functional.py multi_head_attention_forward `dropout_p`), the masks being GIVEN
# --------------------------------------------------------------------------------------------------
def encoder_layer_dropout_ref(h, lp, T, B, nhead, sep, keep, scale):
    """Post-norm layer in training mode.  keep: dict of 0/1 tensors -- "attn" [B*H, T, T] (on the softmax probabilities),
    "out" [T*B, E] (dropout1), "gelu" [T*B, nhid] (after the activation), "mlp" [T*B, E] (dropout2); kept entries are
    multiplied by `scale` = 1 / (1 - p)."""
    E = h.shape[1]
    dh = E // nhead
    qkv = h @ lp["in_w"].T + lp["in_b"]
    q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
    heads = lambda t: t.reshape(T, B, nhead, dh).permute(1, 2, 0, 3)            # [B,H,T,dh]
    scores = heads(q) @ heads(k).transpose(-1, -2) / math.sqrt(dh) + d_q_mask(T, T - sep, dtype=h.dtype)
    probs = torch.softmax(scores, -1) * keep["attn"].reshape(B, nhead, T, T).to(h.dtype) * scale
    a = (probs @ heads(v)).permute(2, 0, 1, 3).reshape(T * B, E)
    a = a @ lp["out_w"].T + lp["out_b"]
    h1 = layernorm_ref(h + a * keep["out"].to(h.dtype) * scale, lp["g1"], lp["be1"])
    g = gelu_erf(h1 @ lp["w1"].T + lp["b1"]) * keep["gelu"].to(h.dtype) * scale
    m = g @ lp["w2"].T + lp["b2"]
    return layernorm_ref(h1 + m * keep["mlp"].to(h.dtype) * scale, lp["g2"], lp["be2"])
