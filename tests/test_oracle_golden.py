"""Pins the CPU oracle (oracle/pfn_oracle.py) against outputs of the UNMODIFIED reference modules
(tests/golden/*.pt, written by oracle/make_golden.py from /root/reference)."""
import os

import pytest
import torch

from oracle import pfn_oracle as O
from oracle.make_golden import MODEL_CASES, build_case_weights, case_inputs, case_borders, checksum

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ref_like_model(case):
    """Same construction path as the reference model, via torch.nn directly (no reference import needed)."""
    from torch import nn

    class Shell(nn.Module):
        def __init__(self, enc, yenc):
            super().__init__()
            layer = nn.TransformerEncoderLayer(case["E"], case["H"], case["nhid"], 0.0, activation='gelu')
            self.transformer_encoder = nn.TransformerEncoder(layer, case["L"], enable_nested_tensor=False)
            self.encoder, self.y_encoder = enc, yenc
            self.decoder = nn.Sequential(nn.Linear(case["E"], case["nhid"]), nn.GELU(), nn.Linear(case["nhid"], case["n_out"]))
            for l in self.transformer_encoder.layers:
                for t in (l.linear2.weight, l.linear2.bias, l.self_attn.out_proj.weight, l.self_attn.out_proj.bias):
                    nn.init.zeros_(t)
    return build_case_weights(Shell, case)


def test_mask_known_answers():
    gold = torch.load(os.path.join(GOLD, "mask.pt"))
    for key, ref in gold.items():
        sz, q = map(int, key.split("_"))
        assert torch.equal(O.d_q_mask(sz, q), ref), key
    m = O.d_q_mask(6, 2)
    ninf = float("-inf")
    assert m[0].tolist() == [0, 0, 0, 0, ninf, ninf] and m[4].tolist() == [0, 0, 0, 0, 0, ninf]
    assert m[5].tolist() == [0, 0, 0, 0, ninf, 0]


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_oracle_model_matches_reference(name):
    gold = torch.load(os.path.join(GOLD, f"model_{name}.pt"))
    case = gold["case"]
    model = _ref_like_model(case)
    cs = checksum(model.state_dict())
    for k, (s, a) in gold["weights_checksum"].items():
        assert abs(cs[k][0] - s) <= 1e-9 * (abs(s) + 1) and abs(cs[k][1] - a) <= 1e-9 * (a + 1), f"weights differ: {k}"
    x, y = case_inputs(case)
    P = O.params_from_state_dict(model.state_dict(), case["L"], torch.float64)
    leaves = [P["enc_w"], P["dec_w2"], P["layers"][0]["in_w"], P["layers"][-1]["w2"], P["layers"][0]["g1"]]
    names = ["encoder.weight", "decoder.2.weight", "transformer_encoder.layers.0.self_attn.in_proj_weight",
             f"transformer_encoder.layers.{case['L'] - 1}.linear2.weight", "transformer_encoder.layers.0.norm1.weight"]
    for t in leaves:
        t.requires_grad_(True)
    logits = O.transformer_forward_ref(P, x.double(), y.double(), case["sep"], case["H"])
    assert (logits - gold["logits"].double()).abs().max().item() <= 2e-4 * gold["logits"].abs().max().item()
    borders = case_borders(case).double()
    nll = O.bar_nll_ref(logits.reshape(-1, case["n_out"]), y[case["sep"]:].flatten().double(), borders, full_support=True)
    loss = nll.mean()
    assert abs(loss.item() - gold["loss"].item()) <= 1e-4 * abs(gold["loss"].item())
    assert (nll.view_as(gold["losses"]) - gold["losses"].double()).abs().max().item() <= 2e-4 * gold["losses"].abs().max().item()
    loss.backward()
    for t, n in zip(leaves, names):
        s, a, nrm = gold["grad_checksum"][n]
        assert abs(t.grad.norm().item() - nrm) <= 2e-3 * nrm + 1e-7, n
        ref_head = gold["grad_samples"][n].double()
        assert (t.grad.flatten()[:16] - ref_head).abs().max().item() <= 2e-3 * (ref_head.abs().max().item() + 1e-6), n


def test_oracle_bar_distribution_matches_reference():
    gold = torch.load(os.path.join(GOLD, "bar.pt"))
    for n_bars in (1, 7, 100, 1000):
        e = gold[n_bars]
        assert torch.equal(O.bucket_idx_ref(e["y"], e["borders"]), e["idx"]), n_bars
        nll = O.bar_nll_ref(e["logits"], e["y"], e["borders"])
        assert (nll - e["nll"]).abs().max().item() <= 1e-5 * (e["nll"].abs().max().item() + 1)
        assert (O.bar_mean_ref(e["logits"], e["borders"]) - e["mean"]).abs().max().item() <= 1e-5
        if n_bars > 1:
            nf = O.bar_nll_ref(e["logits"], e["y_full"], e["borders"], full_support=True)
            assert (nf - e["nll_full"]).abs().max().item() <= 1e-5 * (e["nll_full"].abs().max().item() + 1)
            mf = O.bar_mean_ref(e["logits"], e["borders"], full_support=True)
            assert (mf - e["mean_full"]).abs().max().item() <= 1e-5


# ---- BASELINE.json configurations at model shape (goldens from the unmodified reference, oracle/make_golden.py CONFIG_CASES)
from oracle.make_golden import CONFIG_CASES, case_targets  # noqa: E402


@pytest.mark.parametrize("name", list(CONFIG_CASES))
def test_oracle_matches_reference_at_baseline_config_shapes(name):
    gold = torch.load(os.path.join(GOLD, f"model_{name}.pt"))
    case = gold["case"]
    model = _ref_like_model(case)
    cs = checksum(model.state_dict())
    for k, (s, a) in gold["weights_checksum"].items():
        assert abs(cs[k][0] - s) <= 1e-9 * (abs(s) + 1) and abs(cs[k][1] - a) <= 1e-9 * (a + 1), f"weights differ: {k}"
    x, y = case_inputs(case)
    P = O.params_from_state_dict(model.state_dict(), case["L"], torch.float64)
    with torch.no_grad():
        logits = O.transformer_forward_ref(P, x.double(), y.double(), case["sep"], case["H"])
        assert (logits - gold["logits"].double()).abs().max().item() <= 3e-4 * gold["logits"].abs().max().item()
        t = case_targets(case, y).flatten().double()
        if case["head"] == "bar":
            nll = O.bar_nll_ref(logits.reshape(-1, case["n_out"]), t, case_borders(case).double(), full_support=True)
        else:
            z = logits.flatten()
            nll = torch.clamp(z, min=0) - z * t + torch.log1p(torch.exp(-z.abs()))     # BCE-with-logits, stable form
        assert abs(nll.mean().item() - gold["loss"].item()) <= 1e-4 * abs(gold["loss"].item())
