"""Training-mode dropout (reference train.py:22 default 0.2; tabular checkpoints 0.5): the engine's counter-based masks at
the reference layer's four sites, checked against an fp64 oracle that consumes EXACTLY the masks the kernels use
(pfn_dropout_keep_mask), plus the statistics of the mask generator."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from transformerscandobayesianinference_b200 import _lib as L, bar_distribution, encoders, engine, transformer
from oracle import pfn_oracle as O


def test_keep_mask_statistics_and_elementwise_kernel(cuda_device):
    dev = cuda_device
    for p in (0.2, 0.5):
        thr = L.drop_threshold(p)
        m = torch.empty(4096, 512, device=dev, dtype=torch.uint8)
        L.dropout_keep_mask(m, 1234, thr)
        keep = m.float()
        assert abs(keep.mean().item() - (1 - thr / 256)) < 2e-3
        assert abs(keep.mean(0).std().item()) < 0.02 and abs(keep.mean(1).std().item()) < 0.05       # no row / column structure
        m2 = torch.empty_like(m)
        L.dropout_keep_mask(m2, 1235, thr)
        agree = (m == m2).float().mean().item()
        assert abs(agree - ((1 - thr / 256) ** 2 + (thr / 256) ** 2)) < 5e-3                            # seeds are independent
        for dt in (torch.float32, torch.bfloat16):
            x = torch.randn(4096, 512, device=dev).to(dt)
            r = torch.randn(4096, 512, device=dev).to(dt)
            out = torch.empty_like(x)
            L.dropout(x, out, 1234, thr, residual=r)
            want = (x.float() * keep * (256.0 / (256 - thr)) + r.float()).to(dt)
            assert torch.equal(out, want) or (out.float() - want.float()).abs().max().item() <= 1e-2 * want.float().abs().max().item()
            y = x.clone()
            L.dropout(y, y, 1234, thr)                                                                  # in place, no residual
            assert torch.allclose(y.float(), (x.float() * keep * (256.0 / (256 - thr))).to(dt).float(), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("p", [0.2, 0.5])
def test_training_step_with_dropout_matches_mask_consuming_oracle(cuda_device, p):
    dev = cuda_device
    T, B, F, E, H, nhid, NL, n_out, sep = 24, 3, 2, 64, 2, 128, 2, 12, 14
    torch.manual_seed(17)
    m = transformer.TransformerModel(encoders.Linear(F, E), n_out, E, H, nhid, NL, p, y_encoder=encoders.Linear(1, E)).to(dev)
    with torch.no_grad():
        for l in m.transformer_encoder.layers:
            l.linear2.weight.normal_(0, 0.05); l.self_attn.out_proj.weight.normal_(0, 0.05)
    m.precision = "fp32"
    m.train()
    x, y = torch.rand(T, B, F, device=dev), torch.randn(T, B, device=dev).clamp(-2.5, 2.5)
    crit = bar_distribution.FullSupportBarDistribution(torch.linspace(-3, 3, n_out + 1)).to(dev)
    torch.manual_seed(99)
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())        # what TransformerModel.forward will draw
    torch.manual_seed(99)
    logits = m((x, y), single_eval_pos=sep)
    loss = crit(logits.reshape(-1, n_out), y[sep:].flatten()).mean()
    loss.backward()
    # eval mode is deterministic and differs from the training-mode output
    m.eval()
    with torch.no_grad():
        assert not torch.allclose(m((x, y), single_eval_pos=sep), logits)
    m.train()

    thr = L.drop_threshold(p)
    scale = 256.0 / (256 - thr)

    def mask(rows, cols, li, site):
        out = torch.empty(rows, cols, device=dev, dtype=torch.uint8)
        L.dropout_keep_mask(out, engine.site_seed(seed, li, site), thr)
        return out.cpu().double()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    P = O.params_from_state_dict(sd, NL, torch.float64)
    leaves = {}
    for i, lp in enumerate(P["layers"]):
        pre = f"transformer_encoder.layers.{i}."
        leaves.update({pre + "self_attn.in_proj_weight": lp["in_w"], pre + "self_attn.in_proj_bias": lp["in_b"],
                       pre + "self_attn.out_proj.weight": lp["out_w"], pre + "self_attn.out_proj.bias": lp["out_b"],
                       pre + "linear1.weight": lp["w1"], pre + "linear1.bias": lp["b1"], pre + "linear2.weight": lp["w2"],
                       pre + "linear2.bias": lp["b2"], pre + "norm1.weight": lp["g1"], pre + "norm2.bias": lp["be2"]})
    leaves["encoder.weight"] = P["enc_w"]
    for t in leaves.values():
        t.requires_grad_(True)
    h = O.embed_ref(x.cpu().double(), y.cpu().double(), P["enc_w"], P["enc_b"], P["yenc_w"], P["yenc_b"], sep)
    for li, lp in enumerate(P["layers"]):
        keep = {"attn": mask(B * H * T, T, li, 0).reshape(B * H, T, T), "out": mask(T * B, E, li, 1),
                "gelu": mask(T * B, nhid, li, 2), "mlp": mask(T * B, E, li, 3)}
        h = O.encoder_layer_dropout_ref(h, lp, T, B, H, sep, keep, scale)
    ref_logits = O.gelu_erf(h[sep * B:] @ P["dec_w0"].T + P["dec_b0"]) @ P["dec_w2"].T + P["dec_b2"]
    ref = O.bar_nll_ref(ref_logits, y[sep:].flatten().cpu().double(), torch.linspace(-3, 3, n_out + 1).double(), True).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    named = dict(m.named_parameters())
    for k, t in leaves.items():
        got = named[k].grad.double().cpu()
        assert (got - t.grad).abs().max().item() <= 2e-3 * (t.grad.abs().max().item() + 1e-9), k


def test_default_train_arguments_run_with_dropout(cuda_device):
    """`train()` with the reference's default dropout = 0.2 (train.py:22) must run (it raised in round 1), bf16 engine."""
    from transformerscandobayesianinference_b200 import priors, train as train_mod, utils
    torch.manual_seed(0)
    ys = priors.fast_gp.get_batch(200, 20, 1, device="cuda:0")[1]
    crit = bar_distribution.FullSupportBarDistribution(bar_distribution.get_bucket_limits(50, ys=ys.cpu()))
    loss, pos, model = train_mod.train(priors.fast_gp.DataLoader, crit, encoders.Linear, emsize=256, nhid=256, nlayers=2, nhead=2,
                                       epochs=2, steps_per_epoch=4, batch_size=8, bptt=20, lr=1e-3, warmup_epochs=0,
                                       y_encoder_generator=encoders.Linear, extra_prior_kwargs_dict={"num_features": 1},
                                       single_eval_pos_gen=utils.get_weighted_single_eval_pos_sampler(20), verbose=False)
    assert loss == loss and model.dropout == 0.2
