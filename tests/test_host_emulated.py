"""Host-side orchestration of pase_b200 (geometry, buffer plan, padding / folding index
math, gradient routing) checked on CPU against the reference's golden vectors, with every
kernel launch replaced by its torch emulation (tests/emul_ops.py).  This does NOT exercise
the CUDA kernels -- the `-m gpu` tests do -- it proves the decomposition is exact."""
import pytest
import torch

import emul_ops
from helpers import load_golden, resolve_cfg, fill_state_dict, seeded_randn, assert_close, \
    check_grads
import pase_b200.ops as ops
from pase_b200 import encoder as enc
from pase_b200.frontend import WaveFe


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(ops, "call", emul_ops.call)
    yield


def run_encoder_cpu(model, x):
    """WaveFe.encode minus the CUDA-only guard (emulated kernels run on CPU tensors)."""
    model._sinc_consts(x.device)
    plan = model._plan(x.shape[0], x.shape[2], x.device)
    named = list(model.named_parameters())
    names = tuple(n for n, _ in named)
    tensors = [p for _, p in named]
    if torch.is_grad_enabled():
        return enc._EncoderFn.apply(x, model, plan, model.training, names, *tensors)
    return enc.encoder_forward(plan, model, x, dict(named), model.training, False)


CASES = ["enc_pase_eval_16000", "enc_pasep_eval_3200", "enc_pasep_train_3200",
         "enc_pasep_train_4001", "enc_pase_train_2400", "enc_mini_train_2000",
         "enc_mini_train_1763", "enc_mininornn_train_1600"]


@pytest.mark.parametrize("precision", ["fp32", "3xtf32"])
@pytest.mark.parametrize("name", CASES)
def test_encoder_host_logic(name, precision, emulated):
    """precision='3xtf32' exercises the tensor-core plan (sinc fold 32, hi/lo operand twins,
    folded-row addressing) with the emulated kernels."""
    gold, meta = load_golden(name)
    cfg = resolve_cfg(meta["cfg"])
    model = WaveFe(**cfg)
    model.precision = precision
    model.load_state_dict(fill_state_dict(model.state_dict(), meta["seed"]))
    model.train(meta["training"])
    x = seeded_randn((meta["N"], 1, meta["T"]), meta["seed"] + 1, 0.5)
    assert model.frame_counts(meta["T"]) == gold["frame_counts"].tolist()
    if not meta["training"]:
        with torch.no_grad():
            y, y_ntc = run_encoder_cpu(model, x)
        assert_close(y, gold["y"], 1e-4, 1e-5, name)
        return
    y, y_ntc = run_encoder_cpu(model, x)
    assert_close(y, gold["y"], 1e-4, 1e-5, name)
    N, E, Tq = y.shape
    assert_close(y_ntc.view(N, Tq, E).permute(0, 2, 1), gold["y"], 1e-4, 1e-5, name + " ntc")
    cot = seeded_randn(tuple(y.shape), meta["seed"] + 2)
    # half of the cotangent through each output layout
    loss = (y * (0.5 * cot)).sum() + \
        (y_ntc.view(N, Tq, E) * (0.5 * cot).permute(0, 2, 1)).sum()
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    n = check_grads(grads, gold, 2e-3, 2e-4)
    assert n > 10
    sd = model.state_dict()
    for key, val in gold.items():
        if key.startswith("stat/"):
            assert_close(sd[key[5:]].float(), val.float(), 1e-5, 1e-6, key)


WIDE = ["enc_pase_eval_16000", "enc_pasep_eval_3200", "enc_pasep_train_3200",
        "enc_pasep_train_4001", "enc_pase_train_2400"]       # channel counts multiples of 64


@pytest.mark.parametrize("precision", ["3xf16", "bf16"])
@pytest.mark.parametrize("name", WIDE)
def test_encoder_host_logic_16bit(name, precision, emulated):
    """The 16-bit tensor-core plans (sinc fold 64, bf16 / fp16-pair operand buffers written by
    the producing kernels, power-of-two scaled fp16 gradients) with the emulated kernels.
    3xf16 must meet the fp32 bar; bf16 is compared in relative L2 (bf16 storage)."""
    from helpers import rel_l2
    gold, meta = load_golden(name)
    cfg = resolve_cfg(meta["cfg"])
    model = WaveFe(**cfg)
    model.precision = precision
    model.load_state_dict(fill_state_dict(model.state_dict(), meta["seed"]))
    model.train(meta["training"])
    x = seeded_randn((meta["N"], 1, meta["T"]), meta["seed"] + 1, 0.5)
    exact = precision == "3xf16"
    if not meta["training"]:
        with torch.no_grad():
            y, y_ntc = run_encoder_cpu(model, x)
        if exact:
            assert_close(y, gold["y"], 1e-4, 1e-5, name)
        else:
            assert rel_l2(y, gold["y"]) < 2e-2, rel_l2(y, gold["y"])
        return
    y, y_ntc = run_encoder_cpu(model, x)
    if exact:
        assert_close(y, gold["y"], 1e-4, 1e-5, name)
    else:
        assert rel_l2(y, gold["y"]) < 2e-2, rel_l2(y, gold["y"])
    cot = seeded_randn(tuple(y.shape), meta["seed"] + 2)
    (y * cot).sum().backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    if exact:
        assert check_grads(grads, gold, 2e-3, 2e-4) > 10
    else:
        # bf16 (8-bit mantissas) perturbs pre-activations at the 1e-2 sigma level: ~1 % of
        # the PReLU(init 0) gates flip, so gradients agree with fp32 only to ~1e-1 in relative
        # L2 (measured 0.08..0.18 on these goldens; any bf16 implementation behaves so)
        zero_keys = [k for k in gold if k.startswith(("grad/", "gsample/", "gnorm/")) and
                     (k.endswith("conv.bias") or k.endswith("W.bias"))]
        sub = {k: v for k, v in gold.items() if k not in zero_keys}
        assert check_grads(grads, sub, 2e-3, 2e-4, l2_keys=("",), l2_tol=0.3) > 10
        # conv biases under train-mode BN have an analytically zero gradient; with bf16 du
        # the two BN-backward passes no longer cancel exactly (sum of rounding errors)
        for k, g in grads.items():
            if k.endswith("conv.bias"):
                w = grads[k.replace("conv.bias", "conv.weight")]
                assert float(g.abs().max()) <= 0.1 * float(w.norm()) + 1e-3, k
    plan = model._plan(x.shape[0], x.shape[2], x.device)
    if exact:     # every gradient operand was scaled into fp16's range by a power of two
        for gs in plan.gscale:
            s = float(gs[1])
            assert s > 0 and abs(s * float(gs[0]) - 1.0) < 1e-6 and \
                abs(torch.log2(torch.tensor(s)).item() % 1.0) < 1e-6


def test_weight_batch_table_is_cached_and_rebuilt(emulated, monkeypatch):
    """pase_conv_w_batch job tables: built once per plan, reused on the next step, rebuilt when
    a parameter's storage moves; gradients are views of per-call buffers (never the plan's)."""
    gold, meta = load_golden("enc_pasep_train_3200")
    cfg = resolve_cfg(meta["cfg"])
    model = WaveFe(**cfg)
    model.precision = "3xtf32"
    model.load_state_dict(fill_state_dict(model.state_dict(), meta["seed"]))
    model.train(True)
    x = seeded_randn((meta["N"], 1, meta["T"]), meta["seed"] + 1, 0.5)
    seen = []
    real = emul_ops.call

    def spy(name, *a):
        if name == "pase_conv_w_batch":
            seen.append((a[3], a[0].data_ptr(), a[1]))       # (op, table pointer, njobs)
        return real(name, *a)
    monkeypatch.setattr(ops, "call", spy)

    def step():
        model.zero_grad(set_to_none=True)
        y, _ = run_encoder_cpu(model, x)
        y.square().mean().backward()
        return y.detach().clone(), {k: p.grad for k, p in model.named_parameters()}

    y1, g1 = step()
    first = list(seen)
    # PASE+ has 7 non-sinc conv blocks (all of them feed an input gradient back to the
    # block before); their shapes fit the shared-memory tiled variants (ops 3..5)
    assert [(op, n) for op, _, n in first] == [(3, 7), (4, 7), (5, 7)]
    seen.clear()
    y2, g2 = step()
    assert [t for _, t, _ in seen] == [t for _, t, _ in first]       # same device tables
    plan = model._plan(x.shape[0], x.shape[2], x.device)
    plan_ptrs = {t.untyped_storage().data_ptr() for v in plan.__dict__.values()
                 for t in (v if isinstance(v, list) else [v]) if isinstance(t, torch.Tensor)}
    for k, g in g2.items():
        assert g.untyped_storage().data_ptr() not in plan_ptrs, k
        assert g1[k].untyped_storage().data_ptr() != g.untyped_storage().data_ptr() or \
            g1[k] is g, k
    # moving one weight to new storage invalidates the forward / dgrad tables only
    w = model.blocks[3].conv.weight
    w.data = w.data.clone()
    seen.clear()
    y3, g3 = step()
    new = {op: t for op, t, _ in seen}
    old = {op: t for op, t, _ in first}
    assert new[3] != old[3] and new[4] != old[4] and new[5] == old[5]
    # BatchNorm running statistics advance between the steps, the batch statistics do not:
    # train-mode outputs of the three steps agree
    assert_close(y2, y1, 1e-6, 1e-7, "step 2")
    assert_close(y3, y1, 1e-6, 1e-7, "step 3")


def test_gradient_sink_equals_autograd_path(emulated):
    """WaveFe.grad_sink (FlatAdam.bind_encoder): gradients written by the kernels straight into
    the flat buffer == the gradients autograd would have accumulated, and one FlatAdam step ==
    torch.optim.Adam on the autograd gradients."""
    from pase_b200.optim import FlatAdam
    gold, meta = load_golden("enc_pasep_train_3200")
    cfg = resolve_cfg(meta["cfg"])
    x = seeded_randn((meta["N"], 1, meta["T"]), meta["seed"] + 1, 0.5)
    cot = seeded_randn(tuple(gold["y"].shape), meta["seed"] + 2)
    models = []
    for _ in range(2):
        m = WaveFe(**cfg)
        m.precision = "3xf16"
        m.load_state_dict(fill_state_dict(m.state_dict(), meta["seed"]))
        models.append(m.train(True))
    ref, nat = models
    o_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    o_nat = FlatAdam(list(nat.parameters()), lr=1e-3).bind_encoder(nat)
    for it in range(2):
        o_ref.zero_grad(set_to_none=True)
        o_nat.zero_grad(set_to_none=True)          # must not detach the sunk gradient views
        for m in (ref, nat):
            y, _ = run_encoder_cpu(m, x)
            (y * cot).sum().backward()
        for (k, a), (_, b) in zip(ref.named_parameters(), nat.named_parameters()):
            assert b.grad.data_ptr() == o_nat._gviews[[id(p) for p in o_nat._plist].index(id(b))].data_ptr()
            if it == 0:          # same weights: bit-identical gradients
                assert torch.equal(a.grad, b.grad), k
            elif not k.endswith(("conv.bias", "W.bias")):      # (analytically zero: pure noise)
                # weights differ by the two Adam implementations' round-off
                assert torch.allclose(a.grad, b.grad, rtol=1e-3,
                                      atol=1e-4 * float(a.grad.abs().max())), k
        o_ref.step()
        o_nat.step()
        for (k, a), (_, b) in zip(ref.named_parameters(), nat.named_parameters()):
            if it == 0 or not k.endswith(("conv.bias", "W.bias")):
                assert torch.allclose(a, b, rtol=1e-5, atol=2e-6), k


def test_forward_only_plan_allocates_no_backward_buffers(emulated):
    """Inference at a new (N, T) builds the forward buffers only (feature extraction over
    per-utterance lengths); the first differentiated forward adds the backward set."""
    gold, meta = load_golden("enc_pasep_train_3200")
    cfg = resolve_cfg(meta["cfg"])
    model = WaveFe(**cfg)
    model.precision = "3xf16"
    model.load_state_dict(fill_state_dict(model.state_dict(), meta["seed"]))
    model.train(True)
    x = seeded_randn((meta["N"], 1, meta["T"]), meta["seed"] + 1, 0.5)
    with torch.no_grad():
        y0, _ = run_encoder_cpu(model, x)
    plan = model._plan(x.shape[0], x.shape[2], x.device)
    assert not plan.backward_ready and plan.dyz is None and plan.stats_b is None
    fwd_bytes = plan.nbytes()
    y, _ = run_encoder_cpu(model, x)
    assert plan.backward_ready and plan.nbytes() > 1.5 * fwd_bytes
    assert_close(y, gold["y"], 1e-4, 1e-5, "after the lazy allocation")
    y.square().mean().backward()
    assert all(p.grad is not None for p in model.parameters())
