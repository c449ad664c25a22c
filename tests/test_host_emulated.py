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
