"""Boundary proof (container-only: needs the read-only reference tree at /root/reference).

  (1) The reference's OWN training-step code -- ``backprop_scheduler`` in 'base' mode
      (pase/models/WorkerScheduler/worker_scheduler.py:43-75) called exactly as
      trainer.py:229-244 calls it, with one torch optimizer per worker + one for the
      frontend (trainer.py:86-143) -- drives the native ``pase_b200.pase.pase`` model for two
      optimisation steps; losses are compared step by step with the unmodified reference model
      driven by the same code on the same weights and batch.  (CUDA kernels replaced by their
      torch specs, tests/emul_ops.py: this box has no GPU; the GPU tests cover the kernels.)
  (2) Checkpoints: files written by the reference ``Saver`` / ``Model.save`` load into the
      native modules (``load_pretrained``, ``Model.load``) and vice versa, index files and
      rotation behave identically (pase/models/modules.py:151-373).
"""
import copy
import json
import os
import random

import pytest
import torch

import ref_harness as RH

pytestmark = pytest.mark.skipif(not RH.reference_available(),
                                reason="reference tree not mounted (GPU box)")

import emul_ops  # noqa: E402
from helpers import load_golden, resolve_cfg, fill_state_dict, seeded_randn  # noqa: E402
import pase_b200.ops as ops  # noqa: E402
from pase_b200 import encoder as enc, wf_builder  # noqa: E402
from pase_b200.frontend import WaveFe  # noqa: E402
from pase_b200.pase import pase as native_pase  # noqa: E402
from pase_b200.utils import parse_workers  # noqa: E402
from pase_b200 import modules as NM  # noqa: E402


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(ops, "call", emul_ops.call)

    def encode_cpu(self, x):
        self._sinc_consts(x.device)
        plan = self._plan(x.shape[0], x.shape[2], x.device)
        named = list(self.named_parameters())
        names = tuple(n for n, _ in named)
        return enc._EncoderFn.apply(x.contiguous().float(), self, plan, self.training, names,
                                    *[p for _, p in named])
    monkeypatch.setattr(WaveFe, "encode", encode_cpu)
    yield


def _optimizers(model, fe_lr=5e-4, min_lr=4e-4):
    """trainer.py:86-143: Adam per worker + Adam for the frontend."""
    fe = torch.optim.Adam(model.frontend.parameters(), lr=fe_lr)
    cls = {w.name: torch.optim.Adam(w.parameters(), lr=min_lr)
           for w in model.classification_workers}
    regr = {w.name: torch.optim.Adam(w.parameters(), lr=min_lr)
            for w in model.regression_workers}
    return fe, cls, regr


def test_reference_scheduler_drives_native_model(emulated):
    RH.import_reference()
    from pase.models.WorkerScheduler.worker_scheduler import backprop_scheduler
    gold, meta = load_golden("pase_mini_workers_1600")
    fe_cfg, wcfg = resolve_cfg(meta["fe_cfg"]), meta["workers"]
    B, T, Tq, seed = meta["B"], meta["T"], meta["Tq"], meta["seed"]
    batch = {k: seeded_randn((B, 1, T), seed + 10 + i, 0.5)
             for i, k in enumerate(["chunk", "chunk_ctxt", "chunk_rand", "cchunk"])}
    for i, w in enumerate(wcfg["regr"]):
        if w["name"] != "cchunk":
            batch[w["name"]] = seeded_randn((B, w["num_outputs"], Tq), seed + 100 + i)

    ref = RH.build_ref_pase(fe_cfg, copy.deepcopy(wcfg)).train()
    nat = native_pase(frontend_cfg=fe_cfg, minions_cfg=parse_workers(copy.deepcopy(wcfg))).train()
    assert list(ref.state_dict().keys()) == list(nat.state_dict().keys())
    sd = fill_state_dict(ref.state_dict(), seed)
    ref.load_state_dict(sd)
    nat.load_state_dict(sd)

    trace = {}
    for tag, model in (("ref", ref), ("nat", nat)):
        sched = backprop_scheduler(model, mode="base")           # the reference's class
        fe_opt, cls_opt, regr_opt = _optimizers(model)
        steps = []
        for it in range(2):
            random.seed(seed + it)                               # SPC-style host RNG, if any
            # trainer.py:229: h, chunk, preds, labels = self.model.forward(batch, alpha, device)
            h, chunk, preds, labels = model.forward(dict(batch), 1, "cpu")
            # trainer.py:232-244
            losses, alpha = sched(preds, labels, cls_opt, regr_opt, fe_opt, device="cpu")
            steps.append({k: float(v) for k, v in losses.items()})
        trace[tag] = steps
    for it in range(2):
        r, n = trace["ref"][it], trace["nat"][it]
        assert set(r) == set(n)
        for k in r:
            # step 0: same weights -> fp32 parity; step 1: after one Adam update of every
            # parameter through each implementation's own gradients
            tol = 2e-4 if it == 0 else 2e-3
            assert abs(r[k] - n[k]) <= tol * max(abs(r[k]), 1e-3), (it, k, r[k], n[k])
    # the update really happened (losses moved) and stayed in lock-step
    assert abs(trace["ref"][1]["total"] - trace["ref"][0]["total"]) > 1e-4
    for (k, a), (_, b) in zip(ref.state_dict().items(), nat.state_dict().items()):
        if not a.is_floating_point():
            continue
        if k.startswith("frontend.") and (k.endswith("conv.bias") or k.endswith("W.bias")):
            # biases followed by train-mode BatchNorm: the gradient is analytically zero, both
            # implementations feed Adam rounding noise, and Adam turns ANY gradient into a
            # +-lr step -> the two may differ by up to 2 * lr per step, no more
            assert float((a - b).abs().max()) <= 2 * 5e-4 * 2 + 1e-6, k
            continue
        assert float((a - b).abs().max()) <= 2e-3 * max(float(a.abs().max()), 1e-3), k


def test_saver_checkpoints_are_interchangeable(tmp_path):
    ref_pkg = RH.import_reference()
    from pase.models.modules import Saver as RefSaver
    ref_fe = RH.build_ref_frontend("cfg/frontend/PASE+.cfg")
    nat_fe = wf_builder(resolve_cfg("cfg/frontend/PASE+.cfg"))
    assert list(ref_fe.state_dict().keys()) == list(nat_fe.state_dict().keys())
    sd = fill_state_dict(ref_fe.state_dict(), 7)
    ref_fe.load_state_dict(sd)
    ref_opt = torch.optim.Adam(ref_fe.parameters(), lr=1e-3)
    nat_opt = torch.optim.Adam(nat_fe.parameters(), lr=1e-3)

    # --- reference writes, native reads (Saver directory + bare ckpt file) -------------
    d1 = str(tmp_path / "ref_written")
    rs = RefSaver(ref_fe, d1, max_ckpts=2, optimizer=ref_opt, prefix="PASE-")
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        for step in (10, 20, 30, 40):
            rs.save("FE_e%d" % (step // 10), step)
    ns = NM.Saver(nat_fe, d1, max_ckpts=2, optimizer=nat_opt, prefix="PASE-")
    assert ns.read_latest_checkpoint() == rs.read_latest_checkpoint()
    assert ns.load_weights() is True
    for k, v in ref_fe.state_dict().items():
        assert torch.equal(v, nat_fe.state_dict()[k]), k
    assert ns.load_ckpt_step(ns.read_latest_checkpoint()) == 40
    # README.md:28-33 usage: load_pretrained(path, load_last=True)
    nat2 = wf_builder(resolve_cfg("cfg/frontend/PASE+.cfg"))
    ckpt = os.path.join(d1, "weights_" + rs.read_latest_checkpoint())
    nat2.load_pretrained(ckpt, load_last=True, verbose=False)
    for k, v in ref_fe.state_dict().items():
        assert torch.equal(v, nat2.state_dict()[k]), k
    # load_last=False drops the last two keys -> the reference raises on a key-count mismatch
    with pytest.raises(ValueError, match="LOADING DIFFERENT NUM OF KEYS"):
        wf_builder(resolve_cfg("cfg/frontend/PASE+.cfg")).load_pretrained(ckpt, load_last=False,
                                                                          verbose=False)

    # --- native writes the same sequence: identical index file and rotation -------------
    d2 = str(tmp_path / "nat_written")
    ns2 = NM.Saver(nat_fe, d2, max_ckpts=2, optimizer=nat_opt, prefix="PASE-")
    for step in (10, 20, 30, 40):
        ns2.save("FE_e%d" % (step // 10), step)
    i1 = json.load(open(os.path.join(d1, "PASE-checkpoints")))
    i2 = json.load(open(os.path.join(d2, "PASE-checkpoints")))
    assert i1 == i2
    assert sorted(os.listdir(d1)) == sorted(os.listdir(d2))
    # --- native writes, reference reads ---------------------------------------------------
    ref2 = RH.build_ref_frontend("cfg/frontend/PASE+.cfg")
    rs2 = RefSaver(ref2, d2, max_ckpts=2, optimizer=torch.optim.Adam(ref2.parameters(), lr=1e-3),
                   prefix="PASE-")
    with contextlib.redirect_stdout(io.StringIO()):
        assert rs2.load_weights() is True
    for k, v in nat_fe.state_dict().items():
        assert torch.equal(v, ref2.state_dict()[k]), k
    # Model.save / Model.load (modules.py:311-340) through the model's own saver
    d3 = str(tmp_path / "model_api")
    with contextlib.redirect_stdout(io.StringIO()):
        ref_fe.save(d3, 5)
    nat3 = wf_builder(resolve_cfg("cfg/frontend/PASE+.cfg"))
    nat3.name = ref_fe.name
    nat3.load(d3)
    for k, v in ref_fe.state_dict().items():
        assert torch.equal(v, nat3.state_dict()[k]), k
    assert nat3.get_total_params() == ref_fe.get_total_params() == 7832896
