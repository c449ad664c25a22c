"""cfg/workers/*.cfg -> minion kwargs with loss objects (pase/utils.py:53-90) and the
no-op-in-forward ScaleGrad (pase/utils.py:213-225)."""
import copy
import json

from .losses import ContextualizedLoss


def parse_workers(cfg, pop_transform=True):
    """dict (or path) in the cfg/workers JSON layout -> same dict with each worker's
    ``loss`` string replaced by a ContextualizedLoss (r taken from the worker's ``r``).
    ``transform`` keys describe the CPU data pipeline and are dropped (train.py:64)."""
    if isinstance(cfg, str):
        with open(cfg, "r") as f:
            cfg = json.load(f)
    cfg = copy.deepcopy(cfg)
    for kind, workers in cfg.items():
        for w in workers:
            if isinstance(w.get("loss"), str):
                if w["loss"] in ("LSGAN", "GAN"):
                    raise NotImplementedError("adversarial worker losses are out of scope")
                w["loss"] = ContextualizedLoss(w["loss"], r=w.get("r", None))
            if pop_transform:
                w.pop("transform", None)
    return cfg


def worker_parser(cfg_fname, batch_acum=1, device="cpu", do_losses=True, frontend=None):
    """Reference signature (utils.py:53).  Keeps ``transform`` like the reference does;
    the caller (train.py:64) pops it."""
    return parse_workers(cfg_fname, pop_transform=False)
