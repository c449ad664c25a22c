"""Host-side helpers mirroring the pieces of pase/models/modules.py that sit on
the encoder boundary: batch formatting, output selection and the checkpoint
API (``Model`` / ``Saver``).  File formats are the reference's (checkpoint
index JSON + ``weights_<prefix><name>-<step>.ckpt`` holding ``step``,
``state_dict`` and optionally ``optimizer``; modules.py:151-373) so existing
checkpoints and training directories interoperate.
"""
import json
import os

import torch
import torch.nn as nn

CHUNK_KEYS = ("chunk", "chunk_ctxt", "chunk_rand", "cchunk")


def format_frontend_chunk(batch, device=None):
    """dict of (B,1,T) chunks -> one (k*B,1,T) batch + k; tensor -> (tensor, 0).
    Same key handling as modules.py:16-31 (only ``chunk_rand`` decides whether the
    triplet is concatenated)."""
    if isinstance(batch, dict):
        if "chunk_rand" in batch:
            parts = [batch[k] for k in CHUNK_KEYS if k in batch]
            if device is not None:
                parts = [p.to(device, non_blocking=True) for p in parts]
            return torch.cat(parts, dim=0), len(parts)
        x = batch["chunk"]
        return (x.to(device) if device is not None else x), 1
    return batch, 0


def select_output(h, mode=None):
    """modules.py:62-74."""
    if mode is None:
        return h
    avg = h.mean(dim=2, keepdim=True)
    if mode == "avg_norm":
        return h - avg
    if mode == "avg_concat":
        return torch.cat([h, avg.expand_as(h)], dim=1)
    if mode == "avg_norm_concat":
        return torch.cat([h - avg, avg.expand_as(h)], dim=1)
    return h


def format_frontend_output(y, data_fmt, mode=None):
    """modules.py:33-43."""
    if data_fmt > 1:
        emb = torch.chunk(y, data_fmt, dim=0)
        return emb, emb[0]
    if data_fmt == 1:
        return y, y
    return select_output(y, mode)


class Saver(object):
    """Rolling checkpoint writer/reader, format-compatible with modules.py:151-301."""

    def __init__(self, model, save_path, max_ckpts=5, optimizer=None, prefix=""):
        self.model, self.save_path = model, save_path
        self.max_ckpts, self.optimizer, self.prefix = max_ckpts, optimizer, prefix
        self.ckpt_path = os.path.join(save_path, "%scheckpoints" % prefix)

    def _index(self):
        if os.path.exists(self.ckpt_path):
            with open(self.ckpt_path) as f:
                return json.load(f)
        return {"latest": [], "current": []}

    def save(self, model_name, step, best_val=False):
        os.makedirs(self.save_path, exist_ok=True)
        index = self._index()
        fname = "%s%s%s-%s.ckpt" % (self.prefix, "best_" if best_val else "", model_name, step)
        latest = index["latest"]
        if self.max_ckpts is not None and len(latest) > self.max_ckpts:
            stale = os.path.join(self.save_path, "weights_" + latest[0])
            if os.path.exists(stale):
                os.remove(stale)
            latest = latest[1:]
        latest.append(fname)
        index["latest"], index["current"] = latest, fname
        with open(self.ckpt_path, "w") as f:
            json.dump(index, f, indent=2)
        payload = {"step": step, "state_dict": self.model.state_dict()}
        if self.optimizer is not None:
            payload["optimizer"] = self.optimizer.state_dict()
        torch.save(payload, os.path.join(self.save_path, "weights_" + fname))

    def read_latest_checkpoint(self):
        if not os.path.exists(self.ckpt_path):
            print("[!] No checkpoint found in %s" % self.save_path)
            return None
        return self._index()["current"]

    def load_weights(self):
        cur = self.read_latest_checkpoint()
        if cur is None:
            return False
        st = torch.load(os.path.join(self.save_path, "weights_" + cur), map_location="cpu")
        if "state_dict" in st:
            self.model.load_state_dict(st["state_dict"])
            if self.optimizer is not None and "optimizer" in st:
                self.optimizer.load_state_dict(st["optimizer"])
        else:
            self.model.load_state_dict(st)
        return True

    def load_ckpt_step(self, curr_ckpt):
        return torch.load(os.path.join(self.save_path, "weights_" + curr_ckpt),
                          map_location="cpu")["step"]

    def load_pretrained_ckpt(self, ckpt_file, load_last=False, load_opt=True, verbose=True):
        """Key- and shape-filtered load; raises when the number of matching keys
        differs from the model's (modules.py:267-301)."""
        own = self.model.state_dict()
        st = torch.load(ckpt_file, map_location="cpu")
        src = st["state_dict"] if "state_dict" in st else st
        names = list(src.keys())
        allowed = set(names if load_last else names[:-2])
        picked = {k: v for k, v in src.items()
                  if k in allowed and k in own and tuple(v.shape) == tuple(own[k].shape)}
        if verbose:
            print("model keys: %d, matching checkpoint keys: %d" % (len(own), len(picked)))
        if len(picked) != len(own):
            raise ValueError("WARNING: LOADING DIFFERENT NUM OF KEYS")
        own.update(picked)
        self.model.load_state_dict(own)
        if self.optimizer is not None and "optimizer" in st and load_opt:
            self.optimizer.load_state_dict(st["optimizer"])


class Model(nn.Module):
    """Base class giving every network the reference's persistence / reporting
    API (modules.py:135-149, 304-373)."""

    def __init__(self, max_ckpts=5, name="BaseModel"):
        super().__init__()
        self.name, self.optim, self.max_ckpts = name, None, max_ckpts

    def _own_saver(self, save_path):
        if not hasattr(self, "saver"):
            self.saver = Saver(self, save_path, optimizer=self.optim,
                               prefix=self.name + "-", max_ckpts=self.max_ckpts)
        return self.saver

    def save(self, save_path, step, best_val=False, saver=None):
        (saver or self._own_saver(save_path)).save(self.name, step, best_val=best_val)

    def load(self, save_path):
        if os.path.isdir(save_path):
            self._own_saver(save_path).load_weights()
        else:
            self.load_pretrained(save_path)

    def load_pretrained(self, ckpt_path, load_last=False, verbose=True):
        Saver(self, ".", optimizer=self.optim).load_pretrained_ckpt(
            ckpt_path, load_last, verbose=verbose)

    def parameters(self, recurse=True):
        return (p for p in super().parameters(recurse) if p.requires_grad)

    def get_total_params(self):
        return sum(p.numel() for p in self.parameters())

    def describe_params(self):
        if hasattr(self, "blocks"):
            for b in self.blocks:
                n = sum(p.numel() for p in b.parameters())
                print("-" * 10)
                print(b)
                print("Num params: ", n)
        total = self.get_total_params()
        print("%s total params: %d" % (self.name, total))
        return total
