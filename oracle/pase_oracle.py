"""CPU oracle for the PASE/PASE+ hot path -- TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (plain torch CPU ops on a flat ``state_dict``) of
the reference algorithm for the path named in BASELINE.json: the WaveFe
encoder, the worker heads and the summed multi-task loss.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline -- never from the product package ``pase_b200``.

Pinned: ``tests/test_oracle_golden.py`` checks every function here against
golden vectors produced by the UNMODIFIED reference imported from
/root/reference (script: tests/golden/make_golden.py; harness:
oracle/ref_harness.py).  One part is "parity unpinned": the QRNN layer follows
the published salesforce/pytorch-qrnn algorithm (un-vendored, un-pinned
dependency, requirements.txt:16) as restated in SURVEY.md A.2; the reference
tree holds no test or vector for it.

Reference citations are /root/reference/<file>:<lines>.
"""
import math
import json
import torch
import torch.nn.functional as F

__all__ = [
    "DEFAULT_FE_CFG", "normalize_fe_cfg", "frame_counts", "sinc_filters",
    "encoder_forward", "select_output", "contextualize", "head_mlp",
    "head_decoder", "lim_inputs", "gim_inputs", "pase_forward", "total_loss",
]

# WaveFe.__init__ defaults, pase/models/frontend.py:120-143
DEFAULT_FE_CFG = dict(
    num_inputs=1, sincnet=True,
    kwidths=[251, 10, 5, 5, 5, 5, 5, 5], strides=[1, 10, 2, 1, 2, 1, 2, 2],
    dilations=[1, 1, 1, 1, 1, 1, 1, 1],
    fmaps=[64, 64, 128, 128, 256, 256, 512, 512],
    norm_type="bnorm", pad_mode="reflect", sr=16000, emb_dim=256,
    rnn_dim=None, activation=None, rnn_pool=False, rnn_layers=1,
    rnn_dropout=0, rnn_type="qrnn", vq_K=None, norm_out=False,
    tanh_out=False, resblocks=False, denseskips=False, densemerge="sum",
)


def normalize_fe_cfg(cfg):
    if isinstance(cfg, str):
        with open(cfg) as f:
            cfg = json.load(f)
    out = dict(DEFAULT_FE_CFG)
    out.update(cfg)
    if out["rnn_pool"] and out["rnn_dim"] is None:
        out["rnn_dim"] = out["emb_dim"]          # frontend.py:188-189
    return out


def _pads(k, stride, sinc):
    """Reflect pad (left, right).  Sinc: modules.py:922-928; FeBlock:
    modules.py:1058-1071 (dilation 1)."""
    if sinc:
        return (k // 2 - 1, k // 2) if stride > 1 else (k // 2, k // 2)
    if k <= 1:
        return (0, 0)
    if stride > 1 or k % 2 == 0:
        return (k // 2 - 1, k // 2)
    return (k // 2, k // 2)


def frame_counts(cfg, T):
    """Per-block output lengths (the 'frame indices bit-exact' contract)."""
    cfg = normalize_fe_cfg(cfg)
    out, L = [], T
    for i, (k, s) in enumerate(zip(cfg["kwidths"], cfg["strides"])):
        sinc = cfg["sincnet"] and i == 0
        if sinc and k % 2 == 0:
            k += 1                                # modules.py:835-836
        pl, pr = _pads(k, s, sinc)
        L = (L + pl + pr - k) // s + 1
        out.append(L)
    return out


def sinc_filters(low_hz_, band_hz_, k=251, sr=16000, min_low=50., min_band=50.):
    """(C,1),(C,1) -> (C,1,k) band-pass bank.  modules.py:868-918."""
    half = k // 2
    dev = low_hz_.device
    dt = low_hz_.dtype
    n_lin = torch.linspace(0, (k / 2) - 1, steps=half, device=dev, dtype=dt)
    window = 0.54 - 0.46 * torch.cos(2 * math.pi * n_lin / k)
    n_ = 2 * math.pi * torch.arange(-(k - 1) / 2.0, 0, device=dev, dtype=dt).view(1, -1) / sr
    low = min_low + low_hz_.abs()
    high = torch.clamp(low + min_band + band_hz_.abs(), min_low, sr / 2)
    band = (high - low)[:, 0]
    left = (torch.sin(high @ n_) - torch.sin(low @ n_)) / (n_ / 2) * window
    bp = torch.cat([left, 2 * band.view(-1, 1), left.flip(1)], 1)
    bp = bp / (2 * band[:, None])
    return bp.view(-1, 1, k)


def _bn(h, sd, prefix, training, affine=True, eps=1e-5, momentum=0.1,
        new_stats=None, lib_ops=False):
    """nn.BatchNorm1d over (N,C,T).  modules.py:79 / frontend.py:206-208.
    lib_ops: dispatch to F.batch_norm exactly as nn.BatchNorm1d does (the library call the
    reference issues; used by bench.py's library-baseline leg on CUDA -> cuDNN)."""
    w = sd[prefix + "weight"] if affine else None
    b = sd[prefix + "bias"] if affine else None
    if lib_ops:
        rm = sd[prefix + "running_mean"].clone()
        rv = sd[prefix + "running_var"].clone()
        y = F.batch_norm(h, rm, rv, w, b, training, momentum, eps)
        if training and new_stats is not None:
            new_stats[prefix + "running_mean"], new_stats[prefix + "running_var"] = rm, rv
            new_stats[prefix + "num_batches_tracked"] = sd[prefix + "num_batches_tracked"] + 1
        return y
    if training:
        mean = h.mean(dim=(0, 2))
        var = h.var(dim=(0, 2), unbiased=False)
        if new_stats is not None:
            n = h.shape[0] * h.shape[2]
            new_stats[prefix + "running_mean"] = \
                (1 - momentum) * sd[prefix + "running_mean"] + momentum * mean.detach()
            new_stats[prefix + "running_var"] = \
                (1 - momentum) * sd[prefix + "running_var"] + \
                momentum * var.detach() * n / max(n - 1, 1)
            new_stats[prefix + "num_batches_tracked"] = \
                sd[prefix + "num_batches_tracked"] + 1
    else:
        mean, var = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    y = (h - mean[None, :, None]) / torch.sqrt(var[None, :, None] + eps)
    if affine:
        y = y * w[None, :, None] + b[None, :, None]
    return y


def _prelu(h, a):
    return torch.where(h > 0, h, a[None, :, None] * h)


def _qrnn(h, sd, prefix):
    """(N,C,T) -> (N,H,T); torchqrnn.QRNN(window=2) as called at
    modules.py:52 / frontend.py:256-259 (published algorithm)."""
    X = h.permute(2, 0, 1)                                  # (T,N,C)
    Xm1 = torch.cat([torch.zeros_like(X[:1]), X[:-1]], 0)
    Y = F.linear(torch.cat([X, Xm1], 2), sd[prefix + "linear.weight"],
                 sd[prefix + "linear.bias"])
    Z, Fg, O = Y.chunk(3, dim=2)
    Z, Fg, O = torch.tanh(Z), torch.sigmoid(Fg), torch.sigmoid(O)
    c, cs = None, []
    for t in range(X.shape[0]):
        c = Fg[t] * Z[t] if c is None else Fg[t] * Z[t] + (1 - Fg[t]) * c
        cs.append(c)
    H = O * torch.stack(cs, 0)
    return H.permute(1, 2, 0)


def _pool_skip(skip, Tq):
    """fuse_skip, frontend.py:213-232 (densemerge='sum')."""
    d = skip.shape[2] // Tq
    if d > 1:
        skip = skip[:, :, :Tq * d]
        skip = skip.reshape(skip.shape[0], skip.shape[1], Tq, d).mean(3)
    return skip


def select_output(h, mode=None):
    """modules.py:62-74."""
    if mode == "avg_norm":
        return h - h.mean(2, keepdim=True)
    if mode == "avg_concat":
        return torch.cat([h, h.mean(2, keepdim=True).expand_as(h)], 1)
    if mode == "avg_norm_concat":
        g = h.mean(2, keepdim=True)
        return torch.cat([h - g, g.expand_as(h)], 1)
    return h


def encoder_forward(x, sd, cfg, training=True, new_stats=None,
                    return_blocks=False, lib_ops=False):
    """WaveFe.forward on a (N,1,T) tensor -> (N,emb,T').  frontend.py:234-279;
    FeBlock modules.py:1058-1077.  ``sd`` uses the reference's state_dict keys;
    ``new_stats`` (dict) receives updated BN running buffers in training."""
    cfg = normalize_fe_cfg(cfg)
    for key in ("resblocks", "tanh_out"):
        assert not cfg[key], "oracle covers the PASE.cfg / PASE+.cfg feature set"
    assert cfg["norm_type"] == "bnorm" and cfg["densemerge"] == "sum"
    assert cfg["vq_K"] in (None, 0) and cfg["pad_mode"] == "reflect"
    h = x
    skips, blocks = [], []
    nblk = len(cfg["kwidths"])
    for i, (k, s) in enumerate(zip(cfg["kwidths"], cfg["strides"])):
        p = "blocks.%d." % i
        sinc = cfg["sincnet"] and i == 0
        if sinc:
            if k % 2 == 0:
                k += 1
            w = sinc_filters(sd[p + "conv.low_hz_"], sd[p + "conv.band_hz_"],
                             k, cfg["sr"])
            b = None
        else:
            w, b = sd[p + "conv.weight"], sd[p + "conv.bias"]
        pl, pr = _pads(k, s, sinc)
        if pl + pr > 0:
            h = F.pad(h, (pl, pr), mode="reflect")
        h = F.conv1d(h, w, b, stride=s)
        h = _bn(h, sd, p + "norm.", training, new_stats=new_stats, lib_ops=lib_ops)
        h = F.prelu(h, sd[p + "act.weight"]) if lib_ops else _prelu(h, sd[p + "act.weight"])
        blocks.append(h)
        if cfg["denseskips"] and i + 1 < nblk:
            skips.append(F.conv1d(h, sd["denseskips.%d.weight" % i]))
    if cfg["rnn_pool"]:
        assert cfg["rnn_type"] == "qrnn" and cfg["rnn_layers"] == 1
        h = _qrnn(h, sd, "rnn.layers.0.")
    y = F.conv1d(h, sd["W.weight"], sd["W.bias"])
    for sk in skips:
        y = y + _pool_skip(sk, y.shape[2])
    if cfg["norm_out"]:
        y = _bn(y, sd, "norm_out.", training, affine=False, new_stats=new_stats,
                lib_ops=lib_ops)
    if return_blocks:
        return y, blocks
    return y


# ----------------------------------------------------------------- heads ---

def contextualize(label, r):
    """ContextualizedLoss.contextualize_r, losses.py:15-31: (B,F,T) ->
    (B,F*r,T), out[b, f*r+j, t] = zero-padded label[b, f, t+j-r//2]."""
    if r is None:
        return label
    B, Fd, T = label.shape
    pad = F.pad(label, (r // 2, r // 2))
    win = pad.unfold(2, r, 1)                 # (B,F,T,r)
    return win.permute(0, 1, 3, 2).reshape(B, Fd * r, T)


def head_mlp(x, sd, prefix, hidden_layers=1):
    """MLPMinion forward with context=1, dropout 0, skip False.
    minions.py:452-528, MLPBlock modules.py:527-556."""
    h = x
    for i in range(hidden_layers):
        p = "%sblocks.%d." % (prefix, i)
        h = _prelu(F.conv1d(h, sd[p + "W.weight"], sd[p + "W.bias"]),
                   sd[p + "act.weight"])
    return F.conv1d(h, sd[prefix + "W.weight"], sd[prefix + "W.bias"])


def head_decoder(x, sd, prefix, strides, kwidths, hidden_layers=1):
    """DecoderMinion forward.  minions.py:365-449; GDeconv1DBlock
    modules.py:558-589 (norm None, PReLU)."""
    h = x
    nb = len(strides)
    for i, (s, k) in enumerate(zip(strides, kwidths)):
        p = "%sblocks.%d." % (prefix, i)
        pad = max(0, (s - k) // -2)
        h = F.conv_transpose1d(h, sd[p + "deconv.weight"], sd[p + "deconv.bias"],
                               stride=s, padding=pad)
        if (s % 2 != 0 and k % 2 == 0) or (s % 2 == 0 and k % 2 != 0):
            h = h[:, :, :-1]
        h = _prelu(h, sd[p + "act.weight"])
    for j in range(hidden_layers):
        p = "%sblocks.%d." % (prefix, nb + j)
        h = _prelu(F.conv1d(h, sd[p + "W.weight"], sd[p + "W.bias"]),
                   sd[p + "act.weight"])
    return F.conv1d(h, sd[prefix + "W.weight"], sd[prefix + "W.bias"])


def _pairs(h, augment):
    """make_samples, cls_minions.py:29-43."""
    pos = torch.cat([h[0], h[1]], 1)
    neg = torch.cat([h[0], h[2]], 1)
    if augment:
        pos = torch.cat([pos, torch.cat([h[1], h[0]], 1)], 0)
        neg = torch.cat([neg, torch.cat([h[1], h[2]], 1)], 0)
    return torch.cat([pos, neg], 0)


def lim_inputs(h, augment=False):
    """LIM.forward input assembly, cls_minions.py:69-74."""
    return _pairs(h, augment)


def gim_inputs(h, augment=False):
    """GIM.forward input assembly, cls_minions.py:93-99."""
    return _pairs(h, augment).mean(2, keepdim=True)


def _pair_labels(y):
    """make_labels, cls_minions.py:47-51."""
    half = y.shape[0] // 2
    return torch.cat([torch.ones(half, 1, y.shape[2], device=y.device),
                      torch.zeros(half, 1, y.shape[2], device=y.device)], 0)


_CRIT = {
    "MSELoss": lambda p, t: ((p - t) ** 2).mean(),
    "L1Loss": lambda p, t: (p - t).abs().mean(),
    "BCEWithLogitsLoss": lambda p, t: F.binary_cross_entropy_with_logits(p, t),
}


def pase_forward(batch, sd, fe_cfg, workers_cfg, training=True, new_stats=None):
    """pase.forward (pase.py:310-356) for MLP / decoder regression workers and
    mi / cmi classification workers.  ``sd`` is the reference ``pase`` module's
    state_dict; ``workers_cfg`` the raw JSON dict of cfg/workers/*.cfg.
    Returns (h_tuple, chunk, preds, labels)."""
    keys = ["chunk", "chunk_ctxt", "chunk_rand"]      # cchunk popped, pase.py:314-317
    if "chunk_rand" in batch:
        xs = [batch[k] for k in keys if k in batch]
    else:
        xs = [batch["chunk"]]
    fsd = {k[len("frontend."):]: v for k, v in sd.items() if k.startswith("frontend.")}
    fstats = {} if new_stats is not None else None
    y = encoder_forward(torch.cat(xs, 0), fsd, fe_cfg, training, fstats)
    if new_stats is not None:
        new_stats.update({"frontend." + k: v for k, v in fstats.items()})
    h = torch.chunk(y, len(xs), 0)
    chunk = h[0]
    preds, labels = {}, {}
    for i, w in enumerate(workers_cfg.get("regr", [])):
        p = "regression_workers.%d." % i
        if w.get("type", "mlp") == "decoder":
            preds[w["name"]] = head_decoder(chunk, sd, p, w["strides"], w["kwidths"],
                                            w.get("hidden_layers", 2))
        else:
            preds[w["name"]] = head_mlp(chunk, sd, p, w.get("hidden_layers", 2))
        labels[w["name"]] = batch[w["name"]]
    for i, w in enumerate(workers_cfg.get("cls", [])):
        p = "classification_workers.%d.minion." % i
        aug = bool(w.get("augment", False))
        if w["name"] == "mi":
            xin = lim_inputs(h, aug)
        elif w["name"] == "cmi":
            xin = gim_inputs(h, aug)
        else:
            raise NotImplementedError(w["name"])
        y_ = head_mlp(xin, sd, p, w.get("hidden_layers", 2))
        preds[w["name"]] = y_
        labels[w["name"]] = _pair_labels(y_)
    return h, chunk, preds, labels


def total_loss(preds, labels, workers_cfg):
    """backprop_scheduler 'base' mode, worker_scheduler.py:43-62, with the
    losses wired by worker_parser (utils.py:53-68)."""
    tot, per = 0., {}
    for kind in ("cls", "regr"):
        for w in workers_cfg.get(kind, []):
            tgt = contextualize(labels[w["name"]], w.get("r", None))
            l = w.get("loss_weight", 1.) * _CRIT[w["loss"]](preds[w["name"]], tgt)
            per[w["name"]] = l
            tot = tot + l
    return tot, per
