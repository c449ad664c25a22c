"""``ContextualizedLoss`` (pase/losses.py:6-37) over the fused loss kernels.

Same call protocol as the reference: ``loss(pred, gtruth) -> scalar tensor`` with
``pred`` shaped (B, F*r, T) and ``gtruth`` (B, F, T).  The r-frame context unfold
is performed inside the kernel (no Python loop over time, no r-times label copy).
Predictions produced by pase_b200's heads carry their channel-last buffer
(``_pase_rows``) so no layout conversion is needed; foreign tensors are converted
with the NCT->rows kernel first.
"""
import torch
import torch.nn as nn

from . import functional as Fn


def _rows_of(pred):
    tag = getattr(pred, "_pase_rows", None)
    if tag is not None:
        return tag                      # (rows tensor (B*T, ld), n valid columns)
    rows = Fn.nct_to_rows(pred)
    return rows, pred.shape[1]


class ContextualizedLoss(object):
    def __init__(self, criterion, r=None):
        self.criterion, self.r = criterion, r
        if isinstance(criterion, str):
            self.kind = criterion
        else:
            self.kind = type(criterion).__name__
        if self.kind not in ("MSELoss", "L1Loss", "BCEWithLogitsLoss"):
            raise NotImplementedError("pase_b200 has no fused kernel for loss %r" % self.kind)

    def contextualize_r(self, tensor):
        """Materialised (B,F*r,T) target -- API parity / debugging only; the loss kernels
        never build it."""
        if self.r is None:
            return tensor
        B, F, T = tensor.shape
        pad = torch.nn.functional.pad(tensor, (self.r // 2, self.r // 2))
        return pad.unfold(2, self.r, 1).permute(0, 1, 3, 2).reshape(B, F * self.r, T)

    def __call__(self, pred, gtruth):
        fused = getattr(pred, "_pase_fused_loss", None)
        if fused is not None:             # computed inside the head's output GEMM
            return fused
        if self.kind == "MSELoss":
            rows, ncols = _rows_of(pred)
            r = 1 if self.r is None else int(self.r)
            assert r % 2 == 1, "contextualised loss needs an odd r"
            F = gtruth.shape[1]
            assert ncols == F * r, "prediction has %d channels, label needs %d" % (ncols, F * r)
            return Fn.ctx_mse_rows(rows, gtruth.to(rows.device), F, r)
        if self.kind == "L1Loss":
            assert self.r is None, "L1 with context is not used by the reference cfgs"
            rows, ncols = _rows_of(pred)
            if ncols == 1:                       # (B,1,T) waveform == (B*T,1) rows
                return Fn.l1_loss(rows[:, 0] if rows.shape[1] != 1 else rows.reshape(-1),
                                  gtruth.to(rows.device).reshape(-1))
            return Fn.l1_loss(pred.contiguous(), gtruth.to(pred.device))
        # BCEWithLogitsLoss against the [ones; zeros] pair labels of LIM / GIM / SPC
        n_pos = getattr(gtruth, "_pase_pairs", None)
        if n_pos is None:
            half = gtruth.shape[0] // 2
            if not (bool((gtruth[:half] == 1).all()) and bool((gtruth[half:] == 0).all())):
                raise NotImplementedError("BCE kernel supports the [ones; zeros] pair labels only")
            n_pos = gtruth[:half].numel()
        rows, ncols = _rows_of(pred)
        assert ncols == 1
        logits = rows[:, 0] if rows.shape[1] != 1 else rows.reshape(-1)
        return Fn.bce_pairs(logits, int(n_pos))
