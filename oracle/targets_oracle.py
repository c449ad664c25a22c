"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's LPS regression target
(pase/transforms.py:439-487) and ZNorm (transforms.py:183-202); never imported by the product.

Pinning: the STFT is ``torch.stft`` itself, called as the reference calls it (n_fft, hop, win,
default rectangular window, centre reflect padding; the modern ``return_complex`` form of the
same transform).  ``librosa`` is not installed here (``import pase.transforms`` fails on its
third-party imports), so ``librosa.feature.delta`` is restated as the call it
documents -- ``scipy.signal.savgol_filter(data, width=9, deriv=order, polyorder=order, axis=-1,
mode='interp')`` -- with scipy's own implementation: parity for the delta rows is pinned to
scipy, UNPINNED against librosa itself.
"""
import numpy as np
import torch
from scipy.signal import savgol_filter


def lps(wav, n_fft=2048, hop=160, win=400, der_order=2):
    """wav: 1-D tensor (T,) -> ((1+der_order)*(n_fft/2+1), T // hop) float32 tensor."""
    wav = wav.detach().cpu().float()
    max_frames = wav.shape[0] // hop
    X = torch.stft(wav, n_fft, hop, win, return_complex=True)      # window=None: rectangular
    X = torch.view_as_real(X)
    X = torch.norm(X, 2, dim=2)[:, :max_frames]
    X = 10 * torch.log10(X ** 2 + 10e-20)
    if der_order > 0:
        deltas = [X.numpy()]
        for n in range(1, der_order + 1):
            deltas.append(savgol_filter(X.numpy(), 9, deriv=n, polyorder=n, axis=-1, mode="interp"))
        X = torch.from_numpy(np.concatenate(deltas)).float()
    return X


def znorm(x, mean, std):
    return (x - mean.reshape(-1, 1)) / std.reshape(-1, 1)
