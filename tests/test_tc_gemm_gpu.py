"""tcgen05 / TMA GEMM kernels against the exact fp32 spec.  mode 1 (3xTF32) must be
fp32-equivalent; mode 0 (single TF32) is checked at TF32 tolerance."""
import pytest
import torch

import emul_ops
from pase_b200 import _lib
from test_kernels_gpu import R

pytestmark = pytest.mark.gpu


def _split(x, weights=False):
    """(hi, lo) as the product builds them: activations pass the raw array as 'hi' (the
    tensor core truncates it) with lo = rn(x - trunc(x)); weights get explicit rn hi/lo."""
    lo = torch.zeros_like(x)
    if weights:
        hi = torch.zeros_like(x)
        emul_ops.pase_split_tf32(x, hi, lo, x.numel())
        return hi, lo
    emul_ops.pase_split_tf32(x, None, lo, x.numel())
    return x.clone(), lo


def test_split_kernel_bit_exact():
    x = R(100003, seed=1) * 37.0
    hi, lo = _split(x, weights=True)
    dh, dl = torch.zeros_like(x).cuda(), torch.zeros_like(x).cuda()
    _lib.call("pase_split_tf32", x.cuda(), dh, dl, x.numel())
    assert torch.equal(dh.cpu(), hi) and torch.equal(dl.cpu(), lo)
    assert float((x - hi - lo).abs().max()) <= float(x.abs().max()) * 2.0 ** -22
    _, lo_a = _split(x)
    dl2 = torch.zeros_like(x).cuda()
    _lib.call("pase_split_tf32", x.cuda(), None, dl2, x.numel())       # activation form
    assert torch.equal(dl2.cpu(), lo_a)
    xt = emul_ops._tf32_trunc(x)
    assert float((x - xt - lo_a).abs().max()) <= float(x.abs().max()) * 2.0 ** -21


NT_CASES = [
    # M, N, K, R, rows_in, t_valid, rows_out, fold, bias, stats, acc
    (256, 128, 128, 128, 256, 256, 256, 1, True, True, 0),        # one full tile, plain
    (1000, 64, 1280, 640, 103, 100, 100, 1, True, True, 0),       # block-1 like, k spans 2 rows
    (700, 128, 704, 128, 70, 64, 64, 1, True, True, 0),           # stride-2 k=11 shape
    (300, 256, 1408, 128, 300, 300, 300, 1, False, False, 0),     # BN=256
    (520, 2048, 288, 32, 130, 4100, 129, 32, False, True, 0),     # folded sinc (fold 32)
    (333, 512, 5632, 1024, 111, 100, 100, 1, True, False, 1),     # long K, accumulate
    (260, 1920, 256, 256, 260, 260, 260, 1, False, False, 0),     # dcat shape, wide N
    (300, 128, 200, 64, 300, 300, 300, 1, True, True, 0),         # K % 32 != 0: per-k-block kernel
    (4000, 256, 2816, 256, 1000, 990, 990, 1, True, True, 0),     # many tiles / persistent loop
]


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("M,N,K,Rr,rows_in,t_valid,rows_out,fold,bias,stats,acc", NT_CASES)
def test_tc_gemm_nt(M, N, K, Rr, rows_in, t_valid, rows_out, fold, bias, stats, acc, mode):
    a_rows = M + (K + Rr - 1) // Rr + 2
    A = R(a_rows * Rr, seed=11)
    B = R(N * K, seed=12, scale=0.1)
    groups = (M + rows_in - 1) // rows_in
    C = R(groups * rows_out * N + 8, seed=13)
    bs = R(N, seed=14) if bias else None
    cs = torch.zeros(N, dtype=torch.float64) if stats else None
    cq = torch.zeros(N, dtype=torch.float64) if stats else None
    if mode == 1:
        (Ah, Al), (Bh, Bl) = _split(A), _split(B, weights=True)
    else:
        Ah, Al, Bh, Bl = A, None, B, None
    args = [Ah, Al, a_rows, Rr, Bh, Bl, K, C, N, M, N, K, 0.5, bs, rows_in, t_valid, rows_out,
            fold, cs, cq, acc, mode]
    cpu = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    dev = [a.cuda() if isinstance(a, torch.Tensor) else a for a in args]
    emul_ops.call("pase_tc_gemm_nt", *cpu)
    _lib.call("pase_tc_gemm_nt", *dev)
    torch.cuda.synchronize()
    scale = float((A.abs().mean() * B.abs().mean() * K ** 0.5))
    tol = (2e-6 if mode == 1 else 2e-5) * max(scale, 1e-3) * 8
    out_c, out_d = cpu[7], dev[7].cpu()
    err = float((out_c - out_d).abs().max())
    assert err <= tol, "C max err %.3e > %.3e (mode %d)" % (err, tol, mode)
    if stats:
        for i in (18, 19):
            c, d = cpu[i].float(), dev[i].cpu().float()
            e = float((c - d).abs().max())
            lim = (2e-5 if mode == 1 else 2e-4) * max(float(c.abs().max()), 1.0)
            assert e <= lim, "stats arg %d err %.3e > %.3e" % (i, e, lim)


TN_CASES = [
    # I, J, groups, rpg, lda, pitchA, offA, R, pitchB, acc
    (64, 1280, 3, 100, 64, 103, 1, 640, 102, 0),
    (128, 704, 2, 200, 128, 210, 5, 128, 206, 0),
    (2048, 288, 2, 131, 2048, 131, 0, 32, 140, 0),        # folded sinc wgrad
    (512, 5632, 4, 25, 512, 35, 5, 1024, 30, 0),
    (256, 1920, 1, 300, 256, 300, 0, 1920, 300, 1),
    (100, 512, 1, 77, 100, 77, 0, 512, 77, 0),            # I tail (emb 100)
]


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("I,J,groups,rpg,lda,pitchA,offA,Rr,pitchB,acc", TN_CASES)
def test_tc_gemm_tn(I, J, groups, rpg, lda, pitchA, offA, Rr, pitchB, acc, mode):
    A = R(groups * pitchA * lda + I + 64, seed=21)
    b_rows = groups * pitchB + (J + Rr - 1) // Rr + 2
    B = R(b_rows * Rr, seed=22)
    C = R(I * J, seed=23)
    if mode == 1:
        (Ah, Al), (Bh, Bl) = _split(A), _split(B)
    else:
        Ah, Al, Bh, Bl = A, None, B, None
    args = [Ah, Al, lda, pitchA, offA, Bh, Bl, Rr, pitchB, b_rows, C, J, I, J, groups, rpg, 0.25,
            acc, mode]
    cpu = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    dev = [a.cuda() if isinstance(a, torch.Tensor) else a for a in args]
    emul_ops.call("pase_tc_gemm_tn", *cpu)
    _lib.call("pase_tc_gemm_tn", *dev)
    torch.cuda.synchronize()
    out_c, out_d = cpu[10], dev[10].cpu()
    scale = float(out_c.abs().max())
    tol = (4e-6 if mode == 1 else 4e-5) * max(scale, 1.0)
    err = float((out_c - out_d).abs().max())
    assert err <= tol, "C max err %.3e > %.3e (mode %d)" % (err, tol, mode)
