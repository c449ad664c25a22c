"""GPU debug: NT / TN GEMM accuracy against float64 over a sweep of K (reduction) sizes."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import torch
from pase_b200 import functional as Fn

def rel(a, b):
    return float((a.double() - b).norm() / b.norm())

for prec in ("3xtf32", "3xf16", "tf32"):
    Fn.set_precision(prec)
    out = []
    for M in (800, 400):
        for K in range(64, 1088, 64):
            N = 256
            g = torch.Generator(device="cuda").manual_seed(K)
            A = torch.randn(M, K, device="cuda", generator=g)
            B = torch.randn(N, K, device="cuda", generator=g) * 0.1
            C = torch.empty(M, N, device="cuda")
            Fn.gemm_nt(A.reshape(-1), K, M * K, B.reshape(-1), K, N * K, C.reshape(-1), N, M, N, K, None,
                       a_kind="grad")
            e = rel(C, A.double() @ B.double().t())
            # TN: C2[i,j] = sum_r A[r,i] Bm[r,j]; I = K (columns of A), J = 256
            Bm = torch.randn(M, N, device="cuda", generator=g)
            C2 = torch.empty(K, N, device="cuda")
            Fn.gemm_tn(A.reshape(-1), K, M * K, Bm.reshape(-1), N, M * N, C2.reshape(-1), N, K, N, M)
            e2 = rel(C2, A.double().t() @ Bm.double())
            out.append((M, K, e, e2))
    lim = {"3xtf32": 2e-5, "3xf16": 5e-6, "tf32": 1e-2}[prec]
    print(prec, "worst nt %.1e tn %.1e" % (max(o[2] for o in out), max(o[3] for o in out)))
    for o in out:
        if o[2] > lim or o[3] > lim:
            print("   M=%d K=%d nt %.1e tn %.1e" % o)
