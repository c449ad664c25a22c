"""GPU probe: UMMA K-major SWIZZLE_128B operand starting at an arbitrary row of a TMA-loaded
window (start address not 1024 B aligned), with and without the descriptor base_offset."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pase_b200 import _lib

torch.manual_seed(0)
wrows = 144
W = torch.randn(wrows, 32).cuda()
B = torch.randn(128, 32).cuda()
trunc = lambda t: (t.view(torch.int32) & -8192).view(torch.float32)
for use_bo in (0, 1):
    res = []
    for shift in (0, 1, 2, 3, 4, 7, 8, 9, 10, 15, 16):
        D = torch.zeros(128, 128).cuda()
        _lib.call("pase_tc_probe_rowshift", W.reshape(-1), B.reshape(-1), D.reshape(-1), wrows, shift, use_bo)
        torch.cuda.synchronize()
        ref = trunc(W[shift:shift + 128].contiguous()).double() @ trunc(B).double().t()
        err = float((D.double() - ref).abs().max())
        res.append((shift, "%.1e" % err))
    print("base_offset=%d:" % use_bo, res)
