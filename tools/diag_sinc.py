"""GPU diagnostic: per-channel difference of the raw sinc-layer output (and BN stats) between
the fp32-FFMA plan and the tensor-core plans, same weights / input."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch
from helpers import resolve_cfg, fill_state_dict, seeded_randn
from pase_b200.frontend import WaveFe

cfg = resolve_cfg("cfg/frontend/PASE+.cfg")
seed, N, T = 21, 3, 32000
x = seeded_randn((N, 1, T), seed + 1, 0.5).cuda()
outs = {}
for prec in ("fp32", "3xtf32", "tf32"):
    m = WaveFe(**cfg)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed))
    m.precision = prec
    m = m.cuda().train()
    with torch.no_grad():
        m(x)
    plan = list(m._plans.values())[0]
    g = plan.geoms[0]
    y0 = plan.y[0].view(N, g.rows_out * g.fold, 64)[:, :g.T_out].clone()
    outs[prec] = (y0, plan.bn[0].clone(), plan.Wt[0].clone(), g)
ref = outs["fp32"][0]
std = ref.std(dim=(0, 1))
for prec in ("3xtf32", "tf32"):
    d = (outs[prec][0] - ref).abs()
    per = d.amax(dim=(0, 1)) / std
    top = torch.topk(per, 5)
    print(prec, "y0 max|diff|/std per channel: top", [(int(i), float(v)) for v, i in zip(top.values, top.indices)],
          "median %.2e" % float(per.median()))
    print(prec, "bn mean diff max %.3e  invstd rel diff max %.3e" % (
        float((outs[prec][1][0] - outs["fp32"][1][0]).abs().max()),
        float(((outs[prec][1][1] - outs["fp32"][1][1]) / outs["fp32"][1][1]).abs().max())))
# filter check: polyphase operand row p=0 holds the filter itself
f4 = outs["fp32"][2].view(4, 64, -1)[0][:, :251]
f32 = outs["3xtf32"][2].view(32, 64, -1)[0][:, :251]
print("filter max diff fold4 vs fold32: %.3e" % float((f4 - f32).abs().max()))
print("channel 51 std %.4e, |y0| max %.3e" % (float(std[51]), float(ref[..., 51].abs().max())))
