"""GPU debug: one MLP regression head (fused / unfused) against a float64 torch reference."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")]
import torch
from pase_b200.minions import MLPMinion
from pase_b200.losses import ContextualizedLoss
from pase_b200 import functional as Fn
import pase_oracle as O

def rel(a, b):
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

Fn.set_precision(sys.argv[1] if len(sys.argv) > 1 else "3xf16")
for (F, r) in ((39, 7), (120, 7), (12, 7), (60, 7)):
    B, T, E = 4, 200, 256
    torch.manual_seed(F)
    m = MLPMinion(num_inputs=E, num_outputs=F, dropout=0, hidden_size=256, hidden_layers=1, r=r,
                  skip=False, loss=ContextualizedLoss("MSELoss", r)).cuda()
    x0 = torch.randn(B, E, T, device="cuda")
    lab = torch.randn(B, F, T, device="cuda")
    sd = {k: v.detach().double().cpu() for k, v in m.state_dict().items()}
    xr = x0.detach().double().cpu().requires_grad_(True)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    yr = O.head_mlp(xr, leaves, "", 1)
    lr = ((yr - O.contextualize(lab.double().cpu(), r)) ** 2).mean()
    lr.backward()
    for fused in (False, True):
        m.zero_grad()
        x = x0.clone().requires_grad_(True)
        y = m(x, label=lab if fused else None)
        loss = m.loss(y, lab)
        loss.backward()
        out = ["F=%d fused=%d loss rel %.1e" % (F, fused, abs(float(loss) - float(lr)) / float(lr)),
               "dx %.1e" % rel(x.grad, xr.grad)]
        for k, p in m.named_parameters():
            out.append("%s %.1e" % (k, rel(p.grad, leaves[k].grad)))
        print("  ".join(out))
