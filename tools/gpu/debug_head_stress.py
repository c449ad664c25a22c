"""GPU debug: repeat one regression head's fwd+bwd, look for run-to-run outliers."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")]
import torch
from pase_b200.minions import MLPMinion
from pase_b200.losses import ContextualizedLoss
from pase_b200 import functional as Fn

Fn.set_precision("3xf16")
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 200
noise = torch.empty(64 << 20, device="cuda")
for (F, r, B, T) in ((120, 7, 2, 200), (39, 7, 2, 200), (3075, 7, 2, 200), (120, 7, 6, 200)):
    E = 256
    torch.manual_seed(F)
    m = MLPMinion(num_inputs=E, num_outputs=F, dropout=0, hidden_size=256, hidden_layers=1, r=r,
                  skip=False, loss=ContextualizedLoss("MSELoss", r)).cuda()
    x0 = torch.randn(B, E, T, device="cuda")
    lab = torch.randn(B, F, T, device="cuda")
    for fused in (False, True):
        first, bad = None, {}
        for it in range(REPS):
            m.zero_grad()
            x = x0.clone().requires_grad_(True)
            if it % 3 == 1:
                noise.normal_()          # perturb timing / cache state
            y = m(x, label=lab if fused else None)
            loss = m.loss(y, lab)
            loss.backward()
            cur = {"dx": x.grad.clone()}
            cur.update({k: p.grad.clone() for k, p in m.named_parameters()})
            if first is None:
                first = cur
                continue
            for k in cur:
                d = float((cur[k] - first[k]).abs().max()) / float(first[k].abs().max())
                if d > 2e-6:
                    bad.setdefault(k, []).append((it, d))
        print("F=%d B=%d fused=%d: %s" % (F, B, fused, {k: (len(v), max(d for _, d in v)) for k, v in bad.items()} or "stable"))
