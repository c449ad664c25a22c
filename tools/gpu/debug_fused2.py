"""GPU debug: all regression heads on a fixed encoder output, repeated; find the first
intermediate that deviates between repetitions."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")]
import torch
from helpers import resolve_cfg, fill_state_dict, seeded_randn
from test_heads_gpu import _workers_plus_cfg
from pase_b200.pase import pase as native_pase
from pase_b200.utils import parse_workers
from pase_b200 import functional as Fn

fused = int(sys.argv[1]) if len(sys.argv) > 1 else 1
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fe_cfg, wcfg = resolve_cfg("cfg/frontend/PASE+.cfg"), _workers_plus_cfg()
Bm, Tm, Tq, seed = 2, 32000, 200, 73
Fn.set_precision("3xf16")
model = native_pase(frontend_cfg=fe_cfg, minions_cfg=parse_workers(copy.deepcopy(wcfg)))
model.load_state_dict(fill_state_dict(model.state_dict(), seed))
model = model.cuda().train()
wav = seeded_randn((Bm, 1, Tm), seed + 10, 0.5).cuda()
with torch.no_grad():
    enc = model.frontend(wav).detach().clone()          # (B, 256, 200)
labels = {w["name"]: seeded_randn((Bm, w["num_outputs"], Tq), seed + 100 + i).cuda()
          for i, w in enumerate(wcfg["regr"]) if w["name"] != "cchunk"}
heads = [(i, m) for i, m in enumerate(model.regression_workers) if i > 0]
store = {}
orig = Fn.fused_linear_ctx_mse
cur_head = [None]
def spy(h, weight, bias, label, F, r):
    name = cur_head[0]
    if h.requires_grad:
        h.register_hook(lambda g, name=name: store.__setitem__("dh/" + name, g.detach().clone()))
    return orig(h, weight, bias, label, F, r)
Fn.fused_linear_ctx_mse = spy
import pase_b200.minions as mn
first, bad = None, {}
for it in range(REPS):
    model.zero_grad()
    store.clear()
    x = enc.clone().requires_grad_(True)
    tot = 0
    for i, m in heads:
        cur_head[0] = m.name
        lab = labels[m.name]
        y = m(x, label=lab if fused else None)
        tot = tot + m.loss(y, lab)
    tot.backward()
    cur = dict(store)
    cur["dx"] = x.grad.clone()
    for i, m in heads:
        for k, p in m.named_parameters():
            cur["%s.%s" % (m.name, k)] = p.grad.clone()
    if first is None:
        first = cur
        continue
    for k in cur:
        d = float((cur[k] - first[k]).abs().max()) / float(first[k].abs().max())
        if d > 2e-6:
            nbad = int(((cur[k] - first[k]).abs() > 2e-6 * first[k].abs().max()).sum())
            bad.setdefault(k, []).append((it, d, nbad, cur[k].numel()))
print("fused", fused, "reps", REPS)
for k, v in bad.items():
    print("   ", k, v[:4], len(v))
if not bad:
    print("    stable")
