"""GPU debug: whole-model fused vs unfused regression heads, run-to-run spread."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")]
import torch
from helpers import resolve_cfg, fill_state_dict, rel_l2, seeded_randn
from test_heads_gpu import _workers_plus_cfg
from pase_b200.pase import pase as native_pase, total_loss
from pase_b200.utils import parse_workers
from pase_b200 import functional as Fn

fe_cfg, wcfg = resolve_cfg("cfg/frontend/PASE+.cfg"), _workers_plus_cfg()
Bm, Tm, Tq, seed = 2, 32000, 200, 73
Fn.set_precision("3xf16")
runs = []
for fused in (False, False, True, True):
    model = native_pase(frontend_cfg=fe_cfg, minions_cfg=parse_workers(copy.deepcopy(wcfg)))
    model.load_state_dict(fill_state_dict(model.state_dict(), seed))
    model = model.cuda().train()
    model.fuse_regression_loss = fused
    batch = {k: seeded_randn((Bm, 1, Tm), seed + 10 + i, 0.5).cuda()
             for i, k in enumerate(["chunk", "chunk_ctxt", "chunk_rand", "cchunk"])}
    for i, w in enumerate(wcfg["regr"]):
        if w["name"] != "cchunk":
            batch[w["name"]] = seeded_randn((Bm, w["num_outputs"], Tq), seed + 100 + i).cuda()
    hh_, chunk, preds, labels = model(batch, 1, "cuda")
    tot, per = total_loss(model, preds, labels)
    tot.backward()
    runs.append((float(tot), chunk.detach().cpu(), {k: p.grad.detach().cpu() for k, p in model.named_parameters()}))
names = ["unfused0", "unfused1", "fused0", "fused1"]
for a, b in ((0, 1), (2, 3), (0, 2)):
    ga, gb = runs[a][2], runs[b][2]
    worst = sorted(((rel_l2(gb[k], ga[k]), k) for k in ga if k.startswith("regression_workers") and not k.startswith("regression_workers.0.")), reverse=True)[:6]
    print(names[a], names[b], "tot", runs[a][0], runs[b][0], "chunk rel", rel_l2(runs[b][1], runs[a][1]))
    for r, k in worst:
        print("    %.2e %s" % (r, k))
