"""GPU debug: enc_pasep_train_4001 gradients vs golden, repeated, with the staged / register /
stored-du BatchNorm backward variants."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "oracle")]
import torch
from helpers import load_golden, resolve_cfg, fill_state_dict, seeded_randn, rel_l2, sample_view
from pase_b200.frontend import WaveFe
name = "enc_pasep_train_4001"
gold, meta = load_golden(name)
cfg = resolve_cfg(meta["cfg"])
for prec in ("3xf16", "3xtf32"):
    outs = []
    for rep in range(6):
        model = WaveFe(**cfg)
        model.precision = prec
        model.load_state_dict(fill_state_dict(model.state_dict(), meta["seed"]))
        model = model.cuda().train()
        x = seeded_randn((meta["N"], 1, meta["T"]), meta["seed"] + 1, 0.5).cuda()
        y = model(x)
        cot = seeded_randn(tuple(y.shape), meta["seed"] + 2).cuda()
        (y * cot).sum().backward()
        g = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
        errs = {}
        for k, v in gold.items():
            if k.endswith(("conv.bias", "W.bias")):
                continue
            if k.startswith("grad/"):
                errs[k[5:]] = rel_l2(g[k[5:]], v)
            elif k.startswith("gsample/"):
                errs[k[8:]] = rel_l2(sample_view(g[k[8:]]), v)
        outs.append((errs, g))
    keys = ["blocks.0.conv.low_hz_", "blocks.0.norm.weight", "blocks.1.conv.weight", "blocks.2.conv.weight", "blocks.4.conv.weight", "blocks.7.conv.weight", "W.weight"]
    print(prec, os.environ.get("PASE_B200_BN_STREAM"), os.environ.get("PASE_B200_BN_DU"))
    for errs, _ in outs:
        print("   ", "  ".join("%s %.2e" % (k.replace("blocks.", "b").replace(".conv", "").replace(".weight", ".w"), errs[k]) for k in keys))
    # run-to-run spread of the gradients themselves
    print("    run-to-run:", ["%.1e" % rel_l2(outs[i][1]["blocks.1.conv.weight"], outs[0][1]["blocks.1.conv.weight"]) for i in range(1, 6)])
