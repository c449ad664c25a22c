"""Per-kernel device times of the REAL (warm, back-to-back) encoder step from torch.profiler
(CUPTI activity records; no kernel replay, no cache flush -- unlike the ncu launch list).
Usage: python tools/gpu/kineto_step.py [precision] [eager|graph] > gpurun_out/kineto_<p>.md"""
import sys, os, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import torch
from torch.profiler import profile, ProfilerActivity
from bench import PASE_PLUS
from pase_b200 import wf_builder
from pase_b200.graph import GraphedEncoderStep
from pase_b200.optim import FlatAdam

prec = sys.argv[1] if len(sys.argv) > 1 else "3xf16"
mode = sys.argv[2] if len(sys.argv) > 2 else "graph"
B, T = 32, 32000
dev = torch.device("cuda", 0)
side = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(side)
torch.manual_seed(0)
model = wf_builder(dict(PASE_PLUS)).to(dev).train()
model.precision = prec
opt = FlatAdam(list(model.parameters()), lr=1e-4).bind_encoder(model)
if mode == "graph":
    gs = GraphedEncoderStep(model, opt, lambda y: y.square().mean(), (B, 1, T), dev, stream=side,
                            resident=True, x_init=torch.randn(B, 1, T))
    step = gs.step
else:
    x = torch.randn(B, 1, T, device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        model(x).square().mean().backward()
        opt.step()
for _ in range(5):
    step()
torch.cuda.synchronize()
NS = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
kern = [e for e in evs if not e.name.startswith(("Memcpy", "Memset"))]
per = len(kern) // NS
last = kern[-per:]
t0, t1 = last[0].time_range.start, last[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in last)
print("# torch.profiler (CUPTI) kernel records, last of %d %s steps, %s, B=%d T=%d" % (NS, mode, prec, B, T))
print("\n%d kernels, span %.1f us, summed kernel time %.1f us\n" % (len(last), t1 - t0, busy))
def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "at::")
    cut = name.find("(")
    return (name if cut < 0 else name[:cut])[:64]
agg = collections.OrderedDict()
for e in last:
    nm = short(e.name)
    a = agg.setdefault(nm, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
print("| kernel | launches | us | share of span |\n|---|---:|---:|---:|")
for nm, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% |" % (nm, c, us, 100 * us / (t1 - t0)))
print("\n## in program order\n\n| # | kernel | us | gap before us |\n|---|---|---:|---:|")
prev = None
for i, e in enumerate(last):
    gap = 0.0 if prev is None else e.time_range.start - prev
    prev = e.time_range.end
    print("| %d | `%s` | %.1f | %.1f |" % (i, short(e.name), e.time_range.end - e.time_range.start, gap))
