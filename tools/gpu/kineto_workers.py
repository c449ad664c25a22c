"""Per-kernel device times of the REAL workers+ step (encoder on 3B chunks + 12 heads + flat
Adam, one CUDA graph) from torch.profiler.  Usage: python tools/gpu/kineto_workers.py > out.md"""
import sys, os, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import torch
from torch.profiler import profile, ProfilerActivity
import bench

prec = sys.argv[1] if len(sys.argv) > 1 else "3xf16"
dev = torch.device("cuda", 0)
side = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(side)
hook = {}
orig = bench._event_time
def grab(fn, steps):
    hook["fn"] = fn
    return orig(fn, steps)
bench._event_time = grab
res = bench.time_workers(dev, prec, 32, steps=3, warmup=3, graph=True, stream=side)
fn = hook["fn"]
NS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NS):
        fn()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
kern = [e for e in evs if not e.name.startswith(("Memcpy", "Memset"))]
per = len(kern) // NS
last = kern[-per:]
t0, t1 = last[0].time_range.start, last[-1].time_range.end
def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "at::")
    cut = name.find("(")
    return (name if cut < 0 else name[:cut])[:64]
print("# torch.profiler kernel records, last of %d graph replays of the workers+ step (%s, B=32, T=32000): %.2f ms/step by events"
      % (NS, prec, res["ms_per_step"]))
print("\n%d kernels, span %.1f us\n" % (len(last), t1 - t0))
agg = collections.OrderedDict()
for e in last:
    a = agg.setdefault(short(e.name), [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
print("| kernel | launches | us | share of span |\n|---|---:|---:|---:|")
for nm, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("| `%s` | %d | %.1f | %.1f%% |" % (nm, c, us, 100 * us / (t1 - t0)))
print("\n## launches above 60 us, in program order\n\n| # | kernel | us |\n|---|---|---:|")
for i, e in enumerate(last):
    d = e.time_range.end - e.time_range.start
    if d > 60:
        print("| %d | `%s` | %.1f |" % (i, short(e.name), d))
