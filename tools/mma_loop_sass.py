#!/usr/bin/env python
"""Counts the SASS instructions of the MMA-issue loop (first full-barrier TRYWAIT before the
first UTCHMMA .. the stage-freeing UTCBAR) of every tcgen05 kernel in libpase_b200.so.
The issuing thread's instruction stream bounds the MMA rate (profiles/r01_history.md)."""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "pase_b200/csrc/libpase_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
for f in funcs:
    name = f.split("\n", 1)[0]
    if "tc_gemm" not in name:
        continue
    m = re.search(r"tc_gemm_\w+?_kernelI((?:L[ib]\d+E)+)", name)
    ins = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = [re.sub(r"/\*[0-9a-fx ]+\*/", "", l).strip().rstrip("; ") for l in ins]
    mma = [i for i, o in enumerate(ops) if "UTCHMMA" in o]
    if not mma:
        continue
    first = mma[0]
    start = max(i for i in range(first) if "SYNCS.PHASECHK" in ops[i])
    # fast (converged) path only: ptxas also emits a BRA.DIV fallback copy further down
    ends = [i for i in range(first, len(ops)) if "UTCBAR" in ops[i]]
    end = ends[0] if ends else mma[-1]
    mma = [i for i in mma if i <= end]
    body = ops[start:end + 1]
    kinds = {}
    for o in body:
        k = o.split()[1] if o.startswith("@") else o.split()[0]
        kinds[k] = kinds.get(k, 0) + 1
    top = sorted(kinds.items(), key=lambda kv: -kv[1])[:8]
    targs = re.findall(r"L[ib](\d+)E", m.group(1)) if m else []
    print(re.search(r"tc_gemm_\w+?_kernel", name).group(0), "<%s>" % ",".join(targs), "mma", len(mma),
          "loop instrs", len(body), top)
