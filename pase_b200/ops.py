"""Single funnel for every kernel launch of the product path.

``call(name, *args)`` forwards to the C-ABI library (pase_b200/_lib.py).  There
is deliberately no alternative backend here: without the CUDA library the call
raises.  (tests/ monkeypatches this symbol with a torch emulation ONLY to check
the host-side orchestration on a GPU-less box; the emulation lives in tests/.)
"""
from . import _lib

launch_count = 0


def call(name, *args):
    global launch_count
    launch_count += 1
    return _lib.call(name, *args)
