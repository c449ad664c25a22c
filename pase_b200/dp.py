"""Pure data parallelism over waveform chunks: one process per GPU,
model replicated, per-rank BatchNorm statistics (exactly N independent reference runs),
and ONE NCCL all-reduce per step over a single flat fp32 gradient buffer (31 MB for the
PASE+ encoder, 119 MB with every workers+ head).  No per-parameter collectives.

Two ways to fill the buffer:
  * ``zero()`` ... backward ... ``all_reduce()``: every ``.grad`` is a *view* of the buffer
    and autograd accumulates into it (works with any caller, costs one add per parameter);
  * backward with ``.grad = None`` ... ``pack_and_reduce()``: the gradients autograd produced
    are packed with one multi-tensor copy, reduced, and ``.grad`` is re-pointed at the
    reduced views (no accumulation kernels).
"""
import torch
import torch.distributed as dist


class FlatGradAllReducer(object):
    def __init__(self, params, process_group=None, attach=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        if attach:
            self.attach()
        self.nbytes = n * 4

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def detach(self):
        for p in self.params:
            p.grad = None

    def zero(self):
        self.flat.zero_()

    def _reduce(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / dist.get_world_size(self.group))

    def all_reduce(self):
        """Average the (view-accumulated) gradients over ranks: one collective."""
        self._reduce()
        return self.flat

    def pack(self):
        """Pack freshly produced gradients into the flat buffer and re-point ``.grad`` at
        its views, without the collective (``all_reduce()`` follows; used when the two are
        separated by a CUDA-graph boundary)."""
        srcs, dsts = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                srcs.append(p.grad)
                dsts.append(v)
        if srcs:
            torch._foreach_copy_(dsts, srcs)
        self.attach()
        return self.flat

    def pack_and_reduce(self):
        """Pack freshly produced gradients (``.grad`` not views), reduce, re-point ``.grad``."""
        srcs, dsts = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                srcs.append(p.grad)
                dsts.append(v)
        if srcs:
            torch._foreach_copy_(dsts, srcs)
        self._reduce()
        self.attach()
        return self.flat
