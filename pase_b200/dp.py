"""Pure data parallelism over waveform chunks: one process per GPU,
model replicated, per-rank BatchNorm statistics (exactly N independent reference runs),
and ONE NCCL all-reduce per step over a single flat fp32 gradient buffer that every
parameter's .grad is a view of (no per-parameter collectives, no copies)."""
import torch
import torch.distributed as dist


class FlatGradAllReducer(object):
    def __init__(self, params, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.nbytes = n * 4

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        """Average the gradients over ranks (one collective)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / dist.get_world_size(self.group))
        return self.flat
