"""CUDA-graph capture of one whole training step (H2D of the waveform batch, encoder
forward + backward through the C-ABI kernels, optimizer update, D2H of the loss) so that a
step is ONE launch: the hot loop is launch-bound on the host once the kernels take only a
few ms.  Buffers of the encoder plan are static by construction (EncoderPlan), TMA
descriptors are kernel parameters, so the captured graph is replayable as is.
"""
import torch


class GraphedEncoderStep(object):
    """step(x_host) -> loss (python float).  `x_host` must be a pinned CPU tensor of the shape
    given at construction; `loss_fn(y) -> scalar tensor`; the optimizer must be capturable."""

    def __init__(self, model, optimizer, loss_fn, shape, device, pre_step=None, post_backward=None,
                 warmup=3, stream=None, resident=False, x_init=None,
                 capture_error_mode="global", between=None):
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.x_static = torch.zeros(shape, dtype=torch.float32, device=device)
        self.x_host = torch.zeros(shape, dtype=torch.float32).pin_memory()
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        if x_init is not None:               # batch used by the warm-up steps
            self.x_host.copy_(x_init)
            self.x_static.copy_(x_init)
        self.pre_step, self.post_backward = pre_step, post_backward
        # `between`: a callable that runs EAGERLY between two captured halves of the step
        # ([H2D, forward, backward, post_backward] and [optimizer, D2H]).  Data parallelism
        # puts its NCCL all-reduce there: the collective is never captured, only ordered on
        # the same stream between two graph replays.
        self.between = between
        self.resident = resident        # True: input stays in HBM, no H2D / D2H in the graph
        # capture on the stream the model's autograd nodes already live on: an AccumulateGrad
        # node created on another stream invalidates the capture
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        self.graph = self.graph_b = None
        # "thread_local" when other threads of the process issue CUDA calls during the
        # capture (e.g. an NCCL process group's watchdog)
        self.capture_error_mode = capture_error_mode
        self._loss = None
        self._capture(warmup)

    def _part_a(self):
        if not self.resident:
            self.x_static.copy_(self.x_host, non_blocking=True)
        if self.pre_step is not None:
            self.pre_step()
        y = self.model(self.x_static)
        self._loss = self.loss_fn(y)
        self._loss.backward()
        if self.post_backward is not None:
            self.post_backward()

    def _part_b(self):
        self.opt.step()
        if not self.resident:
            self.loss_host.copy_(self._loss.detach(), non_blocking=True)

    def _one(self):
        self._part_a()
        if self.between is not None:
            self.between()
        self._part_b()

    def _capture(self, warmup):
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.opt.zero_grad(set_to_none=True)
                self._one()
        self.stream.synchronize()
        torch.cuda.current_stream().wait_stream(self.stream)
        self.opt.zero_grad(set_to_none=True)
        mode = self.capture_error_mode
        g = torch.cuda.CUDAGraph()
        if self.between is None:
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode=mode):
                self._one()
        else:
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode=mode):
                self._part_a()
            with torch.cuda.stream(self.stream):
                self.between()
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb, stream=self.stream, pool=g.pool(), capture_error_mode=mode):
                self._part_b()
            self.graph_b = gb
        self.graph = g

    def step(self, x_host=None):
        """Replays the captured step.  The batch is read from the pinned staging buffer
        ``self.x_host`` (a data loader writes batches straight into it); passing ``x_host``
        copies it there first."""
        if x_host is not None:
            self.x_host.copy_(x_host)
        self.graph.replay()
        if self.graph_b is not None:
            with torch.cuda.stream(self.stream):
                self.between()
            self.graph_b.replay()
        if self.resident:
            return None
        self.stream.synchronize()
        return float(self.loss_host)
