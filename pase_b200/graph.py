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
                 capture_error_mode="global"):
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.x_static = torch.zeros(shape, dtype=torch.float32, device=device)
        self.x_host = torch.zeros(shape, dtype=torch.float32).pin_memory()
        self.loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        if x_init is not None:               # batch used by the warm-up steps
            self.x_host.copy_(x_init)
            self.x_static.copy_(x_init)
        self.pre_step, self.post_backward = pre_step, post_backward
        self.resident = resident        # True: input stays in HBM, no H2D / D2H in the graph
        # capture on the stream the model's autograd nodes already live on: an AccumulateGrad
        # node created on another stream invalidates the capture
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        self.graph = None
        # "thread_local" when other threads of the process issue CUDA calls during the
        # capture (e.g. the NCCL process group's watchdog, with a collective in post_backward)
        self.capture_error_mode = capture_error_mode
        self._capture(warmup)

    def _one(self):
        if not self.resident:
            self.x_static.copy_(self.x_host, non_blocking=True)
        if self.pre_step is not None:
            self.pre_step()
        y = self.model(self.x_static)
        loss = self.loss_fn(y)
        loss.backward()
        if self.post_backward is not None:
            self.post_backward()
        self.opt.step()
        if not self.resident:
            self.loss_host.copy_(loss.detach(), non_blocking=True)

    def _capture(self, warmup):
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.opt.zero_grad(set_to_none=True)
                self._one()
        self.stream.synchronize()
        torch.cuda.current_stream().wait_stream(self.stream)
        self.opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream, capture_error_mode=self.capture_error_mode):
            self._one()
        self.graph = g

    def step(self, x_host=None):
        """Replays the captured step.  The batch is read from the pinned staging buffer
        ``self.x_host`` (a data loader writes batches straight into it); passing ``x_host``
        copies it there first."""
        if x_host is not None:
            self.x_host.copy_(x_host)
        self.graph.replay()
        if self.resident:
            return None
        self.stream.synchronize()
        return float(self.loss_host)
