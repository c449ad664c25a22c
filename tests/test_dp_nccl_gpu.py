"""Data parallelism on real GPUs (2 ranks, NCCL over NVLink): the flat-buffer all-reduce
averages the per-rank gradients, and the pipelined step (all-reduce hidden behind the tail
of backward, three CUDA graphs) equals the plain eager step.  Needs >= 2 GPUs:
    gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py -m gpu
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

PASE_PLUS = {"kwidths": [251, 20, 11, 11, 11, 11, 11, 11], "strides": [1, 10, 2, 1, 2, 1, 2, 2],
             "fmaps": [64, 64, 128, 128, 256, 256, 512, 512], "rnn_dim": 512, "denseskips": True,
             "norm_out": True, "rnn_pool": True, "rnn_layers": 1}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from pase_b200 import wf_builder
    from pase_b200.optim import FlatAdam
    from pase_b200.graph import PipelinedDPStep
    from pase_b200.dp import broadcast_parameters_and_buffers
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    B, T, split = 4, 8000, 3
    loss_fn = lambda y: y.square().mean()
    x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(100 + rank)).to(dev)

    def build(seed):
        torch.manual_seed(seed + rank)                  # replicas start different ...
        m = wf_builder(dict(PASE_PLUS)).to(dev).train()
        broadcast_parameters_and_buffers(m)             # ... and are aligned to rank 0
        params, n_lower = PipelinedDPStep.order_params(m, split)
        opt = FlatAdam(params, lr=1e-3).bind_encoder(m)
        return m, opt, n_lower

    # (1) the collective: reduced flat gradient == mean over ranks of the local gradients
    m, opt, n_lower = build(0)
    loss_fn(m(x)).backward()
    local = opt.flat_grad.clone()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    opt.reduce_grads()
    mean = sum(gathered) / world
    assert torch.allclose(opt.flat_grad, mean, rtol=1e-5, atol=1e-8 + 1e-6 * float(mean.abs().max()))
    assert float((gathered[0] - gathered[1]).abs().max()) > 0      # the shards really differ

    # (2) pipelined graphs == eager.  Adam turns ANY gradient into a +-lr step, so parameters
    # are a poor witness (noise-level gradients flip signs between two runs); the reduced flat
    # gradient is compared instead, with lr = 0 so that every step sees the same weights.
    def build0(seed):
        m, _, n_lower = build(seed)
        params, n_lower = PipelinedDPStep.order_params(m, split)
        m.grad_sink = None
        return m, FlatAdam(params, lr=0.0).bind_encoder(m), n_lower
    ma, oa, _ = build0(1)
    mb, ob, n_lower = build0(1)
    for (ka, pa), (kb, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert torch.equal(pa, pb), ka
    loss_fn(ma(x)).backward()
    oa.reduce_grads()
    oa.step()
    ga = oa.flat_grad.clone()
    # 1 eager warm-up step inside the constructor (lazy tables / plans must exist before the
    # capture) + 2 graph replays = 3 optimizer steps
    pipe = PipelinedDPStep(mb, ob, loss_fn, (B, 1, T), dev, split=split, n_lower=n_lower,
                           stream=side, resident=True, warmup=1, x_init=x.cpu().pin_memory())
    worst = 0.0
    for _ in range(2):
        ob.flat_grad.zero_()                       # a stale buffer must not pass the check
        pipe.step()
        torch.cuda.synchronize()
        gb = ob.flat_grad
        # two model instances see BatchNorm statistics that differ in the last bit (fp64
        # atomics), which now and then flips one PReLU gate: at B=4, T=8000 that moves whole
        # gradient rows by ~1e-3 of the norm.  A stale, unreduced or misplaced bucket would be
        # an O(1) error: relative L2 of the whole flat gradient + a loose elementwise bound.
        err = float((ga - gb).norm() / ga.norm())
        worst = max(worst, err)
        assert err <= 5e-3, err
        assert float((ga - gb).abs().max()) <= 5e-2 * float(ga.abs().max())
    assert float(ob.steps[0]) == 3.0 and float(oa.steps[0]) == 1.0
    # every rank holds the same parameters afterwards
    flat = ob.flat_param.clone()
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1])
    if rank == 0:
        open(os.path.join(out, "ok"), "w").write(
            "pipelined vs eager reduced gradient: worst relative L2 = %.3e\n" % worst)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_flat_allreduce_and_pipelined_step_two_gpus(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
    print(open(tmp_path / "ok").read())
