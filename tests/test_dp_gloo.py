"""N>1 host path: flat-gradient all-reduce over 2 ranks with the gloo backend (CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pase_b200.dp import FlatGradAllReducer, broadcast_parameters_and_buffers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # identical replicas
    torch.manual_seed(rank)                    # replicas start DIFFERENT ...
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16),
                                torch.nn.PReLU(16), torch.nn.Linear(16, 4))
    model[1].running_mean.fill_(float(rank))
    broadcast_parameters_and_buffers(model)    # ... and are aligned to rank 0 (incl. buffers)
    torch.manual_seed(0)
    ref0 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16),
                               torch.nn.PReLU(16), torch.nn.Linear(16, 4))
    assert all(torch.equal(a, b) for a, b in zip(model.parameters(), ref0.parameters()))
    assert float(model[1].running_mean.abs().max()) == 0.0
    model = torch.nn.Sequential(model[0], model[2], model[3])       # BN-free for the gradient checks
    red = FlatGradAllReducer(list(model.parameters()))
    g = torch.Generator().manual_seed(100 + rank)     # distinct shard per rank
    x = torch.randn(5, 8, generator=g)
    red.zero()
    model(x).square().mean().backward()
    # every .grad is a view of the single flat buffer
    assert all(p.grad.untyped_storage().data_ptr() == red.flat.untyped_storage().data_ptr()
               for p in model.parameters())
    red.all_reduce()
    first = red.flat.clone()
    # second protocol: fresh gradients packed with one multi-tensor copy
    red.detach()
    model(x).square().mean().backward()
    red.pack_and_reduce()
    assert torch.allclose(red.flat, first, rtol=1e-6, atol=1e-8)
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    # third protocol (CUDA-graph boundary between the two): pack(), then all_reduce()
    red.detach()
    model(x).square().mean().backward()
    red.pack()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    red.reduce()
    assert torch.allclose(red.flat, first, rtol=1e-6, atol=1e-8)
    # fourth protocol: the reference trainer's own sequence (worker_scheduler.py:43-75):
    # optimizer.zero_grad() -- set_to_none=True since torch 2.0, so every .grad stops being
    # a view -- backward, then all_reduce() must still reduce the REAL gradients
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    for set_to_none in (True, False):
        opt.zero_grad(set_to_none=set_to_none)
        model(x).square().mean().backward()
        red.all_reduce()
        assert torch.allclose(red.flat, first, rtol=1e-6, atol=1e-8), set_to_none
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    # a parameter that received no gradient contributes zeros, not stale values
    opt.zero_grad(set_to_none=True)
    model[0](x).square().mean().backward()          # only the first Linear gets gradients
    red.all_reduce()
    assert float(red.views[-1].abs().max()) == 0.0 and float(red.views[0].abs().max()) > 0.0
    torch.save(first.clone(), os.path.join(out, "g%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0 = torch.load(tmp_path / "g0.pt")
    g1 = torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)
    # equals the average of the two single-rank gradients
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.PReLU(16), torch.nn.Linear(16, 4))
    ref = 0
    for r in range(2):
        model.zero_grad()
        x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + r))
        model(x).square().mean().backward()
        ref = ref + torch.cat([p.grad.reshape(-1) for p in model.parameters()]) / 2
    assert torch.allclose(g0, ref, rtol=1e-5, atol=1e-7)
