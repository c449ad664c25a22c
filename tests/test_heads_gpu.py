"""Worker heads + pase wrapper on the GPU (C-ABI kernels) against the reference goldens."""
import pytest
import torch

from test_heads_emulated import build_case, check_case
from helpers import assert_close, seeded_randn
from pase_b200.minions import MLPMinion
from pase_b200.losses import ContextualizedLoss
import pase_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["pase_mini_workers_1600", "pase_plus_workers_3200"])
def test_pase_workers_match_reference_golden(name):
    gold, meta, model, batch = build_case(name, "cuda")
    check_case(gold, meta, model, batch, "cuda", rtol=1e-3, atol=1e-5)


def test_mlp_minion_accepts_reference_layout_tensors():
    """A (B,C,T) tensor that did not come from the encoder goes through the NCT->rows kernel."""
    torch.manual_seed(0)
    m = MLPMinion(num_inputs=64, num_outputs=5, dropout=0, hidden_size=32, hidden_layers=1,
                  r=3, skip=False, loss=ContextualizedLoss("MSELoss", 3)).cuda()
    x = torch.randn(2, 64, 17, device="cuda", requires_grad=True)
    lab = torch.randn(2, 5, 17, device="cuda")
    y = m(x)
    assert tuple(y.shape) == (2, 15, 17)
    loss = m.loss(y, lab)
    loss.backward()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    xc = x.detach().cpu().requires_grad_(True)
    yr = O.head_mlp(xc, sd, "", 1)
    lr = ((yr - O.contextualize(lab.cpu(), 3)) ** 2).mean()
    lr.backward()
    assert_close(y, yr, 1e-4, 1e-5, "mlp pred")
    assert_close(loss, lr, 1e-4, 1e-6, "ctx mse")
    assert_close(x.grad, xc.grad, 1e-3, 1e-6, "dx")


def test_spc_worker_runs_and_matches_oracle_math():
    import random
    from pase_b200.minions import cls_worker_maker
    cfg = {"num_outputs": 1, "dropout": 0, "hidden_size": 32, "hidden_layers": 1, "name": "spc",
           "type": "spc", "loss": ContextualizedLoss("BCEWithLogitsLoss"), "skip": False}
    torch.manual_seed(1)
    w = cls_worker_maker(cfg, 20).cuda()
    x = torch.randn(3, 20, 100, device="cuda")
    random.seed(5)
    y, lab = w(x, 1, device="cuda")
    assert tuple(y.shape) == (6, 1, 1) and tuple(lab.shape) == (6, 1, 1)
    loss = w.loss(y, lab)
    # same draws on the CPU
    random.seed(5)
    N, M = 5, 21
    T = 100
    t = random.choice(list(range(M + 1, T - M)))
    ft = random.choice(list(range(t + 16, T - N)))
    pt = random.choice(list(range(N, t - 16)))
    xc = x.cpu()
    fut, past, cur = xc[:, :, ft:ft + N].reshape(3, -1), xc[:, :, pt - N:pt].reshape(3, -1), xc[:, :, t]
    full = torch.cat([torch.cat([cur, fut], 1), torch.cat([cur, past], 1)], 0).unsqueeze(2)
    sd = {k[len("minion."):]: v.detach().cpu() for k, v in w.state_dict().items()}
    yr = O.head_mlp(full, sd, "", 1)
    assert_close(y, yr, 1e-4, 1e-5, "spc logits")
    ref = torch.nn.functional.binary_cross_entropy_with_logits(
        yr, torch.cat([torch.ones(3, 1, 1), torch.zeros(3, 1, 1)], 0))
    assert_close(loss, ref, 1e-4, 1e-6, "spc bce")
