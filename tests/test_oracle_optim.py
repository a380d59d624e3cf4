import math

import torch

from oracle import optim as oo


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 6):
        g = torch.randn(1000)
        p_ref.grad = g.clone()
        lr = 1e-4 * step
        for grp in opt.param_groups:
            grp["lr"] = lr
        opt.step()
        oo.adamw_step_(p, g, m, v, step, lr)
        torch.testing.assert_close(p, p_ref.data, rtol=1e-6, atol=1e-7)


def test_warmup_lr_sequence():
    seq = oo.lr_sequence(5, warmup_min_lr=0)
    assert seq[0] == 0.0 and seq[1] == 0.0
    assert math.isclose(seq[2], 1e-3 * math.log(2) / math.log(1000), rel_tol=1e-12)
    assert math.isclose(seq[3], 1e-3 * math.log(3) / math.log(1000), rel_tol=1e-12)
    s = oo.WarmupLR()
    for _ in range(1200):
        lr = s.step()
    assert lr == 1e-3
