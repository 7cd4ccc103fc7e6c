"""Runs the reference's own Python (model graph, hardest-contrastive loss) on the oracle.  Needs /root/reference."""
import numpy as np
import pytest
import torch

from oracle import loss_cpu, me_cpu
from tests import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="/root/reference not present")


def test_reference_res16unet34c_builds_on_oracle():
    model_pkg = refload.load_reference_model_module(me_cpu.install)
    Net = model_pkg.load_model("Res16UNet34C")
    net = Net(3, 32, refload.default_config(), D=3)
    n = sum(p.numel() for p in net.parameters())
    assert n == 37_847_808
    convs = [m for m in net.modules() if isinstance(m, me_cpu.MinkowskiConvolution)]
    trs = [m for m in net.modules() if isinstance(m, me_cpu.MinkowskiConvolutionTranspose)]
    bns = [m for m in net.modules() if isinstance(m, me_cpu.MinkowskiBatchNorm)]
    assert (len(convs), len(trs), len(bns)) == (59, 4, 62)
    moms = sorted(m.bn.momentum for m in bns)
    assert moms.count(0.1) == 46 and moms.count(0.05) == 16
    sd = net.state_dict()
    assert sd["conv0p1s1.kernel"].shape == (27, 3, 32)
    assert sd["final.kernel"].shape == (1, 96, 32) and sd["final.bias"].shape == (1, 32)
    assert sd["block2.0.downsample.0.kernel"].shape == (1, 32, 64)
    assert "bn0.bn.running_mean" in sd and "block8.1.norm2.bn.weight" in sd


def test_hardest_loss_oracle_matches_reference_function():
    tr = refload.load_reference_trainer_module(me_cpu.install)
    obj = tr.HardestContrastiveLossTrainer.__new__(tr.HardestContrastiveLossTrainer)
    obj.pos_thresh, obj.neg_thresh = 0.1, 1.4
    g = torch.Generator().manual_seed(0)
    N0, N1, P = 700, 650, 3000
    F0 = torch.nn.functional.normalize(torch.randn(N0, 32, generator=g, dtype=torch.float64), dim=1)
    F1 = torch.nn.functional.normalize(torch.randn(N1, 32, generator=g, dtype=torch.float64), dim=1)
    F1[:300] = F0[:300] + 0.05 * torch.randn(300, 32, generator=g, dtype=torch.float64)
    rng = np.random.default_rng(0)
    i0 = np.sort(rng.integers(0, 300, P))
    pairs = torch.from_numpy(np.stack([i0, np.clip(i0 + rng.integers(-1, 2, P), 0, N1 - 1)], 1))
    for num_pos in (1024, 5000):
        np.random.seed(7)
        ref_pos, ref_neg = obj.contrastive_hardest_negative_loss(F0, F1, pairs, num_pos=num_pos, num_hn_samples=256)
        np.random.seed(7)
        sel0 = np.random.choice(N0, 256, replace=False)
        sel1 = np.random.choice(N1, 256, replace=False)
        pos_sel = np.random.choice(P, num_pos, replace=False) if P > num_pos else None
        pos, neg = loss_cpu.hardest_contrastive_loss(F0, F1, pairs.numpy(), sel0, sel1, pos_sel)
        assert torch.allclose(pos, ref_pos, rtol=1e-12) and torch.allclose(neg, ref_neg, rtol=1e-12)
