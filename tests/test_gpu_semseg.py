"""Semantic-segmentation finetune step (`downstream/semseg/lib/train.py:46-232`) on the GPU against the oracle:
Res16UNet34C(3 -> 13 classes, logits) on a synthetic S3DIS-shaped room, cross-entropy with ignored labels, SGD with dampening
under PolyLR, iter_size = 2 accumulation, and the lenient pretrain -> finetune checkpoint loading."""
import numpy as np
import pytest
import torch

from oracle import me_cpu as OR
from tests import refload
from tests.helpers import det_init, max_rel_err, model_backend, rel_err

pytestmark = pytest.mark.gpu


def _cfg():
    return refload.Cfg(optimizer=dict(optimizer="SGD", lr=0.01, sgd_momentum=0.9, sgd_dampening=0.1, weight_decay=1e-4, iter_size=2,
                                      scheduler="PolyLR", max_iter=100, poly_power=0.9), data=dict(ignore_label=255))


def test_finetune_step_matches_oracle():
    from pointcontrast_b200 import semseg, synth
    from pointcontrast_b200.model import load_model
    mcfg = refload.default_config(); mcfg["net"]["normalize_feature"] = False
    scenes = [synth.synth_scene(s, scale=0.25, voxel=0.05, n_raw=40_000) for s in (0, 1)]
    rng = np.random.default_rng(0)
    subs = []
    for sc in scenes:
        t = rng.integers(0, 13, len(sc["coords"])); t[rng.random(len(t)) < 0.15] = 255
        subs.append((torch.from_numpy(sc["coords"]), torch.from_numpy(sc["feats"]), torch.from_numpy(t)))
    # a "pretraining checkpoint": same backbone, 32-channel head with the trainer's key layout
    pre = load_model("Res16UNet34C")(3, 32, refload.default_config(), D=3)
    det_init(pre, 9)
    net = load_model("Res16UNet34C")(3, 13, mcfg, D=3)
    det_init(net, 4)
    kept = semseg.load_state_with_same_shape(net, {"module." + k: v for k, v in pre.state_dict().items()})
    assert "final.kernel" not in kept and "final.bias" not in kept and "block8.1.conv2.kernel" in kept
    assert torch.equal(net.state_dict()["conv0p1s1.kernel"], pre.state_dict()["conv0p1s1.kernel"])
    state = {k: v.clone() for k, v in net.state_dict().items()}
    tr = semseg.SegmentationTrainer(net, _cfg())
    loss = tr.train_step(subs, shift_coords=False)
    torch.cuda.synchronize()
    assert "_fused_runner" in net.__dict__                       # 13-class head: fused executor with the exact fp32 final layer
    # oracle: same graph in fp64, torch CrossEntropyLoss + SGD(dampening) + the PolyLR lambda
    with model_backend(OR) as mod:
        onet = mod.Res16UNet34C(3, 13, mcfg, D=3).double()
        onet.load_state_dict({k: (v.double() if v.dtype.is_floating_point else v) for k, v in state.items()})
        onet.train()
        oopt = torch.optim.SGD(onet.parameters(), lr=0.01, momentum=0.9, dampening=0.1, weight_decay=1e-4)
        oopt.zero_grad()
        lo = 0.0
        for c, f, t in subs:
            out = onet(OR.SparseTensor(f.double(), coords=c)).F
            l = torch.nn.functional.cross_entropy(out, t.long(), ignore_index=255) / 2
            l.backward()
            lo += float(l.detach())
        grads = {n: p.grad.clone() for n, p in onet.named_parameters()}
        oopt.step()
    assert abs(float(loss) - lo) / abs(lo) < 1e-3
    # the step itself: parameters after SGD vs the oracle's (the update is lr * gradient: compare the update, not the weights)
    for name in ("final.kernel", "final.bias", "block8.1.conv2.kernel", "bn0.bn.weight"):
        upd = dict(net.named_parameters())[name].detach().double().cpu() - state[name].double()
        oupd = dict(onet.named_parameters())[name].detach() - state[name].double()
        assert rel_err(upd, oupd) < (1e-3 if name.startswith("final") else 5e-2), name
    assert abs(tr.scheduler.get_last_lr()[0] - 0.01 * (1 - 1 / 101) ** 0.9) < 1e-12 and tr.curr_iter == 2
