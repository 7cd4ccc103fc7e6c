"""Semantic-segmentation finetune step on the same backbone: mirror of `downstream/semseg/lib/train.py:46-232`
(the optimisation step of the loop, not its logging / validation / tensorboard shell), `lib/solvers.py:27-83` (SGD with
dampening, PolyLR) and `lib/utils.py:19-43` (lenient loading of pretraining checkpoints into a model with a different head).

    model = load_model("Res16UNet34C")(3, num_labels, config, D=3)           # `normalize_feature` False: logits
    load_state_with_same_shape(model, torch.load("weights.pth")["state_dict"])   # PointContrast backbone, fresh `final` layer
    trainer = SegmentationTrainer(model, config)
    loss = trainer.train_step([(coords, feats, target), ...])                 # len == config.optimizer.iter_size

Everything numerical runs on libpcb200: the fused executor (the 13 / 20-class head on the exact fp32 kernels), the cross-entropy
kernels (`pcb_ce_forward_backward`), the flat SGD kernel with dampening.
"""
import logging

import torch

from . import losses, me as ME
from .optim import FlatSGD, PolyLR


def load_state_with_same_shape(model, weights):
    """`downstream/semseg/lib/utils.py:19-43`: keep the checkpoint entries whose name and shape match the model (drops a `final`
    head of another width), stripping the `module.` / `encoder.` prefixes.  Returns the filtered dict AND loads it (strict=False)."""
    state = model.state_dict()
    first = next(iter(weights))
    if first.startswith("module."):
        weights = {k.partition("module.")[2]: v for k, v in weights.items()}
    if next(iter(weights)).startswith("encoder."):
        weights = {k.partition("encoder.")[2]: v for k, v in weights.items()}
    filtered = {k: v for k, v in weights.items() if k in state and v.size() == state[k].size()}
    logging.info("Loading weights:" + ", ".join(filtered.keys()))
    model.load_state_dict(filtered, strict=False)
    ME.bump_weights_epoch()
    return filtered


def initialize_optimizer(params, config):
    """`lib/solvers.py:47-57` (SGD branch; the hot path's optimiser)."""
    if config.optimizer != "SGD":
        raise ValueError("Optimizer type not supported")
    return FlatSGD(params, lr=config.lr, momentum=config.sgd_momentum, dampening=config.sgd_dampening, weight_decay=config.weight_decay)


def initialize_scheduler(optimizer, config, last_step=-1):
    """`lib/solvers.py:66-83`."""
    if config.scheduler == "PolyLR":
        return PolyLR(optimizer, max_iter=config.max_iter, power=config.poly_power, last_step=last_step)
    if config.scheduler == "StepLR":
        return torch.optim.lr_scheduler.StepLR(optimizer, step_size=config.step_size, gamma=config.step_gamma, last_epoch=last_step)
    if config.scheduler == "ExpLR":
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: config.exp_gamma ** (s / config.exp_step_size), last_step)
    raise ValueError("Scheduler not supported")


class SegmentationTrainer:
    def __init__(self, model, config, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model = model.to(self.device)
        self.config = config
        self.optimizer = initialize_optimizer(self.model.parameters(), config.optimizer)
        self.scheduler = initialize_scheduler(self.optimizer, config.optimizer)
        self.ignore_label = config.data.ignore_label
        self.iter_size = config.optimizer.iter_size
        self.curr_iter = 1

    def train_step(self, sub_batches, shift_coords=True):
        """One optimiser step = `iter_size` sub-batches of (coords int32 [N,4], feats fp32 [N,3], target int [N]), gradients
        accumulated (`lib/train.py:97-160`).  Returns the summed (already 1/iter_size-scaled) loss as a device scalar."""
        assert len(sub_batches) == self.iter_size
        self.model.train()
        self.optimizer.zero_grad()
        total = None
        for coords, feats, target in sub_batches:
            if shift_coords:          # `lib/train.py:110`: even/odd-coordinate invariance (shifts the batch column too: SURVEY.md appendix B)
                coords = coords.clone()
                coords[:, :3] += (torch.rand(3) * 100).type_as(coords)
            sinput = ME.SparseTensor(feats, coords).to(self.device)
            soutput = self.model(sinput)
            loss = losses.cross_entropy(soutput.F, target.to(self.device, non_blocking=True), self.ignore_label) / self.iter_size
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        self.optimizer.step()
        self.scheduler.step()
        self.curr_iter += 1
        return total
