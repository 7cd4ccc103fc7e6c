"""Trainer entry mirroring `pretrain/pointcontrast/lib/ddp_trainer.py`: same class names, constructor
(`Trainer(config, data_loader)`), `train()` / `_train_iter()` / `_save_checkpoint()` methods, config keys,
checkpoint layout and loss definitions -- on the libpcb200 kernels.

Intentional differences:
  * no per-iteration `torch.cuda.empty_cache()` and no `set_detect_anomaly(True)` (`ddp_trainer.py:36,321,437`);
  * data parallelism is one flat fp32 gradient buffer all-reduced over NCCL (sum, then 1/world folded into the fused
    SGD kernel) instead of DistributedDataParallel's bucketed hooks -- the result (mean gradient over ranks, per-rank
    BatchNorm statistics, `broadcast_buffers=False`) is the same;
  * the loss-sampling RNG is a per-trainer torch.Generator on the device instead of the process-global numpy RNG.
"""
import logging
import os
import os.path as osp

import torch
import torch.distributed as dist

from . import fused, losses, me as ME
from .model import load_model
from .optim import FlatSGD


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def scaled_all_reduce_dict(res, num_gpus):
    """`lib/distributed.py:260-270`: mean over ranks of a dict of scalar tensors."""
    if get_world_size() == 1:
        return res
    keys = sorted(res)
    buf = torch.stack([res[k].detach().float() for k in keys])
    dist.all_reduce(buf)
    buf /= num_gpus
    return {k: buf[i] for i, k in enumerate(keys)}


def load_state(model, weights, lenient_weight_loading=False):
    """`ddp_trainer.py:54-69`."""
    if lenient_weight_loading:
        model_state = model.state_dict()
        filtered = {k: v for k, v in weights.items() if k in model_state and v.size() == model_state[k].size()}
        logging.info("Load weights:" + ", ".join(filtered.keys()))
        weights = model_state
        weights.update(filtered)
    model.load_state_dict(weights, strict=True)


class ContrastiveLossTrainer:
    def __init__(self, config, data_loader):
        assert config.misc.use_gpu and torch.cuda.is_available(), "DDP mode must support GPU"
        num_feats = 3
        self.world = get_world_size()
        self.is_master = get_rank() == 0
        self.cur_device = torch.cuda.current_device()
        self.device = torch.device("cuda", self.cur_device)
        Model = load_model(config.net.model)
        model = Model(num_feats, config.net.model_n_out, config, D=3).cuda(self.cur_device)
        self.config = config
        self.model = model
        if config.opt.optimizer != "SGD":
            raise NotImplementedError("the hot path uses SGD (`config/defaults.yaml:44`)")
        self.optimizer = FlatSGD(model.parameters(), lr=config.opt.lr, momentum=config.opt.momentum,
                                 weight_decay=config.opt.weight_decay)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, config.opt.exp_gamma)
        self.curr_iter = 0
        self.batch_size = data_loader.batch_size
        self.data_loader = data_loader
        self.neg_thresh = config.trainer.neg_thresh
        self.pos_thresh = config.trainer.pos_thresh
        self.stat_freq = config.trainer.stat_freq
        self.lr_update_freq = config.trainer.lr_update_freq
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(1234 + get_rank())
        self.writer = None

        if config.misc.weight:
            state = torch.load(config.misc.weight, map_location="cpu", weights_only=False)
            load_state(model, state["state_dict"], config.misc.lenient_weight_loading)
        checkpoint_fn = "weights/weights.pth"
        if osp.isfile(checkpoint_fn):
            state = torch.load(checkpoint_fn, map_location="cpu", weights_only=False)
            self.curr_iter = state["curr_iter"]
            load_state(model, state["state_dict"])
            self.optimizer.load_state_dict(state["optimizer"])
            self.scheduler.load_state_dict(state["scheduler"])
            if self.is_master:
                logging.info("=> loaded checkpoint '%s' (curr_iter %d)", checkpoint_fn, state["curr_iter"])
        if self.world > 1:                       # DDP construction semantics: every rank starts from rank 0's state
            dist.broadcast(self.optimizer.flat_param, 0)
            ME.bump_weights_epoch()
            for b in model.buffers():
                dist.broadcast(b, 0)
            self.optimizer.grad_scale = 1.0 / self.world

    # -- checkpoint (`ddp_trainer.py:151-169`)
    def _save_checkpoint(self, curr_iter, filename="checkpoint"):
        if not self.is_master:
            return
        os.makedirs("weights", mode=0o755, exist_ok=True)
        state = {"curr_iter": curr_iter, "state_dict": self.model.state_dict(), "optimizer": self.optimizer.state_dict(),
                 "scheduler": self.scheduler.state_dict(), "config": self.config.to_dict() if hasattr(self.config, "to_dict")
                 else self.config}
        filepath = os.path.join("weights", f"{filename}.pth")
        logging.info("Saving checkpoint: %s ...", filepath)
        torch.save(state, filepath)
        link = "weights/weights.pth"
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(f"{filename}.pth", link)

    # -- shared step pieces
    def _forward_views(self, input_dict):
        dev = self.device
        return fused.forward_pair(self.model, input_dict["sinput0_F"], input_dict["sinput0_C"], input_dict["sinput1_F"],
                                  input_dict["sinput1_C"], dev)

    def _all_reduce_grads(self):
        if self.world > 1:
            dist.all_reduce(self.optimizer.flat_grad)          # sum; the 1/world is folded into the SGD kernel

    def train(self):
        curr_iter = self.curr_iter
        it = iter(self.data_loader)
        while curr_iter < self.config.opt.max_iter:
            curr_iter += 1
            out = self._train_iter(it, None)
            batch_loss = out[0] if isinstance(out, tuple) else out
            if curr_iter % self.lr_update_freq == 0 or curr_iter == 1:
                lr = self.scheduler.get_last_lr()
                self.scheduler.step()
                if self.is_master:
                    logging.info(" Iter: %d, LR: %s", curr_iter, lr)
                    self._save_checkpoint(curr_iter, "checkpoint_" + str(curr_iter))
            if curr_iter % self.stat_freq == 0 and self.is_master:
                logging.info("Train iter %d, Current Loss: %.3e, LR: %s", curr_iter, batch_loss, self.scheduler.get_last_lr())
        self.curr_iter = curr_iter


class HardestContrastiveLossTrainer(ContrastiveLossTrainer):
    """`ddp_trainer.py:171-326`."""

    def contrastive_hardest_negative_loss(self, F0, F1, positive_pairs, num_pos=5192, num_hn_samples=2048, thresh=None):
        N0, N1 = F0.shape[0], F1.shape[0]
        dev, g = F0.device, self.generator
        sel0 = torch.randperm(N0, device=dev, generator=g)[:min(N0, num_hn_samples)]
        sel1 = torch.randperm(N1, device=dev, generator=g)[:min(N1, num_hn_samples)]
        P = positive_pairs.shape[0]
        pos_sel = torch.randperm(P, device=dev, generator=g)[:num_pos] if P > num_pos else None
        return losses.hardest_contrastive_loss(F0, F1, positive_pairs, sel0, sel1, pos_sel, self.pos_thresh, self.neg_thresh)

    def train_step(self, input_dict):
        """One iteration on a batch dict; returns device scalars (loss, pos_loss, neg_loss) without synchronising."""
        self.model.train()
        self.optimizer.zero_grad()
        F0, F1 = self._forward_views(input_dict)
        pos_pairs = input_dict["correspondences"].to(self.device, non_blocking=True)
        pos_loss, neg_loss = self.contrastive_hardest_negative_loss(
            F0, F1, pos_pairs, num_pos=self.config.trainer.num_pos_per_batch * self.batch_size,
            num_hn_samples=self.config.trainer.num_hn_samples_per_batch * self.batch_size)
        loss = pos_loss + neg_loss
        loss.backward()
        self._all_reduce_grads()
        self.optimizer.step()
        return loss.detach(), pos_loss.detach(), neg_loss.detach()

    def _train_iter(self, data_loader_iter, timers):
        input_dict = next(data_loader_iter)
        loss, pos_loss, neg_loss = self.train_step(input_dict)
        result = scaled_all_reduce_dict({"loss": loss, "pos_loss": pos_loss, "neg_loss": neg_loss}, self.world)
        return result["loss"].item(), result["pos_loss"].item(), result["neg_loss"].item()


class PointNCELossTrainer(ContrastiveLossTrainer):
    """`ddp_trainer.py:328-440`."""

    def __init__(self, config, data_loader):
        super().__init__(config, data_loader)
        self.T = config.misc.nceT
        self.npos = config.misc.npos

    def train_step(self, input_dict):
        self.optimizer.zero_grad()
        F0, F1 = self._forward_views(input_dict)
        pos_pairs = input_dict["correspondences"].to(self.device, non_blocking=True)
        q_rows, k_rows = losses.select_positives(pos_pairs, self.npos, self.generator)
        loss = losses.point_nce_loss(F0, F1, q_rows, k_rows, self.T)
        loss.backward()
        self._all_reduce_grads()
        self.optimizer.step()
        return loss.detach()

    def _train_iter(self, data_loader_iter, timers):
        input_dict = next(data_loader_iter)
        loss = self.train_step(input_dict)
        result = scaled_all_reduce_dict({"loss": loss}, self.world)
        return result["loss"].item()


def get_trainer(trainer):
    """`pretrain/pointcontrast/ddp_train.py:33-39`."""
    table = {"HardestContrastiveLossTrainer": HardestContrastiveLossTrainer, "PointNCELossTrainer": PointNCELossTrainer}
    if trainer not in table:
        raise ValueError(f"Trainer {trainer} not found")
    return table[trainer]
