"""Synthetic stand-in for `make_data_loader` (`pretrain/pointcontrast/lib/ddp_data_loaders.py:272-309`): an infinite
iterable with a `batch_size` attribute yielding the reference's batch dict (torch CPU tensors, optionally pinned).
The real ScanNet pair loader (open3d KD-tree on CPU workers) is data preparation, out of scope."""
import torch

from . import synth

_KEYS = ("sinput0_C", "sinput0_F", "sinput1_C", "sinput1_F", "correspondences", "pcd0", "pcd1")


def to_torch(batch, pin=False):
    out = dict(batch)
    for k in _KEYS:
        t = torch.from_numpy(batch[k])
        out[k] = t.pin_memory() if pin else t
    return out


class SyntheticPairLoader:
    """Cycles over `num_batches` pre-generated batches of `batch_size` scene pairs (per rank)."""

    def __init__(self, batch_size, scale=0.9, voxel_size=0.025, num_batches=2, rank=0, pin=True, n_raw=300_000):
        self.batch_size = batch_size
        self.batches = [to_torch(synth.synth_batch(100 * rank + s, batch_size, scale, voxel_size, n_raw),
                                 pin and torch.cuda.is_available()) for s in range(num_batches)]

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        i = 0
        while True:
            yield self.batches[i % len(self.batches)]
            i += 1
