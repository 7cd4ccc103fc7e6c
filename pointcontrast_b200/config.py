"""Configuration with the keys of `pretrain/pointcontrast/config/defaults.yaml` that the hot path reads
(SURVEY.md section 8b), plus Hydra-style dotted overrides (`trainer.batch_size=32 misc.num_gpus=8 ...`,
`scripts/ddp_local.sh:10-26`).  Hydra/OmegaConf themselves are launcher plumbing and out of scope."""
import copy

_DEFAULTS = {
    "trainer": dict(trainer="HardestContrastiveLossTrainer", batch_size=4, num_pos_per_batch=1024,
                    num_hn_samples_per_batch=256, neg_thresh=1.4, pos_thresh=0.1, stat_freq=40, lr_update_freq=1000,
                    positive_pair_search_voxel_size_multiplier=1.5),
    "net": dict(model="Res16UNet34C", model_n_out=32, conv1_kernel_size=3, normalize_feature=True, dist_type="L2"),
    "opt": dict(optimizer="SGD", max_iter=300000, lr=1e-1, momentum=0.8, sgd_momentum=0.9, sgd_dampening=0.1,
                weight_decay=1e-4, bn_momentum=0.05, exp_gamma=0.99, scheduler="ExpLR"),
    "misc": dict(out_dir=".", use_gpu=True, num_gpus=1, weight=None, config=None, lenient_weight_loading=False,
                 nceT=0.07, npos=4096, train_num_thread=2),
    "data": dict(dataset="ScanNetMatchPairDataset", voxel_size=0.025, dataset_root_dir="", scannet_match_dir=""),
}


class Config(dict):
    """Nested dict with attribute access (`config.opt.lr`), like an OmegaConf node."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        if isinstance(v, dict) and not isinstance(v, Config):
            v = Config(v)
            self[k] = v
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self):
        return {k: (Config(v).to_dict() if isinstance(v, dict) else v) for k, v in self.items()}


def _parse(v):
    if v in ("None", "null", ""):
        return None
    if v in ("True", "true"):
        return True
    if v in ("False", "false"):
        return False
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v


def default_config(overrides=()):
    cfg = Config(copy.deepcopy(_DEFAULTS))
    for ov in overrides:
        key, _, val = ov.partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = getattr(node, p)
        node[parts[-1]] = _parse(val)
    return cfg
