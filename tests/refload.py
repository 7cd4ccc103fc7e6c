"""Import pieces of the read-only reference tree (/root/reference) on top of a chosen MinkowskiEngine-compatible
module.  In the build container that is the reference itself; on the GPU box only the model package staged by
`oracle/stage_ref.py` is available (the trainer module is not)."""
import importlib
import os
import sys
import types

REF_PC = "/root/reference/pretrain/pointcontrast"
if not os.path.isdir(REF_PC):       # GPU box: the model package staged by oracle/stage_ref.py (git-ignored, travels with gpurun)
    _staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "pointcontrast")
    if os.path.isdir(os.path.join(_staged, "model")):
        REF_PC = _staged


def available():
    return os.path.isdir(os.path.join(REF_PC, "model"))


def _purge():
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.") or k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]


def load_reference_model_module(me_module_installer):
    """Returns the reference's `model` package imported against the ME surface installed by `me_module_installer()`."""
    _purge()
    me_module_installer()
    sys.path.insert(0, REF_PC)
    try:
        return importlib.import_module("model")
    finally:
        sys.path.remove(REF_PC)


def load_reference_trainer_module(me_module_installer):
    _purge()
    me_module_installer()
    for name in ("tensorboardX", "omegaconf"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.SummaryWriter = object
            m.OmegaConf = object
            sys.modules[name] = m
    sys.path.insert(0, REF_PC)
    try:
        return importlib.import_module("lib.ddp_trainer")
    finally:
        sys.path.remove(REF_PC)


class Cfg(dict):
    """attribute-style nested config, enough for `config.net.x` / `config.opt.y`."""
    def __getattr__(self, k):
        v = self[k]
        return Cfg(v) if isinstance(v, dict) else v


def default_config():
    return Cfg(net=dict(model="Res16UNet34C", model_n_out=32, conv1_kernel_size=3, normalize_feature=True),
               opt=dict(bn_momentum=0.05))
