"""Model zoo of the hot path: `load_model(name)` mirrors `pretrain/pointcontrast/model/__init__.py:20-31`."""
from . import res16unet

MODELS = [getattr(res16unet, a) for a in dir(res16unet) if "Net" in a and isinstance(getattr(res16unet, a), type)]


def get_models():
    return MODELS


def load_model(name):
    table = {m.__name__: m for m in MODELS}
    if name not in table:
        raise KeyError(f"unknown model {name!r}; options: {sorted(table)}")
    return table[name]
