"""`Config`: drop-in for virtex/config.py:6-236 without fvcore/yacs (neither is needed, nor installed).

Same constructor `Config(config_file=None, override_list=[])`, same nested attribute access (`_C.OPTIM.BATCH_SIZE`),
YAML files with `_BASE_` inheritance (relative to the including file), flat `[key, value, ...]` overrides with literal
evaluation of strings and type checking against the defaults, immutability after construction, `dump`, `str`, `repr`.
"""
import ast
import copy
import os
from typing import Any, List, Optional

import yaml


class CfgNode(dict):
    """Attribute-style nested dict that can be frozen."""

    def __init__(self, init: Optional[dict] = None):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        for k, v in (init or {}).items():
            dict.__setitem__(self, k, CfgNode(v) if isinstance(v, dict) else v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError(f"Attempted to set {name} to {value}, but the config is immutable")
        dict.__setitem__(self, name, CfgNode(value) if isinstance(value, dict) and not isinstance(value, CfgNode) else value)

    def __setitem__(self, key, value):
        self.__setattr__(key, value)

    def freeze(self, flag: bool = True):
        object.__setattr__(self, "_frozen", flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze(flag)

    def is_frozen(self):
        return object.__getattribute__(self, "_frozen")

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, CfgNode) else copy.deepcopy(v)) for k, v in self.items()}

    def dump(self, stream=None, **kwargs):
        return yaml.safe_dump(self.to_dict(), stream, **kwargs)

    def __str__(self):
        return yaml.safe_dump(self.to_dict(), default_flow_style=False)

    def __deepcopy__(self, memo):
        return CfgNode(self.to_dict())

    # ------------------------------------------------------------------------------------------------------ merging
    @staticmethod
    def load_yaml_with_base(path: str) -> dict:
        with open(path) as f:
            cfg = yaml.safe_load(f) or {}
        base = cfg.pop("_BASE_", None)
        if base is None:
            return cfg
        if not os.path.isabs(base):
            base = os.path.join(os.path.dirname(path), base)
        merged = CfgNode.load_yaml_with_base(base)
        _merge_dict(cfg, merged)
        return merged

    def merge_from_file(self, path: str):
        self._merge(CfgNode.load_yaml_with_base(path), self, [])

    def merge_from_list(self, lst: List[Any]):
        if len(lst) % 2 != 0:
            raise ValueError(f"override list must hold (key, value) pairs, got {lst}")
        for key, value in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError(f"Non-existent config key: {key}")
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent config key: {key}")
            node.__setattr__(parts[-1], _coerce(_decode(value), node[parts[-1]], key))

    @staticmethod
    def _merge(src: dict, dst: "CfgNode", path):
        for k, v in src.items():
            full = ".".join(path + [k])
            if k not in dst:
                raise KeyError(f"Non-existent config key: {full}")
            if isinstance(v, dict) and isinstance(dst[k], CfgNode):
                CfgNode._merge(v, dst[k], path + [k])
            else:
                dst.__setattr__(k, _coerce(_decode(v), dst[k], full))


def _merge_dict(src: dict, dst: dict):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_dict(v, dst[k])
        else:
            dst[k] = v


def _decode(value):
    if isinstance(value, str):
        try:
            return ast.literal_eval(value)
        except (ValueError, SyntaxError):
            return value
    return value


def _coerce(value, original, key):
    """yacs-style type check: the replacement must have the default's type (int->float and tuple<->list allowed)."""
    if original is None or value is None or type(value) is type(original):
        return value
    if isinstance(original, float) and isinstance(value, int) and not isinstance(value, bool):
        return float(value)
    if isinstance(original, (list, tuple)) and isinstance(value, (list, tuple)):
        return type(original)(value)
    raise ValueError(f"Type mismatch ({type(original).__name__} vs. {type(value).__name__}) for config key: {key}")


_DEFAULTS = {
    "RANDOM_SEED": 0,
    "AMP": True,
    "CUDNN_DETERMINISTIC": False,
    "CUDNN_BENCHMARK": True,
    "DATA": {
        "ROOT": "datasets/coco",
        "TOKENIZER_MODEL": "datasets/vocab/coco_10k.model",
        "VOCAB_SIZE": 10000, "UNK_INDEX": 0, "SOS_INDEX": 1, "EOS_INDEX": 2, "MASK_INDEX": 3,
        "IMAGE_CROP_SIZE": 224,
        "MAX_CAPTION_LENGTH": 30,
        "IMAGE_TRANSFORM_TRAIN": ["random_resized_crop", "horizontal_flip", "color_jitter", "normalize"],
        "IMAGE_TRANSFORM_VAL": ["smallest_resize", "center_crop", "normalize"],
        "MASKED_LM": {"MASK_PROPORTION": 0.15, "MASK_PROBABILITY": 0.85, "REPLACE_PROBABILITY": 0.10},
    },
    "MODEL": {
        "NAME": "virtex",
        "VISUAL": {"NAME": "torchvision::resnet50", "FEATURE_SIZE": 2048, "PRETRAINED": False, "FROZEN": False},
        "TEXTUAL": {"NAME": "transdec_postnorm::L1_H2048_A32_F8192", "DROPOUT": 0.1},
        "DECODER": {"NAME": "beam_search", "BEAM_SIZE": 5, "NUCLEUS_SIZE": 0.9, "MAX_DECODING_STEPS": 30},
    },
    "OPTIM": {
        "OPTIMIZER_NAME": "sgd",
        "SGD_MOMENTUM": 0.9,
        "WEIGHT_DECAY": 0.0001,
        "NO_DECAY": ".*textual.(embedding|transformer).*(norm.*|bias)",
        "CLIP_GRAD_NORM": 10.0,
        "LOOKAHEAD": {"USE": True, "ALPHA": 0.5, "STEPS": 5},
        "BATCH_SIZE": 256,
        "CNN_LR": 0.2,
        "LR": 0.001,
        "NUM_ITERATIONS": 500000,
        "WARMUP_STEPS": 10000,
        "LR_DECAY_NAME": "cosine",
        "LR_STEPS": [],
        "LR_GAMMA": 0.1,
    },
}

CONFIG_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class Config:
    """Package-wide configuration: defaults <- YAML file <- override list, then frozen (virtex/config.py:36-219)."""

    def __init__(self, config_file: Optional[str] = None, override_list: List[Any] = []):
        _C = CfgNode(copy.deepcopy(_DEFAULTS))
        if config_file is not None:
            if not os.path.exists(config_file) and os.path.exists(os.path.join(CONFIG_ROOT, config_file)):
                config_file = os.path.join(CONFIG_ROOT, config_file)  # names relative to the shipped configs/
            _C.merge_from_file(config_file)
        _C.merge_from_list(list(override_list))
        _C.freeze()
        self._C = _C

    def dump(self, file_path: str):
        with open(file_path, "w") as f:
            self._C.dump(stream=f)

    def __getattr__(self, attr: str):
        return self._C.__getattr__(attr)

    def __str__(self):
        return self._C.__str__()

    def __repr__(self):
        return self._C.__repr__()
