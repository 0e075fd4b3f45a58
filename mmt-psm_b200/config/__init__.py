"""cfg surface of the hot path — mirrors `maskrcnn_benchmark.config.cfg` (reference
config/defaults.py:21-411; yacs is not available in this image, so a compatible CfgNode ships here).

    from mmt_psm_b200.config import cfg
    cfg.merge_from_file("configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml"); cfg.merge_from_list([...]); cfg.freeze()
"""
import ast
import copy
import os

import yaml


class CfgNode(dict):
    """yacs-compatible subset: attribute access, merge_from_file / merge_from_list / merge_from_other_cfg,
    freeze / defrost, clone; python-literal decoding of string values ("(4, 8, 16)", "True", "'bce'");
    list<->tuple and int->float coercion against the default's type; unknown keys are errors."""

    _FROZEN = "__frozen__"

    def __init__(self, init=None):
        super().__init__()
        self.__dict__[CfgNode._FROZEN] = False
        for k, v in (init or {}).items():
            if isinstance(v, dict) and not isinstance(v, CfgNode):
                v = tuple(v["__tuple__"]) if set(v) == {"__tuple__"} else CfgNode(v)
            dict.__setitem__(self, k, v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__[CfgNode._FROZEN]:
            raise AttributeError("cfg is frozen; cannot set %s" % name)
        self[name] = value

    def _freeze(self, flag):
        self.__dict__[CfgNode._FROZEN] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._freeze(flag)

    def freeze(self):
        self._freeze(True)

    def defrost(self):
        self._freeze(False)

    def is_frozen(self):
        return self.__dict__[CfgNode._FROZEN]

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__[CfgNode._FROZEN] = self.is_frozen()
        return out

    @staticmethod
    def _decode(v):
        if isinstance(v, str):
            try:
                return ast.literal_eval(v)
            except (ValueError, SyntaxError):
                return v
        return v

    @staticmethod
    def _coerce(new, old, key):
        if old is None or new is None or type(new) is type(old) or isinstance(old, str):
            return new
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            return float(new)
        raise ValueError("cfg type mismatch for %s: %s vs default %s" % (key, type(new).__name__, type(old).__name__))

    def _merge_dict(self, other, path):
        for k, v in other.items():
            full = ".".join(path + [k])
            if k not in self:
                raise KeyError("Non-existent config key: " + full)
            if isinstance(v, dict) and isinstance(self[k], CfgNode):
                self[k]._merge_dict(v, path + [k])
            else:
                dict.__setitem__(self, k, self._coerce(self._decode(v), self[k], full))

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge_dict(yaml.safe_load(f) or {}, [])

    def merge_from_other_cfg(self, other):
        self._merge_dict(other, [])

    def merge_from_list(self, kv):
        if len(kv) % 2:
            raise ValueError("merge_from_list needs KEY VAL pairs")
        for full, v in zip(kv[0::2], kv[1::2]):
            node, parts = self, full.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError("Non-existent config key: " + full)
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: " + full)
            dict.__setitem__(node, parts[-1], self._coerce(self._decode(v), node[parts[-1]], full))


def _load_defaults():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "defaults.yaml")) as f:
        return CfgNode(yaml.safe_load(f))


_C = _load_defaults()
cfg = _C


def get_cfg_defaults():
    return _load_defaults()


# `scripts/train_mt.sh` overrides (reference scripts/train_mt.sh:4-20) — the configuration the headline
# metric is quoted on, as a KEY VAL list for merge_from_list.
TRAIN_MT_OVERRIDES = [
    "MODEL.RELATION_NMS.WEIGHT", 1., "MODEL.RELATION_NMS.ALPHA", 0., "MODEL.RELATION_NMS.LOSS", 1.,
    "MODEL.RELATION_NMS.D_LOSS", 0., "MODEL.RELATION_NMS.GAMMA", 1., "MODEL.RELATION_NMS.REG_IOU", True,
    "MODEL.RELATION_NMS.REG_IOU_MSK", False,
    "MT.ALPHA", 0.99, "MT.ALPHA_RAMPUP", 0.99, "MT.LAMBDA", 5., "MT.RAMPUP_STEP", 250, "MT.RAMPDOWN_STEP", 250,
    "MT.TEMP", 0.5, "MT.SHARPEN", True, "MT.START_MT", 1000, "MT.AUG_S", 1, "MT.CLS_LOSS_TYPE", "bce",
    "SOLVER.IMS_PER_BATCH", 4, "MT.CLS_LOSS", 0.2, "MT.MEAN_TARGET", True, "MT.FLIP", True, "MT.CONF_PAIR", True,
    "MODEL.META_ARCHITECTURE", "GeneralizedRCNN", "MODEL.ROI_BOX_HEAD.DO", 0.5, "MODEL.RELATION_NMS.DO", 0.5,
    "SOLVER.CHECKPOINT_PERIOD", 50, "MT.AUG", True, "MT.CLS_NEG", True, "MT.HARD_NEG", True,
    "MT.CLS_BALANCE_WEIGHT", 1.5, "MT.RANK_FILTER", 0.2, "MT.FG_HINT", 1., "MT.T_ADAPT", True,
    "DATASETS.NO_LABEL", True, "DATASETS.GEN_TRUE", 0, "DATASETS.TUNE_SOURCE", False, "SOLVER.BASE_LR", 0.005,
]
