"""Loading a state dict whose key names differ from the model's by a prefix (the reference's
utils/model_serialization.py:9-80 contract, pinned by its tests/checkpoint.py): every model entry takes the checkpoint
entry whose name is the LONGEST suffix of its own name (`block1.layer1.weight` <- `layer1.weight`; `res.layer2.weight`
prefers `res.layer2.weight` over `layer2.weight`), `module.` prefixes of DataParallel wrappers are ignored on either
side, unmatched model entries keep their values.  For this package's GeneralizedRCNN the tensors additionally cross the
layout boundary (reference OIHW / IOHW / fc6 column order -> kernel layouts, modeling/detector.py)."""
import logging
from collections import OrderedDict

import torch


def strip_prefix_if_present(state_dict, prefix):
    keys = sorted(state_dict.keys())
    if not keys or not all(k.startswith(prefix) for k in keys):
        return state_dict
    return OrderedDict((k[len(prefix):], v) for k, v in state_dict.items())


def align_state_dict(model_keys, loaded):
    """{model key: checkpoint tensor} by longest-suffix match; model keys without a match are absent."""
    out = OrderedDict()
    names = list(loaded.keys())
    log = logging.getLogger(__name__)
    for mk in model_keys:
        best = ""
        for lk in names:
            if len(lk) > len(best) and (mk == lk or mk.endswith("." + lk) or mk.endswith(lk)):
                best = lk
        if best:
            out[mk] = loaded[best]
            if best != mk:
                log.debug("%s loaded from %s", mk, best)
    return out


def _unwrap(model):
    return model.module if isinstance(model, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else model


def load_state_dict(model, loaded_state_dict):
    """Returns (missing model keys, unused checkpoint keys)."""
    core = _unwrap(model)
    loaded = strip_prefix_if_present(loaded_state_dict, "module.")
    if hasattr(core, "box_heads"):                      # this package's detector: reference layout -> kernel layouts
        from ..modeling.detector import load_reference_state_dict, reference_state_dict
        aligned = align_state_dict(list(reference_state_dict(core).keys()), loaded)
        missing, _ = load_reference_state_dict(core, aligned)
    else:
        own = core.state_dict()
        aligned = align_state_dict(list(own.keys()), loaded)
        with torch.no_grad():
            for k, v in aligned.items():
                own[k].copy_(v)
        missing = [k for k in own if k not in aligned]
    used = {id(v) for v in aligned.values()}
    unused = [k for k, v in loaded.items() if id(v) not in used]
    return missing, unused
