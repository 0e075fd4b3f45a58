"""Checkpoint / resume for the mean-teacher trainer — the surface of reference utils/checkpoint.py:13-160 that
`MTtrainer.save_model` and `tools/train_mean_teacher.py:36-43` use (`DetectronCheckpointer(cfg, model, optimizer, scheduler,
save_dir, save_to_disk)`, `.save(name, **extra)`, `.load(f)`, `.has_checkpoint()`, the `last_checkpoint` tag file).

Weights are written in the REFERENCE's layout and key names (OIHW convs, IOHW deconv, fc6 columns in (c, h, w) order:
modeling/detector.py `reference_state_dict`), so a file saved here loads into the reference and a reference `.pth`
(e.g. e2e_mask_rcnn_R_50_FPN_1x.pth) loads here; `module.` prefixes of DataParallel checkpoints are stripped, keys the
model does not build (the IR-Net modules when switched off in cfg) are reported, not fatal.  Optimizer state = the flat momentum arena."""
import logging
import os

import torch

from .model_serialization import _unwrap, load_state_dict


class Checkpointer(object):
    def __init__(self, model, optimizer=None, scheduler=None, save_dir="", save_to_disk=None, logger=None):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.save_dir, self.save_to_disk = save_dir, save_to_disk
        self.logger = logger if logger is not None else logging.getLogger(__name__)

    def save(self, name, **kwargs):
        if not self.save_dir or not self.save_to_disk:
            return None
        data = {"model": {k: v.cpu() for k, v in self._model_state().items()}}
        if self.optimizer is not None:
            data["optimizer"] = {k: (v.cpu() if isinstance(v, torch.Tensor) else v)
                                 for k, v in self.optimizer.state_dict().items()}
        if self.scheduler is not None:
            data["scheduler"] = self.scheduler.state_dict()
        data.update(kwargs)
        os.makedirs(self.save_dir, exist_ok=True)
        save_file = os.path.join(self.save_dir, "{}.pth".format(name))
        self.logger.info("Saving checkpoint to %s", save_file)
        torch.save(data, save_file)
        self.tag_last_checkpoint(save_file)
        return save_file

    def _model_state(self):
        """The detector in the reference's layout and key names; any other module as its own state dict (wrapper
        prefixes included, like the reference — the loader ignores them)."""
        if hasattr(_unwrap(self.model), "box_heads"):
            from ..modeling.detector import reference_state_dict
            return reference_state_dict(_unwrap(self.model))
        return self.model.state_dict()

    # the COCO Mask R-CNN file the reference starts from: loaded as TRANSFER LEARNING (reference utils/checkpoint.py:79-117,
    # 148-160) — the class-count-dependent predictors are dropped, optimizer / scheduler state is not taken, the
    # iteration counter restarts, and an existing `last_checkpoint` does not override it
    TRANSFER_FILE = "e2e_mask_rcnn_R_50_FPN_1x.pth"
    TRANSFER_DROPPED = ("cls_score", "bbox_pred", "mask_fcn_logits")

    def load(self, f=None, test=False):
        """Resumes from `last_checkpoint` if the run directory has one (like the reference), else from `f`; with
        neither, the model keeps its initialisation.  Returns the extra data stored next to the weights."""
        transfer = bool(f) and not test and self.TRANSFER_FILE in f
        if self.has_checkpoint() and not transfer:
            f = self.get_checkpoint_file()
        if not f:
            self.logger.info("No checkpoint found. Initializing model from scratch")
            return {}
        self.logger.info("Loading checkpoint from %s", f)
        checkpoint = torch.load(f, map_location="cpu", weights_only=False)
        sd = checkpoint.pop("model", checkpoint)
        if transfer:
            sd = {k: v for k, v in sd.items() if not any(name in k for name in self.TRANSFER_DROPPED)}
            checkpoint.pop("optimizer", None)
            checkpoint.pop("scheduler", None)
            checkpoint["iteration"] = -1
        missing, unexpected = load_state_dict(self.model, sd)     # longest-suffix key matching, `module.` ignored
        if missing:
            self.logger.warning("keys not in the checkpoint (kept at their initialisation): %s", missing[:8])
        if unexpected:
            self.logger.info("checkpoint keys this path does not build: %s", unexpected[:8])
        if not test:
            if "optimizer" in checkpoint and self.optimizer is not None:
                self.optimizer.load_state_dict(checkpoint.pop("optimizer"))
            if "scheduler" in checkpoint and self.scheduler is not None:
                self.scheduler.load_state_dict(checkpoint.pop("scheduler"))
        return checkpoint

    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self):
        try:
            with open(os.path.join(self.save_dir, "last_checkpoint")) as f:
                return f.read().strip()
        except IOError:
            return ""

    def tag_last_checkpoint(self, last_filename):
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as f:
            f.write(last_filename)


class DetectronCheckpointer(Checkpointer):
    """Reference signature (utils/checkpoint.py:162-176): the cfg comes first.  Caffe2 / model-zoo URL loading is not
    part of this path (no network)."""

    def __init__(self, cfg, model, optimizer=None, scheduler=None, save_dir="", save_to_disk=None, logger=None):
        super().__init__(model, optimizer, scheduler, save_dir, save_to_disk, logger)
        self.cfg = cfg.clone()
