"""IoU-matrix -> per-prediction ground-truth assignment (reference modeling/matcher.py:6-138, top_k < 2 path)."""
import torch


class Matcher(object):
    BELOW_LOW_THRESHOLD = -1
    BETWEEN_THRESHOLDS = -2

    def __init__(self, high_threshold, low_threshold, allow_low_quality_matches=False, top_k=1):
        assert low_threshold <= high_threshold
        if top_k >= 2:
            raise NotImplementedError("K_HEAD > 1 (top-k matching) is not on the MMT-PSM path")
        self.high_threshold, self.low_threshold = high_threshold, low_threshold
        self.allow_low_quality_matches = allow_low_quality_matches

    def match(self, target, prediction):
        """`self(boxlist_iou(target, prediction))` in one fused launch pair (libmmtpsm_b200 `mmt_match_boxes_f32`):
        the [G, A] IoU matrix — 40 x 159 882 per image at RPN scale, ~10 elementwise passes in the reference — is
        never materialised.  Same values: per pair the same fp32 IoU expression, first maximum on ties."""
        if target.size != prediction.size:
            raise RuntimeError("boxlists should have same image size, got %s, %s" % (target, prediction))
        if len(target) == 0 or len(prediction) == 0:
            which = "ground-truth" if len(target) == 0 else "proposal"
            raise ValueError("No %s boxes available for one of the images during training" % which)
        from .. import kernels as K
        return K.match_boxes(target.bbox, prediction.bbox, self.high_threshold, self.low_threshold,
                             self.allow_low_quality_matches)

    def __call__(self, quality):
        if quality.numel() == 0:
            which = "ground-truth" if quality.shape[0] == 0 else "proposal"
            raise ValueError("No %s boxes available for one of the images during training" % which)
        vals, matches = quality.max(dim=0)
        all_matches = matches
        # same assignments as the reference's masked writes (matcher.py:56-66), as selects: no index lists, no host sync
        below = vals < self.low_threshold
        between = (vals >= self.low_threshold) & (vals < self.high_threshold)
        matches = torch.where(below, matches.new_full((), Matcher.BELOW_LOW_THRESHOLD), matches)
        matches = torch.where(between, matches.new_full((), Matcher.BETWEEN_THRESHOLDS), matches)
        if self.allow_low_quality_matches:
            # every prediction that attains some ground truth's best IoU gets its arg-max match back (matcher.py:68-138)
            best_per_gt, _ = quality.max(dim=1)
            restore = (quality == best_per_gt[:, None]).any(dim=0)
            matches = torch.where(restore, all_matches, matches)
        return matches
