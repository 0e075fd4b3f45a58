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

    def __call__(self, quality):
        if quality.numel() == 0:
            which = "ground-truth" if quality.shape[0] == 0 else "proposal"
            raise ValueError("No %s boxes available for one of the images during training" % which)
        vals, matches = quality.max(dim=0)
        all_matches = matches.clone() if self.allow_low_quality_matches else None
        below = vals < self.low_threshold
        between = (vals >= self.low_threshold) & (vals < self.high_threshold)
        matches[below] = Matcher.BELOW_LOW_THRESHOLD
        matches[between] = Matcher.BETWEEN_THRESHOLDS
        if self.allow_low_quality_matches:
            best_per_gt, _ = quality.max(dim=1)
            pairs = torch.nonzero(quality == best_per_gt[:, None])
            upd = pairs[:, 1]
            matches[upd] = all_matches[upd]
        return matches
