"""Ramp schedules (reference utils/miscellaneous.py:233-247)."""
import math


def sigmoid_rampup(current, rampup_length):
    if rampup_length == 0:
        return 1.0
    current = min(max(current, 0.0), rampup_length)
    phase = 1.0 - current / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def sigmoid_rampdown(gap_time, rampdown_length):
    if rampdown_length == 0:
        return 1.0
    phase = 1.0 - gap_time / rampdown_length
    return float(math.exp(-12 * phase * phase))
