"""Windowed training meters behind the trainer's log line (`loss: 1.8430 (1.9021)  time: 0.0301 (0.0304)`):
`MetricLogger.update(name=value, ...)`, attribute access to a meter, `str()` = name: window median (run average).
Same surface as the reference's utils/metric_logger.py; values stay Python floats (no tensor round trips)."""
import collections
import statistics


class SmoothedValue(object):
    """Last `window_size` values plus the running sum / count of everything seen."""

    def __init__(self, window_size=20):
        self.window = collections.deque(maxlen=window_size)
        self.total = 0.0
        self.count = 0

    def update(self, value):
        value = float(value)
        self.window.append(value)
        self.total += value
        self.count += 1

    @property
    def median(self):
        # lower middle value for an even count, like the torch.median the reference's meter uses
        return statistics.median_low(self.window) if self.window else 0.0

    @property
    def avg(self):
        return sum(self.window) / len(self.window) if self.window else 0.0

    @property
    def global_avg(self):
        return self.total / self.count if self.count else 0.0


class MetricLogger(object):
    def __init__(self, delimiter="\t"):
        self.delimiter = delimiter
        self.meters = collections.defaultdict(SmoothedValue)

    def update(self, **named_values):
        for name, value in named_values.items():
            if hasattr(value, "item"):      # 0-dim tensor
                value = value.item()
            if not isinstance(value, (int, float)):
                raise TypeError("meter %s: expected a number, got %s" % (name, type(value).__name__))
            self.meters[name].update(value)

    def __getattr__(self, name):
        meters = self.__dict__.get("meters", {})
        if name in meters:
            return meters[name]
        raise AttributeError("%s has no meter or attribute %r" % (type(self).__name__, name))

    def __str__(self):
        return self.delimiter.join("%s: %.4f (%.4f)" % (k, m.median, m.global_avg) for k, m in self.meters.items())
