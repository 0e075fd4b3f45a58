"""Optimizer + LR schedule — mirrors `maskrcnn_benchmark.solver` (reference solver/build.py:5-33,
lr_scheduler.py:10-53).  `make_optimizer` returns a FusedSGDEMA: the student's parameters, gradients and momentum
(and the EMA teacher's parameters) live in flat fp32 arenas, and ONE kernel launch per iteration performs
SGD(momentum, per-tensor lr / weight decay) followed by the teacher EMA — instead of torch.optim.SGD over 130 param
groups plus 264 mul_/add_ launches (engine/MTtrainer.py:277-281)."""
from bisect import bisect_right

import torch

from .. import kernels as K


class FusedSGDEMA(object):
    def __init__(self, cfg, student, teacher=None):
        s = cfg.SOLVER
        self.momentum = s.MOMENTUM
        self.student, self.teacher = student, teacher
        params = list(student.named_parameters())
        device = params[0][1].device
        self.entries, total = [], 0
        for name, p in params:
            n = p.numel()
            lr_mult, wd = (s.BIAS_LR_FACTOR, s.WEIGHT_DECAY_BIAS) if "bias" in name else (1.0, s.WEIGHT_DECAY)
            self.entries.append(dict(name=name, begin=total, end=total + n, lr_mult=lr_mult, wd=wd,
                                     trainable=p.requires_grad))
            total += (n + 3) // 4 * 4
        self.total = total
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=device)
        # gradient arena + one "this tensor received a gradient" flag per tensor behind it: data-parallel ranks all-reduce
        # BOTH in the one exchange of the path (`reduce_buffer`), so the set of tensors SGD updates is the same on every
        # rank even when one rank's iteration skipped a branch (e.g. its teacher produced no boxes)
        self._g_full = torch.zeros(total + (len(params) + 3) // 4 * 4, dtype=torch.float32, device=device)
        self.flat_g = self._g_full[:total]
        self._flags = self._g_full[total:total + len(params)]
        self._agreed = False
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=device)
        self.touched = [False] * len(params)
        self._index = {id(p): i for i, (_, p) in enumerate(params)}
        for i, ((name, p), e) in enumerate(zip(params, self.entries)):
            view = self.flat_p[e["begin"]:e["end"]].view_as(p)
            view.copy_(p.data)
            p.data = view
            if p.requires_grad:
                p.grad = self.flat_g[e["begin"]:e["end"]].view_as(p)
                p.register_post_accumulate_grad_hook(self._mark(i))
        self.flat_t = None
        if teacher is not None:
            tparams = list(teacher.named_parameters())
            assert [n for n, _ in tparams] == [n for n, _ in params], "teacher/student parameter lists differ"
            self.flat_t = torch.zeros(total, dtype=torch.float32, device=device)
            for (name, p), e in zip(tparams, self.entries):
                view = self.flat_t[e["begin"]:e["end"]].view_as(p)
                view.copy_(p.data)
                p.data = view
        from .. import functional as Fk
        Fk.set_fused_grad_sink(self.mark_touched)   # bias-free tcgen05 convs write dW into flat_g directly
        self.base_lr = s.BASE_LR
        self.lr = s.BASE_LR
        self.param_groups = [{"lr": self.lr}]  # for the reference's log line (MTtrainer.py:218)
        self.first_step = True
        self._seg_cache = {}

    def _mark(self, i):
        def hook(_):
            self.touched[i] = True
        return hook

    def mark_touched(self, params):
        """Parameters whose gradient was accumulated by a replayed CUDA graph (no autograd hook fires on replay)."""
        for p in params:
            i = self._index.get(id(p))
            if i is not None:
                self.touched[i] = True

    def set_lr(self, lr):
        self.lr = lr
        self.param_groups[0]["lr"] = lr

    def zero_grad(self):
        self._g_full.zero_()
        self._agreed = False
        self.touched = [False] * len(self.entries)

    def _write_flags(self):
        """This rank's touched flags -> the slot behind the gradient arena.  On CUDA through pinned staging: a copy from
        pageable memory synchronises the stream first — here that is the trunk's backward graph, enqueued a moment ago, and
        the host would lose its whole lead over the device at every iteration of a data-parallel run."""
        host = torch.tensor(self.touched, dtype=torch.float32)
        if self._flags.is_cuda:
            host = host.pin_memory()
        self._flags.copy_(host, non_blocking=True)
        self._agreed = True

    def reduce_buffer(self):
        """The tensor a data-parallel trainer all-reduces (SUM): the gradient arena followed by this rank's per-tensor
        touched flags.  After the reduction `step` updates every tensor ANY rank accumulated a gradient into."""
        self._write_flags()
        return self._g_full

    def _segments(self, touched):
        key = (self.lr, tuple(touched))
        if key not in self._seg_cache:
            if len(self._seg_cache) > 64:
                self._seg_cache.clear()
            ent = [(e["begin"], e["end"], self.lr * e["lr_mult"], e["wd"], e["trainable"] and t)
                   for e, t in zip(self.entries, touched)]
            self._seg_cache[key] = K.make_segments(ent, self.flat_p.device)
        return self._seg_cache[key]

    def reduce_buckets(self, head_prefixes=("box_heads.", "mask_heads.", "relation_nms.")):
        """(early, late): two lists of contiguous views of the all-reduce buffer.  `early` = the gradient ranges of the
        parameters whose names start with `head_prefixes` (final once the heads' backward has run: they can be reduced
        UNDER the trunk's backward); `late` = everything else — the trunk's ranges and the touched flags.  Together they
        tile the whole buffer, so reducing both is reducing `reduce_buffer()`."""
        cuts, on = [], False
        for e in self.entries:
            head = e["name"].startswith(tuple(head_prefixes))
            if head != on:
                cuts.append(e["begin"])
                on = head
        if on:
            cuts.append(self.total)
        bounds = [0] + cuts + [self._g_full.numel()]
        early = [self._g_full[bounds[i]:bounds[i + 1]] for i in range(1, len(bounds) - 1, 2)]
        late = [self._g_full[bounds[i]:bounds[i + 1]] for i in range(0, len(bounds) - 1, 2)]
        return [t for t in early if t.numel()], [t for t in late if t.numel()]

    def publish_touched(self):
        """Write this rank's touched flags behind the gradient arena (they ride in the LAST bucket's all-reduce)."""
        self._write_flags()

    def _agreed_segments(self):
        """Segment table whose update mask comes from the all-reduced flags: built on the device, no read-back."""
        segs = self._segments([True] * len(self.entries)).clone()
        segs.view(torch.int32).view(-1, 8)[:, 6] *= (self._flags > 0).to(torch.int32)   # mmt_param_segment.trainable
        return segs

    def step(self, ema_alpha=None, grad_scale=1.0):
        """SGD on every tensor that received a gradient this iteration (torch.optim.SGD skips grad-less params),
        then, if `ema_alpha` is given, teacher <- alpha * teacher + (1 - alpha) * student over ALL parameters."""
        segs = self._agreed_segments() if self._agreed else self._segments(self.touched)
        teacher = self.flat_t if ema_alpha is not None else None
        K.sgd_ema_(self.flat_p, teacher, self.flat_g, self.flat_m, segs, len(self.entries), self.momentum,
                   0.0 if ema_alpha is None else ema_alpha, self.first_step, grad_scale)
        self.first_step = False

    def ema_only(self, alpha):
        K.ema_(self.flat_t, self.flat_p, alpha)

    def state_dict(self):
        return {"momentum": self.flat_m.clone(), "first_step": self.first_step, "lr": self.lr}

    def load_state_dict(self, sd):
        self.flat_m.copy_(sd["momentum"])
        self.first_step = sd["first_step"]
        self.set_lr(sd["lr"])


def make_optimizer(cfg, model, teacher=None):
    return FusedSGDEMA(cfg, model, teacher)


class WarmupMultiStepLR(object):
    """Same schedule as reference solver/lr_scheduler.py:10-53 (linear / constant warm-up x multi-step decay)."""

    def __init__(self, optimizer, milestones, gamma=0.1, warmup_factor=1.0 / 3, warmup_iters=500,
                 warmup_method="linear", last_epoch=-1):
        if list(milestones) != sorted(milestones):
            raise ValueError("Milestones should be a list of increasing integers. Got %s" % (milestones,))
        if warmup_method not in ("constant", "linear"):
            raise ValueError("Only 'constant' or 'linear' warmup_method accepted, got %s" % warmup_method)
        self.optimizer, self.milestones, self.gamma = optimizer, list(milestones), gamma
        self.warmup_factor, self.warmup_iters, self.warmup_method = warmup_factor, warmup_iters, warmup_method
        self.base_lr = optimizer.base_lr
        self.last_epoch = last_epoch
        self.step()

    def get_lr(self):
        f = 1
        if self.last_epoch < self.warmup_iters:
            if self.warmup_method == "constant":
                f = self.warmup_factor
            else:
                a = self.last_epoch / self.warmup_iters
                f = self.warmup_factor * (1 - a) + a
        return self.base_lr * f * self.gamma ** bisect_right(self.milestones, self.last_epoch)

    def step(self):
        self.last_epoch += 1
        self.optimizer.set_lr(self.get_lr())

    def state_dict(self):
        return {"last_epoch": self.last_epoch}

    def load_state_dict(self, sd):
        self.last_epoch = sd["last_epoch"]
        self.optimizer.set_lr(self.get_lr())


def make_lr_scheduler(cfg, optimizer):
    s = cfg.SOLVER
    return WarmupMultiStepLR(optimizer, s.STEPS, s.GAMMA, warmup_factor=s.WARMUP_FACTOR, warmup_iters=s.WARMUP_ITERS,
                             warmup_method=s.WARMUP_METHOD)
