"""CUDA-graph capture of the static-shape trunk of the iteration (backbone + FPN [+ RPN head / MGD adaptors]).

One MT iteration issues ~4000 kernel launches, ~1500 of them from the conv trunks whose shapes never change; driven
from Python each costs 5-30 us of host time and the iteration is host-bound (profiles/r01_profile_step_v1.log:
enqueue 66 ms vs 53 ms of device work).  `GraphedCallable` records such a trunk once — forward, and optionally the
backward through it — and replays it with two `cudaGraphLaunch` calls.

Training variant (`backward=True`), same scheme as `torch.cuda.make_graphed_callables` but accumulating parameter
gradients IN PLACE: the parameters' `.grad` are persistent views into the optimizer's flat gradient arena
(solver.FusedSGDEMA), so the captured AccumulateGrad nodes are `grad_view += dW` kernels inside the backward graph
and no gradient leaves the graph except d(inputs) (none here: the image needs no gradient).

Constraints honoured by the callers: fixed input shapes per instance (one instance per shape and per call site —
an instance's activations are overwritten by its next replay, so the supervised and the unlabeled student passes
own separate instances), no host synchronisation inside `fn`, parameters updated in place only.
"""
import torch

from .. import lib as _l


class GraphedCallable(object):
    def __init__(self, fn, example_inputs, backward=False, on_backward=None, warmup=2, pool=None):
        """fn(*tensors) -> tuple of tensors.  `on_backward(touched_params)` is called after every backward replay
        with the parameters whose .grad the captured backward accumulates into."""
        assert all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs)
        self.fn, self.backward, self.on_backward = fn, backward, on_backward
        self.static_in = [t.detach().clone() for t in example_inputs]
        self.shapes = [tuple(t.shape) for t in example_inputs]
        self.pool = pool if pool is not None else torch.cuda.graph_pool_handle()
        self.touched = []
        dev = self.static_in[0].device

        # warm-up on a side stream (lazy library state, allocator), like torch's own graphed callables
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                with torch.set_grad_enabled(backward):
                    outs = self._as_tuple(fn(*self.static_in))
                if backward:
                    live = [o for o in outs if o.requires_grad]
                    torch.autograd.backward(live, [torch.zeros_like(o) for o in live])
                del outs
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

        n0 = _l.launch_count()
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd, pool=self.pool):
            with torch.set_grad_enabled(backward):
                outs = self._as_tuple(fn(*self.static_in))
        self.fwd_launches = _l.launch_count() - n0
        self.static_out = outs
        self.g_bwd, self.static_gout, self.bwd_launches = None, None, 0
        if backward:
            self.grad_idx = [i for i, o in enumerate(outs) if o.requires_grad]
            live = [outs[i] for i in self.grad_idx]
            self.static_gout = [torch.zeros_like(o) for o in live]
            seen, hooks = [], []
            params = self._leaf_params(live)
            for p in params:
                hooks.append(p.register_post_accumulate_grad_hook(lambda q, seen=seen: seen.append(q)))
            # weight gradients written straight into the arena (functional._fused_grad_sink) bypass AccumulateGrad:
            # record those parameters too, they must be reported on every replay
            from .. import functional as Fk
            orig_sink = Fk._fused_grad_sink
            if orig_sink is not None:
                Fk.set_fused_grad_sink(lambda ps: (seen.extend(ps), orig_sink(ps)))
            n0 = _l.launch_count()
            self.g_bwd = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(self.g_bwd, pool=self.pool):
                    torch.autograd.backward(live, self.static_gout)
            finally:
                Fk.set_fused_grad_sink(orig_sink)
            self.bwd_launches = _l.launch_count() - n0
            for h in hooks:
                h.remove()
            self.touched = list(seen)
            self.static_out = tuple(o.detach() for o in outs)
            self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        torch.cuda.synchronize(dev)

    @staticmethod
    def _as_tuple(o):
        return tuple(o) if isinstance(o, (tuple, list)) else (o,)

    @staticmethod
    def _leaf_params(outs):
        """Leaf tensors with requires_grad reachable from `outs` (the AccumulateGrad nodes of the autograd graph)."""
        seen, leaves, stack = set(), [], [o.grad_fn for o in outs if o.grad_fn is not None]
        while stack:
            fn = stack.pop()
            if fn is None or fn in seen:
                continue
            seen.add(fn)
            if hasattr(fn, "variable"):
                leaves.append(fn.variable)
            for nxt, _ in fn.next_functions:
                if nxt is not None:
                    stack.append(nxt)
        return leaves

    def matches(self, inputs):
        return len(inputs) == len(self.shapes) and all(tuple(t.shape) == s for t, s in zip(inputs, self.shapes))

    def __call__(self, *inputs, prefetched=False):
        """Replay on the current stream, or — `prefetched` — only wait for the replay `launch_forward` enqueued."""
        if not self.matches(inputs):
            raise ValueError("GraphedCallable captured for input shapes %s, got %s" % (self.shapes, [tuple(t.shape) for t in inputs]))
        self._skip_replay = bool(prefetched)
        if self.backward and torch.is_grad_enabled():
            return _Replay.apply(self, self._anchor, *inputs)
        return self._replay_forward(inputs)

    def launch_forward(self, inputs, stream):
        """Enqueue input copy + forward replay on `stream` (which the caller has ordered after the producers of
        `inputs` and after every consumer of the previous replay's outputs) and record the completion event."""
        with torch.cuda.stream(stream):
            for s, x in zip(self.static_in, inputs):
                s.copy_(x, non_blocking=True)
            self.g_fwd.replay()
            self._done = torch.cuda.Event()
            self._done.record(stream)
        _l.add_launch_count(self.fwd_launches)

    def _replay_forward(self, inputs):
        if getattr(self, "_skip_replay", False):
            torch.cuda.current_stream().wait_event(self._done)
            self._skip_replay = False
            return self.static_out
        for s, x in zip(self.static_in, inputs):
            if s.data_ptr() != x.data_ptr():
                s.copy_(x, non_blocking=True)
        self.g_fwd.replay()
        _l.add_launch_count(self.fwd_launches)
        return self.static_out

    # ---- direct (autograd-free) use: the caller owns the boundary -------------------------------------------
    def forward_detached(self, inputs, prefetched=False):
        """Replay (or pick up the prefetched replay) and return the static output buffers as plain tensors: no autograd
        node is created; the caller feeds `grad_buffer`s and calls `backward_direct` itself."""
        self._skip_replay = bool(prefetched)
        return self._replay_forward(inputs)

    def grad_buffer(self, out_index):
        """The static gradient buffer of output `out_index` (None if that output does not require grad)."""
        return self.static_gout[self.grad_idx.index(out_index)] if out_index in self.grad_idx else None

    def backward_direct(self):
        """Replay the captured backward over whatever the grad buffers hold."""
        self.g_bwd.replay()
        _l.add_launch_count(self.bwd_launches)
        if self.on_backward is not None:
            self.on_backward(self.touched)

    def _replay_backward(self, gouts):
        for s, g in zip(self.static_gout, gouts):
            if g is None:
                s.zero_()
            else:
                s.copy_(g, non_blocking=True)
        self.g_bwd.replay()
        _l.add_launch_count(self.bwd_launches)
        if self.on_backward is not None:
            self.on_backward(self.touched)


class _Replay(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gc, anchor, *inputs):
        ctx.gc = gc
        outs = gc._replay_forward(inputs)
        ctx.n_in = len(inputs)
        # the static buffers are handed out as they are: valid until this instance's next replay
        ret = tuple(o.view_as(o) for o in outs)
        ctx.mark_non_differentiable(*[o for i, o in enumerate(ret) if i not in gc.grad_idx])
        return ret

    @staticmethod
    def backward(ctx, *gouts):
        gc = ctx.gc
        gc._replay_backward([gouts[i] for i in gc.grad_idx])
        return (None, None) + (None,) * ctx.n_in
