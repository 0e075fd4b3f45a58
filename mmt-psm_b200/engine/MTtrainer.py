"""Mean-teacher training loop — mirrors reference engine/MTtrainer.py:16-284 (`MTtrainer.train`,
`forward_source`, `forward_unlabel`, `update_teacher`, `weight_sum_losses`, `reduce_loss_dict`).

One iteration = supervised student forward -> (no-grad) teacher forward on K colour views x {orig, flip} ->
student forward on the unlabeled view (PSM + MGD) -> weighted sum -> backward -> [NCCL all-reduce of the flat
student gradient arena] -> ONE fused SGD + teacher-EMA launch.  Data parallelism is the north star's addition
(the reference never wraps the student): every rank holds a student+teacher replica, gradients are
summed with a single `all_reduce` over the flat fp32 arena and scaled by 1/world inside the optimizer kernel;
teachers stay bit-identical across ranks without any traffic because EMA is a function of the synced student."""
import datetime
import gc
import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor

import torch
import torch.distributed as dist

from ..modeling.balanced_positive_negative_sampler import BalancedPositiveNegativeSampler as _Sampler
from ..utils.comm import get_world_size
from ..utils.miscellaneous import sigmoid_rampdown, sigmoid_rampup


def reduce_loss_dict(loss_dict, keys=None):
    """Logging-only reduction to rank 0 (reference MTtrainer.py:16-41).  `keys`: the fixed set of terms every rank
    contributes to the collective — a rank whose iteration skipped a branch (teacher without boxes: no mt_* terms)
    sends zeros for the terms it lacks instead of a shorter tensor."""
    world = get_world_size()
    if world < 2:
        return loss_dict
    with torch.no_grad():
        names = sorted(keys) if keys is not None else sorted(loss_dict.keys())
        ref = next(iter(loss_dict.values()))
        vals = torch.stack([loss_dict[k].reshape(()).to(ref.dtype) if k in loss_dict else ref.new_zeros(()) for k in names], dim=0)
        dist.reduce(vals, dst=0)
        if dist.get_rank() == 0:
            vals /= world
        return {k: v for k, v in zip(names, vals)}


def weight_sum_losses(loss_dict, step, rampup_length, rampdown_length, total_length, l=1, balanced=None, start_mt=1000):
    """Consistency-loss ramp and per-term weights (reference MTtrainer.py:67-109; note the ramp-DOWN also uses
    `rampup_length` as its length, :92)."""
    if 0 < (step - start_mt) < rampup_length:
        weight = l * sigmoid_rampup(step - start_mt, rampup_length)
    elif (total_length - step) < rampdown_length:
        weight = l * sigmoid_rampdown(total_length - step, rampup_length)
    else:
        weight = l
    out = {k: (weight * v if "mt" in k else v) for k, v in loss_dict.items()}
    for k in list(out):
        if balanced is not None and k in balanced:
            out[k] = out[k] * balanced[k]
    return out


class MTtrainer(object):
    def __init__(self, model_s, model_t, data_loader, optimizer, scheduler, ckpt_s, ckpt_t, checkpoint_period, cfg,
                 cuda_graphs=False):
        self.cfg = cfg
        self.logger = logging.getLogger("maskrcnn_benchmark.trainer")
        self.scheduler = scheduler
        self.student, self.teacher = model_s, model_t
        self.student_bs, self.teacher_bs = cfg.MT.AUG_S, cfg.MT.AUG_K
        self.device = next(model_s.parameters()).device
        self.checkpoint_period = checkpoint_period
        self.ckpt_s, self.ckpt_t = ckpt_s, ckpt_t
        self.optimizer = optimizer
        self.lambda_value, self.alpha = cfg.MT.LAMBDA, cfg.MT.ALPHA
        self.rampup_step, self.rampdown_step, self.start_mt = cfg.MT.RAMPUP_STEP, cfg.MT.RAMPDOWN_STEP, cfg.MT.START_MT
        self.balanced_weight = {"mt_classifier": cfg.MT.CLS_LOSS, "nms_loss": cfg.MODEL.RELATION_NMS.LOSS,
                                "mt_fg_loss": cfg.MT.FG_HINT}
        self.dataloader_s = data_loader["source"] if data_loader else None
        self.dataloader_u = data_loader.get("no_label") if (data_loader and cfg.DATASETS.NO_LABEL) else None
        self._iter_u = None
        self.max_iter = len(self.dataloader_s) if self.dataloader_s is not None else cfg.SOLVER.MAX_ITER
        self.start_iter = 0
        self.n_step_unlabel = cfg.MT.N_STEP_UNLABEL
        self.world = get_world_size()
        self.last_losses = None
        # Python's generational GC pauses (100+ ms when a full collection walks the autograd / BoxList object graphs)
        # land at random inside iterations that are ~30 ms long: pace it instead — automatic collection off, a young-
        # generation collection every `gc_every` iterations (0 = leave the interpreter's GC alone)
        self.gc_every, self._gc_steps = 50, 0
        # Data-dependent ROI counts keep producing block sizes the caching allocator has not seen: each miss is a
        # cudaMalloc, which blocks the host for 15-100 ms while the device is busy (profiles/r02_stall_trace.md: every
        # isolated 40-235 ms step of round 1 coincides with one).  Expandable segments grow by mapping 2 MB pages
        # (cuMemMap) instead — same trace, no stall.
        if self.device.type == "cuda":
            try:
                torch.cuda.memory._set_allocator_settings("expandable_segments:True")
            except Exception as e:   # older torch: keep going, only the stalls stay
                self.logger.warning("could not switch the CUDA caching allocator to expandable segments: %s", e)
        # Second host thread for the (no-grad, independent) teacher pass: the supervised pass and the teacher pass are each
        # a string of ~300 small launches separated by ~8 device read-backs (data-dependent ROI counts); run one after the
        # other the host sits in cudaStreamSynchronize ~10 ms per iteration with an empty device queue behind every one of
        # them (profiles/r02_timeline.md).  On two threads / two streams one pass enqueues while the other waits.
        # Only with CUDA graphs on (the production / bench configuration); parity tests run the passes in order.
        self.teacher_thread = os.environ.get("MMT_TEACHER_THREAD", "1") == "1"
        # (measured: no gain — 24.6 vs 24.3 ms, profiles/r02_ab_backward_thread.log: the backward's Python-level nodes contend
        # for the GIL with the unlabeled pass — kept as a knob, off by default)
        self.backward_thread = os.environ.get("MMT_BACKWARD_THREAD", "0") == "1"
        self.early_rpn_targets = os.environ.get("MMT_EARLY_RPN_TARGETS", "1") == "1"
        self._bw_pool = None
        # (measured: no effect on the step time, profiles/r02_ab_teacher_first.log — kept as a knob, off by default)
        self.teacher_first = os.environ.get("MMT_TEACHER_FIRST", "0") == "1"
        if self.teacher_first:
            self.teacher.side_priority = -1
        # Head / glue work of both host threads on HIGH-priority streams, and the teacher's trunk graph captured with a few SMs
        # left free (`teacher_trunk_ctas` of 148).  The trunk graphs are chains of persistent conv launches that own every SM
        # for 50-500 us at a time; their kernel nodes keep the capture stream's priority 0.  A small head kernel of the OTHER
        # pass (they run concurrently with the teacher's trunk for most of the heads phase) is then picked first at every
        # kernel boundary, or starts at once on a free SM, instead of queueing behind the next conv wave — and every one of
        # the ~10 read-backs of a pass waits for such a kernel.  Measured together 24.0 -> 23.4 ms per iteration (either one
        # alone: within noise; profiles/r02_ab_head_priority.log).  Only with CUDA graphs on.
        self.head_priority = int(os.environ.get("MMT_HEAD_PRIORITY", "-1"))
        self.teacher_trunk_ctas = int(os.environ.get("MMT_TEACHER_TRUNK_CTAS", "132"))
        self._head_stream = None
        self._loss_slot, self._loss_host, self._loss_event = None, None, None
        self._stage_probe = None     # list: host timestamps of stage()'s phases are appended (bench.py, MMT_BENCH_TRACE=2)
        self._t_stream, self._t_pool = None, None
        self._copy_stream, self._staged, self._pregrow_copy, self._prefetched = None, None, False, None
        self.pool_gb = tuple(float(v) for v in os.environ.get("MMT_POOL_GB", "6,4,0.5").split(","))   # main, teacher, copy stream
        self._pregrown_for, self._pregrow_teacher = 0, False
        self.loss_keys = ["loss_classifier", "loss_box_reg", "loss_seg", "loss_objectness", "loss_rpn_box_reg"]
        if cfg.MODEL.RELATION_NMS.USE_RELATION_NMS:
            self.loss_keys.append("nms_loss")
        if cfg.MT.LAMBDA > 0 and cfg.DATASETS.NO_LABEL:
            self.loss_keys += [k for k, on in (("mt_classifier", cfg.MT.CLS_LOSS), ("mt_fg_loss", cfg.MT.FG_HINT)) if on]
        self.set_cuda_graphs(cuda_graphs)

    def set_cuda_graphs(self, on):
        """Replay the fixed-shape conv trunks of both networks from CUDA graphs (fixed-size crops only)."""
        self.cuda_graphs = bool(on)
        self.student.enable_cuda_graphs(on, getattr(self.optimizer, "mark_touched", None))
        self.teacher.enable_cuda_graphs(on)
        self.teacher.trunk_max_ctas = self.teacher_trunk_ctas if (on and 0 < self.teacher_trunk_ctas < 148) else None

    def weight_sum_loss(self, loss_dict, step):
        return weight_sum_losses(loss_dict, step, self.rampup_step, self.rampdown_step, self.max_iter,
                                 l=self.lambda_value, balanced=self.balanced_weight, start_mt=self.start_mt)

    # ---- forward passes -----------------------------------------------------------------------
    def forward_source(self, image, target):
        image = image.to(self.device)
        target = [t.to(self.device) for t in target]
        return self.student(image, target)

    def _next_unlabeled(self):
        # the reference builds a NEW loader iterator (worker respawn) every iteration (MTtrainer.py:252);
        # a persistent iterator yields the same stream of batches without the respawn
        if self._iter_u is None:
            self._iter_u = iter(self.dataloader_u)
        try:
            return next(self._iter_u)
        except StopIteration:
            self._iter_u = iter(self.dataloader_u)
            return next(self._iter_u)

    def _teacher_pass(self, teacher_list):
        """The no-grad teacher forward; None when the teacher produced no boxes for an image (the pair is skipped)."""
        with torch.no_grad():
            try:
                return self.teacher.forward_teacher(teacher_list)
            except ValueError as e:  # the reference swallows everything; only "no boxes" is expected
                if "boxes available" not in str(e):
                    raise
                self.logger.warning("teacher produced no boxes for one image, skip this pair (%s)", e)
                return None

    # ---- input staging ------------------------------------------------------------------------
    def stage(self, data_s=None, data_u_list=None, iteration=None, targets=None):
        """Start the host -> device copies of the NEXT iteration's image tensors (pinned host memory) on a dedicated copy
        stream and return at once: called right after `train_step` has enqueued the current iteration, the uploads
        (61 MB per iteration on the bench workload: ~2.4 ms of PCIe time) run under that iteration's compute instead of in
        front of the next one's first kernel.  `train_step` picks the staged copies up by the identity of the host tensors
        it is later handed; anything not staged is copied the usual way.  Upload order = use order: labeled batch, the
        student's unlabeled view, the teacher's views."""
        if self.device.type != "cuda":
            return targets
        probe = self._stage_probe
        if probe is not None:
            probe.append([time.perf_counter()])
        dev_targets = None
        if targets is not None:
            # the next iteration's ground truth: boxes / labels uploaded and the polygons packed into their device arrays
            # now, while the host would otherwise wait for this iteration's loss; hand the returned list to `train_step`
            dev_targets = [self._upload_boxlist(t) for t in targets]
            for t in dev_targets:
                if t.has_field("masks") and hasattr(t.get_field("masks"), "packed"):
                    t.get_field("masks").packed(self.device)
        if probe is not None:
            probe[-1].append(time.perf_counter())
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._pregrow_copy = True
        order = []
        if data_s is not None:
            order.append(data_s.tensors if hasattr(data_s, "tensors") else data_s)
        if data_u_list is not None:
            views = list(data_u_list)
            order += views[len(views) - self.student_bs:] + views[:len(views) - self.student_bs]
        staged = {}
        with torch.cuda.stream(self._copy_stream):
            if self._pregrow_copy:      # the copy stream has its own allocator pool: two generations of staged batches
                self._pregrow_copy = False
                self._pregrow_pool(self.pool_gb[2] if len(self.pool_gb) > 2 else 0.5, small_blocks=0)
            for t in order:
                if isinstance(t, torch.Tensor) and not t.is_cuda and id(t) not in staged:
                    staged[id(t)] = (t, t.to(self.device, non_blocking=True))
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._staged = (staged, done)
        if probe is not None:
            probe[-1].append(time.perf_counter())
        # With the iteration number at hand the NEXT iteration's conv trunks are enqueued as well (they only need the
        # staged inputs and this iteration's SGD + EMA, already in the main stream's queue): a caller that now blocks on this
        # iteration's loss finds the device busy with them when it comes back, instead of idle until `train_step` has
        # re-filled the queue.
        if (iteration is not None and self.cuda_graphs and data_s is not None and data_u_list is not None
                and self.n_step_unlabel == 1 and iteration > self.start_mt and self.lambda_value > 0
                and bool(self.cfg.DATASETS.NO_LABEL) and getattr(self.student, "_graphs", None)
                and getattr(self.teacher, "_graphs", None)):
            views = list(data_u_list)
            if all(isinstance(v, torch.Tensor) for v in views):
                host_img = data_s.tensors if (hasattr(data_s, "tensors") and not isinstance(data_s, torch.Tensor)) else data_s
                img_t = self._take(host_img)
                dev_views = [self._take(v) for v in views]
                self._prefetched = {id(host_img): img_t}
                self._prefetched.update({id(h): d for h, d in zip(views, dev_views)})
                if self.student.prefetch_student(img_t, dev_views[len(views) - self.student_bs:]):
                    if probe is not None:
                        probe[-1].append(time.perf_counter())
                    with torch.no_grad():
                        self.teacher.prefetch_trunk("aug", dev_views[:self.teacher_bs])
        if probe is not None:
            probe[-1].append(time.perf_counter())
        return dev_targets

    def _upload_boxlist(self, t):
        """`t.to(device)` with every host tensor going through pinned staging and a non-blocking copy.  A copy from pageable
        memory synchronises the stream it is enqueued on before it starts: `stage` — called while the iteration it follows
        is still running — blocked 6.4 ms per call in its ten small target uploads (bench.py, MMT_BENCH_TRACE=2)."""
        from ..structures.bounding_box import BoxList

        def up(v):
            if isinstance(v, torch.Tensor):
                if v.device.type == "cpu" and v.numel():
                    v = v if v.is_pinned() else v.pin_memory()
                return v.to(self.device, non_blocking=True)
            return v.to(self.device) if hasattr(v, "to") else v

        out = BoxList(up(t.bbox), t.size, t.mode)
        for k, v in t.extra_fields.items():
            out.add_field(k, up(v))
        return out

    def _take(self, t):
        """The device copy of host tensor / ImageList `t`: the staged one if `stage` uploaded it, else a copy enqueued now."""
        if hasattr(t, "tensors") and not isinstance(t, torch.Tensor):
            from ..structures.image_list import ImageList
            return ImageList(self._take(t.tensors), t.image_sizes)
        if self._prefetched and id(t) in self._prefetched:
            return self._prefetched.pop(id(t))          # uploaded AND its trunk already enqueued by `stage`
        st = self._staged
        if st is not None and id(t) in st[0]:
            host, dev_t = st[0].pop(id(t))
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(st[1])
            dev_t.record_stream(cur)     # allocated on the copy stream, consumed on this one
            return dev_t
        return t.to(self.device, non_blocking=True)

    def _pregrow_pool(self, gigabytes, small_blocks=2048):
        """Grow the caching allocator's pools OF THE CURRENT STREAM once, up front: one large block and a few thousand
        small ones, allocated and released again, stay mapped in the stream's (expandable) segments, so the data-dependent
        block sizes of later iterations are carved out of them instead of mapping fresh pages — an operation that stalls
        the calling thread AND every other thread waiting for the allocator lock for ~10 ms each while the device is busy
        (profiles/r02_stall_trace.md, second half)."""
        if self.device.type != "cuda" or gigabytes <= 0:
            return
        big = torch.empty(int(gigabytes * (1 << 30)), dtype=torch.uint8, device=self.device)
        small = [torch.empty(256 << 10, dtype=torch.uint8, device=self.device) for _ in range(small_blocks)]
        del big, small

    def _backward_job(self, loss, main_stream):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(main_stream):
            loss.backward()

    def _teacher_job(self, teacher_list, main_stream):
        """Body of the second host thread: the teacher pass on its own CUDA stream (ordered after everything the main
        stream had enqueued when the job was submitted: the previous iteration's SGD + EMA, this iteration's uploads)."""
        torch.cuda.set_device(self.device)
        if self._t_stream is None:
            self._t_stream = torch.cuda.Stream(device=self.device, priority=-1 if self.teacher_first else min(self.head_priority, 0))
        self._t_stream.wait_stream(main_stream)
        with torch.cuda.stream(self._t_stream):
            if self._pregrow_teacher:
                self._pregrow_teacher = False
                self._pregrow_pool(self.pool_gb[1])
            return self._teacher_pass(teacher_list)

    def forward_unlabel(self, data_u_list=None, teacher_future=None):
        unlabel, loss_dict = [], {}
        for _ in range(self.n_step_unlabel):
            if data_u_list is None:
                data_u_list, _ = self._next_unlabeled()
            student = [s.to(self.device) for s in data_u_list[-self.student_bs:]]
            if teacher_future is not None:
                teacher_results = teacher_future.result()          # re-raises whatever the teacher thread raised
                torch.cuda.current_stream(self.device).wait_stream(self._t_stream)
                teacher_future = None
            else:
                teacher_results = self._teacher_pass([f.to(self.device) for f in data_u_list[:self.teacher_bs]])
            if teacher_results is None:
                data_u_list = None
                continue
            unlabel.append(self.student.forward_student(student, teacher_results))
            data_u_list = None
        if not unlabel:
            return loss_dict
        keys = unlabel[0].keys()
        return {k: sum(d[k] for d in unlabel) / len(unlabel) for k in keys}

    def close(self):
        """Stop the teacher thread (idempotent).  `train()` calls it; callers that drive `train_step` themselves may."""
        for attr in ("_t_pool", "_bw_pool"):
            pool = getattr(self, attr, None)
            setattr(self, attr, None)
            if pool is not None:
                pool.shutdown(wait=True)

    def update_teacher(self, step):
        alpha = min(1 - 1 / (step + 1), self.alpha)
        self.optimizer.ema_only(alpha)

    # ---- one iteration ------------------------------------------------------------------------
    def train_step(self, iteration, data_s, target_s, data_u_list=None):
        """One training iteration; returns (weighted loss dict, total loss).  With CUDA graphs on and `head_priority` < 0 the
        iteration is enqueued from a high-priority stream that starts after everything on the caller's current stream and
        that the caller's stream waits for on return (the trunk graphs run on their own side streams either way)."""
        if not (self.cuda_graphs and self.head_priority < 0 and self.device.type == "cuda"):
            return self._train_step(iteration, data_s, target_s, data_u_list)
        ambient = torch.cuda.current_stream(self.device)
        if self._head_stream is None:
            self._head_stream = torch.cuda.Stream(device=self.device, priority=self.head_priority)
        self._head_stream.wait_stream(ambient)
        with torch.cuda.stream(self._head_stream):
            losses_dict, losses = self._train_step(iteration, data_s, target_s, data_u_list)
        ambient.wait_stream(self._head_stream)
        for t in [losses] + list(losses_dict.values()):      # allocated on the head stream, read by the caller on its own
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(ambient)
        return losses_dict, losses

    def _train_step(self, iteration, data_s, target_s, data_u_list=None):
        if self.gc_every > 0:
            if self._gc_steps == 0:
                gc.collect()
                gc.disable()
            elif self._gc_steps % self.gc_every == 0:
                gc.collect(1)
            self._gc_steps += 1
        # First iteration after a CUDA-graph capture (torch.cuda.graph() empties the allocator's cache when it starts, so
        # anything grown earlier is gone): grow the per-stream pools now, and park everything alive — models, captured
        # graphs, cached anchors, the config tree: as long-lived as the trainer — in the GC's permanent generation, so
        # that the paced collections above only walk the objects of the last iterations (17-20 ms -> < 1 ms each).
        n_graphs = len(getattr(self.student, "_graphs", ())) + len(getattr(self.teacher, "_graphs", ()))
        if self.cuda_graphs and n_graphs and n_graphs != self._pregrown_for:
            self._pregrown_for = n_graphs
            if self.gc_every > 0:
                gc.collect()
                gc.freeze()
            self._pregrow_pool(self.pool_gb[0])
            self._pregrow_teacher = True
            self._pregrow_copy = self._copy_stream is not None
        reduced = False
        mt_on = iteration > self.start_mt and self.lambda_value > 0 and bool(self.cfg.DATASETS.NO_LABEL)
        stacked = False
        if self.cuda_graphs:
            # the conv trunks of the iteration are independent: enqueue them all on the side stream first, so the
            # host-side glue of each pass overlaps the conv work of the next (detector.prefetch_trunk / prefetch_student)
            data_s = self._take(data_s)   # pinned host batches: staged on the copy stream, or a non-blocking copy now
            img_s = data_s.tensors if hasattr(data_s, "tensors") else data_s
            if mt_on and data_u_list is not None and self.n_step_unlabel == 1 and all(
                    isinstance(d, torch.Tensor) for d in data_u_list):
                # upload order = use order: the student's view first, so its (stacked) trunk starts while the copy
                # engine is still bringing in the teacher's views
                data_u_list = list(data_u_list)
                n_u = len(data_u_list)
                if self.teacher_first:
                    # The teacher's trunk ([2K*B] views, the larger one) heads the iteration's critical path — teacher trunk
                    # -> teacher heads (read-backs) -> student heads on the teacher's ROIs -> trunk backward — while the
                    # supervised pass has slack until the two join: enqueue it first, on a high-priority stream.
                    for i in range(n_u):
                        data_u_list[i] = self._take(data_u_list[i])
                    with torch.no_grad():
                        self.teacher.prefetch_trunk("aug", data_u_list[:self.teacher_bs])
                    stacked = self.student.prefetch_student(img_s, data_u_list[-self.student_bs:])
                    if not stacked:
                        self.student.prefetch_trunk("sup", img_s)
                        self.student.prefetch_trunk("unl", data_u_list[-self.student_bs:])
                else:
                    for i in range(n_u - self.student_bs, n_u):
                        data_u_list[i] = self._take(data_u_list[i])
                    stacked = self.student.prefetch_student(img_s, data_u_list[-self.student_bs:])
                    if not stacked:
                        self.student.prefetch_trunk("sup", img_s)
                    for i in range(n_u - self.student_bs):
                        data_u_list[i] = self._take(data_u_list[i])
                    with torch.no_grad():
                        self.teacher.prefetch_trunk("aug", data_u_list[:self.teacher_bs])
                    if not stacked:
                        self.student.prefetch_trunk("unl", data_u_list[-self.student_bs:])
            else:
                self.student.prefetch_trunk("sup", img_s)
        if self.cuda_graphs and self.early_rpn_targets and getattr(self.student, "_pending", None):
            target_s = [t.to(self.device) for t in target_s]
            self.student.precompute_rpn_targets(data_s, target_s)
        self.scheduler.step()
        self.optimizer.zero_grad()
        teacher_future = None
        if stacked and self.teacher_thread and self.n_step_unlabel == 1 and _Sampler.rng_device is None:
            # (parity mode draws from torch's global CPU generator in the reference's order: passes stay sequential there)
            if self._t_pool is None:
                self._t_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mmt-teacher")
            teacher_future = self._t_pool.submit(self._teacher_job, list(data_u_list[:self.teacher_bs]),
                                                 torch.cuda.current_stream(self.device))
        loss_dict = self.forward_source(data_s, target_s)
        # The supervised terms carry weight 1 (no 'mt' in their keys, MTtrainer.py:99 of the reference), so their
        # backward does not depend on the consistency ramp: run it NOW — its device work overlaps the teacher pass's
        # host-side glue — and back-propagate the consistency terms after the unlabeled pass.  Gradients of both
        # passes accumulate in the flat arena: d(sum)/dw = sum of d/dw, the reference's single backward.  With the
        # stacked student trunk the early backward stops at the trunk boundary (detached leaves) and the backbone's
        # backward runs once, at the end, with both branches' gradients (detector.backward_deferred).
        sup_weighted = self.weight_sum_loss(loss_dict, iteration)
        sup_total = sum(v for v in sup_weighted.values())
        if mt_on:
            bw_future = None
            if teacher_future is not None and self.backward_thread:
                # The main thread is the iteration's longest host path (it reaches this point ~21.5 ms into a 24 ms
                # iteration with the teacher's results long waiting, profiles/r02_join_probe.log): the supervised heads'
                # backward — ~2 ms of enqueueing — moves to a worker while this thread starts on the unlabeled pass.
                # Autograd runs the backward kernels on the streams of the forward ops whichever thread calls it; the
                # worker is joined before the consistency terms' backward touches the same parameters' gradients.
                if self._bw_pool is None:
                    self._bw_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mmt-backward")
                bw_future = self._bw_pool.submit(self._backward_job, sup_total, torch.cuda.current_stream(self.device))
            else:
                sup_total.backward()
            mt_dict = self.forward_unlabel(data_u_list, teacher_future)
            if bw_future is not None:
                bw_future.result()
            mt_weighted = self.weight_sum_loss(mt_dict, iteration)
            losses_dict = dict(sup_weighted)
            losses_dict.update(mt_weighted)
            mt_total = sum(v for v in mt_weighted.values()) if mt_weighted else None
            losses = sup_total.detach() + (mt_total.detach() if mt_total is not None else 0)
            self._publish_loss(losses)         # the iteration's result exists now: its read-back need not wait for the backward
            if stacked:
                # heads first; their gradient bucket (fc6 alone is 51 MB of the 185 MB arena) is then final and its
                # all-reduce runs on NCCL's stream UNDER the trunk's backward graph; the trunk bucket follows it
                self.student.backward_deferred_heads(mt_total)
                if self.world > 1:
                    early, late = self.optimizer.reduce_buckets()
                    pending = [dist.all_reduce(t, async_op=True) for t in early]
                self.student.backward_deferred_trunk()
                if self.world > 1:
                    self.optimizer.publish_touched()
                    pending += [dist.all_reduce(t, async_op=True) for t in late]
                    for w in pending:
                        w.wait()
                    reduced = True
            elif mt_total is not None:
                mt_total.backward()
        else:
            losses_dict, losses = sup_weighted, sup_total
            self._publish_loss(losses)
            losses.backward()
        if self.world > 1 and not reduced:
            # one NCCL all-reduce: the whole student gradient arena + the per-tensor "touched" flags behind it
            dist.all_reduce(self.optimizer.reduce_buffer())
        ema = None
        if self.lambda_value > 0 and iteration > (self.start_mt - 10):
            ema = min(1 - 1 / (iteration - (self.start_mt - 10) + 1), self.alpha)
        self.optimizer.step(ema_alpha=ema, grad_scale=1.0 / self.world)
        self.last_losses = {k: v.detach() for k, v in losses_dict.items()}
        return losses_dict, losses

    def _publish_loss(self, total):
        """Enqueue the 4-byte device -> host copy of the iteration's total loss into a pinned slot, right behind the kernels
        that produce it (the forward passes and the loss sums) and AHEAD of the backward / all-reduce / optimizer launches."""
        if self.device.type != "cuda":
            self._loss_host, self._loss_event = total.detach().reshape(()).clone(), None
            return
        if self._loss_slot is None:
            self._loss_slot = torch.empty((), dtype=torch.float32).pin_memory()
        self._loss_slot.copy_(total.detach().reshape(()), non_blocking=True)
        self._loss_event = torch.cuda.Event()
        self._loss_event.record(torch.cuda.current_stream(self.device))
        self._loss_host = self._loss_slot

    def loss_value(self):
        """The last `train_step`'s total loss as a Python float.  Waits for the copy `_publish_loss` enqueued — i.e. for the
        iteration's forward passes — not for its backward and optimizer step, so a caller that reads every iteration's loss
        (bench.py's end-to-end loop, a logger) does not drain the device at every iteration boundary.  Call it before the
        next `train_step` (one pinned slot)."""
        if self._loss_host is None:
            raise RuntimeError("loss_value() before the first train_step()")
        if self._loss_event is not None:
            self._loss_event.synchronize()
        return float(self._loss_host)

    def train(self):
        from ..utils.metric_logger import MetricLogger
        meters = MetricLogger(delimiter="  ")
        self.student.train()
        self.teacher.eval()
        t0 = end = time.time()
        iteration = self.start_iter
        for iteration, (data_s, target_s, _) in enumerate(self.dataloader_s, self.start_iter):
            data_u = None
            if (self.dataloader_u is not None and iteration > self.start_mt and self.lambda_value > 0
                    and self.n_step_unlabel == 1):
                data_u, _ = self._next_unlabeled()   # fetched up front so that its trunks can be prefetched
            losses_dict, losses = self.train_step(iteration, data_s, target_s, data_u)
            if iteration % 20 == 0 or iteration == self.max_iter:  # the only host syncs of the loop
                red = reduce_loss_dict(losses_dict, self.loss_keys)
                meters.update(loss=sum(v for v in red.values()), **red)
                meters.update(time=(time.time() - end) / (20 if iteration else 1))
                end = time.time()
                eta = str(datetime.timedelta(seconds=int(meters.time.global_avg * (self.max_iter - iteration))))
                self.logger.info("eta: %s  iter: %d  %s  lr: %.6f  max mem: %.0f" % (
                    eta, iteration, str(meters), self.optimizer.param_groups[0]["lr"],
                    torch.cuda.max_memory_allocated() / 1024.0 / 1024.0))
            if iteration % self.checkpoint_period == 0:
                self.save_model(iteration)
            if iteration % 1000 == 999:
                gc.collect()          # full collection (generation 2) at a point where a pause costs nothing measurable
        self.save_model(iteration + 1, final=True)
        self.close()
        if self.gc_every > 0:
            gc.enable()
        total = time.time() - t0
        self.logger.info("Total training time: %s (%.4f s / it)" % (datetime.timedelta(seconds=total), total / max(self.max_iter, 1)))

    def save_model(self, iteration=0, final=False):
        if self.ckpt_s is None:
            return
        # `iteration` stored in the file = number of COMPLETED steps (the step the resumed run starts with): the periodic
        # save runs after 0-based step `iteration` has been applied, the final one is called with iteration + 1 already.
        # The teacher is saved whenever its EMA has started (MTtrainer.train_step: iteration > START_MT - 10).
        ema_on = self.lambda_value > 0 and iteration > (self.start_mt - 10)
        if iteration > 0 and not final:
            self.ckpt_s.save("model_{:07d}".format(iteration), iteration=iteration + 1)
            if ema_on:
                self.ckpt_t.save("t_model_{:07d}".format(iteration), iteration=iteration + 1)
        if final:
            self.ckpt_s.save("model_final", iteration=iteration)
            if ema_on:
                self.ckpt_t.save("t_model_final", iteration=iteration)
