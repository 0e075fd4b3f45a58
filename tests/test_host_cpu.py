"""CPU suite: host-side logic of the operator surface the reference's launcher drives — cfg semantics (yacs quirks
of SURVEY App. D), BoxList container contract, loss weighting / ramps, LR schedule, registry, synthetic data contract.
No CUDA, no oracle."""
import math

import pytest
import torch

from conftest import ROOT  # noqa: F401  (puts the repo on sys.path)


def test_cfg_yacs_semantics():
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.config import TRAIN_MT_OVERRIDES, get_cfg_defaults
    c = get_cfg_defaults()
    # (ii) flat KEY VAL list from the command line, literal-eval'd strings; (iii) list<->tuple, int->float coercion
    c.merge_from_list(["MT.CLS_LOSS_TYPE", "'bce'", "MT.SHARPEN", "True", "SOLVER.BASE_LR", "0.005", "MT.LAMBDA", 5,
                       "MODEL.RPN.ANCHOR_STRIDE", "(4, 8, 16, 32, 64)"])
    assert c.MT.CLS_LOSS_TYPE == "bce" and c.MT.SHARPEN is True and c.SOLVER.BASE_LR == 0.005
    assert isinstance(c.MT.LAMBDA, float) and c.MT.LAMBDA == 5.0
    assert c.MODEL.RPN.ANCHOR_STRIDE == (4, 8, 16, 32, 64)
    c.merge_from_list(TRAIN_MT_OVERRIDES)
    assert c.MT.START_MT == 1000 and c.MT.T_ADAPT is True and c.MODEL.ROI_BOX_HEAD.DO == 0.5
    # (iv) unknown keys are errors
    with pytest.raises(KeyError):
        c.merge_from_list(["MT.NO_SUCH_KEY", 1])
    with pytest.raises(ValueError):
        c.merge_from_list(["MT.START_MT", "'soon'"])
    # (v) freeze makes assignment raise; clone deep-copies and keeps the frozen flag
    c.freeze()
    with pytest.raises(AttributeError):
        c.MT.LAMBDA = 1.0
    d = c.clone()
    d.defrost()
    d.MT.LAMBDA = 1.0
    assert c.MT.LAMBDA == 5.0 and d.MT.LAMBDA == 1.0
    # (vi) `if cfg.MODEL.RELATION_MASK:` on a node is always truthy (reference mask_head.py:48)
    assert bool(c.MODEL.RELATION_MASK)


def test_boxlist_contract():
    from mmt_psm_b200.structures.bounding_box import BoxList
    from mmt_psm_b200.structures.boxlist_ops import boxlist_iou, cat_boxlist
    b = BoxList(torch.tensor([[10., 20., 29., 59.], [0., 0., 99., 49.]]), (100, 60), "xyxy")
    b.add_field("labels", torch.tensor([1, 2]))
    assert b.area().tolist() == [20. * 40., 100. * 50.]                     # legacy +1 extents
    w = b.convert("xywh")
    assert w.bbox.tolist() == [[10., 20., 20., 40.], [0., 0., 100., 50.]]
    assert torch.equal(w.convert("xyxy").bbox, b.bbox)
    f = b.transpose(0)                                                       # horizontal flip: x' = W - x - 1
    assert f.bbox.tolist() == [[70., 20., 89., 59.], [0., 0., 99., 49.]]
    assert torch.equal(f.transpose(0).bbox, b.bbox)
    s = b[torch.tensor([1])]
    assert len(s) == 1 and s.get_field("labels").tolist() == [2] and s.size == (100, 60)
    c = cat_boxlist([b, s])
    assert len(c) == 3 and c.get_field("labels").tolist() == [1, 2, 2]
    iou = boxlist_iou(b, s)
    assert abs(float(iou[1, 0]) - 1.0) < 1e-7 and 0 < float(iou[0, 0]) < 1
    with pytest.raises(ValueError):
        BoxList(torch.zeros(3, 5), (10, 10))
    clipped = BoxList(torch.tensor([[-5., -5., 120., 70.]]), (100, 60)).clip_to_image(remove_empty=False)
    assert clipped.bbox.tolist() == [[0., 0., 99., 59.]]


def test_loss_weighting_and_ramps():
    from mmt_psm_b200.engine.MTtrainer import weight_sum_losses
    from mmt_psm_b200.utils.miscellaneous import sigmoid_rampdown, sigmoid_rampup
    losses = {"loss_classifier": torch.tensor(2.0), "mt_classifier": torch.tensor(3.0), "mt_fg_loss": torch.tensor(4.0)}
    bal = {"mt_classifier": 0.2, "nms_loss": 1.0, "mt_fg_loss": 1.0}
    # plateau: keys containing 'mt' x lambda, then the per-term balance (reference MTtrainer.py:67-109)
    out = weight_sum_losses(losses, 5000, 250, 250, 10 ** 6, l=5.0, balanced=bal, start_mt=1000)
    assert float(out["loss_classifier"]) == 2.0
    assert abs(float(out["mt_classifier"]) - 3.0 * 5.0 * 0.2) < 1e-6 and abs(float(out["mt_fg_loss"]) - 20.0) < 1e-6
    # ramp-up right after START_MT
    out = weight_sum_losses(losses, 1100, 250, 250, 10 ** 6, l=5.0, balanced=bal, start_mt=1000)
    w = 5.0 * sigmoid_rampup(100, 250)
    assert abs(float(out["mt_fg_loss"]) - 4.0 * w) < 1e-5 and w < 5.0
    # ramp-down near the end uses the ramp-UP length (reference quirk, MTtrainer.py:92)
    out = weight_sum_losses(losses, 999900, 250, 400, 10 ** 6, l=5.0, balanced=bal, start_mt=1000)
    assert abs(float(out["mt_fg_loss"]) - 4.0 * 5.0 * sigmoid_rampdown(100, 250)) < 1e-5
    assert sigmoid_rampup(0, 250) == pytest.approx(math.exp(-5.0)) and sigmoid_rampup(250, 250) == 1.0


def test_warmup_multistep_lr():
    from mmt_psm_b200.solver import WarmupMultiStepLR

    class Opt(object):
        base_lr = 0.02

        def set_lr(self, lr):
            self.lr = lr
    o = Opt()
    s = WarmupMultiStepLR(o, (30, 40), gamma=0.1, warmup_factor=1.0 / 3, warmup_iters=10, warmup_method="linear")
    assert o.lr == pytest.approx(0.02 / 3)                       # last_epoch 0
    for _ in range(5):
        s.step()
    assert o.lr == pytest.approx(0.02 * (1 / 3 * 0.5 + 0.5))     # half-way through the linear warm-up
    for _ in range(10):
        s.step()
    assert o.lr == pytest.approx(0.02)
    for _ in range(20):
        s.step()
    assert o.lr == pytest.approx(0.002)                          # epoch 35: one milestone passed
    with pytest.raises(ValueError):
        WarmupMultiStepLR(o, (40, 30))


def test_registry_and_model_surface():
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.modeling import registry
    from mmt_psm_b200.modeling.detector import build_detection_model, reference_state_dict
    assert "R-50-FPN" in registry.BACKBONES and "SingleConvRPNHead" in registry.RPN_HEADS
    assert "FPN2MLPFeatureExtractor" in registry.ROI_BOX_FEATURE_EXTRACTORS
    cfg = mt_psm_cfg(start_mt=0)
    m = build_detection_model(cfg, is_student=True)
    sd = reference_state_dict(m)
    # the reference's state-dict names and layouts at the boundary (OIHW convs, [out, in] linears, IOHW deconv)
    assert tuple(sd["backbone.body.stem.conv1.weight"].shape) == (64, 3, 7, 7)
    assert tuple(sd["backbone.body.layer1.0.conv1.weight"].shape) == (64, 64, 1, 1)
    assert tuple(sd["backbone.fpn.fpn_layer1.weight"].shape) == (256, 256, 3, 3)
    assert tuple(sd["rpn.head.cls_logits.weight"].shape) == (3, 256, 1, 1)
    assert tuple(sd["box_heads.box.feature_extractor.fc6.weight"].shape) == (1024, 256 * 7 * 7)
    assert tuple(sd["mask_heads.mask.predictor.conv5_mask.weight"].shape) == (256, 256, 2, 2)
    assert tuple(sd["hint_adaptor.adapter_1.weight"].shape) == (128, 256, 1, 1)
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert any(n.startswith("backbone.body.stem") for n in frozen) and any(n.startswith("backbone.body.layer1") for n in frozen)
    assert not any(n.startswith("backbone.body.layer2") for n in frozen)          # FREEZE_CONV_BODY_AT = 2


def test_synthetic_batch_contract():
    from mmt_psm_b200.data.synthetic import make_batch
    images, targets, views = make_batch(7, batch=2, size=128, n_views=3, n_cyto=3, n_nuc=2)
    assert tuple(images.shape) == (2, 3, 128, 128) and images.dtype == torch.float32
    assert len(views) == 3 and all(tuple(v.shape) == (2, 3, 128, 128) for v in views)
    for t in targets:
        assert len(t) == 5 and t.mode == "xyxy" and t.size == (128, 128)
        assert sorted(t.get_field("labels").tolist()) == [1, 1, 1, 2, 2]
        assert len(t.get_field("masks").polygons) == 5
        b = t.bbox
        assert (b[:, 2] > b[:, 0]).all() and (b[:, 3] > b[:, 1]).all() and b.min() >= 0 and b.max() <= 127
    again = make_batch(7, batch=2, size=128, n_views=3, n_cyto=3, n_nuc=2)
    assert torch.equal(images, again[0]) and torch.equal(views[2], again[2][2])   # seeded: identical on every machine


def test_checkpoint_roundtrip_reference_layout(tmp_path):
    """save -> load restores weights bit for bit through the reference-layout boundary; the file holds the reference's
    key names / layouts; `last_checkpoint` resumes; scheduler state travels."""
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.modeling.detector import build_detection_model
    from mmt_psm_b200.solver import WarmupMultiStepLR
    from mmt_psm_b200.utils.checkpoint import DetectronCheckpointer
    from mmt_psm_b200.utils.synthetic_init import init_model
    cfg = mt_psm_cfg(start_mt=0)
    a = build_detection_model(cfg, is_student=True)
    init_model(a, 3)

    class Opt(object):
        base_lr = 0.01

        def set_lr(self, lr):
            self.lr = lr
    sched = WarmupMultiStepLR(Opt(), (30, 40))
    for _ in range(7):
        sched.step()
    ck = DetectronCheckpointer(cfg, a, None, sched, str(tmp_path), True)
    path = ck.save("model_0000007", iteration=7)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert tuple(raw["model"]["backbone.body.stem.conv1.weight"].shape) == (64, 3, 7, 7)       # OIHW in the file
    assert raw["iteration"] == 7 and raw["scheduler"]["last_epoch"] == 7
    b = build_detection_model(cfg, is_student=True)
    sched_b = WarmupMultiStepLR(Opt(), (30, 40))
    extra = DetectronCheckpointer(cfg, b, None, sched_b, str(tmp_path), True).load()            # via last_checkpoint
    assert extra["iteration"] == 7 and sched_b.last_epoch == 7
    sa, sb = a.state_dict(), b.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa if "cell_anchors" not in k)
    assert DetectronCheckpointer(cfg, b, save_dir=str(tmp_path / "empty"), save_to_disk=True).load() == {}



def test_metric_logger_contract_of_the_reference_test():
    """The assertions of the reference's own tests/test_metric_logger.py, on this package's meters."""
    from mmt_psm_b200.utils.metric_logger import MetricLogger
    meter = MetricLogger()
    for i in range(10):
        meter.update(metric=float(i))
    m = meter.meters["metric"]
    assert (m.count, m.total, m.median, m.avg) == (10, 45, 4, 4.5)
    assert meter.metric is m and meter.delimiter == "\t"
    with pytest.raises(AttributeError):
        _ = meter.not_existent
    meter.update(loss=torch.tensor(2.0), time=0.5)
    assert "loss: 2.0000 (2.0000)" in str(meter) and abs(meter.time.global_avg - 0.5) < 1e-12


def test_batch_samplers_known_answers_of_the_reference_tests():
    """Known answers of the reference's tests/test_data_samplers.py for the batch samplers its train loader stacks."""
    import itertools
    from torch.utils.data.sampler import BatchSampler, RandomSampler, Sampler, SequentialSampler
    from mmt_psm_b200.data.samplers import GroupedBatchSampler, IterationBasedBatchSampler

    class Subset(Sampler):
        def __init__(self, ids):
            self.ids = ids

        def __iter__(self):
            return iter(self.ids)

        def __len__(self):
            return len(self.ids)

    data = list(range(40))
    for bs in (1, 3, 5, 6):   # groups of 10 consecutive ids: order is preserved
        out = list(GroupedBatchSampler(SequentialSampler(data), [i // 10 for i in data], bs, False))
        assert list(itertools.chain.from_iterable(out)) == data
    gid = [0, 0, 1, 0, 1, 1, 0, 1, 1, 0]
    seq = SequentialSampler(list(range(10)))
    assert list(GroupedBatchSampler(seq, gid, 1, False)) == [[i] for i in range(10)]
    assert list(GroupedBatchSampler(seq, gid, 3, False)) == [[0, 1, 3], [2, 4, 5], [6, 9], [7, 8]]
    assert list(GroupedBatchSampler(seq, gid, 4, False)) == [[0, 1, 3, 6], [2, 4, 5, 7], [8], [9]]
    assert list(GroupedBatchSampler(seq, gid, 3, True)) == [[0, 1, 3], [2, 4, 5]]
    assert list(GroupedBatchSampler(Subset([0, 3, 5, 6, 7, 8]), gid, 3, False)) == [[0, 3, 6], [5, 7, 8]]
    assert list(GroupedBatchSampler(Subset([5, 0, 6, 1, 3, 8]), gid, 3, False)) == [[5, 8], [0, 6, 1], [3]]
    assert list(GroupedBatchSampler(Subset([5, 0, 6, 1, 3, 8]), gid, 3, True)) == [[0, 6, 1]]
    g = GroupedBatchSampler(RandomSampler(list(range(10))), gid, 3, True)
    n = len(g)
    assert len(list(g)) == n == len(g)
    for bs in (2, 3, 4):
        for iters in (4, 10, 20):
            for drop_last in (False, True):
                inner = BatchSampler(SequentialSampler(list(range(10))), bs, drop_last=drop_last)
                it = IterationBasedBatchSampler(inner, iters)
                assert len(it) == iters
                for i, batch in enumerate(it):
                    start = (i % len(inner)) * bs
                    assert batch == list(range(start, min(start + bs, 10)))


def test_checkpointer_contract_of_the_reference_test(tmp_path):
    """The cases of the reference's own tests/checkpoint.py: save / resume via `last_checkpoint` or by file name across
    DataParallel prefixes, and longest-suffix key matching into a nested model."""
    import os
    from collections import OrderedDict
    from torch import nn
    from mmt_psm_b200.utils.checkpoint import Checkpointer
    from mmt_psm_b200.utils.model_serialization import load_state_dict
    make = lambda: nn.Sequential(nn.Linear(2, 3), nn.Linear(3, 1))
    pairs = [(make(), make()), (nn.DataParallel(make()), make()), (make(), nn.DataParallel(make())),
             (nn.DataParallel(make()), nn.DataParallel(make()))]
    for i, (trained, fresh) in enumerate(pairs):
        d = str(tmp_path / ("same%d" % i))
        Checkpointer(trained, save_dir=d, save_to_disk=True).save("checkpoint_file")
        ck = Checkpointer(fresh, save_dir=d)
        assert ck.has_checkpoint() and ck.get_checkpoint_file() == os.path.join(d, "checkpoint_file.pth")
        ck.load()
        for a, b in zip(trained.parameters(), fresh.parameters()):
            assert a is not b and a.equal(b)
    for i, (trained, fresh) in enumerate([(make(), make()), (nn.DataParallel(make()), make()), (make(), nn.DataParallel(make()))]):
        d, other = str(tmp_path / ("src%d" % i)), str(tmp_path / ("dst%d" % i))
        os.makedirs(other)
        Checkpointer(trained, save_dir=d, save_to_disk=True).save("checkpoint_file")
        ck = Checkpointer(fresh, save_dir=other)
        assert not ck.has_checkpoint() and ck.get_checkpoint_file() == ""
        ck.load(os.path.join(d, "checkpoint_file.pth"))
        for a, b in zip(trained.parameters(), fresh.parameters()):
            assert a is not b and a.equal(b)
    for wrap in (False, True):
        m = nn.Module()
        m.block1 = nn.Module()
        m.block1.layer1 = nn.Linear(2, 3)
        m.layer2 = nn.Linear(3, 2)
        m.res = nn.Module()
        m.res.layer2 = nn.Linear(3, 2)
        sd = OrderedDict([("layer1.weight", torch.rand(3, 2)), ("layer1.bias", torch.rand(3)), ("layer2.weight", torch.rand(2, 3)),
                          ("layer2.bias", torch.rand(2)), ("res.layer2.weight", torch.rand(2, 3)), ("res.layer2.bias", torch.rand(2))])
        model = nn.DataParallel(m) if wrap else m
        missing, unused = load_state_dict(model, sd)
        assert not missing and not unused
        for loaded, stored in zip(model.state_dict().values(), sd.values()):
            assert loaded is not stored and loaded.equal(stored)


def test_checkpointer_transfer_learning_file(tmp_path):
    """The COCO start file is loaded the reference's way (utils/checkpoint.py:79-117,148-160): predictors dropped, no
    optimizer / scheduler state, iteration -1, and it wins over an existing last_checkpoint."""
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.modeling.detector import build_detection_model, reference_state_dict
    from mmt_psm_b200.solver import WarmupMultiStepLR
    from mmt_psm_b200.utils.checkpoint import DetectronCheckpointer
    from mmt_psm_b200.utils.synthetic_init import init_model
    cfg = mt_psm_cfg(start_mt=0)
    src = build_detection_model(cfg, is_student=True)
    init_model(src, 5)
    f = str(tmp_path / "e2e_mask_rcnn_R_50_FPN_1x.pth")
    torch.save({"model": {"module." + k: v for k, v in reference_state_dict(src).items()}, "iteration": 89999,
                "scheduler": {"last_epoch": 89999}}, f)

    class Opt(object):
        base_lr = 0.01

        def set_lr(self, lr):
            self.lr = lr
    dst = build_detection_model(cfg, is_student=True)
    init_model(dst, 1)
    before = {k: v.clone() for k, v in dst.state_dict().items()}
    sched = WarmupMultiStepLR(Opt(), (30, 40))
    run = str(tmp_path / "run")
    DetectronCheckpointer(cfg, dst, None, sched, run, True).save("model_0000003", iteration=3)     # a last_checkpoint exists
    extra = DetectronCheckpointer(cfg, dst, None, sched, run, True).load(f)
    assert extra["iteration"] == -1 and sched.last_epoch != 89999
    a, b = src.state_dict(), dst.state_dict()
    for k in b:
        if "cell_anchors" in k:
            continue
        if any(n in k for n in ("cls_score", "bbox_pred", "mask_fcn_logits")):
            assert torch.equal(b[k], before[k]), k          # kept at its own initialisation
        else:
            assert torch.equal(b[k], a[k]), k


def test_run_logger_contract(tmp_path, capsys):
    from mmt_psm_b200.utils.logger import setup_logger
    log = setup_logger("mmt_test_run", str(tmp_path), 0)
    setup_logger("mmt_test_run", str(tmp_path), 0)                      # idempotent: one line per message
    log.getChild("trainer").info("iter: 20  loss: 1.5")
    for h in log.handlers:
        h.flush()
    text = open(str(tmp_path / "log.txt")).read()
    assert text.count("iter: 20  loss: 1.5") == 1 and "mmt_test_run.trainer INFO" in text
    assert "iter: 20  loss: 1.5" in capsys.readouterr().out
    quiet = setup_logger("mmt_test_rank1", str(tmp_path / "r1"), 1)
    quiet.info("nothing")
    assert not (tmp_path / "r1").exists()


def test_split_nonzero_equals_per_mask_nonzero():
    """modeling.box_head.split_nonzero: one `nonzero` over the concatenated masks == a `nonzero` per mask (what the reference
    runs, one host sync per image: box_head/loss.py:107-111, mask_head/mask_head.py:24-31)."""
    import torch
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.modeling.box_head import split_nonzero
    g = torch.Generator().manual_seed(0)
    masks = [torch.rand(n, generator=g) > 0.7 for n in (17, 1, 40, 0, 5, 512)]
    out = split_nonzero(masks)
    assert all(torch.equal(o, m.nonzero().squeeze(1)) for o, m in zip(out, masks))
    assert torch.equal(split_nonzero(masks[:1])[0], masks[0].nonzero().squeeze(1))
    assert [len(o) for o in split_nonzero([torch.zeros(4, dtype=torch.bool)] * 2)] == [0, 0]
