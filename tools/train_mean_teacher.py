#!/usr/bin/env python
"""Mean-teacher training launcher with the command line of the reference's tools/train_mean_teacher.py:110-175
(`--config-file FILE [--gpuid ID] [--local_rank R] [--skip-test] KEY VAL ...`, what scripts/train_mt.sh:1-20 invokes), on
the B200 path: this package's cfg / models / FusedSGDEMA / MTtrainer, `torchrun` for multi-GPU.

    python tools/train_mean_teacher.py --config-file configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml --synthetic \
        SOLVER.MAX_ITER 200 MT.START_MT 20 OUTPUT_DIR runs/demo

The reference's dataset classes are out of scope: `--synthetic` feeds the seeded synthetic crop stream the
benchmark uses (labeled batches + unlabeled uint8 crops whose colour views are rendered on the device).  The yaml and
scripts/train_mt.sh switch the IR-Net modules on exactly as the reference's do (BASELINE config 4); append
`MODEL.RELATION_NMS.USE_RELATION_NMS False MODEL.RELATION_MASK.USE_RELATION False` for the plain mean-teacher iteration the
headline metric is quoted on (config 3).  `--init-weights FILE` loads a reference-layout checkpoint (e.g. e2e_mask_rcnn_R_50_FPN_1x.pth);
without it the deterministic synthetic initialisation is used.  Resumes from OUTPUT_DIR/last_checkpoint."""
import argparse
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser(description="MMT-PSM mean-teacher training on libmmtpsm_b200")
    ap.add_argument("--config-file", default="", metavar="FILE")
    ap.add_argument("--local_rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)))
    ap.add_argument("--skip-test", dest="skip_test", action="store_true")
    ap.add_argument("--gpuid", default=None)
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--synthetic", action="store_true", help="train on the seeded synthetic crop stream")
    ap.add_argument("--init-weights", default="", help="reference-layout .pth to start both networks from")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--preset", default="", choices=["", "train_mt"],
                    help="train_mt: apply the option list of the reference's scripts/train_mt.sh before KEY VAL")
    ap.add_argument("opts", nargs=argparse.REMAINDER, default=None)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.config import cfg
    from mmt_psm_b200.data.loaders import SyntheticSourceLoader, UnlabeledViewLoader, synthetic_unlabeled_crops
    from mmt_psm_b200.engine import MTtrainer
    from mmt_psm_b200.modeling.detector import build_detection_model
    from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
    from mmt_psm_b200.utils.checkpoint import DetectronCheckpointer
    from mmt_psm_b200.utils.synthetic_init import init_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(args.local_rank)
    device = torch.device("cuda", args.local_rank)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=device)
    rank = dist.get_rank() if world > 1 else 0
    if args.config_file:
        cfg.merge_from_file(args.config_file)
    if args.preset == "train_mt":
        from mmt_psm_b200.config import TRAIN_MT_OVERRIDES
        cfg.merge_from_list(TRAIN_MT_OVERRIDES + ["MODEL.RELATION_NMS.USE_RELATION_NMS", True])   # train_mt.sh:7
    cfg.merge_from_list(args.opts or [])
    cfg.freeze()
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format="%(asctime)s %(name)s: %(message)s")
    from mmt_psm_b200.utils.logger import setup_logger
    log = setup_logger("maskrcnn_benchmark", cfg.OUTPUT_DIR, rank).getChild("trainer")   # stdout + OUTPUT_DIR/log.txt on rank 0

    torch.manual_seed(0)
    model_s = build_detection_model(cfg, is_student=True).to(device)
    model_t = build_detection_model(cfg, is_teacher=True).to(device)
    init_model(model_s, 0)                                   # identical start on every rank
    out_dir = cfg.OUTPUT_DIR or "."
    pre = DetectronCheckpointer(cfg, model_s, save_dir="", save_to_disk=False)
    if args.init_weights:
        pre.load(args.init_weights)     # the COCO start file goes through the transfer-learning rules (utils/checkpoint.py)
    model_t.load_state_dict(model_s.state_dict())
    optimizer = make_optimizer(cfg, model_s, model_t)
    scheduler = make_lr_scheduler(cfg, optimizer)
    ck_s = DetectronCheckpointer(cfg, model_s, optimizer, scheduler, out_dir, rank == 0)
    ck_t = DetectronCheckpointer(cfg, model_t, None, None, os.path.join(out_dir, "teacher"), rank == 0)
    extra = ck_s.load()
    start_iter = int(extra.get("iteration", 0))
    if start_iter:
        if ck_t.has_checkpoint():
            ck_t.load()
        else:
            model_t.load_state_dict(model_s.state_dict())   # in place: the teacher's parameters are views of the EMA arena
    torch.manual_seed(100 + rank)

    if not args.synthetic:
        raise SystemExit("the reference's Pap-smear dataset classes are out of scope of this path (SURVEY §8): run with "
                         "--synthetic, or build MTtrainer with your own {'source': ..., 'no_label': ...} loaders")
    per_gpu = max(1, cfg.SOLVER.IMS_PER_BATCH // (2 * world))     # img_ratio = 1/2: half labeled, half unlabeled
    size = cfg.INPUT.MIN_SIZE_TRAIN if isinstance(cfg.INPUT.MIN_SIZE_TRAIN, int) else cfg.INPUT.MIN_SIZE_TRAIN[0]
    loaders = {"source": SyntheticSourceLoader(cfg.SOLVER.MAX_ITER - start_iter, per_gpu, size, rank=rank)}
    if cfg.DATASETS.NO_LABEL:
        loaders["no_label"] = UnlabeledViewLoader(synthetic_unlabeled_crops(8, seed=99 + rank), cfg, device, per_gpu, seed=rank)
    trainer = MTtrainer(model_s, model_t, loaders, optimizer, scheduler, ck_s, ck_t, cfg.SOLVER.CHECKPOINT_PERIOD, cfg,
                        cuda_graphs=not args.no_graphs)
    trainer.start_iter = start_iter
    trainer.max_iter = cfg.SOLVER.MAX_ITER
    log.info("training %d -> %d iterations on %d GPU(s), %d labeled + %d unlabeled crops per GPU", start_iter,
             cfg.SOLVER.MAX_ITER, world, per_gpu, per_gpu if cfg.DATASETS.NO_LABEL else 0)
    trainer.train()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
