"""Loss trajectory of BASELINE config 4 (MT + IR-Net) on the CPU oracle trainer, same two synthetic batches, initialisation and
schedule position as bench.py --workload mt_relation: does the reference's own arithmetic stay finite over N iterations?

    python scripts/oracle_relation_trace.py [iterations] [size]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmt_psm_b200  # noqa: E402,F401
from mmt_psm_b200.config.presets import mt_psm_cfg  # noqa: E402
from mmt_psm_b200.data.synthetic import make_batch  # noqa: E402
from mmt_psm_b200.modeling.detector import build_detection_model, reference_state_dict  # noqa: E402
from mmt_psm_b200.utils.synthetic_init import synthetic_state_dict  # noqa: E402
from oracle import model_cpu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
size = int(sys.argv[2]) if len(sys.argv) > 2 else 800
relation = os.environ.get("MMT_RELATION", "1") == "1"
torch.set_num_threads(os.cpu_count() or 1)
c = mt_psm_cfg(start_mt=0, relation=relation)
m = build_detection_model(c, is_student=True)
shapes = {k: tuple(v.shape) for k, v in reference_state_dict(m).items()}
del m
oc = model_cpu.Cfg()
oc.start_mt = 0
oc.relation = relation
trainer = model_cpu.Trainer(synthetic_state_dict(shapes, 0), oc)
trainer.last_epoch = 2000
pool = []
for j in range(2):
    images, targets, views = make_batch(1234 + 1000 * j, 2, size, 3, 20, 20)
    tg = [dict(bbox=t.bbox, labels=t.get_field("labels"),
               polygons=[[p.numpy() for p in inst.polygons] for inst in t.get_field("masks").polygons]) for t in targets]
    pool.append((images, tg, views))
torch.manual_seed(7)
for it in range(n):
    images, tg, views = pool[it % 2]
    t0 = time.perf_counter()
    ld, total = trainer.step(2000 + it, images, tg, views)
    print("it %d %.0fs total %.4f | %s" % (it, time.perf_counter() - t0, float(total),
                                          " ".join("%s=%.4f" % (k, float(v)) for k, v in ld.items())), flush=True)
