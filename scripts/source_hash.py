"""sha256 over the kernel sources and the C ABI header: identifies the BUILD a measured artefact (ncu capture, traffic
table) belongs to.  bench.py refuses to quote an ncu traffic figure whose recorded hash is not the current one."""
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_hash():
    h = hashlib.sha256()
    files = []
    for d in ("mmt-psm_b200/csrc", "include"):
        for f in sorted(os.listdir(os.path.join(ROOT, d))):
            if f.endswith((".cu", ".cuh", ".h")):
                files.append(os.path.join(d, f))
    for rel in files:
        h.update(rel.encode())
        with open(os.path.join(ROOT, rel), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(source_hash())
