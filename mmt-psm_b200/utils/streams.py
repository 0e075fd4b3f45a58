"""Fork / join over CUDA side streams for INDEPENDENT per-level work (the five pyramid levels of the RPN head and of the
MGD adaptors, the FPN's lateral and output convolutions).

Each of those is a string of small launches — at 800x800 the three coarsest levels have 3 .. 40 pixel tiles, a fraction of
the 148 SMs, and cost 12-37 us apiece whatever their size (profiles/r02_conv_tile_cfg.log) — issued one after the other on
one stream.  Issued on parallel streams they overlap each other and the tail of the finest level's launch.  Inside a CUDA
graph capture the fork / join events become parallel branches of the graph; under autograd the backward of each level runs
on the stream its forward ran on (PyTorch's stream semantics), so the backward gets the same overlap."""
import os

import torch

_ENABLED = os.environ.get("MMT_LEVEL_STREAMS", "1") == "1"
_pools = {}


def set_enabled(on):
    global _ENABLED
    _ENABLED = bool(on)


def _side_streams(cur, n):
    key = (cur.device.index, cur.cuda_stream)      # per calling stream: the trainer's two host threads never share one
    pool = _pools.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=cur.device))
    return pool[:n]


def run_parallel(fns, device=None):
    """fns: zero-argument callables, mutually independent.  Returns their results in order.  fns[0] runs on the current
    stream, the others on side streams that start after everything enqueued on the current stream so far; the current
    stream then waits for all of them."""
    if not _ENABLED or len(fns) <= 1 or device is None or device.type != "cuda":
        return [f() for f in fns]
    cur = torch.cuda.current_stream(device)
    sides = _side_streams(cur, len(fns) - 1)
    fork = torch.cuda.Event()
    fork.record(cur)
    outs = [None] * len(fns)
    for i, s in enumerate(sides, 1):
        s.wait_event(fork)
        with torch.cuda.stream(s):
            outs[i] = fns[i]()
    outs[0] = fns[0]()
    for s in sides:
        cur.wait_stream(s)
    return outs
