"""CPU suite, world_size 2 over gloo: the host-side data-parallel logic of the path — rank sharding of the
streams, the single all-reduce over the flat student-gradient arena, and the logging reduce."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.config.presets import mt_psm_cfg
    from mmt_psm_b200.data.samplers import DistributedSampler
    from mmt_psm_b200.engine.MTtrainer import reduce_loss_dict
    from mmt_psm_b200.solver import FusedSGDEMA
    from mmt_psm_b200.utils.comm import get_rank, get_world_size, synchronize
    assert get_world_size() == world and get_rank() == rank
    # identical replicas (seeded), different gradients per rank
    torch.manual_seed(0)
    student = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    teacher = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    teacher.load_state_dict(student.state_dict())
    opt = FusedSGDEMA(mt_psm_cfg(), student, teacher)
    for p in student.parameters():
        assert p.data_ptr() >= opt.flat_p.data_ptr() and p.grad.data_ptr() >= opt.flat_g.data_ptr()
    opt.zero_grad()
    x = torch.full((4, 5), float(rank + 1))
    student(x).sum().backward()                      # grads accumulate INTO the arena views
    local = opt.flat_g.clone()
    dist.all_reduce(opt.flat_g)                      # the one exchange of the path
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(opt.flat_g, sum(gathered)), "arena all-reduce != sum of the local arenas"
    assert torch.equal(student[0].weight.grad, opt.flat_g[:35].view(7, 5))   # parameters see the reduced gradient
    assert all(opt.touched)
    # a rank whose iteration skipped a branch (rank 1: no gradient for the second layer) must still update the same
    # tensors as everybody else: the touched flags travel with the gradient arena in the one all-reduce
    opt.zero_grad()
    (student(x).sum() if rank == 0 else student[0](x).sum()).backward()
    assert opt.touched == ([True] * 4 if rank == 0 else [True, True, False, False])
    dist.all_reduce(opt.reduce_buffer())
    mask = opt._agreed_segments().view(torch.int32).view(-1, 8)[:, 6]
    assert mask.tolist() == [1, 1, 1, 1], mask.tolist()                       # identical on both ranks
    masks = [None] * world
    dist.all_gather_object(masks, mask.tolist())
    assert masks[0] == masks[1]
    opt.zero_grad()
    student[0](x).sum().backward()                                            # nobody touches layer 2: nobody updates it
    dist.all_reduce(opt.reduce_buffer())
    assert opt._agreed_segments().view(torch.int32).view(-1, 8)[:, 6].tolist() == [1, 1, 0, 0]
    # bucketed exchange (heads' bucket reduced under the trunk's backward, trunk bucket + flags after it): the buckets tile
    # the buffer exactly, and reducing them one by one equals the single all-reduce of the whole buffer
    opt.zero_grad()
    student(x * (rank + 2)).sum().backward()
    early, late = opt.reduce_buckets(head_prefixes=("1.",))
    spans = sorted((t.data_ptr(), t.numel()) for t in early + late)
    assert spans[0][0] == opt._g_full.data_ptr() and sum(n for _, n in spans) == opt._g_full.numel()
    assert all(a + 4 * n == b for (a, n), (b, _) in zip(spans, spans[1:]))
    assert sum(t.numel() for t in early) == 28 and early[0].data_ptr() == student[1].weight.grad.data_ptr()
    whole = opt._g_full.clone()
    pending = [dist.all_reduce(t, async_op=True) for t in early]
    opt.publish_touched()
    whole[opt.total:opt.total + 4] = 1.0
    pending += [dist.all_reduce(t, async_op=True) for t in late]
    for w in pending:
        w.wait()
    dist.all_reduce(whole)
    assert torch.equal(opt._g_full, whole)
    # logging reduce (reference MTtrainer.py:16-41): rank 0 receives the mean; a rank lacking a term sends zero for it
    red = reduce_loss_dict({"loss_a": torch.tensor(float(rank)), "mt_b": torch.tensor(2.0 * rank)})
    if rank == 0:
        assert abs(float(red["loss_a"]) - 0.5) < 1e-6 and abs(float(red["mt_b"]) - 1.0) < 1e-6
    mine = {"loss_a": torch.tensor(1.0)} if rank == 1 else {"loss_a": torch.tensor(1.0), "mt_b": torch.tensor(4.0)}
    red = reduce_loss_dict(mine, keys=["loss_a", "mt_b"])
    if rank == 0:
        assert abs(float(red["loss_a"]) - 1.0) < 1e-6 and abs(float(red["mt_b"]) - 2.0) < 1e-6
    # sharding: disjoint, covering, equal-sized
    ds = list(range(11))
    mine = list(DistributedSampler(ds, shuffle=True))
    all_idx = [None] * world
    dist.all_gather_object(all_idx, mine)
    assert len(mine) == 6 and sorted(set(sum(all_idx, []))) == ds
    synchronize()
    out.put((rank, True))
    dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(out.get(timeout=5)[0] for _ in range(world)) == [0, 1]


def test_shard_indices_padding_and_determinism():
    sys.path.insert(0, ROOT)
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200.data.samplers import IterationBasedBatchSampler, shard_indices
    a = [shard_indices(10, 4, r, epoch=3) for r in range(4)]
    assert all(len(x) == 3 for x in a) and set(sum(a, [])) == set(range(10))
    assert a == [shard_indices(10, 4, r, epoch=3) for r in range(4)]
    assert shard_indices(10, 4, 0, epoch=4) != a[0]
    bs = torch.utils.data.BatchSampler(torch.utils.data.SequentialSampler(range(5)), 2, False)
    it = IterationBasedBatchSampler(bs, 7, start_iter=2)
    assert len(list(it)) == 5 and len(it) == 7
