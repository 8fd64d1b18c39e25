"""Host-side logic of the N>1 path on CPU: world_size-2 gloo runs of the flat-arena gradient all-reduce, the header-flag
agreement, parameter broadcast and clip sharding (the same code drives NCCL on the GPUs), plus the dataset drop-in."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeArena:
    """The four flat buffers of arena.FlatArena, on the CPU (the real one needs CUDA parameters)."""

    def __init__(self, n, rank):
        self.g = torch.full((n,), float(rank + 1))
        self.g[:16] = 0.0
        self.p = torch.full((n,), float(10 + rank))
        self.m = torch.full((n,), float(20 + rank))
        self.v = torch.full((n,), float(30 + rank))
        self.state = torch.full((96,), rank + 1, dtype=torch.uint8)      # device-side optimizer state (step counters, skip counter)


class _FakeOpt:
    pre_step_hook = None
    grad_scale = 1.0


def _worker(rank, world, port, buckets, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "cvad_parallel", os.path.join(ROOT, "causal-learning-based-video-anomaly-detection_paper_code_raw_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    r, l, w = par.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    dp = par.DataParallel(buckets=buckets)
    arena = _FakeArena(4096 + 1024, rank)
    if rank == 1:
        arena.g[0] = 1.0          # rank 1 saw a non-finite loss
        arena.g[2] = 1.0          # ... and its structure-learner group received a gradient
    opt = dp.attach(_FakeOpt())
    assert opt.grad_scale == 1.0 / world
    opt.pre_step_hook(arena)
    ok = bool((arena.g[16:] == sum(range(1, world + 1))).all())
    ok &= float(arena.g[0]) == 1.0 and float(arena.g[2]) == 1.0 and float(arena.g[1]) == 0.0   # every rank sees the flags
    # a model with one arena-managed parameter (a view into arena.p), one frozen parameter outside the arena (M-A's stem) and
    # BatchNorm buffers: after the broadcast the FULL state_dict must agree across ranks (what torch DDP guarantees too)
    model = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4))
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            t.fill_(rank + 1)
    model[0].weight.data = arena.p[1024:1040].view(4, 4)
    model[0].bias.requires_grad = False
    dp.broadcast_parameters(arena, src=0, model=model)
    ok &= bool((arena.p == 10).all() and (arena.m == 20).all() and (arena.v == 30).all() and (arena.state == 1).all())
    ok &= all(bool((v == (10 if k == "0.weight" else 1)).all()) for k, v in model.state_dict().items())
    lo, hi = dp.shard(11)
    q.put((rank, ok, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def _run(buckets):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, buckets, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 6, 6, 11)      # disjoint, covering clip shards


def test_gloo_world2_single_bucket():
    _run(1)


def test_gloo_world2_bucketed():
    _run(3)


def test_dataset_dropin_synthetic():
    sys.path.insert(0, ROOT)
    import avenue_dataset_usage as a
    tr, te = a.create_avenue_dataloaders("synthetic:10", batch_size=4, clip_length=8, frame_size=(64, 64))
    x, y = next(iter(tr))
    assert x.shape == (4, 3, 8, 64, 64) and x.dtype == torch.float32 and y.shape == (4,)
    assert 0.0 <= float(x.min()) and float(x.max()) <= 1.0
    assert sum(b[0].shape[0] for b in tr) == 10 and len(te.dataset) >= 4


def test_numa_binding_reads_sysfs_topology(tmp_path, monkeypatch):
    """bind_to_gpu_numa_node: the GPU's PCI address -> numa_node -> that node's CPU list (intersected with what the process may use);
    unreadable topology or a single-node box leaves the process unbound."""
    import types

    import torch
    from cvad_b200 import parallel

    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 2:
        pytest.skip("needs two usable CPUs")
    half = allowed[len(allowed) // 2:]
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    for n, cpus in ((0, allowed[:len(allowed) // 2]), (1, half)):
        d = tmp_path / f"devices/system/node/node{n}"
        d.mkdir(parents=True)
        (d / "cpulist").write_text(",".join(str(c) for c in cpus) + "\n")
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0))
    try:
        assert parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) == 1
        assert sorted(os.sched_getaffinity(0)) == half
    finally:
        os.sched_setaffinity(0, allowed)
    (dev / "numa_node").write_text("-1\n")                       # no NUMA information for the device
    assert parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    monkeypatch.setenv("CVAD_NUMA_BIND", "0")
    (dev / "numa_node").write_text("1\n")
    assert parallel.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert sorted(os.sched_getaffinity(0)) == allowed
