"""Data parallelism over the 8 GPUs of one box: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

Clips are independent units: inference shards clips with no communication; training has exactly one
exchange step per iteration -- the sum all-reduce of the flat gradient arena (M-B 0.76 MB, M-C 73 KB, M-A 18 MB fp32).
The arena is contiguous, so there is no pack/unpack; its 16-float header rides along, which all-reduces the
"non-finite loss" flag and the group-activity flags so every rank takes the same skip decisions.  The 1/world scaling
is folded into the fused clip+Adam kernel (``grad_scale``).  BatchNorm statistics and M-B's batch-coupled loss terms
are rank-local (DDP semantics): an 8x4 run is not numerically a 1x32 run.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, group=None, buckets: int = 1):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.buckets = max(1, int(buckets))

    def attach(self, optimizer):
        optimizer.pre_step_hook = self.all_reduce_grads
        optimizer.grad_scale = 1.0 / self.world
        self._early_work = None
        return optimizer

    # -- overlapped exchange (graphs.graphed_optimizer_step): the gradient arena is [header | late range | early range]; "early" = the
    # tensors whose gradients are complete first (M-A: everything behind the backbone -- detector, tail, direct classifier: 26 of the
    # 31 MB), "late" = the rest plus the header flags.  The early bucket is reduced on NCCL's stream while the late backward still runs.
    def set_split(self, early_offset: int):
        """Arena element offset where the early range starts (a multiple of the arena block)."""
        self.early_offset = int(early_offset)

    def start_early(self, arena):
        if self.world == 1:
            return
        off = getattr(self, "early_offset", 0)
        if 0 < off < arena.g.numel():
            self._early_work = dist.all_reduce(arena.g[off:], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish_late(self, arena):
        if self.world == 1:
            return
        off = getattr(self, "early_offset", 0)
        if self._early_work is None:          # no split configured: one collective over everything
            dist.all_reduce(arena.g, op=dist.ReduceOp.SUM, group=self.group)
            return
        dist.all_reduce(arena.g[:off], op=dist.ReduceOp.SUM, group=self.group)
        self._early_work.wait()
        self._early_work = None

    def all_reduce_grads(self, arena):
        if self.world == 1:
            return
        g = arena.g
        if self.buckets == 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            return
        n = g.numel()
        step = (n + self.buckets - 1) // self.buckets
        works = [dist.all_reduce(g[i:i + step], op=dist.ReduceOp.SUM, group=self.group, async_op=True) for i in range(0, n, step)]
        for w in works:
            w.wait()

    def broadcast_parameters(self, arena, src: int = 0, model=None):
        """Make every replica start from rank ``src``'s state: the parameter / Adam-moment arenas, the device-side optimizer state
        (per-group Adam step counters, skip counter: they matter after a load_state_dict on one rank) and -- when ``model`` is given --
        every parameter and buffer that does NOT live in the arena (frozen tensors such as M-A's stem, BatchNorm running statistics
        and num_batches_tracked).  torch's DistributedDataParallel broadcasts buffers the same way."""
        dist.broadcast(arena.p, src=src, group=self.group)
        dist.broadcast(arena.m, src=src, group=self.group)
        dist.broadcast(arena.v, src=src, group=self.group)
        dist.broadcast(arena.state, src=src, group=self.group)
        arena._device_lr = None              # the broadcast overwrote lr_device: re-publish on the next sync_lr_to_device()
        if model is not None:
            lo, hi = arena.p.data_ptr(), arena.p.data_ptr() + arena.p.numel() * arena.p.element_size()
            for t in list(model.parameters()) + list(model.buffers()):
                if lo <= t.data_ptr() < hi:
                    continue                 # a view into the arena: already sent
                dist.broadcast(t.data, src=src, group=self.group)

    def shard(self, n_items: int):
        """Contiguous shard [lo, hi) of n_items for this rank (inference: disjoint clip shards, no communication)."""
        per = (n_items + self.world - 1) // self.world
        lo = min(self.rank * per, n_items)
        return lo, min(lo + per, n_items)


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local: int, sysfs: str = "/sys") -> int | None:
    """Pin this process (and the pinned host buffers it allocates afterwards: first touch) to the NUMA node its GPU hangs off.

    With one process per GPU every rank streams its own batch from pinned host memory each step (177 MB of fp32 frames for M-A);
    when all eight ranks' buffers land on one socket, that socket's memory and the inter-socket link become the bottleneck.  Reads
    the node from sysfs (``/sys/bus/pci/devices/<bus id>/numa_node``); does nothing when the topology cannot be read, the box has
    one node, or ``CVAD_NUMA_BIND=0``.  Returns the node it bound to, or None."""
    if os.environ.get("CVAD_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        props = torch.cuda.get_device_properties(local)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0 or not os.path.isdir(os.path.join(sysfs, "devices/system/node/node1")):
            return None
        with open(os.path.join(sysfs, f"devices/system/node/node{node}/cpulist")) as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:      # topology files missing / unreadable: stay unbound
        return None


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            bind_to_gpu_numa_node(local)
            # binding the group to its device creates the communicator right here (all ranks together) and lets barrier()
            # use it instead of guessing a device
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world
