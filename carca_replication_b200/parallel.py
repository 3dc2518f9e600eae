"""Data parallelism over users: one process per GPU, torch.distributed (NCCL over
NVLink; gloo in the CPU tests) for the only two exchanges the path has:

  * training — ONE all-reduce per step of a flat fp32 bucket holding every parameter gradient,
    plus the 2-float BCE partial sums so that each rank normalises its loss by the GLOBAL mask
    count (the reference divides by sum(mask) of the whole batch, src/carca.py:443; per-rank means
    would weight users differently);
  * evaluation — ONE all-reduce of the 5 metric accumulators at the end of evaluate().

Users are independent units (nothing in CARCA.forward mixes batch rows), so no other collective
exists: weights, the item table and the attribute table are replicated.  The reference has no
counterpart (single device); this wraps the model so the reference-style loop stays unchanged:

    dp = UserDataParallel(model)                 # after dist.init_process_group
    train(dp.module, dp.shard_loader(...), ..., loss_fn=dp.loss_fn, is_main=dp.rank == 0)
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from .carca import BinaryCrossEntropy


class PeerAllReduce:
    """Sum of a flat fp32 CUDA tensor over the ranks of one NVLink / NVSwitch box in ONE kernel on peer-mapped memory
    (csrc/peer.cu: copy-in, per-block handshake, owner-reduces-and-pushes, handshake, copy-out) — the gradient
    exchange of the data-parallel step without NCCL in the step: one graph node, bit-identical sums on every rank.
    torch.distributed is only the bootstrap here (the CUDA IPC handles travel through all_gather_object).

    Collective constructor: every rank of `group` must build it with the same capacity.  `available` is False (and
    all_reduce_ raises) when the buffers could not be mapped, e.g. ranks on different nodes; callers fall back to
    dist.all_reduce then."""

    def __init__(self, capacity_floats: int, device, process_group=None):
        import ctypes as C

        from . import _native as N

        self.group = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        self.capacity = int(capacity_floats)
        self.device = torch.device(device)
        self.available = False
        self.error: Optional[str] = None
        self._base = C.c_void_p()
        self._peers: List[Optional[int]] = [None] * self.world
        self._bases = None
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        handle = None
        try:
            if self.world > 8 or self.device.type != "cuda":
                raise RuntimeError("needs 2..8 CUDA ranks of one box")
            with torch.cuda.device(self.device):
                nbytes = int(N.lib().carca_peer_buffer_bytes(self.capacity))
                N.call("carca_peer_alloc", C.byref(self._base), nbytes)
                buf = C.create_string_buffer(64)
                N.call("carca_peer_export", self._base, buf)
                handle = bytes(buf.raw)
        except (RuntimeError, OSError) as ex:
            self.error = f"{type(ex).__name__}: {ex}"[:200]
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, handle, group=process_group)        # collective even when a rank failed
        ok = self.error is None and all(h is not None for h in handles)
        if ok:
            try:
                with torch.cuda.device(self.device):
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            self._peers[r] = self._base.value
                            continue
                        ptr = C.c_void_p()
                        N.call("carca_peer_open", C.create_string_buffer(h, 64), C.byref(ptr))
                        self._peers[r] = ptr.value
                self._bases = (C.c_void_p * self.world)(*self._peers)
            except (RuntimeError, OSError) as ex:
                self.error = f"{type(ex).__name__}: {ex}"[:200]
                ok = False
        flags: List[Optional[bool]] = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=process_group)         # all ranks or none (also the setup barrier)
        self.available = all(flags)
        if not self.available and self.error is None:
            self.error = "a peer rank could not map the buffers"

    def all_reduce_(self, t: Tensor) -> Tensor:
        from . import _native as N

        if not self.available:
            raise RuntimeError(f"PeerAllReduce is not available: {self.error}")
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda or t.numel() > self.capacity:
            raise RuntimeError("PeerAllReduce.all_reduce_: needs a contiguous float32 CUDA tensor within the capacity")
        N.call("carca_peer_allreduce", t.data_ptr(), t.numel(), self._bases, self.rank, self.world, N.i32p(self.status),
               N.stream())
        return t

    def buffer(self, n_floats: int) -> Tensor:
        """A float32 tensor of n elements that LIVES in this rank's communication buffer: all_reduce_ of it is zero-copy
        (peers read it and write the sums straight into it).  One tensor at a time: every call returns the same memory."""
        if not self.available or n_floats > self.capacity:
            raise RuntimeError("PeerAllReduce.buffer: not available or beyond the capacity")
        from . import _native as N

        class _Region:          # torch.as_tensor aliases foreign device memory through the CUDA array interface
            pass

        reg = _Region()
        reg.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "version": 2, "strides": None,
                                        "data": (self._base.value + int(N.lib().carca_peer_data_offset()), False)}
        reg.owner = self        # the buffer outlives every tensor carved from it
        return torch.as_tensor(reg, device=self.device)

    def data_ptr(self) -> int:
        from . import _native as N

        return self._base.value + int(N.lib().carca_peer_data_offset())

    def timed_out(self) -> bool:
        """True if a wait inside an all-reduce gave up (a peer did not arrive within ~20 s).  Host sync."""
        return bool(int(self.status.item()) & 8)

    def close(self) -> None:
        from . import _native as N

        if self._base.value is None:
            return
        torch.cuda.synchronize(self.device)
        for r, ptr in enumerate(self._peers):
            if ptr is not None and r != self.rank:
                N.lib().carca_peer_close(ptr)
        N.lib().carca_peer_free(self._base)
        self._base.value, self._peers, self.available = None, [None] * self.world, False


class ShardedLoader:
    """Wraps a loader of GLOBAL batches; yields this rank's slice of each (keeps len() for train())."""

    def __init__(self, dp: "UserDataParallel", loader: Iterable):
        self.dp, self.loader = dp, loader

    def __len__(self) -> int:
        return len(self.loader)

    def __iter__(self):
        for batch in self.loader:
            yield self.dp.shard(batch)


class UserDataParallel:
    def __init__(self, module: torch.nn.Module, process_group=None, broadcast: bool = True, peer_allreduce: bool = True):
        if not dist.is_initialized():
            raise RuntimeError("UserDataParallel needs torch.distributed to be initialised")
        self.module = module
        self.group = process_group
        self.rank = dist.get_rank(process_group)
        self.world = dist.get_world_size(process_group)
        self.params: List[Tensor] = [p for p in module.parameters() if p.requires_grad]
        self._sizes = [p.numel() for p in self.params]
        self._bucket: Optional[Tensor] = None
        self._queued = False
        # gradient exchange over peer memory (one kernel, no NCCL in the step) where the ranks share an NVLink box
        self.peer: Optional[PeerAllReduce] = None
        self._peer_small: Optional[PeerAllReduce] = None
        dev = self.params[0].device if self.params else torch.device("cpu")
        if peer_allreduce and self.world > 1 and dev.type == "cuda":
            from . import _native

            if not _native.is_emulated():
                n = sum(self._sizes)
                self.peer = PeerAllReduce(2 * n + 64, dev, process_group)
                self._peer_small = PeerAllReduce(256, dev, process_group)
                if not (self.peer.available and self._peer_small.available):
                    self.peer = self._peer_small = None
                else:
                    from . import ops

                    ops.FLAT_GRAD_ALLOC = self._grad_buffer_in_peer_memory
        self.loss_fn = BinaryCrossEntropy()
        self.loss_fn.reduce_sums = self._allreduce_sum
        from . import ops

        ops.set_seed_salt(self.rank)          # independent dropout masks per rank
        if broadcast:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._on_grad)

    # -------------------------------------------------------------- gradient exchange
    def _on_grad(self, _param: Tensor) -> None:
        """First gradient of a backward pass queues the end-of-backward bucket all-reduce."""
        if not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._reduce_grads)

    def _reduce_grads(self) -> None:
        self._queued = False
        if self.world == 1:
            return
        dev = self.params[0].device
        flat = self._shared_grad_buffer()
        if flat is not None:
            # the fused training step (ops.TrainCoreFn) produces every gradient inside one flat buffer:
            # all-reduce it in place — no bucket copies
            self._sum(flat)
            return
        if self._bucket is None or self._bucket.device != dev:
            self._bucket = torch.empty(sum(self._sizes), dtype=torch.float32, device=dev)
        views = self._bucket.split(self._sizes)
        for p, v in zip(self.params, views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad.reshape(-1))
        self._sum(self._bucket)                                                  # the one collective
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.reshape(p.shape).clone()
            else:
                p.grad.copy_(v.reshape(p.shape))

    def _shared_grad_buffer(self) -> Optional[Tensor]:
        """The flat tensor spanning all gradients when they are views of one storage (else None)."""
        grads = [p.grad for p in self.params]
        if any(g is None or not g.is_contiguous() or g.dtype != torch.float32 for g in grads):
            return None
        storage = grads[0].untyped_storage()
        base = storage.data_ptr()
        if any(g.untyped_storage().data_ptr() != base for g in grads[1:]):
            return None
        lo = min(g.storage_offset() for g in grads)
        hi = max(g.storage_offset() + g.numel() for g in grads)
        if hi - lo > 2 * sum(self._sizes):                  # mostly foreign data: not worth reducing
            return None
        return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(storage, lo, (hi - lo,))

    def _grad_buffer_in_peer_memory(self, n_floats: int, device, params=()) -> Optional[Tensor]:
        """ops._FlatZeros hook: the fused training step's flat gradient buffer is carved out of the communication
        buffer, so its all-reduce needs no staging copy.  Not while a parameter still holds gradients of an earlier
        backward in that memory (gradient accumulation over several backwards): then the step gets ordinary memory."""
        if self.peer is None or torch.device(device) != self.peer.device or n_floats > self.peer.capacity:
            return None
        mine = {p.data_ptr() for p in self.params}
        if not params or any(t.data_ptr() not in mine for t in params):
            return None                       # a backward of some other model in this process
        base = self.peer.data_ptr()
        for p in self.params:
            if p.grad is not None and p.grad.untyped_storage().data_ptr() == base:
                return None
        return self.peer.buffer(n_floats)

    def _sum(self, t: Tensor) -> None:
        """In-place sum over ranks: the peer-memory kernel when the ranks share an NVLink box, else NCCL / gloo."""
        if self.peer is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0:
            comm = self._peer_small if t.numel() <= self._peer_small.capacity else self.peer
            if t.numel() <= comm.capacity:
                comm.all_reduce_(t)
                return
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _allreduce_sum(self, t: Tensor) -> None:
        if self.world > 1:
            self._sum(t)

    # -------------------------------------------------------------- sharding helpers
    def shard(self, batch: Sequence[Optional[Tensor]]) -> tuple:
        """This rank's contiguous slice of a global batch (dim 0), B/G users each."""
        out = []
        for t in batch:
            if t is None:
                out.append(None)
                continue
            n = t.shape[0]
            if n < self.world:
                # an empty slice would put B = 0 into the kernels of one rank while the others wait in the
                # gradient all-reduce: refuse instead (use drop_last / a batch size >= the number of ranks)
                raise ValueError(f"UserDataParallel.shard: a batch of {n} users cannot be split over {self.world} ranks")
            lo, hi = (n * self.rank) // self.world, (n * (self.rank + 1)) // self.world
            out.append(t[lo:hi])
        return tuple(out)

    def shard_loader(self, loader: Iterable) -> "ShardedLoader":
        return ShardedLoader(self, loader)

    # -------------------------------------------------------------- evaluation
    def evaluate(self, loader: Iterable, device, k: int, sharded: bool = False):
        """evaluate() over the global loader: each rank scores its slice of every batch, the
        accumulators are all-reduced once.  Returns the same triple as src/train.py:35-53 except
        that the loss is the mean over (rank, batch) pairs."""
        from .train import evaluate

        it = loader if sharded else self.shard_loader(loader)
        return evaluate(self.module, it, device, k, reduce_fn=self._allreduce_sum)


def bind_host_to_gpu_numa_node(device_index: int) -> dict:
    """One process per GPU: keep this process (and the pinned host buffers it allocates afterwards, first-touch) on the
    NUMA node the GPU's PCIe root hangs off, so that its per-step host->device copies do not cross the socket
    interconnect.  Reads the node from sysfs; a machine that does not expose one (numa_node = -1, e.g. a VM) is left
    alone.  Returns what was found / done."""
    import os

    info = {"gpu": int(device_index), "numa_node": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"], info["cpus"] = True, len(cpus)
    except (OSError, ValueError, AttributeError, RuntimeError) as ex:
        info["error"] = f"{type(ex).__name__}: {ex}"[:120]
    return info
