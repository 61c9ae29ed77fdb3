"""All-reduce of the batch-sharded evaluation over NVLink peer memory (``csrc/collective.cu``).

One process per GPU (``torch.distributed``, NCCL for the rendezvous).  Each rank's evaluation writes its partial
``[gradient ; objective]`` vector into a symmetric-memory buffer (``torch.distributed._symmetric_memory``: the same
allocation mapped into every peer's address space).  With NVSwitch multicast (NVLS) ``opl_allreduce_sum_nvls_device``
reduces each rank's 1/world slice in the switch (``multimem.ld_reduce``) and broadcasts it (``multimem.st``); without
it ``opl_allreduce_sum_device`` pulls every rank's vector over NVLink and adds them in rank order.  One kernel either way.  If the symmetric allocation is not available (no peer access,
a backend without it) the caller falls back to ``torch.distributed.all_reduce`` -- also a device path, never a host one.
"""
import ctypes
import os

import torch

from . import _device as dev
from . import _native as nat

FLAG_DOUBLES = 16         # 2 x 16 uint32 flag slots behind the vector buffers
MAX_WORLD = 16
MAX_COUNT = 1 << 20       # one-shot pull: every rank reads every peer's whole vector; beyond this a ring is cheaper


class PeerAllReduce(object):
    """Double-buffered symmetric vectors of ``count`` doubles plus the per-writer flag slots."""

    def __init__(self, count):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > MAX_WORLD:
            raise RuntimeError("peer all-reduce supports up to %d ranks" % MAX_WORLD)
        self.count = int(count)
        self.pad = (self.count + 1) & ~1                      # keeps the second buffer 16-byte aligned
        total = 4 * self.pad + FLAG_DOUBLES                   # in[2], result[2], flags
        self.sym = symm_mem.empty(total, dtype=torch.float64, device="cuda")
        self.sym.zero_()
        self.handle = symm_mem.rendezvous(self.sym, dist.group.WORLD)
        self.peers = [self.handle.get_buffer(r, (total,), torch.float64) for r in range(self.world)]
        self.base = [t.data_ptr() for t in self.peers]
        self.seq = 0
        want = os.environ.get("OPL_ALLREDUCE", "peer")
        multicast = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.mode = "nvls" if (multicast and want in ("peer", "nvls")) else "pull"
        if want == "nvls" and not multicast:
            raise RuntimeError("OPL_ALLREDUCE=nvls but this system offers no multicast address")
        # multicast address of element 0 of the symmetric tensor (the tensor may sit at an offset inside its allocation)
        self.mc_base = multicast + (self.sym.data_ptr() - int(self.handle.buffer_ptrs[self.rank])) if multicast else 0
        self.counter = torch.zeros(1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        dist.barrier()                                        # every rank's flags are zero before anyone announces

    def next_slot(self):
        """View of this rank's buffer for the NEXT reduction; the evaluation writes into it."""
        parity = (self.seq + 1) & 1
        return self.sym[parity * self.pad: parity * self.pad + self.count]

    def reduce(self, out):
        """out = sum over ranks (rank order) of the vectors written into ``next_slot()``."""
        self.seq += 1
        parity = self.seq & 1
        u64 = ctypes.c_uint64 * self.world
        flags = u64(*[b + 8 * 4 * self.pad for b in self.base])
        if self.mode == "nvls":
            res_off = 8 * (2 + parity) * self.pad
            nat.check(nat.lib.opl_allreduce_sum_nvls_device(
                self.world, self.rank, self.mc_base + 8 * parity * self.pad, self.mc_base + res_off,
                ctypes.c_void_p(self.base[self.rank] + res_off), flags, self.count, self.seq & 0xffffffff, dev.ptr(out),
                dev.ptr(self.counter), dev.stream_ptr()))
            return out
        bufs = u64(*[b + 8 * parity * self.pad for b in self.base])
        nat.check(nat.lib.opl_allreduce_sum_device(self.world, self.rank, bufs, flags, self.count, self.seq & 0xffffffff,
                                                   dev.ptr(out), dev.stream_ptr()))
        return out


def make(count):
    """A PeerAllReduce for vectors of ``count`` doubles, or None when NCCL should do the reduction."""
    want = os.environ.get("OPL_ALLREDUCE", "peer")       # peer (NVLS if available, else pull) | nvls | pull | nccl
    if want not in ("peer", "nvls", "pull") or count > MAX_COUNT:
        return None
    try:
        return PeerAllReduce(count)
    except Exception as exc:       # no symmetric memory on this system: NCCL does it
        import warnings
        warnings.warn("peer-memory all-reduce unavailable (%s); using NCCL" % (exc,))
        return None
