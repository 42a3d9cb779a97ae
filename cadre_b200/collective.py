"""In-switch gradient all-reduce over NVLink / NVSwitch (cadre_allreduce_* in include/cadre_b200.h).

Replaces Shared_grad_buffers.add_gradient + the chief installing the summed gradient (ppo_agent/models.py:231-239,
ppo_agent/chief.py:13-16) for one process per GPU. `torch.distributed._symmetric_memory` is used for what it is:
plumbing - it allocates the gradient buffer so that every rank can map every other rank's copy (and one multicast
address for all of them) and exchanges the handles; the reduction itself is the library's `allreduce_kernel`
(multimem.ld_reduce / multimem.st through the switch, csrc/allreduce.cu).
"""
import ctypes

import torch

from . import _lib
from ._lib import CadreError


class SwitchAllReduce:
    """Symmetric fp32 buffer of `count` floats + the in-place sum over the ranks of `group`.

    All ranks must construct it collectively and issue the same sequence of `sum_()` calls."""

    def __init__(self, count, device, group=None, multicast=True):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_available() and dist.is_initialized()):
            raise CadreError("SwitchAllReduce needs an initialised torch.distributed process group")
        self.device = torch.device(device)
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._lib = _lib.lib()
        flag_words = self._lib.cadre_allreduce_flag_bytes() // 4
        with torch.cuda.device(self.device):
            self.buffer = symm_mem.empty(int(count), dtype=torch.float32, device=self.device)
            self._flags = symm_mem.empty(flag_words, dtype=torch.int32, device=self.device)
            self.buffer.zero_()
            self._flags.zero_()
            torch.cuda.synchronize(self.device)
            self._hdl = symm_mem.rendezvous(self.buffer, group)
            self._flag_hdl = symm_mem.rendezvous(self._flags, group)
            dist.barrier(group)      # every rank has zeroed its flags before anyone signals
            mc = int(self._hdl.multicast_ptr) if multicast else 0
            self.multicast = mc != 0
            bufs = (ctypes.c_void_p * self.world)(*[int(p) for p in self._hdl.buffer_ptrs])
            flags = (ctypes.c_void_p * self.world)(*[int(p) for p in self._flag_hdl.buffer_ptrs])
            h = ctypes.c_void_p()
            _lib.check(self._lib.cadre_allreduce_create(ctypes.byref(h), self.rank, self.world, bufs,
                                                        ctypes.c_void_p(mc or None), flags,
                                                        ctypes.c_int64(int(count))))
        self._h = h

    def sum_(self, offset=0, count=None, stream=None):
        """buffer[offset:offset+count] <- sum over ranks (in place, asynchronous on `stream` / the current stream)."""
        count = self.buffer.numel() - offset if count is None else count
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_allreduce_sum(self._h, ctypes.c_int64(int(offset)), ctypes.c_int64(int(count)),
                                                     1 if self.multicast else 0, ctypes.c_void_p(s.cuda_stream)))

    def set_blocks(self, blocks):
        """Thread blocks per launch (the same value on every rank)."""
        _lib.check(self._lib.cadre_allreduce_set_blocks(self._h, int(blocks)))

    def check(self):
        """Synchronise and raise if a barrier ever timed out."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.cadre_allreduce_check(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.cadre_allreduce_destroy(h)
            except Exception:
                pass
