"""The step's single collective as ONE kernel over NVLink peer memory (include/fvae_b200.h: fvae_p2p_*).

The reference has no data parallelism, hence no counterpart; the baseline for this op is one `ncclAllReduce` of the flat
gradient (0.25 - 2.2 MB fp32: pure latency, ~45 us on 8 x B200 against a ~1 ms step).  `P2PAllReduce` maps every rank's
communication buffer into every other rank (CUDA IPC handles exchanged once through torch.distributed) and then reduces with
a single launch of `fvae_p2p_allreduce`: push to the peers, flag, wait, sum in rank order (bit-identical on all ranks).
PyTorch is plumbing here (process group for the one-time handle exchange, stream handles)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _cabi


class P2PAllReduce:
    def __init__(self, n: int, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("P2PAllReduce needs an initialised torch.distributed process group (for the handle exchange)")
        self.group, self.n = group, int(n)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        L = _cabi.lib()
        self.bytes = int(L.fvae_p2p_buffer_bytes(self.n, self.world))
        self._own = C.c_void_p()
        handle = C.create_string_buffer(64)
        with torch.cuda.device(self.device):
            _cabi.check(L.fvae_p2p_alloc(self.bytes, C.byref(self._own), handle), "fvae_p2p_alloc")
            # (device index, handle) of every rank; all ranks must sit on ONE node with peer access (NVLink / NVSwitch)
            mine = (torch.cuda.current_device(), bytes(handle.raw))
            everyone: List[Optional[tuple]] = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            self._peers: List[int] = []
            self._opened: List[C.c_void_p] = []
            for r, (dev_r, h) in enumerate(everyone):
                if r == self.rank:
                    self._peers.append(self._own.value)
                    continue
                if not torch.cuda.can_device_access_peer(self.device.index, dev_r):
                    raise RuntimeError(f"GPU {self.device.index} cannot access GPU {dev_r} as a peer: use the NCCL all-reduce")
                p = C.c_void_p()
                _cabi.check(L.fvae_p2p_open(h, C.byref(p)), "fvae_p2p_open")
                self._opened.append(p)
                self._peers.append(p.value)
            self.peer_table = torch.tensor(self._peers, dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=group)            # nobody reduces before every mapping exists
        self.epoch = 0

    def all_reduce(self, buf: torch.Tensor, scale: float = 1.0) -> None:
        """buf (fp32, contiguous, 16-byte aligned, numel <= n) <- scale * sum over ranks, on the current stream."""
        if buf.dtype != torch.float32 or not buf.is_contiguous() or buf.numel() > self.n or buf.device != self.device:
            raise ValueError("buf must be a contiguous fp32 CUDA tensor of at most n elements on this rank's device")
        self.epoch += 1
        with torch.cuda.device(self.device):
            rc = _cabi.lib().fvae_p2p_allreduce(buf.data_ptr(), buf.numel(), self.peer_table.data_ptr(), self.world, self.rank,
                                                self.epoch & 0xFFFFFFFF or 1, float(scale), self.n,
                                                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _cabi.check(rc, "fvae_p2p_allreduce")

    def close(self) -> None:
        L = _cabi.lib()
        try:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)    # no peer still reads or writes my buffer
        except Exception:
            pass
        for p in self._opened:
            L.fvae_p2p_close(p)
        self._opened = []
        if self._own:
            L.fvae_p2p_free(self._own)
            self._own = C.c_void_p()
