"""Multi-GPU plumbing: one process per GPU, documents sharded d -> d % world, one all-reduce of the
sufficient statistics per outer iteration.  The collective itself is
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests); this module only wraps the
library's device buffers as tensors so they can be reduced in place on the handle's stream.
"""
from __future__ import annotations

from typing import Optional


class _DevBuf:
    """A raw device pointer exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Reducer:
    """Sum-reduces (stats fp32, small fp64) device buffers across the ranks of a torch process group."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist

        assert dist.is_initialized(), "torch.distributed must be initialised (torchrun)"
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._views = {}

    def stream_ptr(self) -> int:
        """cudaStream_t of torch's current stream; the legacy default stream (0) is passed as cudaStreamLegacy (0x1)
        because a NULL stream asks the library for a private one."""
        p = int(self.torch.cuda.current_stream().cuda_stream)
        return p if p != 0 else 1

    def view(self, ptr: int, n: int, typestr: str, device: int):
        key = (ptr, n, typestr)
        if key not in self._views:
            self._views[key] = self.torch.as_tensor(_DevBuf(ptr, n, typestr), device="cuda:%d" % device)
        return self._views[key]

    def allreduce_device(self, bufs, device: int) -> None:
        """bufs: [(ptr, n, typestr)] -- in-place sum over ranks, enqueued on torch's current stream."""
        for ptr, n, typestr in bufs:
            if n:
                self.dist.all_reduce(self.view(ptr, n, typestr, device), op=self.dist.ReduceOp.SUM, group=self.group)

    def all_gather_bytes(self, blob: bytes) -> list:
        """Every rank's `blob` (same length on all ranks), in rank order."""
        t = self.torch.frombuffer(bytearray(blob), dtype=self.torch.uint8)
        nccl = self.dist.get_backend(self.group) == "nccl"
        if nccl:
            t = t.cuda()
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]

    def all_agree(self, ok: bool) -> bool:
        """True iff `ok` on every rank."""
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32)
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    def connect_peers(self, export_fn, connect_fn, blob_bytes: int) -> bool:
        """The handshake of the peer-memory exchange (tmvb_*_comm_export / _comm_connect): all-gather the IPC handle
        blobs, map the peers.  Returns False (on every rank) when any rank cannot map its peers -- the caller then
        keeps the NCCL all-reduce path."""
        import ctypes as C

        ok, blobs = True, None
        try:
            buf = C.create_string_buffer(blob_bytes)
            export_fn(buf, blob_bytes)
            blobs = self.all_gather_bytes(buf.raw)
        except Exception:
            ok = False
            blobs = self.all_gather_bytes(bytes(blob_bytes)) if blobs is None else blobs
        if ok:
            try:
                connect_fn(self.rank, self.world, b"".join(blobs), blob_bytes)
            except Exception:
                ok = False
        return self.all_agree(ok)

    def barrier(self) -> None:
        """All ranks have reached this point (host side)."""
        self.all_agree(True)

    def allreduce_host(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64)
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return float(t.item())


def default_reducer() -> Optional[Reducer]:
    """A Reducer when running under torchrun with world_size > 1, else None."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return Reducer()
    return None


def connect_model_peers(model, prefix: str) -> bool:
    """Peer-memory handshake of a sharded model's handle (tmvb_<prefix>_comm_export / _connect): True when the statistics will be
    summed by the one-kernel peer all-reduce (tmvb_<prefix>_peer_reduce), False when the run keeps torch.distributed all-reduces
    (one GPU, TMVB_P2P=0, or a rank that cannot map its peers)."""
    import os

    from . import _lib

    red = model.reducer
    if red is None or red.world <= 1 or os.environ.get("TMVB_P2P", "1") == "0":
        return False
    lib, h = _lib.load(), model._h
    exp, con = getattr(lib, "tmvb_%s_comm_export" % prefix), getattr(lib, "tmvb_%s_comm_connect" % prefix)
    return red.connect_peers(lambda buf, n: _lib.check(exp(h, buf, n)), lambda rank, world, blobs, n: _lib.check(con(h, rank, world, blobs, n)),
                             _lib.COMM_BLOB_BYTES)
