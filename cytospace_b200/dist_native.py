"""NCCL data plane through the C ABI (``cyb_dist_*``, ``cyb_gather_columns``): the transport
``chunking.solve_chunks`` uses on GPUs.  ``torch.distributed`` is only the bootstrap (it carries the 128-byte
communicator id and the pickled chunk plan, host metadata); every expression block and index vector moves with
the library's own NCCL calls on the engine's stream.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native


class _Done:
    """Completion handle of a stream-ordered NCCL call."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        self.event.synchronize()


class NativeTransport:
    def __init__(self, engine, rank: int, world: int, id_bytes: bytes, bootstrap=None):
        self.engine, self.rank, self.world, self.bootstrap = engine, rank, world, bootstrap
        self.lib, self.ffi = _native.load(), _native.ffi()
        comm = self.ffi.new("void **")
        with torch.cuda.device(engine.device):
            _native.check(self.lib.cyb_dist_init(self.ffi.from_buffer(id_bytes), world, rank, comm))
        self.comm = comm[0]
        self.send_stream = None

    @classmethod
    def from_torch_group(cls, engine, group=None):
        """Bootstrap over an initialised ``torch.distributed`` group: rank 0 creates the id, everybody joins."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        lib, ffi = _native.load(), _native.ffi()
        box = [None]
        if rank == 0:
            buf = bytearray(lib.CYB_DIST_ID_BYTES)
            _native.check(lib.cyb_dist_unique_id(ffi.from_buffer(buf)))
            box[0] = bytes(buf)
        src = 0 if group is None else dist.get_global_rank(group, 0)
        dist.broadcast_object_list(box, src=src, group=group)
        return cls(engine, rank, world, bytearray(box[0]), bootstrap=(dist, group))

    def _stream(self):
        return _native.stream_ptr(torch.cuda.current_stream(self.engine.device))

    def _event(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.engine.device))
        return _Done(ev)

    def bcast_object(self, obj, root=0):
        dist, group = self.bootstrap
        box = [obj]
        dist.broadcast_object_list(box, src=root if group is None else dist.get_global_rank(group, root), group=group)
        return box[0]

    def broadcast(self, t, root=0):
        assert t.is_contiguous()
        with torch.cuda.device(self.engine.device):
            _native.check(self.lib.cyb_dist_broadcast(self.comm, _native.ptr("void *", t), t.numel() * t.element_size(),
                                                      root, self._stream()))

    def isend(self, t, dst):
        """Asynchronous with respect to the compute stream: the send runs on a side stream behind everything queued
        so far, so that the column gather of the next block overlaps it."""
        assert t.is_contiguous()
        cur = torch.cuda.current_stream(self.engine.device)
        if self.send_stream is None:
            self.send_stream = torch.cuda.Stream(device=self.engine.device)
        self.send_stream.wait_stream(cur)
        with torch.cuda.device(self.engine.device):
            _native.check(self.lib.cyb_dist_send(self.comm, _native.ptr("void *", t), t.numel() * t.element_size(), dst,
                                                 _native.stream_ptr(self.send_stream)))
        t.record_stream(self.send_stream)
        ev = torch.cuda.Event()
        ev.record(self.send_stream)
        return _Done(ev)

    def send_group(self, blocks):
        """``blocks`` = [(tensor, dst), ...] sent as ONE NCCL group on the side stream: transfers to different owners
        progress concurrently (a single peer-to-peer pair does not fill rank 0's NVLink egress)."""
        cur = torch.cuda.current_stream(self.engine.device)
        if self.send_stream is None:
            self.send_stream = torch.cuda.Stream(device=self.engine.device)
        self.send_stream.wait_stream(cur)
        with torch.cuda.device(self.engine.device):
            _native.check(self.lib.cyb_dist_group_start())
            for t, dst in blocks:
                assert t.is_contiguous()
                _native.check(self.lib.cyb_dist_send(self.comm, _native.ptr("void *", t), t.numel() * t.element_size(), dst,
                                                     _native.stream_ptr(self.send_stream)))
            _native.check(self.lib.cyb_dist_group_end())
        for t, _dst in blocks:
            t.record_stream(self.send_stream)
        ev = torch.cuda.Event()
        ev.record(self.send_stream)
        return _Done(ev)

    def recv(self, t, src):
        assert t.is_contiguous()
        with torch.cuda.device(self.engine.device):
            _native.check(self.lib.cyb_dist_recv(self.comm, _native.ptr("void *", t), t.numel() * t.element_size(), src,
                                                 self._stream()))

    def all_gather(self, send):
        out = torch.empty((self.world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        with torch.cuda.device(self.engine.device):
            _native.check(self.lib.cyb_dist_all_gather(self.comm, _native.ptr("void *", send), _native.ptr("void *", out),
                                                       send.numel() * send.element_size(), self._stream()))
        return [out[r] for r in range(self.world)]

    def gather_columns(self, x, idx):
        """x[:, idx] on the device through ``cyb_gather_columns``."""
        cols = torch.from_numpy(np.asarray(idx).astype(np.int32)).to(x.device)
        out = torch.empty((x.shape[0], cols.numel()), dtype=x.dtype, device=x.device)
        dt = self.lib.CYB_F64 if x.dtype == torch.float64 else self.lib.CYB_F32
        with torch.cuda.device(self.engine.device):
            _native.check(self.lib.cyb_gather_columns(_native.ptr("void *", x), dt, x.shape[0], x.stride(0),
                                                      _native.ptr("int32_t *", cols), cols.numel(),
                                                      _native.ptr("void *", out), out.stride(0), self._stream()))
        return out

    def close(self):
        if self.comm is not None:
            torch.cuda.synchronize(self.engine.device)
            _native.check(self.lib.cyb_dist_destroy(self.comm))
            self.comm = None
