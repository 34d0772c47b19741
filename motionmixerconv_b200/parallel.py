"""Batch-sharded data parallelism: host-side logic (device-agnostic, so the world_size-2 gloo tests can run it on CPU).

One process per GPU; rank r owns sequences [r*B/W, (r+1)*B/W) of the global batch; every rank holds all parameters,
gradients travel as ONE flat fp32 bucket with a single all-reduce(SUM) per step; the 1/W of the mean is folded into the
fused Adam (``grad_scale``).  There is no other collective on the path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ALIGN = 4   # floats: every parameter view starts 16-byte aligned in the flat buffers


def flat_offsets(numels):
    """Offsets (in floats) of each tensor inside the flat bucket, and the bucket length."""
    offs, o = [], 0
    for n in numels:
        offs.append(o)
        o += (int(n) + ALIGN - 1) // ALIGN * ALIGN
    return offs, o


def shard_bounds(batch, rank, world):
    """[lo, hi) of rank's shard; the global batch must divide evenly (equal shards keep mean-of-means == global mean)."""
    if batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (batch, world))
    per = batch // world
    return rank * per, (rank + 1) * per


def shard(t, rank, world):
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def allreduce_bucket(flat_grads: torch.Tensor, group=None):
    """SUM the flat gradient bucket over the group (NCCL on GPUs, gloo in the CPU tests).  Returns grad_scale = 1/W,
    the factor the optimiser applies."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class _RawDeviceArray:
    """__cuda_array_interface__ shim: lets torch view memory that libmmx allocated (mmx_peer_alloc) as a tensor."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerGradBucket:
    """The flat gradient bucket in IPC-shared device memory + everything ``mmx_adam_step_peer`` needs: the all-reduce of the
    bucket and the Adam update run as ONE kernel over NVLink peer memory (csrc/mmx_api_peer.cu) instead of an NCCL call
    between two CUDA graphs.  One process per GPU, all ranks on one node with peer access.

    ``torch.distributed`` is used once, at construction, to exchange the 64-byte IPC handles (``all_gather_object``);
    the data path has no collective library call.  Raises if a peer cannot be mapped (the caller falls back to NCCL).
    """

    def __init__(self, numel, device, group):
        import ctypes as C
        from . import _lib as L
        self.lib = lib = L.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel = int(numel)
        g_bytes = (self.numel * 4 + 255) // 256 * 256
        flag_bytes = int(lib.mmx_peer_flag_bytes(self.world))
        with torch.cuda.device(device):
            me = torch.cuda.current_device()
            # every step of the set-up is collective: a rank that fails locally still takes part in the exchanges, and all ranks
            # take the same decision at the end (a rank raising alone would leave its peers blocked in a collective)
            err, handle_bytes, self.ptr = None, b"", None
            ptr = C.c_void_p()
            if lib.mmx_peer_alloc(g_bytes + flag_bytes, C.byref(ptr)) != 0:
                err = "mmx_peer_alloc: %s" % (lib.mmx_last_error() or b"?").decode()
            else:
                self.ptr = ptr.value
                handle = C.create_string_buffer(64)
                if lib.mmx_ipc_export(self.ptr, handle) != 0:
                    err = "mmx_ipc_export: %s" % (lib.mmx_last_error() or b"?").decode()
                else:
                    handle_bytes = bytes(handle.raw)
            handles = [None] * self.world
            dist.all_gather_object(handles, (handle_bytes, me, err), group=group)
            peers, self._opened = [], []
            if err is None and any(h[2] for h in handles):
                err = "a peer could not allocate / export its bucket"
            for r, (h, peer_dev, _) in enumerate(handles):
                if err:
                    break
                if r == self.rank:
                    peers.append(self.ptr)
                    continue
                if peer_dev == me:
                    err = "rank %d shares device %d with this rank" % (r, me)
                    break
                q = C.c_void_p()
                if lib.mmx_ipc_open(h, C.byref(q)) != 0:
                    err = "rank %d: %s" % (r, (lib.mmx_last_error() or b"?").decode())
                    break
                peers.append(q.value)
                self._opened.append(q.value)
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.close()
                if self.ptr is not None:
                    lib.mmx_peer_free(self.ptr)
                    self.ptr = None
                raise RuntimeError("PeerGradBucket: peer mapping unavailable (%s)" % (err or "on another rank"))
            self.peer_g = torch.tensor(peers, dtype=torch.int64, device=device)
            self.peer_flags = torch.tensor([q + g_bytes for q in peers], dtype=torch.int64, device=device)
            self.epoch = torch.zeros(2, dtype=torch.int32, device=device)
            self._raw = _RawDeviceArray(self.ptr, self.numel)
            self.g = torch.as_tensor(self._raw, device=device)
            dist.barrier(group=group, device_ids=[me] if dist.get_backend(group) == "nccl" else None)

    def close(self):
        lib = self.lib
        for q in getattr(self, "_opened", []):
            lib.mmx_ipc_close(q)
        self._opened = []

    def adam_step(self, p, m, v, hyper, stream):
        """all-reduce(SUM) of every rank's bucket + Adam on this rank's replica, one kernel (collective)."""
        from . import _lib as L
        L.check(self.lib, self.lib.mmx_adam_step_peer(p.data_ptr(), m.data_ptr(), v.data_ptr(), self.peer_g.data_ptr(), self.peer_flags.data_ptr(),
                                                      self.rank, self.world, self.numel, hyper.data_ptr(), self.epoch.data_ptr(), stream),
                "mmx_adam_step_peer")
