"""Latency of the data-parallel gradient exchange + optimiser step, alone (no forward / backward): the fused peer-memory kernel
(mmx_adam_step_peer) vs an NCCL all-reduce followed by mmx_adam_step, for the K2 (30 K floats) and K4 (183 K floats) buckets.
Two regimes: back-to-back launches (device-resident training loop) and one host synchronisation per iteration (the e2e loop).
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_bench.py"""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200 import _lib as L
from motionmixerconv_b200 import parallel as P_

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = L.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bench(n, iters=200):
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-5, 0.5, 0.5, 1.0 / world, 0.1, 0.001], dtype=torch.float32, device=dev)
    p, m, v = (torch.zeros(n, device=dev) for _ in range(3))
    peer = P_.PeerGradBucket(n, dev, dist.group.WORLD)
    g_nccl = torch.zeros(n, device=dev)

    def run_peer():
        peer.adam_step(p, m, v, hyper, st)

    def run_nccl():
        dist.all_reduce(g_nccl)
        L.check(lib, lib.mmx_adam_step(p.data_ptr(), g_nccl.data_ptr(), m.data_ptr(), v.data_ptr(), n, hyper.data_ptr(), st), "adam")

    out = {}
    for name, fn in (("peer", run_peer), ("nccl", run_nccl)):
        for sync_each in (False, True):
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            dist.barrier(device_ids=[local])
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
                if sync_each:
                    torch.cuda.synchronize()
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e) / iters * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["%s_%s_us" % (name, "hostsync" if sync_each else "backtoback")] = round(float(t), 2)
    peer.close()
    return out


for n in (30048, 182960):
    r = bench(n)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bucket_floats": n, **r, "aborts": lib.mmx_tc5_abort_count()}), flush=True)
dist.barrier(device_ids=[local])
import threading
threading.Timer(15.0, lambda: os._exit(0)).start()
dist.destroy_process_group()
os._exit(0)
