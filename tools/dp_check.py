"""Data-parallel parity on real GPUs (torchrun, NCCL): after a few steps every rank must hold bit-identical parameters, and
they must equal (to fp32 rounding) a single-GPU run over the full batch.  Prints progress lines so a hang is attributable.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionmixerconv_b200.mlp_mixer import MlpMixer
from motionmixerconv_b200.train import TrainStep
from tests.synthetic import synthetic_pose_windows


def log(*a):
    print("[rank %s %.1fs]" % (os.environ.get("RANK", "0"), time.time() - T0), *a, file=sys.stderr, flush=True)


T0 = time.time()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
log("process group up")
cfg = dict(num_classes=66, num_blocks=2, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10, pred_len=10, activation="mish",
           regularization=0, input_size=66, r_se=8, use_se=True)
prec = os.environ.get("PREC", "tf32")
Bper = int(os.environ.get("BPER", 96))
x, gt = synthetic_pose_windows(Bper * world, 10, 10, 66, scale="h36m", seed=5)
xs, gts = torch.from_numpy(x).to(dev), torch.from_numpy(gt).to(dev)
steps = 4
ok = True
for use_graph in (False, True):
    torch.manual_seed(100 + rank)                    # DIFFERENT initial weights per rank: the constructor broadcast must fix that
    model = MlpMixer(**cfg).to(dev).train().set_precision(prec)
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, process_group=dist.group.WORLD, use_cuda_graph=use_graph)
    log("TrainStep built (graph=%s, gradient exchange: %s)" % (use_graph, "peer-memory all-reduce fused into Adam" if ts.peer is not None else "NCCL all-reduce"))
    lo, hi = rank * Bper, (rank + 1) * Bper
    for i in range(steps):
        loss = ts.step(xs[lo:hi], gts[lo:hi])
        torch.cuda.synchronize()
        log("step %d loss %.4f" % (i, float(loss)))
    pmax, pmin = ts.flat.p.clone(), ts.flat.p.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    ident = bool(torch.equal(pmax, pmin))
    # single-GPU run over the full batch from rank 0's initial weights
    torch.manual_seed(100)
    ref = MlpMixer(**cfg).to(dev).train().set_precision(prec)
    ts1 = TrainStep(ref, lr=1e-3, weight_decay=1e-5, use_cuda_graph=False)
    for i in range(steps):
        ts1.step(xs, gts)
    torch.cuda.synchronize()
    upd = (ts.flat.p - ts1.flat.p).abs().max().item()
    moved = steps * 1e-3
    log("graph=%s identical_across_ranks=%s max|p_dp - p_single| = %.3e (weights moved by ~%.0e)" % (use_graph, ident, upd, moved))
    from motionmixerconv_b200 import _lib as L_
    aborts = L_.load().mmx_tc5_abort_count()
    log("timed-out waits: %d" % aborts)
    ok = ok and ident and upd < 0.02 * moved and aborts == 0
    ts.release_graphs()
dist.barrier(device_ids=[local])
if rank == 0:
    print("DP_CHECK", "OK" if ok else "FAILED")
import threading
killer = threading.Timer(20.0, lambda: os._exit(0 if ok else 1))
killer.daemon = True
killer.start()
dist.destroy_process_group()
killer.cancel()
log("process group destroyed")
sys.exit(0 if ok else 1)
