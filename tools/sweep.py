"""Large-batch throughput sweep (BASELINE.json configs[4]): batch 16K-256K sequences per GPU x hidden dim 64-512, MlpMixer and
ConvMixer, one JSON line per cell (throughput, block-granular HBM roofline fraction, the kernels that served it, or the error
an unsupported cell returns).  Single GPU: `python tools/sweep.py`; N GPUs: under torchrun (weak scaling, per-GPU batch fixed,
gradient exchange fused into the Adam kernel over NVLink peer memory).  env: SWEEP_H, SWEEP_B, SWEEP_E (comma lists), SWEEP_FAMILIES=mlp,conv, SWEEP_CPU=1
adds the reference's CPU step (oracle/_ref, 1024-sequence slice, all host threads) once per width.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motionmixerconv_b200.conv_mixer_model import ConvMixer
from motionmixerconv_b200.mlp_mixer import MlpMixer
from motionmixerconv_b200.train import TrainStep

PEAK = 6539.2
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pg = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD


def ints(name, dflt):
    return [int(v) for v in os.environ.get(name, dflt).split(",") if v]


def mlp_cfg(H):
    return dict(num_classes=66, num_blocks=4, hidden_dim=H, tokens_mlp_dim=20, channels_mlp_dim=H, seq_len=10, pred_len=10,
                activation="mish", regularization=0.1, input_size=66, r_se=8, use_se=True)


def conv_cfg(E):
    return dict(num_blocks=4, dimPosIn=66, dimPosEmb=E, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1, conv1_kernel_shape=(1, 3),
                conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation="mish", regularization=0.1, use_se=True, r_se=8,
                encoder_n_harmonic_functions=0, encoder_omega0=0)


def cpu_reference(family, cfg, To):
    from oracle import make_ref
    if not make_ref.available():
        return None
    Mlp, Conv, mpjpe = make_ref.import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    m = (Mlp if family == "mlp" else Conv)(**cfg).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    x, gt = torch.randn(1024, 10, 66) * 0.3, torch.randn(1024, To, 66) * 300
    n, t0 = 0, None
    for i in range(4):
        if i == 1:
            t0 = time.perf_counter()
        opt.zero_grad()
        mpjpe(m(x), gt).backward()
        opt.step()
        n += i >= 1
    return {"value": 1024 * n / (time.perf_counter() - t0), "unit": "sequences/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": "3 steps of a 1024-sequence slice (unmodified reference modules, torch CPU fp32)"}


def run(family, width, B, steps):
    torch.manual_seed(0)
    cfg = mlp_cfg(width) if family == "mlp" else conv_cfg(width)
    To = 10 if family == "mlp" else 25
    cell = {"family": family, "width": width, "per_gpu_batch": B, "n_gpus": world}
    try:
        model = (MlpMixer(**cfg) if family == "mlp" else ConvMixer(**cfg)).to(dev).train()
        if family == "mlp":
            model.set_precision("tf32")
        ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, process_group=pg)
        x = torch.randn(B, 10, 66, device=dev) * 0.3
        gt = torch.randn(B, To, 66, device=dev) * 300
        for _ in range(3):
            loss = ts.step(x, gt)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            loss = ts.step(x, gt)
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        assert torch.isfinite(loss)
        tile = 10 * width * 4
        alg = B * cfg["num_blocks"] * (5 if family == "mlp" else 10) * tile       # block-granular floor: 5 tiles per block (ConvMixer: per half)
        kern = "fp32 SIMT kernels"
        if family == "mlp":
            kern = "tcgen05 family" if all(ts.plan.saves) else "fp32 SIMT kernels (shape not served by the tcgen05 family)"
        cell.update({"ms_per_step": ms, "value": world * B / ms * 1e3, "unit": "sequences/s", "kernels": kern, "loss": float(loss),
                     "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": PEAK, "unit": "GB/s", "frac": alg / ms / 1e6 / PEAK,
                                  "basis": "whole step vs the block-granular algorithmic floor (5 activation tiles per fused block / half)"}})
        ts.release_graphs()
        del ts, model, x, gt
    except Exception as ex:          # unsupported cells are part of the result
        cell["error"] = "%s: %s" % (type(ex).__name__, str(ex)[:300])
    torch.cuda.empty_cache()
    return cell


if __name__ == "__main__":
    fams = os.environ.get("SWEEP_FAMILIES", "mlp,conv").split(",")
    Bs = ints("SWEEP_B", "16384,65536,262144")
    out = []
    for fam in fams:
        widths = ints("SWEEP_H", "50,64,128,256,512") if fam == "mlp" else ints("SWEEP_E", "64,128,256,512")
        for wd in widths:
            cpu = None
            if os.environ.get("SWEEP_CPU") == "1" and rank == 0 and world == 1:
                cpu = cpu_reference(fam, mlp_cfg(wd) if fam == "mlp" else conv_cfg(wd), 10 if fam == "mlp" else 25)
            for B in Bs:
                if B * wd > 262144 * 128:          # keep the slow fp32-SIMT cells (wide models) to a few seconds each
                    out.append({"family": fam, "width": wd, "per_gpu_batch": B, "n_gpus": world, "skipped": "time budget of the sweep"})
                    continue
                cell = run(fam, wd, B, steps=5 if B * wd <= 65536 * 128 else 2)
                if cpu:
                    cell["cpu_baseline"] = cpu
                out.append(cell)
                if rank == 0:
                    print(json.dumps(cell), flush=True)
    if world > 1:
        import threading
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)
