#!/usr/bin/env python
"""Benchmark of the hot path: train sequences/sec (fwd + MPJPE + bwd + Adam), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload k2|k4]

N=1 workload = BASELINE.json configs[1]: MotionMixer MlpMixer with squeeze-excitation, H36M xyz
10 -> 10 frames, batch 4096 per GPU (SURVEY.md §8d "K2": hidden 50, tokens 20, channels 50, 4 blocks,
mish, dropout 0.1, r_se 8).  N>1: the driver launches this file under torchrun; each rank trains on
its own shard of the global batch (weak scaling: 4096 sequences per GPU) with one flat-bucket NCCL
all-reduce per step.  One JSON line is printed by rank 0.

``--impl reference`` times the reference's CPU implementation of the same step on the host cores
(the torch-CPU port in oracle/mixer_torch.py — the reference itself is Python and does not exist on
the GPU box), all threads, same config / metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # SURVEY.md §8d K2 / BASELINE.json configs[1]
    "k2": dict(name="MotionMixer MlpMixer+SE, H36M xyz 10->10 frames (22 joints x 3), batch 4096 per GPU",
               family="mlp", B=4096, scale="h36m", loss_scale=1.0,
               cfg=dict(num_classes=66, num_blocks=4, hidden_dim=50, tokens_mlp_dim=20, channels_mlp_dim=50, seq_len=10,
                        pred_len=10, activation="mish", regularization=0.1, input_size=66, r_se=8, use_se=True)),
    # SURVEY.md §8d K1 / BASELINE.json configs[0]: ConvMixer __main__ config (harmonic encoder 64), batch 256
    "k1": dict(name="ConvMixer (train_mixer_h36m.py __main__ config), H36M xyz 10->25 frames (22 joints x 3), batch 256 per GPU",
               family="conv", B=256, scale="h36m", loss_scale=1.0,
               cfg=dict(num_blocks=4, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1,
                        conv1_kernel_shape=(1, 3), conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation="mish",
                        regularization=0.1, use_se=True, r_se=8)),
    # the same ConvMixer at the K2 batch size (large-batch sweep, BASELINE.json configs[4])
    "k1_b4096": dict(name="ConvMixer (train_mixer_h36m.py __main__ config), H36M xyz 10->25 frames, batch 4096 per GPU",
                     family="conv", B=4096, scale="h36m", loss_scale=1.0,
                     cfg=dict(num_blocks=4, dimPosIn=66, dimPosEmb=50, dimPosOut=66, in_nTP=10, out_nTP=25, conv_nChan=1,
                              conv1_kernel_shape=(1, 3), conv1_stride=(1, 1), conv1_padding=(0, 1), mode_conv="twice", activation="mish",
                              regularization=0.1, use_se=True, r_se=8)),
    # SURVEY.md §8d K3 / BASELINE.json configs[2]: ConvMixer AIS autoregressive config (BatchNorm, C=4, E=192, 5x9 / 9x5
    # kernels, 6 blocks), ONE 10 -> 5 pass of the rollout, batch 256
    "k3": dict(name="ConvMixer AIS autoregressive config (BatchNorm, C=4, E=192, k=(5,9)), 10->5 frames (11 joints x 3), batch 256 per GPU",
               family="conv", B=256, scale="ais", loss_scale=1.0,
               cfg=dict(num_blocks=6, dimPosIn=33, dimPosEmb=192, dimPosOut=33, in_nTP=10, out_nTP=5, conv_nChan=4,
                        conv1_kernel_shape=(5, 9), mode_conv="twice", activation="mish", regularization=-1.0, use_se=True, r_se=8,
                        encoder_n_harmonic_functions=0, encoder_omega0=0)),
    # SURVEY.md §8d K4 / BASELINE.json configs[3] (AMASS-shaped), per-GPU batch 4096
    "k4": dict(name="MotionMixer MlpMixer+SE, AMASS-shaped 18 joints 10->25 frames, batch 4096 per GPU",
               family="mlp", B=4096, scale="amass", loss_scale=1000.0,
               cfg=dict(num_classes=54, num_blocks=5, hidden_dim=128, tokens_mlp_dim=20, channels_mlp_dim=128, seq_len=10,
                        pred_len=25, activation="gelu", regularization=0.1, input_size=54, r_se=8, use_se=True)),
}


# stdout carries exactly ONE line (the JSON result): file descriptor 1 is pointed at stderr for the whole run, so that banners
# printed by native libraries (NCCL's "NCCL version ..." when NCCL_DEBUG is set, ...) cannot precede it
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms.  The sampler is started before the warm-up (nvidia-smi takes
    a few hundred ms to produce its first line: longer than a short timed region) and `mark()` brackets the timed region;
    `stop()` reports the samples that fall inside the marks (all samples, and says so, if none did)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.marks = index, [], None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def mark(self):
        self.marks.append(time.time())

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo, hi = (self.marks[0], self.marks[-1] + 0.03) if len(self.marks) >= 2 else (0.0, float("inf"))

        def collect(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1]))
                    mx.append(float(r[2]))
                except (ValueError, IndexError):
                    continue
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sorted(sm), mx, reasons

        inside = [row for row in self.rows if lo <= row[0] <= hi]
        sm, mx, reasons = collect(inside)
        window = "timed region"
        if not sm:
            sm, mx, reasons = collect(self.rows)
            window = "warm-up + timed region (no sample fell inside the timed region)"
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def make_data(w, n_batches, seed):
    from tests.synthetic import synthetic_pose_windows
    c = w["cfg"]
    T, To, D = (c["seq_len"], c["pred_len"], c["input_size"]) if w["family"] == "mlp" else (c["in_nTP"], c["out_nTP"], c["dimPosIn"])
    return [synthetic_pose_windows(w["B"], T, To, D, scale=w["scale"], seed=seed + i) for i in range(n_batches)]


# ---------------------------------------------------------------------------------------------------
def run_reference(args, w):
    """CPU arm: the oracle port on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    from oracle import mixer_torch as MT
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: each step is a 1024-sequence slice of the 4096-sequence batch
    Bs = min(w["B"], int(os.environ.get("MMX_CPU_SAMPLE_B", 1024)))
    wl = dict(w, B=Bs)
    data = [(torch.from_numpy(x), torch.from_numpy(g)) for x, g in make_data(wl, 2, 1234)]
    tr = MT.CpuTrainer(w["family"], w["cfg"], MT.random_params(w["family"], w["cfg"], 0), loss_scale=w["loss_scale"])
    for i in range(args.warmup):
        tr.step(*data[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        tr.step(*data[i % 2])
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    sample = "%d-sequence slice of the %d-sequence batch per step, fp32, torch %s CPU, %d threads" % (Bs, w["B"], torch.__version__, cores)
    emit({
        "impl": "reference", "metric": "train sequences/sec", "value": val, "unit": "sequences/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "model": w["cfg"], "per_gpu_batch": w["B"]},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def cpu_baseline(w, budget_s=15.0):
    import torch
    from oracle import mixer_torch as MT
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(w["B"], 1024)
    wl = dict(w, B=Bs)
    data = [(torch.from_numpy(x), torch.from_numpy(g)) for x, g in make_data(wl, 2, 1234)]
    tr = MT.CpuTrainer(w["family"], w["cfg"], MT.random_params(w["family"], w["cfg"], 0), loss_scale=w["loss_scale"])
    tr.step(*data[0])
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 40):
        tr.step(*data[n % 2])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": "sequences/s", "cores": cores, "kind": "port",
            "sample": "%d steps of a %d-sequence slice of the %d-sequence batch (oracle/mixer_torch.py, torch CPU fp32, %d threads)" % (n, Bs, w["B"], cores)}


# ---------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    from motionmixerconv_b200 import _lib as L
    from motionmixerconv_b200 import functional as F_
    from motionmixerconv_b200.conv_mixer_model import ConvMixer
    from motionmixerconv_b200.mlp_mixer import MlpMixer
    from motionmixerconv_b200.train import TrainStep
    from oracle import mixer_torch as MT   # only for random_params (weight layout) and the cpu_baseline leg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's own banner ("NCCL version ...", printed when NCCL_DEBUG is set) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    model = MlpMixer(**w["cfg"]) if w["family"] == "mlp" else ConvMixer(**w["cfg"])
    model.load_state_dict(MT.random_params(w["family"], w["cfg"], 0), strict=True)
    model = model.to(dev).train()
    prec = args.precision if w["family"] == "mlp" else "fp32"      # the tensor-core MixerBlock kernels serve the MlpMixer path
    if w["family"] == "mlp":
        model.set_precision(prec)
    B = w["B"]
    n_batches = 4
    host = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(g).pin_memory()) for x, g in make_data(w, n_batches, 1234 + 1000 * rank)]
    devd = [(x.to(dev), g.to(dev)) for x, g in host]
    ts = TrainStep(model, lr=1e-3, weight_decay=1e-5, loss_scale=w["loss_scale"], process_group=pg)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 512 MB > 126 MB L2

    def timed(n_steps, data, read_loss, ts=ts):
        evs = []
        for i in range(n_steps):
            flush.add_(1.0)                       # evict L2 between timed iterations (outside the events)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if read_loss:      # end-to-end arm: this step's host inputs were prefetched during the previous step; prefetch the next
                loss = ts.step(*data[i % n_batches], prefetch=data[(i + 1) % n_batches])
            else:
                loss = ts.step(*data[i % n_batches])
            if read_loss:
                loss_host = loss.to("cpu", non_blocking=False)   # D2H read of the step's result inside the region
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize(dev)
        return sum(s.elapsed_time(e) for s, e in evs), float(loss)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    # ---- device-resident arm (value) ----
    sampler = ClockSampler(local)
    sampler.start()
    timed(args.warmup, devd, False)
    barrier()
    sampler.mark()
    t_ms, last_loss = timed(args.steps, devd, False)
    barrier()
    # ---- end-to-end arm: pinned host inputs, H2D inside the region, loss read back every step ----
    timed(max(1, args.warmup // 2), host, True)
    barrier()
    t_e2e_ms, _ = timed(args.steps, host, True)
    barrier()
    sampler.mark()
    clocks = sampler.stop()
    # ---- the other arithmetic mode of the same step (device-resident), for the record ----
    t_alt_ms, alt = 0.0, None
    if w["family"] == "mlp" and not args.no_alt_precision:
        alt = "fp32" if prec == "tf32" else "tf32"
        model2 = MlpMixer(**w["cfg"])
        model2.load_state_dict(MT.random_params(w["family"], w["cfg"], 0), strict=True)
        model2 = model2.to(dev).train().set_precision(alt)
        ts2 = TrainStep(model2, lr=1e-3, weight_decay=1e-5, loss_scale=w["loss_scale"], process_group=pg)
        timed(args.warmup, devd, False, ts2)
        barrier()
        t_alt_ms, _ = timed(args.steps, devd, False, ts2)
        barrier()
    tt = torch.tensor([t_ms, t_e2e_ms, t_alt_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t_e2e_ms, t_alt_ms = tt.tolist()

    # ---- dominant kernel (block backward) timed alone for the roofline ----
    roof = None
    if rank == 0:
        import ctypes as C
        lib = L.load()
        pl = ts.plan
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        c = w["cfg"]
        if w["family"] == "mlp":
            mb, tw, tg = pl.blocks[1]
            d = pl._desc(mb, True)
            call = lambda: L.check(lib, lib.mmx_mlp_block_bwd(C.byref(d), C.byref(tw), C.byref(tg), pl.acts[1].data_ptr(), pl.dact[0].data_ptr(),
                                                              pl.dact[1].data_ptr(), st), "mmx_mlp_block_bwd")
            tile = c["seq_len"] * c["hidden_dim"] * 4
            kname = "mlp_block_bwd (MixerBlock backward, forward recomputed in-kernel)"
            flops = B * 4 * 2 * c["seq_len"] * (2 * c["tokens_mlp_dim"] * c["hidden_dim"] + 2 * c["hidden_dim"] * c["channels_mlp_dim"])
        else:
            kind, mb, half, tw, tg = pl.ops[2]
            d = pl._desc(mb, half, True)
            if 2 in pl.bn:          # BatchNorm half: time the second backward pass (BN + conv + LN backward)
                b = pl.bn[2]
                coef = torch.zeros(3 * c["conv_nChan"], device=dev)
                call = lambda: L.check(lib, lib.mmx_conv_half_bn_bwd2(C.byref(d), C.byref(tw), C.byref(tg), b["bn"].data_ptr(), coef.data_ptr(),
                                                                      pl.acts[2].data_ptr(), b["z"].data_ptr(), pl.dact[0].data_ptr(),
                                                                      b["gd"].data_ptr(), pl.dact[1].data_ptr(), st), "mmx_conv_half_bn_bwd2")
            else:
                call = lambda: L.check(lib, lib.mmx_conv_half_bwd(C.byref(d), C.byref(tw), C.byref(tg), pl.acts[2].data_ptr(), pl.dact[0].data_ptr(),
                                                                  pl.dact[1].data_ptr(), st), "mmx_conv_half_bwd")
            tile = c["conv_nChan"] * c["in_nTP"] * c["dimPosEmb"] * 4
            kname = "conv_half_bwd (one ConvMixerBlock half backward, forward recomputed in-kernel)"
            kt, kp = mb.conv1.kernel
            flops = B * 4 * 2 * c["conv_nChan"] ** 2 * kt * kp * c["in_nTP"] * c["dimPosEmb"]
        reps, tot = 20, 0.0
        for i in range(reps + 3):
            flush.add_(1.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            call()
            e.record()
            torch.cuda.synchronize(dev)
            if i >= 3:
                tot += s.elapsed_time(e)
        t_k = tot / reps
        alg = B * 3 * tile                      # read saved input + upstream grad, write input grad
        peak, which = peaks()
        ach = alg / (t_k * 1e-3) / 1e9
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # fp32 FMA pipe: 148 SMs x 128 lanes x 2 flop x 1.965 GHz (nominal)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
                ent = tj.get(args.workload + "_" + prec) or tj.get(args.workload)
            if ent:
                traffic, traffic_src = ent["bytes"], ent["source"]
        roof = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": which, "algorithmic_bytes_per_launch": alg, "kernel_ms": t_k, "traffic": traffic,
                "traffic_source": traffic_src,
                "fp32_tflops": flops / (t_k * 1e-3) / 1e12, "fp32_frac": flops / (t_k * 1e-3) / 1e12 / fp32_peak,
                "note": ("TF32 tensor-core contractions (2e-3 parity mode): the kernel is bound by instruction issue of the fp32 "
                         "elementwise work between the MMAs (LayerNorm, Mish, dropout, SE) at 6 warps / SM, not by HBM or the tensor pipe"
                         if prec == "tf32" else
                         "fp32 SIMT (1e-5 parity mode): the kernel is bound by the fp32 pipe / latency, not HBM") + " — see DESIGN.md §4"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    c = w["cfg"]
    x0, g0 = host[0]
    out = {
        "metric": "train sequences/sec", "value": world * B * args.steps / (t_ms * 1e-3), "unit": "sequences/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if prec == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": w["name"], "model": c, "per_gpu_batch": B, "global_batch": B * world,
                   "precision": ("tf32: MixerBlock contractions on the tensor cores with TF32 operands and fp32 accumulation; LayerNorm, "
                                 "activations, SE, residuals, loss, Adam and the embed / head kernels fp32 (north star: 2e-3 parity mode)")
                                if prec == "tf32" else "fp32 everywhere (north star: 1e-5 parity mode)",
                   "parallelism": "dp%d" % world, "optimizer": "Adam lr 1e-3 wd 1e-5 (fused, flat buffers)",
                   "l2": "512 MB L2 flush between timed iterations (outside the CUDA-event pairs)",
                   "step": "CUDA graph A: memset + embed/encoder + block kernels fwd + head + mpjpe + head bwd + block kernels bwd + embed/encoder bwd; [NCCL all-reduce of the flat bucket]; CUDA graph B: fused adam"},
        "e2e": {"value": world * B * args.steps / (t_e2e_ms * 1e-3), "unit": "sequences/s",
                "h2d_bytes_per_step": x0.numel() * 4 + g0.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": t_e2e_ms / args.steps,
                "note": "TrainStep.step(x_host, gt_host, prefetch=next): every step's inputs cross PCIe from pinned memory inside the "
                        "timed region (double-buffered on a copy stream, overlapping the previous step's kernels) and the loss is read back"},
        "gpu_launches": ts.kernel_launches_per_step * args.steps,
        "final_loss": last_loss,
        "clocks": clocks,
        "roofline": roof,
    }
    if alt:
        out["other_precision"] = {"precision": alt, "value": world * B * args.steps / (t_alt_ms * 1e-3), "unit": "sequences/s",
                                  "ms_per_step": t_alt_ms / args.steps}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(w)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MMX_WORKLOAD", "k2"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default=os.environ.get("MMX_BENCH_PRECISION", "fp32"), choices=["fp32", "tf32"],
                    help="arithmetic of the MixerBlock contractions (north star: fp32 = 1e-5 parity mode, tf32 = 2e-3 parity mode)")
    ap.add_argument("--no-alt-precision", action="store_true", help="skip the extra timing of the other precision mode")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
