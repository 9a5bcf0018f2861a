#!/usr/bin/env python
"""Contract benchmark: audio samples/sec (fwd+bwd) of the dasp hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # CPU reference arm

Workload (BASELINE.json configs[4], the config the headline metric is quoted on): the chain
parametric_eq -> compressor -> noise_shaped_reverberation (12 bands, 1023 taps, IR 96000) -> distortion
on batch 1024 x 2 ch x 48000 samples @ 44.1 kHz, forward + backward of loss = mean(y^2) with gradients
to x and to every parameter.  One "step" = one such pass over one synthetic batch.  With --gpus N each
rank owns an independent batch of the same size (items are independent: no data-path collective;
"scaling": "weak"); value = samples processed by all ranks / max-over-ranks device time.

The JSON line carries: value (inputs resident in HBM), e2e (same metric through the public API with
pinned-host inputs copied H2D and the loss/parameter gradients read back every step), roofline of the
dominant stage (algorithmic bytes / CUDA-event time / measured HBM peak), per-stage breakdown,
cpu_baseline (oracle port of the reference algorithm timed on the host cores, N=1 only), clocks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100
N_SAMPLES = 48000
CHS = 2
IR_LEN = 96000
TAPS = 1023
METRIC = "audio samples/sec (fwd+bwd) @ batch=1024x2chx48k"
UNIT = "samples/s"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------

def eq_ranges(sr=SR):
    g, q = (-20.0, 20.0), (0.1, 6.0)
    hi = (sr // 2) - 1000
    fr = [(20, 2000), (80, 2000), (2000, 8000), (8000, 12000), (12000, hi), (4000, hi)]
    out = []
    for f in fr:
        out += [g, f, q]
    return out


COMP_RANGES = [(-60.0, 0.0), (1.0, 20.0), (5.0, 100.0), (5.0, 100.0), (0.0, 12.0), (0.0, 12.0)]


def make_inputs(bs, seed):
    """seeded synthetic batch: x ~ U(-1,1), parameters ~ U(0,1) mapped through the reference Processor
    ranges (modules.py:136-155, 179-186, 204-230); distortion drive 0..24 dB per (item, channel) row."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(bs, CHS, N_SAMPLES, generator=g) * 2 - 1
    p01 = torch.rand(bs, 18 + 6 + 25, generator=g)
    eq = [p01[:, i] * (hi - lo) + lo for i, (lo, hi) in enumerate(eq_ranges())]
    c01 = p01[:, 18:24].clone()
    c01[:, 4].clamp_(min=0.05)     # knee_db > 0: knee == 0 yields NaN gradients in the reference too
    comp = [c01[:, i] * (hi - lo) + lo for i, (lo, hi) in enumerate(COMP_RANGES)]
    rev = [p01[:, 24 + i].clone() for i in range(25)]
    drive = torch.rand(bs * CHS, generator=g) * 24.0
    return x, eq, comp, rev, drive


def chain(mod, x, eq, comp, rev, drive, **rev_kw):
    y = mod.parametric_eq(x, SR, *eq)
    y = mod.compressor(y, SR, *comp)
    y = mod.noise_shaped_reverberation(y, SR, *rev, num_samples=IR_LEN, num_bandpass_taps=TAPS, **rev_kw)
    return mod.distortion(y, SR, drive)


# ------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi polled in the background from before the warm-up; only the samples whose timestamps fall
    inside the timed region [t0, t1] are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self, t0, t1):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, sm_all, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            sm_all.append(clk)
            if not (t0 - 0.02 <= ts <= t1 + 0.02):
                continue
            sm.append(clk); mx.append(cmax)
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples_in_timed_region": len(sm), "samples_total": len(sm_all), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (time-domain conv1d reverb, FFT-grid IIRs)
# ------------------------------------------------------------------------------------------

def cpu_chain_seconds(bs, reps=1):
    """seconds per fwd+bwd chain step of the CPU oracle port (reference algorithm, all host threads)"""
    import torch
    import oracle
    x, eq, comp, rev, drive = make_inputs(bs, seed=1)
    leaves = [x] + eq + comp + rev + [drive]
    best = None
    for _ in range(reps):
        for t in leaves:
            t.requires_grad_(True)
            t.grad = None
        t0 = time.perf_counter()
        y = chain(oracle, x, eq, comp, rev, drive, method="direct")
        y.pow(2).mean().backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def make_config(bs, world, chunk):
    """the `config` object shared by both arms (the reference arm times a bounded SAMPLE of this workload)"""
    return {"workload": "configs[4]: chain eq->comp->reverb(12 bands, IR 96000, 1023 taps, device Philox noise)->dist, "
                        f"batch {bs}/GPU x 2ch x 48000 @44.1k, fwd+bwd of mean(y^2), grads to x and all params",
            "global_batch": bs * world, "per_gpu_batch": bs, "parallelism": f"dp{world} (independent items, no collective)",
            "l2": "inputs (393 MB/tensor) exceed the 126 MB L2: no flush needed", "reverb_chunk_items": chunk}


def run_reference_arm(args, rank):
    import torch
    if rank != 0:
        return
    bs = 4                                      # the reference's per-item CPU rate improves with the batch: give it 4
    for _ in range(args.warmup):
        cpu_chain_seconds(bs)
        break                                   # one warm-up pass is enough on the CPU (tens of seconds each)
    # each step = the full chain fwd+bwd on FOUR items (~30 s on 8 cores): the run is capped at ~3 minutes so that
    # any --steps K finishes "within a few minutes"; `steps` in the JSON line is the number actually timed
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        cpu_chain_seconds(bs)
        done += 1
        if time.perf_counter() - t0 > 180.0:
            break
    dt = (time.perf_counter() - t0) / done
    args.steps = done
    val = bs * CHS * N_SAMPLES / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.batch, max(args.gpus, 1), None),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{bs} items x 2ch x 48000 per step (of the {args.batch}-item batch), full chain fwd+bwd; "
                                   "oracle port of the reference algorithm (FFT-grid IIRs, time-domain conv1d reverb, "
                                   "CPU mt19937 noise), all host threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="items per GPU (BASELINE config: 1024)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    import dasp_pytorch_b200 as D
    from dasp_pytorch_b200 import functional as F

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    bs = args.batch
    x_h, eq_h, comp_h, rev_h, drive_h = make_inputs(bs, seed=1000 + rank)
    x_pin = x_h.pin_memory()
    p_pin = torch.stack(eq_h + comp_h + rev_h, 1).contiguous().pin_memory()       # (bs, 49)
    d_pin = drive_h.pin_memory()

    def to_leaves(xd, pd, dd):
        # ONE leaf per tensor kind: the 49 per-item parameters stay packed as (bs, 49); the processors get
        # column views, so the parameter gradients arrive packed as well (pd.grad)
        xd.requires_grad_(True)
        pd.requires_grad_(True)
        dd.requires_grad_(True)
        cols = list(pd.unbind(1))
        return xd, cols[:18], cols[18:24], cols[24:49], dd

    p_dev = p_pin.to(dev)
    x, eq, comp, rev, drive = to_leaves(x_pin.to(dev), p_dev, d_pin.to(dev))
    leaves = [x, drive, p_dev]

    def step():
        for t in leaves:
            t.grad = None
        y = chain(D, x, eq, comp, rev, drive)
        loss = y.pow(2).mean()
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region (device time, CUDA events on the launching stream) ----
    F.STAGE_TIMING = []                       # per-stage CUDA events, see functional._timed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(wall0, wall1)
    stage_events = F.STAGE_TIMING
    F.STAGE_TIMING = None
    stages = {}
    for name, a, b in stage_events:
        stages.setdefault(name, []).append(a.elapsed_time(b))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    samples_per_step = bs * CHS * N_SAMPLES
    value = world * samples_per_step * args.steps / (ms_max * 1e-3)

    # ---- end-to-end: pinned host inputs -> H2D -> fwd+bwd -> D2H of loss and parameter gradients ----
    # Every step copies ITS inputs from pinned host memory and reads its loss + parameter gradients back.
    # Like any input pipeline, the copy of step k+1 is issued on a side stream while step k computes
    # (double-buffered device tensors); nothing is skipped or cached across steps.
    copy_stream = torch.cuda.Stream(device=dev)

    def upload():
        with torch.cuda.stream(copy_stream):
            bufs = (x_pin.to(dev, non_blocking=True), p_pin.to(dev, non_blocking=True), d_pin.to(dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return bufs, ev

    def e2e_loop(n_steps):
        nxt = upload()
        out = None
        for k in range(n_steps):
            (xd, pd, dd), ev = nxt
            torch.cuda.current_stream(dev).wait_event(ev)
            for t in (xd, pd, dd):
                t.record_stream(torch.cuda.current_stream(dev))
            if k + 1 < n_steps:
                nxt = upload()
            xx, e_, c_, r_, d_ = to_leaves(xd, pd, dd)
            y = chain(D, xx, e_, c_, r_, d_)
            loss = y.pow(2).mean()
            loss.backward()
            out = (float(loss.item()), pd.grad.cpu(), d_.grad.cpu())      # D2H read of this step's results
        return out

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * samples_per_step * args.steps / float(t.item())
    h2d = x_pin.numel() * 4 + p_pin.numel() * 4 + d_pin.numel() * 4
    d2h = 4 + bs * 49 * 4 + bs * CHS * 4

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        e = samples_per_step
        leff = min(IR_LEN, N_SAMPLES)       # only the first min(L, N) IR taps can reach the N outputs
        item_rev_fwd = (2 * N_SAMPLES * 4) * 2 + 2 * (2 * leff * 4)                # x in, y out, IR write + read
        item_rev_bwd = 3 * (2 * N_SAMPLES * 4) + 2 * leff * 4 + 2 * (2 * leff * 4) + 2 * 12 * leff * 4
        alg = {  # algorithmic bytes per launch: 8 B/sample fwd, 12 B/sample bwd for the streaming ops
            "eq_fwd": 8 * e, "eq_bwd": 12 * e, "comp_fwd": 8 * e, "comp_bwd": 12 * e, "dist_fwd": 8 * e,
            "dist_bwd": 12 * e, "reverb_fwd": item_rev_fwd * bs, "reverb_bwd": item_rev_bwd * bs,
        }
        kern = {"eq_fwd": "eq_fwd_kernel", "eq_bwd": "eq_bwd_kernel", "comp_fwd": "dynamics_fwd_kernel",
                "comp_bwd": "dynamics_bwd_kernel", "dist_fwd": "pointwise_fwd_kernel", "dist_bwd": "pointwise_bwd_kernel",
                "reverb_fwd": "reverb fwd pipeline: spectral_gen_kernel, ifft_shape_kernel, x_fft_kernel, partition_mac_kernel, "
                              "ifft_mix_kernel (own in-shared-memory 8192-point FFT fused with the element-wise stages), "
                              "cuFFT C2C(8192) x1 (IR partitions)",
                "reverb_bwd": "reverb bwd pipeline: g_blocks_kernel, cuFFT C2C(8192) x3, partition_mac_kernel x2, "
                              "finish_dx_blocks_kernel, ir_grad_pp_kernel"}
        # DRAM bytes per item actually moved by the two reverb pipelines: dram__bytes_read.sum + dram__bytes_write.sum
        # summed over their kernels in one `ncu --set full` capture of a 128-item chunk at this geometry
        # (profiles/r01_reverb_kernels_b128_full.md: fwd 1110 + 1346 MB, bwd 1636 + 358 MB per 128 items)
        traffic_item = {"reverb_fwd": 19.18e6, "reverb_bwd": 15.58e6} if (N_SAMPLES, IR_LEN) == (48000, 96000) else {}
        breakdown = {}
        for name, v in stages.items():
            m = statistics.mean(v)
            breakdown[name] = {"ms": round(m, 4), "alg_GBps": round(alg[name] / (m * 1e-3) / 1e9, 1),
                               "frac": round(alg[name] / (m * 1e-3) / 1e9 / peak, 4)}
        dom = max(breakdown, key=lambda k: breakdown[k]["ms"]) if breakdown else None
        roofline = None
        if dom:
            roofline = {"bound": "hbm", "kernel": kern[dom], "achieved": breakdown[dom]["alg_GBps"], "peak": peak,
                        "unit": "GB/s", "frac": breakdown[dom]["frac"],
                        "traffic": traffic_item[dom] * bs if dom in traffic_item else None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": breakdown[dom]["ms"]}
        chunk_items = F.reverb_chunk_items(dev)
        chunks = -(-bs // chunk_items)
        own_launches_per_step = 1 + 2 + 1 + 1 + 1 + 2 + chunks * (5 + 6)   # eq f/b, comp f/b, dist f/b, reverb per chunk
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(bs, world, chunk_items),
            "roofline": roofline, "stages": breakdown,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": own_launches_per_step * args.steps, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            sec = cpu_chain_seconds(4)
            line["cpu_baseline"] = {"value": 4 * CHS * N_SAMPLES / sec, "unit": UNIT, "cores": torch.get_num_threads(),
                                    "kind": "port", "sample": "4 items x 2ch x 48000 (of the 1024-item batch), full chain "
                                                              "fwd+bwd, once; oracle port of the reference algorithm"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
