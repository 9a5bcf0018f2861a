#!/usr/bin/env python
"""Markdown results table from the committed bench lines profiles/r02_bench_n{1,2,4,8}.json (README.md / DESIGN.md)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(n):
    p = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    d1 = load(1)
    out = []
    out.append("| GPUs | scaling | global batch | ms / step | G samples/s (device) | G samples/s (e2e: H2D + D2H every step) | vs 1 GPU | weak scaling (1024 items/GPU) | NCCL edges: scatter / gather (ms) |")
    out.append("|---:|---|---:|---:|---:|---:|---:|---|---|")
    for n in (1, 2, 4, 8):
        d = load(n)
        if d is None:
            continue
        eff = d["value"] / (n * d1["value"]) if d1 else float("nan")
        weak = d.get("weak")
        edges = d.get("edges")
        out.append(f"| {n} | {d['scaling']} | {d['config']['global_batch']} | {d['ms_per_step']:.3f} | {d['value'] / 1e9:.2f} | "
                   f"{d['e2e']['value'] / 1e9:.2f} | {eff:.3f} | "
                   + (f"{weak['value'] / 1e9:.2f} G samples/s at {weak['ms_per_step']:.2f} ms" if weak else "—") + " | "
                   + (f"{edges['scatter_ms']:.2f} / {edges['gather_ms']:.2f}" if edges else "—") + " |")
    if d1:
        st = d1["stages"]
        out.append("")
        out.append("Stage breakdown at 1 GPU (one C-ABI call each, CUDA events, eager pass of the same step; frac = algorithmic "
                   "bytes / time / 6561.3 GB/s measured HBM peak):")
        out.append("")
        out.append("| stage | ms | frac of HBM roofline | round 1 (ms / frac) |")
        out.append("|---|---:|---:|---|")
        r1 = {"eq_fwd": "0.41 / 0.29", "eq_bwd": "1.57 / 0.115", "comp_fwd": "0.15 / 0.79", "comp_bwd": "0.36 / 0.50",
              "reverb_fwd": "6.42 / 0.037", "reverb_bwd": "3.89 / 0.277", "dist_fwd": "0.12 / 1.005", "dist_bwd": "0.18 / 1.015"}
        for k in ("eq_fwd", "eq_bwd", "comp_fwd", "comp_bwd", "reverb_fwd", "reverb_bwd", "dist_fwd", "dist_bwd"):
            if k in st:
                out.append(f"| {k} | {st[k]['ms']:.3f} | {st[k]['frac']:.3f} | {r1[k]} |")
        out.append(f"| whole step (graph replay) | {d1['ms_per_step']:.2f} | — | 14.02 |")
        cfg = d1.get("configs", {})
        if cfg:
            out.append("")
            out.append("BASELINE configs 2-4 on one GPU (device-timed, L2 flushed): " + "; ".join(
                f"`{k}` fwd {v['fwd_ms']:.3f} ms, fwd+bwd {v['fwdbwd_ms']:.3f} ms ({v['gsamples_per_s']:.1f} G samples/s)"
                for k, v in cfg.items()) + ".")
        sh = []
        for n in (2, 4, 8):
            d = load(n)
            c4 = (d or {}).get("c4_reverb_256x2x48000_ir96000_sharded")
            if c4:
                sh.append(f"{n} GPUs ({c4['items_per_gpu'][0]} items each) {c4['fwdbwd_ms']:.3f} ms = {c4['gsamples_per_s']:.1f} G samples/s")
        if sh and "c4_reverb_256x2x48000_ir96000" in cfg:
            out.append("")
            out.append(f"Config 4 (reverb 256 x 2 x 48000, IR 96000) split over the GPUs, fwd+bwd: 1 GPU "
                       f"{cfg['c4_reverb_256x2x48000_ir96000']['fwdbwd_ms']:.3f} ms; " + "; ".join(sh) + ".")
        rg, cb = d1.get("reference_gpu"), d1.get("cpu_baseline")
        if rg and "value" in rg:
            out.append("")
            out.append(f"Reference's own CUDA path on the same B200 (`reference_gpu`): {rg['value'] / 1e6:.2f} M samples/s at batch "
                       f"{rg['batch']} -> this repo is {d1['value'] / rg['value']:.0f}x faster on the device-timed metric.")
        if cb:
            out.append(f"Reference's CPU path (`cpu_baseline`, {cb['kind']}, {cb['cores']} threads): {cb['value'] / 1e3:.1f} k samples/s.")
    print("\n".join(out))


def update_readme():
    """replace the block between the results markers of README.md"""
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        main()
    path = os.path.join(ROOT, "README.md")
    text = open(path).read()
    a = text.index("<!-- results:begin")
    a = text.index("\n", a) + 1
    b = text.index("<!-- results:end -->")
    head = "### Measured on B200 (round 2; `profiles/r02_bench_n{1,2,4,8}.json`, all four from the final build)\n\n"
    open(path, "w").write(text[:a] + head + buf.getvalue() + text[b:])


if __name__ == "__main__":
    if "--update-readme" in sys.argv:
        update_readme()
    else:
        main()
