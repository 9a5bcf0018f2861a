"""Turn the raw ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_chain_r01.csv profiles/r01_chain_launches.md
    python tools/summarize_profiles.py full gpurun_out/x.ncu-rep profiles/r01_x.md
"""
import collections
import csv
import re
import subprocess
import sys


def short(n):
    m = re.search(r"(\w+_kernel|regular_fft\w*|vector_fft\w*|preprocess\w*|postprocess\w*|elementwise_kernel|reduce_kernel)", n)
    return m.group(1) if m else n[:50]


def launches(src, dst, note=""):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    idx = [i for i, n in enumerate(names) if "eq_fwd_kernel" in n]
    step = rows[idx[-2]: idx[-1]] if len(idx) >= 2 else rows
    # the slice runs from the timed step's first kernel to the next step's first kernel: drop the host-side
    # parameter-prep kernels of the following (e2e) step at the tail
    tot, cnt = {}, collections.Counter()
    for x in step:
        n = short(x["Kernel Name"]) + " grid=" + x["Grid Size"].replace(" ", "")
        v = float(x["Metric Value"])
        tot[n] = tot.get(n, 0) + v
        cnt[n] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none): one chain step\n\n{note}\n\n")
        f.write("Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.\n\n")
        f.write("| total us | share | launches | avg us | kernel |\n|---:|---:|---:|---:|---|\n")
        for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| {v/1e3:.1f} | {100*v/s:.1f}% | {cnt[n]} | {v/cnt[n]/1e3:.1f} | `{n}` |\n")
        f.write(f"\ntotal {s/1e3:.1f} us over {len(step)} launches\n")


def full(src, dst, note=""):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def g(d, k):
        try:
            return float(d[idx[k]].replace(",", ""))
        except Exception:
            return float("nan")

    want = [("time", "gpu__time_duration.sum"), ("dram rd", "dram__bytes_read.sum"), ("dram wr", "dram__bytes_write.sum"),
            ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("FMA pipe inst %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            ("FMA pipe cycles %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            ("smem/shuffle pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ("regs", "launch__registers_per_thread"), ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed")]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary\n\n{note}\n\n")
        f.write("| kernel | grid x block | " + " | ".join(f"{l} ({units[idx[k]]})" if units[idx[k]] else l for l, k in want) + " | top stalls (warps per issue) |\n")
        f.write("|---|---|" + "---:|" * len(want) + "---|\n")
        seen = set()
        for d in data:
            name = short(d[idx["Kernel Name"]])
            tmpl = re.search(r"<([^>]*)>", d[idx["Kernel Name"]])
            key = (name, tmpl.group(1) if tmpl else "", d[idx["launch__grid_size"]])
            if key in seen:
                continue
            seen.add(key)
            st = sorted([(g(d, h), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in hdr
                         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")],
                        reverse=True)[:4]
            f.write(f"| `{name}<{key[1]}>` | {d[idx['launch__grid_size']]} x {d[idx['launch__block_size']]} | " +
                    " | ".join(f"{g(d, k):.1f}" for _, k in want) + " | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |\n")


def traffic(src, dst, note=""):
    """profiles/r02_traffic.json: DRAM bytes per item of the reverb's forward and backward pipelines, summed over the
    kernels of one `ncu --set full` capture of tools/debug/reverb_step.py <items> (bench.py reads this file)"""
    import json
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units = rows[1]
    fwd_names = ("spectral_gen", "ifft_shape", "x_fft", "ifft_mix", "vector_fft", "regular_fft")
    tot = {"reverb_fwd": 0.0, "reverb_bwd": 0.0}
    per_kernel = []
    seen_bwd = False
    for d in data:
        name = d[idx["Kernel Name"]]
        b = sum(float(d[idx[k]].replace(",", "")) * scale.get(units[idx[k]], 1.0)
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        if "g_fft" in name:
            seen_bwd = True
        is_fwd = (not seen_bwd) and (any(f in name for f in fwd_names) or "partition_mac" in name)
        tot["reverb_fwd" if is_fwd else "reverb_bwd"] += b
        per_kernel.append({"kernel": short(name), "dir": "fwd" if is_fwd else "bwd", "dram_bytes": b,
                           "us": float(d[idx["gpu__time_duration.sum"]].replace(",", ""))})
    items = int(note) if note else 148
    out = {"geometry": [48000, 96000, 1023], "items_in_capture": items,
           "dram_bytes_per_item": {k: v / items for k, v in tot.items()},
           "source": f"profiles/r02_reverb_kernels_b148_full.md / {src.split('/')[-1]} (ncu --set full, dram__bytes_read.sum + "
                     "dram__bytes_write.sum per launch, summed over the kernels of one 148-item chunk of "
                     "tools/debug/reverb_step.py; tools/summarize_profiles.py traffic)", "kernels": per_kernel}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["dram_bytes_per_item"]))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": launches, "full": full, "traffic": traffic}[mode](src, dst, note)
