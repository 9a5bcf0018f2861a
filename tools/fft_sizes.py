"""cuFFT (via torch.fft) timing of candidate transform lengths for the reverb pipeline."""
import sys
import torch
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for batch, sizes in ((192, [49022 + 2, 49152, 49280, 50000, 51200, 53248, 57344, 65536]), (768, [49152, 65536]),
                     (16, [96000, 96040, 96228, 96768, 98000, 98304, 100352, 102400, 106496, 114688, 131072]),
                     (64, [96000, 98304, 131072]), (384, [97200, 98304, 131072]), (32, [144000, 147456, 163840, 262144])):
    for n in sizes:
        x = torch.randn(batch, n, device=dev)
        X = torch.fft.rfft(x)
        r = t(lambda: torch.fft.rfft(x)); c = t(lambda: torch.fft.irfft(X, n))
        mb = batch * n * 4 / 1e6
        print(f"batch {batch:4d} n {n:7d}  r2c {r*1e3:8.1f} us ({2*mb/r/1e3:6.2f} TB/s eff)  c2r {c*1e3:8.1f} us ({2*mb/c/1e3:6.2f} TB/s eff)", flush=True)
