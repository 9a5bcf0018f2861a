"""Small forward of the device-noise reverb for compute-sanitizer / debugging (one item, short clip)."""
import sys
import torch
sys.path.insert(0, ".")
import dasp_pytorch_b200 as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
x = torch.rand(bs, 2, n, device=dev)
p = [torch.rand(bs, device=dev) for _ in range(25)]
y = D.noise_shaped_reverberation(x, 44100, *p, num_samples=L, num_bandpass_taps=1023)
torch.cuda.synchronize()
print("ok", n, L, float(y.abs().max()))
