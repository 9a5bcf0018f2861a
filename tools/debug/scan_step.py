"""One forward+backward of parametric_eq and compressor at 1024 x 2 x 48000 (for ncu captures of the scan kernels)."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import dasp_pytorch_b200 as D
from helpers import COMP_RANGES, SR, denorm, eq_ranges

dev = torch.device("cuda:0")
bs, n = 1024, 48000
torch.manual_seed(0)
x = (torch.rand(bs, 2, n, device=dev) * 2 - 1).requires_grad_(True)
pe = [q.to(dev).requires_grad_(True) for q in denorm(torch.rand(bs, 18), eq_ranges())]
pc = [q.to(dev).requires_grad_(True) for q in denorm(torch.rand(bs, 6).clamp(min=0.05), COMP_RANGES)]
for _ in range(2):
    D.parametric_eq(x, SR, *pe).square().mean().backward()
    D.compressor(x, SR, *pc).square().mean().backward()
torch.cuda.synchronize()
