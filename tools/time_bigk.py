"""Per-dilation timing of one DynConv2d(64,64,k=20,d) forward on the config-2 layer shape (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_gcns_torch_b200.gcn_lib import dense as D  # noqa: E402

torch.manual_seed(0)
x = torch.randn(16, 64, 4096, 1, device="cuda")
for d in (1, 2, 3, 4, 8, 16, 27):
    mod = D.DynConv2d(64, 64, 20, d, "edge", "relu", "batch", True).cuda().eval()
    with torch.no_grad():
        for _ in range(2):
            mod(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            mod(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"d={d:2d} K={20*d:4d}  {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
