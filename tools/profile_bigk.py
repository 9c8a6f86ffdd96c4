"""ncu target: one DynConv2d forward per dilation on the config-2 layer shape (large-K slab path)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_gcns_torch_b200.gcn_lib import dense as D  # noqa: E402

torch.manual_seed(0)
x = torch.randn(16, 64, 4096, 1, device="cuda")
for d in (4, 16, 27):
    mod = D.DynConv2d(64, 64, 20, d, "edge", "relu", "batch", True).cuda().eval()
    with torch.no_grad():
        for _ in range(2):
            mod(x)
torch.cuda.synchronize()
