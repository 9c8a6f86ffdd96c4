"""A/B timing of the tensor-core selection kernels on the headline layer (DynConv2d(64,64,k=20) forward, B=16
N=4096): 'tc1' = one query tile per CTA (knn_tc_kernel), 'tc' = default routing (knn_tc4_kernel where it applies).
CUDA events on the launch stream, inputs rotate over 8 batches (134 MB > L2).  Prints one JSON line."""
import json
import sys

import torch

sys.path.insert(0, ".")
from deep_gcns_torch_b200 import _native  # noqa: E402
from deep_gcns_torch_b200.gcn_lib import dense as D  # noqa: E402


def main():
    torch.manual_seed(0)
    B, C, N, k = 16, 64, 4096, 20
    mod = D.DynConv2d(C, C, k, 1, "edge", "relu", "batch", True).cuda().eval()
    xs = [torch.randn(B, C, N, 1, device="cuda") for _ in range(8)]
    out = {}
    ref = None
    for path in ("tc1", "tc", "tc1", "tc"):
        _native.set_knn_path(path)
        with torch.no_grad():
            for i in range(5):
                y = mod(xs[i % 8])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 40
            e0.record()
            for i in range(steps):
                y = mod(xs[i % 8])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            y0 = mod(xs[0])
        if ref is None:
            ref = y0.clone()
        out.setdefault(path, []).append({"ms_per_step": ms, "edges_per_s": B * N * k / ms * 1e3,
                                         "equal_to_first": bool(torch.equal(y0, ref))})
    _native.set_knn_path("auto")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
