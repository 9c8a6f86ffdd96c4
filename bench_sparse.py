#!/usr/bin/env python
"""Secondary benchmark (not the driver's contract line): fused GENConv aggregate on the
ogbn-arxiv-shaped synthetic CSR of BASELINE config 3 (169,343 nodes, ~2.5 M edges after
to_undirected + self loops, C=128) and, with --products, the ogbn-products-shaped one
(2.449 M nodes, 61.9 M edges).  Reports edges/s and the HBM roofline of SURVEY.md 8d
(gather model 4C+4 B/edge + (8C+4) B/node)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--products", action="store_true")
    ap.add_argument("--zipf", action="store_true", help="power-law in-degrees (hub rows)")
    ap.add_argument("--no-hubs", action="store_true", help="disable the CTA-per-hub-row kernel")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--aggr", default="softmax_sg")
    ap.add_argument("--cpu", action="store_true", help="also time the oracle port on the host")
    a = ap.parse_args()
    from deep_gcns_torch_b200 import _native
    from deep_gcns_torch_b200.gcn_lib import sparse as S
    from oracle import sparse as osp
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    C = 128
    if a.products:
        N, E0 = 2449029, 61859140
        ei = torch.stack((torch.randint(0, N, (E0,), generator=g), torch.randint(0, N, (E0,), generator=g)))
    else:
        N = 169343
        s, d = torch.randint(0, N, (1166243,), generator=g), torch.randint(0, N, (1166243,), generator=g)
        ei = osp.to_undirected_with_self_loops(s, d, N)
    if a.zipf:
        u = torch.rand(ei.shape[1], generator=g).clamp_min(1e-9)
        ei[1] = (u.pow(-2.0) - 1).clamp(max=N - 1).long()
    E = ei.shape[1]
    x = torch.randn(N, C, generator=g).to(dev)
    eic = ei.to(dev)
    t0 = time.perf_counter()
    csr = _native.csr_build(eic, N)
    if a.no_hubs:
        csr = csr[:3]
    torch.cuda.synchronize()
    csr_ms = (time.perf_counter() - t0) * 1e3
    conv = S.GENConv(C, C, aggr=a.aggr, t=0.1, mlp_layers=1).to(dev).eval()
    prm, keep = _native.genconv_params(a.aggr, 0.1, 1.0, 0.0, 1e-7, None, add_residual=True)
    for _ in range(5):
        _native.genconv_aggregate(x, x, csr, prm)
    torch.cuda.synchronize()
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for _ in range(a.steps):
        flush.zero_()                      # evict L2 between iterations (arxiv x = 87 MB would otherwise sit in L2)
        beg.record()
        _native.genconv_aggregate(x, x, csr, prm)
        end.record()
        torch.cuda.synchronize()
        times.append(beg.elapsed_time(end))
    times.sort()
    ms = times[len(times) // 2]
    peak = 6563.9
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    gather_bytes = E * (4 * C + 4) + N * (8 * C + 4)
    compulsory = 2 * N * C * 4 + E * 4 + N * 4
    out = {"workload": "GENConv aggregate %s, N=%d E=%d C=%d" % (a.aggr, N, E, C), "ms": ms,
           "edges_per_s": E / (ms * 1e-3), "csr_build_ms_one_time": csr_ms,
           "roofline": {"bound": "hbm", "gather_model_GBps": gather_bytes / (ms * 1e-3) / 1e9,
                        "compulsory_GBps": compulsory / (ms * 1e-3) / 1e9, "peak": peak,
                        "frac_gather_model": gather_bytes / (ms * 1e-3) / 1e9 / peak}}
    if a.cpu:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        xc = x.cpu()
        osp.genconv_pre_mlp(xc, ei, None, a.aggr, 0.1)
        t0 = time.perf_counter()
        osp.genconv_pre_mlp(xc, ei, None, a.aggr, 0.1)
        out["cpu_baseline"] = {"edges_per_s": E / (time.perf_counter() - t0), "cores": torch.get_num_threads(),
                               "kind": "port"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
