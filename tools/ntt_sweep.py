"""BASELINE configs[4]: batched negacyclic NTT / INTT throughput sweep, degree 2^10 .. 2^16, achieved GB/s
against the HBM roofline.  Algorithmic bytes per transform = 2 * N * batch * sizeof(word) (read once, write once).

  python tools/ntt_sweep.py [--field 0|1|both] [--logs 10,12,16] [--mb 2048] [--iters 10] [--check]

Timing: CUDA events recorded on the library's own stream (lf_ctx_profile brackets every launch); buffers are far larger than the
126 MB L2 and the input is re-read from HBM every iteration.  One JSON line at the end."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import latticefold_b200 as lf
from latticefold_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--field", default="both"); ap.add_argument("--logs", default="10,11,12,13,14,15,16")
ap.add_argument("--mb", type=int, default=2048, help="bytes per buffer (MiB); batch = mb / (N * word)")
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--check", action="store_true")
args = ap.parse_args()
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
P = {0: 0xFFFFFFFF00000001, 1: 2013265921}; NAME = {0: "goldilocks", 1: "babybear"}
ctx = lf.Context(synth.RING_GOLDILOCKS, 0)
fields = [0, 1] if args.field == "both" else [int(args.field)]
rows = []
for field in fields:
    word = 8 if field == 0 else 4; tdt = torch.int64 if field == 0 else torch.int32
    for lg in [int(x) for x in args.logs.split(",")]:
        n = 1 << lg; batch = max(1, (args.mb << 20) // (n * word))
        g = torch.Generator(device="cuda"); g.manual_seed(lg)
        a = torch.randint(0, P[field] if field == 1 else (1 << 62), (batch, n), dtype=tdt, device="cuda", generator=g)   # canonical: < p
        f = torch.empty_like(a); b = torch.empty_like(a)
        torch.cuda.synchronize()
        plan = lf.NttPlan(ctx, field, lg)
        for _ in range(3):
            plan.forward_device(a.data_ptr(), f.data_ptr(), batch); plan.inverse_device(f.data_ptr(), b.data_ptr(), batch)
        ctx.sync()
        if args.check:
            assert torch.equal(a, b), "round trip failed"
        res = {}
        for name, fn, src, dst in (("fwd", plan.forward_device, a, f), ("inv", plan.inverse_device, f, b)):
            ctx.profile(True)
            for _ in range(args.iters):
                fn(src.data_ptr(), dst.data_ptr(), batch)
            rep = ctx.profile_report(); ctx.profile(False)
            ms = sum(v[1] for v in rep.values()) / args.iters
            res[name] = dict(ms=ms, GBps=2 * n * batch * word / 1e9 / (ms / 1e3), launches=sum(v[0] for v in rep.values()) // args.iters)
        rows.append(dict(field=NAME[field], log_n=lg, batch=batch, fwd=res["fwd"], inv=res["inv"]))
        print(f"{NAME[field]:10s} N=2^{lg:<2d} batch={batch:<8d} fwd {res['fwd']['ms']:7.3f} ms {res['fwd']['GBps']:7.1f} GB/s ({100 * res['fwd']['GBps'] / peak:4.1f}%)   "
              f"inv {res['inv']['ms']:7.3f} ms {res['inv']['GBps']:7.1f} GB/s ({100 * res['inv']['GBps'] / peak:4.1f}%)", flush=True)
        plan.close(); del a, f, b
print(json.dumps(dict(metric="negacyclic NTT throughput", unit="GB/s", peak_gbs=peak, bytes_per_transform="2*N*batch*word", rows=rows)))
