"""BASELINE configs[4] (reference analogue): elementwise CRT / ICRT ("INTT->NTT" / "NTT->INTT" of
crates/latticefold/build.rs:467-499) over N = 2^15 .. 2^20 ring elements, achieved GB/s vs the HBM roofline.
Algorithmic bytes = 2 * N * E (read + write, E = 192 B for Goldilocks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import latticefold_b200 as lf
from latticefold_b200 import synth

ring = int(sys.argv[1]) if len(sys.argv) > 1 else 0
R = synth.RINGS[ring]; E = R["d"] * 8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
ctx = lf.Context(ring, 0)
rows = []
for lg in range(15, 21):
    N = 1 << lg
    a = synth.uniform_field(R["p"], N * R["d"], lg).reshape(N, R["d"])
    v = ctx.upload(a, 1)
    for _ in range(3):
        w = ctx.crt(v); u = ctx.icrt(w)
    ctx.profile(True)
    outs = []
    for _ in range(10):
        w = ctx.crt(v); u = ctx.icrt(w); outs.append((w, u))
    rep = ctx.profile_report(); ctx.profile(False)
    cnt, ms = rep["k_matrix_apply"]
    gbs = 2 * N * E / 1e9 / (ms / cnt / 1e3)
    assert np.array_equal(u.download(), a)
    rows.append(dict(N=N, us_per_transform=1e3 * ms / cnt, GBps=gbs, frac_of_hbm=gbs / peak))
    print(f"{R['name']} N=2^{lg}: {1e3 * ms / cnt:8.1f} us per transform, {gbs:7.1f} GB/s = {100 * gbs / peak:5.1f}% of {peak:.0f} GB/s", flush=True)
    del outs
print(json.dumps(dict(ring=R["name"], peak_gbs=peak, rows=rows)))
