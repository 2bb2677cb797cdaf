"""A/B timing of the batched commit kernel variants (LF_DOT_CT x LF_DOT_WPB) at kappa=26, n=2^18, 15 pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latticefold_b200 as lf
from latticefold_b200 import synth
R = 0; p = synth.RINGS[R]["p"]
kappa, n, cnt = 26, 1 << 18, 15
ctx = lf.Context(R, 0)
A = synth.uniform_field(p, kappa * n * 24, 1).reshape(kappa, n, 24)
sch = lf.AjtaiCommitmentScheme(ctx, A)
fs = [ctx.upload(synth.uniform_field(p, n * 24, 10 + i).reshape(n, 24)) for i in range(cnt)]
ref = None
for ct in (4, 2, 1):
    for wpb in (4, 8, 13, 16):
        os.environ["LF_DOT_CT"], os.environ["LF_DOT_WPB"] = str(ct), str(wpb)
        out = sch.commit_batch(fs)
        if ref is None: ref = out
        assert np.array_equal(out, ref)
        ctx.profile(True)
        for _ in range(3): sch.commit_batch(fs)
        rep = ctx.profile_report(); ctx.profile(False)
        c, ms = rep["k_dot"]
        print(f"CT={ct} WPB={wpb}: {ms / c:.3f} ms per batched commit", flush=True)
