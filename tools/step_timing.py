"""Diagnostic: per-phase wall times of resident prover steps at a given size (LF_TIMING_DETAIL=1)."""
import os, sys, time
os.environ["LF_TIMING_DETAIL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import latticefold_b200 as lf
from latticefold_b200 import synth
import bench
logw = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wl = bench.workload(logw); R = 0
ctx = lf.Context(R, 0)
prob = synth.make_instance(R, wl["W"], wl["B"], wl["L"], wl["b"], wl["K"], wl["kappa"], kind=wl["kind"], config_id=100, ops=None)
pr = lf.NIFSProver(ctx, prob)
f = ctx.witness_f_from_w_ccs(R, prob["w_ccs"], wl["B"], wl["L"])
prob["w_i_f"] = prob["w_acc_f"] = f
prob["cm_i_cm"] = np.ascontiguousarray(bench._commit_with_prover(ctx, pr, lf, prob, f))
lc, _ = pr.linearize(prob, lf.Transcript(R)); prob["acc"] = synth.split_lcccs(R, prob, lc)
wa, wi = pr.upload_witness(f), pr.upload_witness(f)
for i in range(steps):
    t = time.time(); pr.prove_resident(prob, wa, wi, lf.Transcript(R)); dt = (time.time() - t) * 1e3
    print(f"step {i}: {dt:.1f} ms", " ".join(f"{n}={v:.1f}" for n, v in pr.timing_detail()))
