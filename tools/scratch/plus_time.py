import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import latticefold_b200 as lf
from latticefold_b200 import plus
from tests import plus_cases as pc
for lg in (17, 19):
    n = 1 << lg
    fs, A = pc.range_check_inputs(n, 2, seed=1, k=2)
    ctx = lf.Context(2, 0)
    Ad = plus.Matrix(ctx, A)
    f = torch.from_numpy(fs[0].view(np.int64)).pin_memory().numpy().view(np.uint64)
    for it in range(3):
        t0 = time.perf_counter(); inst = plus.RgInstance.from_f(ctx, f, Ad, 8, 2, pc.frog_l()); ctx.sync(); print("n=2^%d from_f total %.3f ms" % (lg, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
        t0 = time.perf_counter(); out = plus.Rg(ctx, lg, [inst]).range_check([], plus.PoseidonTranscript()); print("   range_check %.3f ms" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
        t0 = time.perf_counter(); del inst; print("   free %.3f ms" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
