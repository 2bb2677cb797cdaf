"""Every LatticeFold+ entry point once at small sizes (for compute-sanitizer memcheck / racecheck runs)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import latticefold_b200 as lf
from latticefold_b200 import plus
from tests import plus_cases as pc
ctx = lf.Context(pc.RING_FROG, 0)
T = plus.PoseidonTranscript
for name in ("test_set_check_mix", "rect_with_M", "short_rows"):
    nvars, sets, M, acc = pc.set_check_cases()[name]
    out = plus.In(ctx, nvars, sets).set_check(M, T()); assert plus.set_check_verify(out, T()) is acc
n, kappa, k, l = 1 << 14, 1, 2, pc.frog_l()
fs, A = pc.range_check_inputs(n, kappa, seed=7, L=2)
m = pc.identity(n); m["val"] = m["val"].copy(); m["val"][0, 0] = 2
Ad = plus.Matrix(ctx, A)
inst = [plus.RgInstance.from_f(ctx, fs[i], Ad, 8, k, l) for i in range(2)]
d = plus.Rg(ctx, 14, inst).range_check([m], T()); assert plus.range_check_verify(d, T())
pf, cx, g = plus.Cm(plus.Rg(ctx, 14, inst)).prove([m], T()); assert plus.cm_verify(pf, 1, T())[0]
abc, f = pc.r1cs_instance(96, 3)
_, lp = plus.ComR1CS(ctx, abc, f).linearize(T()); assert plus.r1cs_linearize_verify(lp, T())
abc, f = pc.r1cs_instance(n, 4)
pr = plus.PlusProver(ctx, Ad, abc, 8, k, l, 1 << 11, T())
p1 = pr.prove([plus.ComR1CS(ctx, abc, f), plus.ComR1CS(ctx, abc, f)])
assert plus.PlusVerifier(kappa, 3, 1 << 11, T()).verify(p1)
dp, F = plus.decompose(ctx, Ad, g[0], np.stack([cx[2 * 16: 2 * 16 + 28][0::2], cx[2 * 16: 2 * 16 + 28][1::2]], axis=1), [m], 1 << 12)
print("ok", ctx.launches())
