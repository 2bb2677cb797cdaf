"""bench.py --config plus: the LatticeFold+ range check at the reference's own benchmark rows --
crates/latticefold-plus/benches/utils/mod.rs `range_check::WITNESS_SCALING` (n, k, kappa) = (2^15 .. 2^19, 2, 2), benches/rgchk.rs
(`RangeCheck-Prover`: Rg::range_check with the instance built outside the timed closure) and benches/double_commitment.rs
(`RgInstance::from_f`).  One GPU; a "step" is one range_check of one instance.  `value` = witness ring elements per second through
range_check with the instance resident; `e2e` = from_f from a HOST witness + range_check + the proof image back on the host.
After timing, the proof of the last step is checked by the product's verifier and (at sizes the oracle finishes) compared bit for
bit with the CPU oracle's.  `--impl reference` times the oracle (the reference's dense ring-valued algorithm) on the host cores."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, KAPPA, B = 2, 2, 8
METRIC = "LatticeFold+ range-check prover witness elements/sec (set-check sumcheck + evaluations)"
METRIC_CM = "LatticeFold+ commitment-transformation prover witness elements/sec (range check + two cm sumchecks + g)"


def config_of(n, op="rgchk", L=1):
    if op == "cm":
        return {"workload": f"latticefold-plus benches commitment_transform::FOLDING_ARITY row (L, n, k, kappa) = ({L}, {n}, {K}, {KAPPA}): frog ring (X^16 + 1, coefficient form), "
                            f"Cm::prove with M = [] (range check + two sumchecks + g)", "n": n, "L": L, "k": K, "kappa": KAPPA, "b": B,
                "l2": "sumcheck tables exceed L2 (range check 69 L x n x 8 B, cm 2 x n x 128 B)"}
    return {"workload": f"latticefold-plus benches range_check::WITNESS_SCALING row (n, k, kappa) = ({n}, {K}, {KAPPA}): frog ring (X^16 + 1, coefficient form), "
                        f"one instance, Rg::range_check with M = []", "n": n, "k": K, "kappa": KAPPA, "b": B,
            "l2": "set-check tables 69 x n x 8 B exceed L2 from n = 2^18; smaller rows are L2-resident"}


METRIC_FOLD = "LatticeFold+ prover witness elements/sec (PlusProver::prove: L x R1CS linearization + mlin + decomposition)"


def fold_bench(args, reference_arm):
    """--plus-op fold: benches/e2e.rs `E2E-Prover` (n, L, k, kappa): a fresh PlusProver folds L committed R1CS instances.  Every step starts from
    host witnesses (the orchestration of plus.rs lives on the host side of the C ABI), so `value` and `e2e` are the same measurement."""
    sys.path.insert(0, ROOT)
    from tests import plus_cases as pc
    L = args.plus_instances or 3
    n = 1 << (min(args.log_w, 15) if reference_arm else args.log_w); l = pc.frog_l(); Bfold = 1 << 11
    _, A = pc.range_check_inputs(n, KAPPA, seed=1, k=K)
    abc, f0 = pc.r1cs_instance(n, 5)
    fs = [f0] + [np.ascontiguousarray(np.roll(f0, i + 1, axis=0)) for i in range(L - 1)]
    for f in fs:
        f[0, 0] = 1
    cfg = {"workload": f"latticefold-plus benches/e2e.rs E2E-Prover row (n, L, k, kappa) = ({n}, {L}, {K}, {KAPPA}): frog ring (X^16 + 1), identity-shaped R1CS with 0/1 witnesses, "
                       f"one PlusProver::prove from an empty accumulator, B = {Bfold}", "n": n, "L": L, "k": K, "kappa": KAPPA, "B": Bfold, "l2": "witnesses and sumcheck tables exceed L2"}
    from oracle.pyoracle import Oracle
    if reference_arm:
        orc = Oracle(); ts = []
        for i in range(args.warmup + args.steps):
            flow = pc.OraclePlus(orc, A, abc, B, K, l, Bfold)
            t0 = time.perf_counter(); flow.prove([(abc, f) for f in fs]); dt = time.perf_counter() - t0
            if i >= args.warmup:
                ts.append(dt)
        ms = 1e3 * float(np.mean(ts)); v = L * n / (ms / 1e3)
        print(json.dumps(dict(metric=METRIC_FOLD, value=v, unit="elements/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                              dtype="u64 (mod 15912092521325583641)", data="synthetic", impl="reference", config=cfg,
                              cpu_baseline=dict(value=v, unit="elements/s", cores=orc.threads(), kind="port", sample=f"PlusProver::prove at n = {n}, L = {L} on the oracle"),
                              e2e=dict(value=v, unit="elements/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return
    import torch
    import latticefold_b200 as lf
    from latticefold_b200 import plus
    from bench import ClockSampler
    ctx = lf.Context(pc.RING_FROG, 0); Ad = plus.Matrix(ctx, A)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))
    comps = [plus.ComR1CS(ctx, abc, f) for f in fs]

    def step():
        return plus.PlusProver(ctx, Ad, abc, B, K, l, Bfold, plus.PoseidonTranscript()).prove(comps)

    def step_e2e():
        pr = plus.PlusProver(ctx, Ad, abc, B, K, l, Bfold, plus.PoseidonTranscript())
        out = pr.prove([plus.ComR1CS(ctx, abc, f) for f in fs])
        pr.acc_download()
        return out

    for _ in range(max(args.warmup, 3)):
        step()
    l0 = ctx.launches(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as cs:
        torch.cuda.synchronize(); a.record(stream)
        for _ in range(args.steps):
            proof = step()
        b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps; launches = (ctx.launches() - l0) // args.steps
    step_e2e(); torch.cuda.synchronize(); a.record(stream)
    for _ in range(args.steps):
        step_e2e()
    b.record(stream); torch.cuda.synchronize(); ms_e2e = a.elapsed_time(b) / args.steps
    ctx.profile(True); step(); rep = ctx.profile_report(); ctx.profile(False)
    kern = sorted(rep.items(), key=lambda kv: -kv[1][1])
    ok = plus.PlusVerifier(KAPPA, len(abc), Bfold, plus.PoseidonTranscript()).verify(proof)
    verify = dict(product_verifier="accept" if ok else "REJECT")
    cpu = None
    if not args.no_cpu_baseline:
        orc = Oracle(); sn = 1 << min(args.log_w, 15)
        if sn == n:
            flow = pc.OraclePlus(orc, A, abc, B, K, l, Bfold); t0 = time.perf_counter(); want = flow.prove([(abc, f) for f in fs]); dt = time.perf_counter() - t0
            verify["oracle_bit_exact"] = bool(np.array_equal(want["cmproof"], proof["cmproof"]) and np.array_equal(want["dproof"], proof["dproof"]) and all(np.array_equal(x, y) for x, y in zip(want["lproof"], proof["lproof"])))
            verify["oracle_verifier"] = "accept" if flow.verify(proof) else "REJECT"
        else:
            _, sA = pc.range_check_inputs(sn, KAPPA, seed=1, k=K); sabc, sf0 = pc.r1cs_instance(sn, 5)
            flow = pc.OraclePlus(orc, sA, sabc, B, K, l, Bfold); t0 = time.perf_counter(); flow.prove([(sabc, sf0)] * L); dt = time.perf_counter() - t0
        cpu = dict(value=L * sn / dt, unit="elements/s", cores=orc.threads(), kind="port", sample=f"one PlusProver::prove at n = {sn}, L = {L} on the oracle ({dt:.1f} s)")
    verify["verified"] = ok and verify.get("oracle_bit_exact", True) and verify.get("oracle_verifier", "accept") == "accept"
    v = L * n / (ms * 1e-3)
    print(json.dumps(dict(metric=METRIC_FOLD, value=v, unit="elements/s", n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="u64 (mod 15912092521325583641, Montgomery on the device)", data="synthetic", config=cfg, clocks=cs.summary(), gpu_launches=int(launches),
                          value_note="value: the instances' witnesses resident on the device (uploaded once), the accumulated witnesses stay there; e2e: fresh ComR1CS objects per step (witness upload inside) and the two accumulated witnesses read back",
                          e2e=dict(value=L * n / (ms_e2e * 1e-3), unit="elements/s", ms_per_step=ms_e2e, h2d_bytes_per_step=int(sum(f.nbytes for f in fs)), d2h_bytes_per_step=int(2 * fs[0].nbytes)),
                          roofline=dict(bound="latency", kernel=kern[0][0], note="host-paced sumcheck rounds (L + 3 sumchecks per step); per-kernel device time below",
                                        kernels=[dict(kernel=k_, launches=c, total_ms=round(t, 4)) for k_, (c, t) in kern if t >= 0.05]),
                          cpu_baseline=cpu, verify=verify, verified=verify["verified"])))


def reference(args, rank):
    if rank != 0:
        return
    if args.plus_op == "fold":
        return fold_bench(args, True)
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle
    from tests import plus_cases as pc
    orc = Oracle(); threads = orc.threads()
    op = args.plus_op; L = args.plus_instances or (2 if op == "cm" else 1)
    n = 1 << min(args.log_w, 15)      # bounded sample: the dense ring-valued sumcheck of the reference algorithm is ~ 3 s per 2^15 elements
    fs, A = pc.range_check_inputs(n, KAPPA, seed=1, k=K, L=L)
    l = pc.frog_l()
    ts = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if op == "cm":
            orc.plus_cm_prove(2, n.bit_length() - 1, fs, A, B, K, l, want_g=False)
        else:
            orc.plus_range_check(2, n.bit_length() - 1, fs, A, B, K, l)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            ts.append(dt)
    # the oracle entry point runs from_f inside; time it alone and subtract
    t0 = time.perf_counter(); orc.plus_rg_from_f(2, fs[0], A, B, K, l); t_from = time.perf_counter() - t0
    ms = max(1e3 * (float(np.mean(ts)) - L * t_from), 1e-3); v = L * n / (ms / 1e3)
    print(json.dumps(dict(metric=METRIC_CM if op == "cm" else METRIC, value=v, unit="elements/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                          scaling="weak", vs_baseline=None, dtype="u64 (mod 15912092521325583641)", data="synthetic", impl="reference", config=config_of(n, op, L),
                          cpu_baseline=dict(value=v, unit="elements/s", cores=threads, kind="port", sample=f"{op} at n = {n}, L = {L} (from_f {1e3 * t_from:.0f} ms per instance subtracted)"),
                          e2e=dict(value=v, unit="elements/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main(args, rank, world, local):
    if args.impl == "reference":
        return reference(args, rank)
    if rank != 0:
        return      # one instance does not shard: replicas only
    if args.plus_op == "fold":
        return fold_bench(args, False)
    import torch
    import latticefold_b200 as lf
    from latticefold_b200 import plus
    sys.path.insert(0, ROOT)
    from tests import plus_cases as pc
    from bench import ClockSampler, peaks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --config plus: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    n = 1 << args.log_w; nvars = args.log_w; l = pc.frog_l()
    op = args.plus_op; L = args.plus_instances or (2 if op == "cm" else 1)
    fs, A = pc.range_check_inputs(n, KAPPA, seed=1, k=K, L=L)
    ctx = lf.Context(pc.RING_FROG, local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    Ad = plus.Matrix(ctx, A)
    f_pinned = torch.from_numpy(fs.view(np.int64)).pin_memory().numpy().view(np.uint64)
    insts = [plus.RgInstance.from_f(ctx, f_pinned[i], Ad, B, K, l) for i in range(L)]
    g_pinned = torch.empty((L, n, 16), dtype=torch.int64).pin_memory().numpy().view(np.uint64) if op == "cm" else None
    comx_box = [None]

    def prove(instances, with_g):
        if op == "cm":
            pf, comx, _ = plus.Cm(plus.Rg(ctx, nvars, instances)).prove([], plus.PoseidonTranscript(), want_g=with_g, g_out=g_pinned if with_g else None)
            comx_box[0] = comx
            return pf
        return plus.Rg(ctx, nvars, instances).range_check([], plus.PoseidonTranscript())

    def step_resident():      # instances resident; the folded witness g stays on the device side of the boundary
        return prove(insts, False)

    def step_e2e():           # witnesses from host memory, proof (and g) back on the host
        return prove([plus.RgInstance.from_f(ctx, f_pinned[i], Ad, B, K, l) for i in range(L)], True)

    def timed(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(stream)
        for _ in range(steps):
            out = fn()
        b.record(stream); torch.cuda.synchronize()
        return a.elapsed_time(b) / steps, out

    for _ in range(max(args.warmup, 3)):
        step_resident(); step_e2e()
    l0 = ctx.launches()
    with ClockSampler(local) as cs:
        ms_res, proof = timed(step_resident, args.steps)
        launches = (ctx.launches() - l0) // args.steps
        ms_e2e, proof2 = timed(step_e2e, args.steps)
    t0 = time.perf_counter(); plus.RgInstance.from_f(ctx, f_pinned[0], Ad, B, K, l); ctx.sync(); ms_from_f = 1e3 * (time.perf_counter() - t0)
    # per-kernel device time of one more step
    ctx.profile(True); step_resident(); rep = ctx.profile_report(); ctx.profile(False)
    kern = sorted(rep.items(), key=lambda kv: -kv[1][1])
    peak, peak_src = peaks()
    top, (cnt, tot) = kern[0]
    n_tables = L * (K * 33 + 3)
    # k_plus_round reads every table once per round (n, n/2, ...: 2 n words per table in all), k_plus_fold reads n and writes n / 2 per round
    alg = {"k_plus_round": 2 * n_tables * n * 8, "k_plus_fold": 3 * n_tables * n * 8 + (2 * 3 * 2 * n * 136 if op == "cm" else 0),
           "k_plus_cm_round": 2 * 2 * 2 * n * 136, "k_plus_lin4": 2 * L * n * (2 * 128 + 2 + 128), "k_plus_h": L * n * (K * 16 + 128), "k_plus_g": L * n * (2 + 3 * 128)}
    roof = dict(bound="hbm", kernel=top, launches_per_step=cnt, avg_launch_ms=tot / cnt, achieved=None, peak=peak, unit="GB/s", frac=None, traffic=None, peak_source=peak_src,
                note="algorithmic bytes of k_plus_round = 2 x 69 tables x n x 8 B over the rounds of one step (every table entry is one base-field word; the reference holds 128 B per entry)")
    if top in alg:
        roof["achieved"] = alg[top] / (tot * 1e-3) / 1e9; roof["frac"] = roof["achieved"] / peak; roof["algorithmic_bytes_per_step"] = alg[top]
    roof["kernels"] = [dict(kernel=k, launches=c, total_ms=round(t, 4), algorithmic_bytes=alg.get(k), achieved_gbs=(alg[k] / (t * 1e-3) / 1e9 if k in alg else None)) for k, (c, t) in kern if t >= 0.02]
    if op == "cm":
        okv, comx_v = plus.cm_verify(proof, 0, plus.PoseidonTranscript(), nvars=nvars, L=L, kappa=KAPPA)
        verify = dict(product_verifier="accept" if okv and np.array_equal(comx_v, comx_box[0]) else "REJECT", entry_points_agree=bool(np.array_equal(proof, proof2)))
    else:
        verify = dict(product_verifier="accept" if plus.range_check_verify(proof, plus.PoseidonTranscript()) else "REJECT", entry_points_agree=bool(np.array_equal(proof, proof2)))
    cpu = None
    if not args.no_cpu_baseline or not args.no_verify:
        from oracle.pyoracle import Oracle
        orc = Oracle()
        sn = 1 << min(args.log_w, 15)
        sf, sA = (fs, A) if sn == n else pc.range_check_inputs(sn, KAPPA, seed=1, k=K, L=L)
        t0 = time.perf_counter()
        if op == "cm":
            want, want_x, want_g = orc.plus_cm_prove(2, sn.bit_length() - 1, sf, sA, B, K, l)
        else:
            want = orc.plus_range_check(2, sn.bit_length() - 1, sf, sA, B, K, l)
        t_all = time.perf_counter() - t0
        t0 = time.perf_counter(); orc.plus_rg_from_f(2, sf[0], sA, B, K, l); t_from = L * (time.perf_counter() - t0)
        cpu = dict(value=L * sn / max(t_all - t_from, 1e-6), unit="elements/s", cores=orc.threads(), kind="port", sample=f"one {op} at n = {sn}, L = {L} on the oracle ({t_all - t_from:.2f} s; from_f {t_from:.2f} s not counted)")
        verify["oracle_verifier"] = "accept" if (orc.plus_cm_verify(2, proof, [])[0] if op == "cm" else orc.plus_range_check_verify(2, proof)) else "REJECT"
        if sn == n:
            verify["oracle_bit_exact"] = bool(np.array_equal(want, proof)) and (op != "cm" or (bool(np.array_equal(want_x, comx_box[0])) and bool(np.array_equal(want_g, g_pinned))))
    verify["verified"] = verify["product_verifier"] == "accept" and verify["entry_points_agree"] and verify.get("oracle_verifier", "accept") == "accept" and verify.get("oracle_bit_exact", True)
    line = dict(metric=METRIC_CM if op == "cm" else METRIC, value=L * n / (ms_res * 1e-3), unit="elements/s", n_gpus=1, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_res, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="u64 (mod 15912092521325583641, Montgomery on the device)", data="synthetic", config=config_of(n, op, L), clocks=cs.summary(),
                gpu_launches=int(launches), e2e=dict(value=L * n / (ms_e2e * 1e-3), unit="elements/s", ms_per_step=ms_e2e, h2d_bytes_per_step=int(fs.nbytes), d2h_bytes_per_step=int(proof.nbytes) + (int(g_pinned.nbytes) if op == "cm" else 0)),
                roofline=roof, cpu_baseline=cpu, phases_ms=dict(from_f_ms=ms_from_f, **{("cm_prove_ms" if op == "cm" else "range_check_ms"): ms_res}), verify=verify, verified=verify["verified"])
    print(json.dumps(line))
