#!/usr/bin/env python3
"""bench.py -- LatticeFold prover step throughput (R1CS constraints / second) on B200.

  python bench.py --gpus N --steps K --warmup W            product arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the C++ restatement of the reference algorithm
                                                           (oracle/, kind "port": the reference itself is Rust with an
                                                           un-vendored dependency and cannot be built in this image)
  python bench.py --config c3 [--log-w 20]                 BASELINE configs[2]: BabyBear ring, degree-three CCS, one GPU
  python bench.py --config ntt [--gpus N]                  BASELINE configs[4]: negacyclic NTT sweep (GB/s per size, JSON line)
  python bench.py --config plus [--log-w 17]               LatticeFold+ range check at the reference's benchmark rows

A step = one NIFSProver::prove (linearization + 2 decompositions + folding; crates/latticefold/benches/utils.rs:619-680)
on the configuration BASELINE.json quotes the metric on (configs[1]: Goldilocks ring, 2^16-constraint R1CS, 1 GPU).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from latticefold_b200 import synth  # noqa: E402

METRIC = "prover constraints/sec (commit+decomp+sumcheck)"
DTYPE = {"c2": "u64 (mod 2^64-2^32+1)", "c3": "u32 (mod 15*2^27+1)"}


def workload(log_w, config="c2"):
    """configs[1] of BASELINE.json = SURVEY 8 row C2 (see synth.bench_workload); c3 = configs[2]."""
    return synth.bench_workload(config, log_w)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:        # NVML in-process: no nvidia-smi process per sample, far less intrusive for the timed region
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active" for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                self.stop.wait(0.1)
            return
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.5)

    def __enter__(self):
        self.t.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def cpu_step(orc, prob, threads):
    orc.set_threads(threads)
    _, _, _, ms = orc.nifs_prove(prob, orc.transcript(prob["ring"]), want_f=True)
    return ms


def run_reference(args, rank, world):
    """CPU arm: C++ restatement of the reference algorithm on the host cores, on the SAME configuration as the product arm
    (same instance generator, same W).  A full-size step takes about a minute on 16 cores, so the number of timed steps is capped
    by a wall-clock budget (--cpu-budget-s, default 240 s; at least one step) and no untimed warm-up step is spent: the line's
    `steps` / `warmup` are the counts actually run, the requested ones are kept under `requested`."""
    if rank != 0:
        return
    from oracle.pyoracle import Oracle
    from tests.helpers import OracleOps
    from tools.make_bench_golden import oracle_bench_problem
    orc = Oracle()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    import math
    log_w = args.log_w + int(math.log2(max(args.gpus, 1)))
    if args.cpu_sample_log_w is not None:
        log_w = min(log_w, args.cpu_sample_log_w)
    t_setup = time.time()
    wl, prob = oracle_bench_problem(orc, args.config, log_w)
    t_setup = time.time() - t_setup
    t, t0 = [], time.time()
    while len(t) < max(args.steps, 1) and (not t or (time.time() - t0) + t[-1] / 1e3 < args.cpu_budget_s):
        t.append(cpu_step(orc, prob, cores))
    ms = float(np.mean(t))
    value = prob["constraints"] / (ms / 1e3)
    full = wl["W"] == synth.bench_workload(args.config, args.log_w + int(math.log2(max(args.gpus, 1))))["W"]
    sample = (f"{len(t)} full-size step(s) at W=2^{log_w} ({'the same configuration as the product arm' if full else 'a slice of the product arm configuration'}; "
              f"requested steps={args.steps}, warmup={args.warmup}; capped by a {args.cpu_budget_s:.0f} s budget; instance set-up {t_setup:.0f} s untimed)")
    line = dict(metric=METRIC, value=value, unit="constraints/s", n_gpus=args.gpus, steps=len(t), warmup=0,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=DTYPE[args.config], data="synthetic", impl="reference",
                requested=dict(steps=args.steps, warmup=args.warmup),
                config=config_of(wl, args, "cpu"), cpu_baseline=dict(value=value, unit="constraints/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="constraints/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def config_of(wl, args, where):
    R = synth.RINGS[wl["ring"]]
    circuit = "dummy R1CS" if wl["degree"] == 2 else "degree-three CCS (arith/ccs.rs:14-43)"
    which = "configs[1]; weak-scaled to W = gpus * 2^%d when gpus > 1, as configs[3]" % args.log_w if wl["config"] == "c2" else "configs[2], as a whole prover step"
    return dict(workload=f"{R['name']} ring (d = {R['d']}), {circuit} ({wl['kind']} witness) with {wl['W']}+2 constraints, one NIFSProver::prove step "
                         f"(BASELINE.json {which})", W=wl["W"], B=wl["B"], L=wl["L"], b=wl["b"], K=wl["K"], kappa=wl["kappa"], n=wl["W"] * wl["L"],
                parallelism=("cpu threads" if where == "cpu" else ("1 GPU" if args.gpus == 1 else
                             f"{args.gpus} GPUs: witness columns / hypercube sharded, one all-reduce per commit batch, sumcheck round and evaluation")),
                l2="inputs larger than L2: Ajtai matrix + 2K witness pieces per step are GBs vs 126 MB L2")


def algorithmic_bytes(wl, ccs, world):
    """SURVEY 8(d) per-unit figures in the REFERENCE layout (E bytes per ring element), summed over each kernel's launches in one
    step, per rank.  These are the bytes the reference's data structures would move for the same work; kernels that read int8
    digits or packed limbs move fewer real bytes, which is the point of those layouts."""
    R = synth.RINGS[wl["ring"]]
    E = R["d"] * (4 if R["p"] < (1 << 32) else 8)      # SURVEY 8(d) sizes the 31-bit BabyBear ring with 4-byte limbs (288 B per element)
    s, t, d_ccs = ccs["s"], ccs["t"], ccs["d"]
    n, K, kappa, L = wl["W"] * wl["L"] // world, wl["K"], wl["kappa"], wl["L"]
    m = (1 << s) // world
    rows = wl["W"] + 2                                  # effective rows of every Mz table (the dummy circuits have W + 2 non-empty rows)
    tau = R["tau"]
    M_fold, M_lin = 5 + 2 * K * tau, t + 1
    commit = 2 * (kappa * n + (K - 1) * n + (K - 1) * kappa) * E
    return {
        "k_commit_mma": commit, "k_dot_commit": commit,
        # K9 with the fold fused into the next round's evaluation: round 1 reads M 2^s, round r >= 2 reads M 2^(s-r+2) and writes M 2^(s-r+1)
        "k_fold_sc_round1": M_fold * m * E,
        "k_fold_sc_round": M_fold * (3 * m // 2) * E,                       # sum over r >= 2 of (2^(s-r+2) + 2^(s-r+1)) ~ 3 * 2^(s-1)
        "k_fold_sc_round2": M_fold * (m + m // 4) * E,
        "k_sc_generic": M_lin * (m + 3 * m // 2) * E, "k_sc_wide": M_lin * 2 * m * E,
        "k_fold": None, "k_fold_digits": M_fold * (m + m // 2) * E,
        "k_matrix_apply": 2 * (2 * K + 1) * n * E,                           # CRT of the 2K pieces + ICRT of the folded witness
        "k_gadget_recompose": (2 * K + 1) * (n + n // L) * E,
        "k_digit_split": 2 * (n + K * n) * E,
        "k_coeff_eval": (2 * K + 1 + 3) * n * E,                            # v_s of 2K pieces + v of the instance, each against one eq table
        "k_dot_eval": (4 * K * t + t + 3) * rows * E,                       # u_s, eta (2K t rows each), u
        "k_spmv": (2 * K + 1) * t * (rows * (E + 8) + 2 * rows * E),
        "k_lincomb": (2 * K + 1) * n * E + 2 * (K * t + 1) * rows * E,       # f_0 over 2K pieces + the two zeta-Horner combinations
        "k_digit_lincomb": 2 * (K * tau + 1) * n * E,
        "k_eq_table": None, "k_eq_combine": 5 * m * E,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "ntt", "plus"])
    ap.add_argument("--log-w", type=int, default=None, dest="log_w")
    ap.add_argument("--cpu-sample-log-w", type=int, default=None, dest="cpu_sample_log_w", help="reference arm: prove a smaller slice instead of the full W")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, dest="cpu_budget_s")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--plus-op", default="rgchk", choices=["rgchk", "cm", "fold"], dest="plus_op", help="--config plus: range check (benches/rgchk.rs), commitment transformation (benches/cm.rs) or a whole PlusProver::prove (benches/e2e.rs)")
    ap.add_argument("--plus-instances", type=int, default=None, dest="plus_instances", help="--config plus: number of RgInstances (the folding arity L of benches/utils/mod.rs)")
    ap.add_argument("--full-step", action="store_true", dest="full_step", help="--config c3: time a whole NIFSProver::prove step instead of commit + linearization")
    args = ap.parse_args()
    if args.log_w is None:
        args.log_w = {"c2": 16, "c3": 20, "ntt": 0, "plus": 17}[args.config]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.config == "ntt":
        from tools import ntt_bench
        return ntt_bench.main(args, rank, world, local)
    if args.config == "plus":
        from tools import plus_bench
        return plus_bench.main(args, rank, world, local)
    if args.config == "c3" and not args.full_step:
        return run_c3(args, rank, world, local)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import hashlib
    import torch
    import torch.distributed as dist
    import latticefold_b200 as lf
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(local)      # pinned staging buffers and the transcript thread next to the GPU's PCIe root
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Weak scaling: N GPUs prove ONE instance of N * 2^log_w constraints, witness columns / hypercube sharded over the
    # ranks with one small all-reduce per commit batch, sumcheck round and evaluation (BASELINE configs[3]).
    import math
    wl = workload(args.log_w + int(math.log2(world)), args.config)
    RING = wl["ring"]
    assert world & (world - 1) == 0, "rank count must be a power of two"
    ctx = lf.Context(RING, local)
    if world > 1:
        ctx.set_shard(rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    prob = synth.bench_instance(wl, rank, world, ops=ctx)
    pr = lf.NIFSProver(ctx, prob)                 # static inputs (Ajtai matrix slice, CCS) go to HBM once, outside the timed region
    A_host = prob["A"] if prob["A"].nbytes < (4 << 30) else None      # the commitment of the instance below reuses the prover's device copy for large matrices
    if A_host is None:
        prob.pop("A")
    W_loc = wl["W"] // world
    f = ctx.witness_f_from_w_ccs(RING, prob["w_ccs"][rank * W_loc:(rank + 1) * W_loc], wl["B"], wl["L"])     # elementwise: local slice
    # pinned host copies of the per-step inputs / outputs for the end-to-end leg
    def pinned_like(a):
        t_ = torch.empty(a.shape, dtype=torch.int64, pin_memory=True)
        arr = t_.numpy().view(np.uint64); arr[...] = a
        return arr, t_
    keep = []
    f_pin, k_ = pinned_like(f); keep.append(k_)
    del f
    prob["w_i_f"], prob["w_acc_f"] = f_pin, f_pin
    if A_host is not None:
        prob["cm_i_cm"] = np.ascontiguousarray(_commit_with_prover(ctx, pr, lf, prob, f_pin))
    else:
        w_tmp = pr.upload_witness(f_pin); prob["cm_i_cm"] = np.ascontiguousarray(pr.witness_commit(w_tmp)); pr.free_witness(w_tmp)
    lc, _ = pr.linearize(prob, lf.Transcript(RING))
    prob["acc"] = synth.split_lcccs(RING, prob, lc)
    ccs = prob["ccs"]
    w_acc, w_i = pr.upload_witness(f_pin), pr.upload_witness(f_pin)
    out_proof, k1 = pinned_like(np.zeros(pr.proof_words, dtype=np.uint64)); out_lc, k2 = pinned_like(np.zeros(pr.lcccs_words, dtype=np.uint64))
    out_f, k3 = pinned_like(np.zeros((pr.n, synth.RINGS[RING]["d"]), dtype=np.uint64)); keep += [k1, k2, k3]
    last = {}

    def step_resident():
        last["resident"] = pr.prove_resident(prob, w_acc, w_i, lf.Transcript(RING))

    def step_e2e():
        return pr.prove(prob, lf.Transcript(RING), out=(out_proof, out_lc, out_f))

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms

    for _ in range(args.warmup):
        step_resident()
    l0, c0 = ctx.launches(), ctx.collectives()
    with ClockSampler(local) as clk:
        ms_res = timed(step_resident, args.steps)
    launches = (ctx.launches() - l0) // max(args.steps, 1)
    collectives = (ctx.collectives() - c0) // max(args.steps, 1)
    phases = pr.timings()
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    constraints = prob["constraints"]
    value = constraints / (ms_res / 1e3)          # one sharded instance: its constraints are the whole job's
    e2e_value = constraints / (ms_e2e / 1e3)

    # ---- checker, after the timed regions: the proof of the last timed step is (i) identical on every rank and
    # to the proof the host-buffer entry point produced, (ii) accepted by the product's own NIFSVerifier and by the oracle's, and
    # (iii) at sizes that have a committed golden digest (tests/golden/bench_digests.json: oracle outputs on the same instance),
    # proof, folded LCCCS and folded witness hash to it.
    verify = None
    if not args.no_verify:
        from tests import helpers
        proof_res, lc_res = last["resident"]
        dg = helpers.step_digests(out_proof, out_lc, out_f)
        same_entry = bool(np.array_equal(proof_res, out_proof) and np.array_equal(lc_res, out_lc))
        ranks_agree = True
        if world > 1:
            h8 = int.from_bytes(hashlib.sha256(np.ascontiguousarray(proof_res).tobytes() + np.ascontiguousarray(lc_res).tobytes()).digest()[:7], "little")
            tmax = torch.tensor([h8], dtype=torch.int64, device="cuda"); tmin = tmax.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            ranks_agree = bool(tmax.item() == tmin.item())
        verify = dict(proof_sha256=dg["proof"], lcccs_sha256=dg["lcccs"], entry_points_agree=same_entry, ranks_agree=ranks_agree)
        if rank == 0:
            try:
                lc_v = lf.nifs_verify(prob, lf.Transcript(RING), proof_res)
                verify["product_verifier"] = "accept" if np.array_equal(lc_v, lc_res) else "accept, but folded instance differs"
            except lf.LfError as e:
                verify["product_verifier"] = f"REJECT: {e}"
            try:
                from oracle.pyoracle import Oracle
                orc_v = Oracle()
                light = {k: v for k, v in prob.items() if k not in ("A", "w_i_f", "w_acc_f")}
                lc_o = orc_v.nifs_verify(light, orc_v.transcript(RING), proof_res)
                verify["oracle_verifier"] = "accept" if np.array_equal(lc_o, lc_res) else "accept, but folded instance differs"
            except Exception as e:      # noqa: BLE001
                verify["oracle_verifier"] = f"REJECT: {e}"
            gold = helpers.bench_golden().get(helpers.bench_case_key(args.config, args.log_w)) if world == 1 and os.path.exists(helpers.BENCH_GOLDEN_PATH) else None
            if gold is not None:
                verify["golden"] = "match" if all(gold[k] == dg[k] for k in ("proof", "lcccs", "witness")) else "MISMATCH"
                verify["witness_sha256"] = dg["witness"]
            else:
                verify["golden"] = "none committed for this size" if world == 1 else "n/a (sharded: verifier acceptance + rank agreement)"
            verify["verified"] = bool(same_entry and ranks_agree and verify["product_verifier"] == "accept" and verify["oracle_verifier"] == "accept"
                                      and verify["golden"] != "MISMATCH")

    # per-kernel device time of one extra step (events around every launch; not part of the timed region above)
    ctx.profile(True)
    step_resident()
    prof = ctx.profile_report()
    ctx.profile(False)
    total_kernel_ms = sum(v[1] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    hbm, peak_src = peaks()
    alg = algorithmic_bytes(wl, ccs, world)
    s, K, kappa = ccs["s"], wl["K"], wl["kappa"]
    n = wl["W"] * wl["L"] // world
    R = synth.RINGS[RING]
    # 64-bit multiply-accumulates per step of the integer-bound sumcheck kernel against the measured
    # issue bound of IMAD.WIDE: one warp instruction per 4 cycles per scheduler = 32 lanes / clk / SM, 4 IMAD.WIDE per lazily reduced 64-bit MAC:
    # 148 SMs x 32 x 1.965 GHz / 4 = 2.33e12 MAC / s.  (The register-only microbenchmark tools/microbench/imad_peak.cu reaches 24.8 of those 32
    # lanes, 1.785e12 MAC / s -- the figure this ratio was quoted against until the kernel itself, with its late rounds no longer latency bound, passed it.)
    MAC_PEAK = 148 * 32 * 1.965e9 / 4
    pairs_r2 = max(((1 << s) // world) // 2 - 1, 0) + max(world - 1, 0)                # sum over rounds >= 2 of the pair count
    macs = {"k_fold_sc_round": pairs_r2 * R["S"] * (2 * K * R["tau"]) * 66} if RING == synth.RING_GOLDILOCKS else {}
    top_name, (top_cnt, top_ms) = top
    a_bytes = alg.get(top_name)
    # DRAM bytes per launch from the committed `ncu --set full` captures (profiles/): see profiles/README.md
    ncu_traffic = NCU_TRAFFIC if (world == 1 and args.config == "c2" and args.log_w == 16) else {}
    kernels = []
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        if ms < 0.5 and name != top_name:
            continue
        b = alg.get(name)
        kernels.append(dict(kernel=name, launches=cnt, total_ms=round(ms, 4), algorithmic_bytes=b,
                            achieved_gbs=(b / 1e9) / (ms / 1e3) if b else None, frac=((b / 1e9) / (ms / 1e3) / hbm) if b else None,
                            bound=KERNEL_BOUND.get(name, "hbm")))
    roofline = dict(bound="hbm", kernel=top_name, launches_per_step=top_cnt, avg_launch_ms=top_ms / top_cnt, share_of_kernel_time=top_ms / total_kernel_ms,
                    achieved=(a_bytes / 1e9) / (top_ms / 1e3) if a_bytes else None, peak=hbm, unit="GB/s",
                    frac=((a_bytes / 1e9) / (top_ms / 1e3) / hbm) if a_bytes else None,
                    traffic=ncu_traffic.get(top_name),
                    traffic_note="dram read+write per launch of this kernel's largest launch in profiles/ (ncu --set full)",
                    peak_source=peak_src, algorithmic_bytes_per_step=a_bytes, kernels=kernels,
                    int_pipe={k: dict(macs_per_step=v, achieved_mac_per_s=v / (prof[k][1] / 1e3), peak_mac_per_s=MAC_PEAK,
                                      frac=v / (prof[k][1] / 1e3) / MAC_PEAK) for k, v in macs.items() if k in prof and prof[k][1] > 0},
                    note="achieved = algorithmic bytes in the reference layout (SURVEY 8d) / CUDA-event time of the kernel's launches in one step; "
                         "kernels[] lists every kernel >= 0.5 ms per step with the same arithmetic and the resource that bounds it")
    line = dict(metric=METRIC, value=value, unit="constraints/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_res, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=DTYPE[args.config], data="synthetic",
                config=config_of(wl, args, "gpu"), clocks=clk.summary(), gpu_launches=int(launches), collectives_per_step=collectives,
                e2e=dict(value=e2e_value, unit="constraints/s", ms_per_step=ms_e2e, h2d_bytes_per_step=int(2 * f_pin.nbytes),
                         d2h_bytes_per_step=int(out_proof.nbytes + out_lc.nbytes + out_f.nbytes)),
                roofline=roofline, phases_ms=phases,
                kernels_ms={k: dict(launches=v[0], total_ms=round(v[1], 4)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})
    if verify is not None:
        line["verified"] = verify.get("verified"); line["proof_sha256"] = verify["proof_sha256"]; line["verify"] = verify
    line["host"] = dict(poseidon=lf.Transcript(RING).backend(), cpus=os.cpu_count())   # dense-layer code path of the host transcript
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import Oracle
        from tools.make_bench_golden import oracle_bench_problem
        orc = Oracle(); cores = os.cpu_count() or 1; orc.set_threads(cores)
        sl = min(args.log_w, {"c2": 12, "c3": 10}[args.config])
        _, sprob = oracle_bench_problem(orc, args.config, sl)
        ms_cpu = cpu_step(orc, sprob, cores)
        line["cpu_baseline"] = dict(value=sprob["constraints"] / (ms_cpu / 1e3), unit="constraints/s", cores=cores, kind="port",
                                    sample=f"one step of the same workload at W=2^{sl} (same ring, DP, kappa, circuit) on {cores} host threads: {ms_cpu:.0f} ms; "
                                           f"the full-size CPU run is `bench.py --impl reference`")
    if rank == 0:
        print(json.dumps(line), flush=True)
    pr.free_witness(w_acc); pr.free_witness(w_i); pr.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_c3(args, rank, world, local):
    """BASELINE configs[2] = SURVEY 8 row C3 as BASELINE.md states it: BabyBear ring (d = 72, packed 4-byte limbs on the device), W = 2^20
    constraints of the reference's degree-three CCS, BabyBearDP (256, 4, 2, 8), kappa = 8, ONE GPU; a step = Witness::commit (A f, the
    witness-sized Ajtai commitment, arith.rs:357-362) + LFLinearizationProver::prove (linearization.rs:145-189: 4 sparse mat-vecs,
    eq table, the degree-4 sumcheck over 5 MLEs of 2^22 entries, evaluations).  `--impl reference`: the CPU restatement on the same
    configuration (bounded step count)."""
    if rank != 0:
        return
    wl = workload(args.log_w, "c3")
    RING = wl["ring"]; R = synth.RINGS[RING]
    if args.impl == "reference":
        from oracle.pyoracle import Oracle
        from tools.make_bench_golden import oracle_c3lin
        orc = Oracle(); cores = os.cpu_count() or 1; orc.set_threads(cores)
        log_w = args.log_w if args.cpu_sample_log_w is None else min(args.log_w, args.cpu_sample_log_w)
        t, t0 = [], time.time()
        while len(t) < max(args.steps, 1) and (not t or (time.time() - t0) * (len(t) + 1) / len(t) < args.cpu_budget_s):
            wl_s, prob, _, _, ms = oracle_c3lin(orc, log_w); t.append(ms)
        ms = float(np.mean(t)); value = prob["constraints"] / (ms / 1e3)
        sample = (f"{len(t)} step(s) of commit + linearization at W=2^{log_w} ({'the same configuration as the product arm' if log_w == args.log_w else 'a slice of the product arm configuration'}; "
                  f"requested steps={args.steps}, warmup={args.warmup}; {args.cpu_budget_s:.0f} s budget)")
        print(json.dumps(dict(metric=METRIC, value=value, unit="constraints/s", n_gpus=1, steps=len(t), warmup=0, ms_per_step=ms, higher_is_better=True, scaling="weak",
                              vs_baseline=None, dtype=DTYPE["c3"], data="synthetic", impl="reference", requested=dict(steps=args.steps, warmup=args.warmup),
                              config=config_of(wl_s, args, "cpu"), cpu_baseline=dict(value=value, unit="constraints/s", cores=cores, kind="port", sample=sample),
                              e2e=dict(value=value, unit="constraints/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)
        return
    import torch
    import latticefold_b200 as lf
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    ctx = lf.Context(RING, local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    t_setup = time.time()
    prob = synth.bench_instance(wl, 0, 1, ops=ctx)
    pr = lf.NIFSProver(ctx, prob)
    prob.pop("A")                                   # 19 GB of host limbs at 2^20: the device copy is all the step needs
    f = ctx.witness_f_from_w_ccs(RING, prob["w_ccs"], wl["B"], wl["L"])
    f_pin_t = torch.empty(f.shape, dtype=torch.int64, pin_memory=True); f_pin = f_pin_t.numpy().view(np.uint64); f_pin[...] = f; del f
    prob["w_i_f"] = f_pin
    w_i = pr.upload_witness(f_pin)
    cm = pr.witness_commit(w_i); prob["cm_i_cm"] = np.ascontiguousarray(cm)
    out_lc, out_pf, out_cm = np.empty(pr.lcccs_words, dtype=np.uint64), np.empty(pr.lin_proof_words(prob), dtype=np.uint64), np.empty_like(cm)
    t_setup = time.time() - t_setup

    def step_resident():
        pr.witness_commit(w_i, out=out_cm)
        prob["cm_i_cm"] = out_cm
        return pr.linearize_resident(prob, w_i, lf.Transcript(RING), out=(out_lc, out_pf))

    def step_e2e():
        w = pr.upload_witness(f_pin)
        pr.witness_commit(w, out=out_cm); prob["cm_i_cm"] = out_cm
        r = pr.linearize_resident(prob, w, lf.Transcript(RING), out=(out_lc, out_pf))
        pr.free_witness(w)
        return r

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream); torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps

    for _ in range(args.warmup):
        step_resident()
    l0 = ctx.launches()
    with ClockSampler(local) as clk:
        ms_res = timed(step_resident, args.steps)
    launches = (ctx.launches() - l0) // max(args.steps, 1)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    constraints = prob["constraints"]
    # checker: the product's LFLinearizationVerifier accepts and reproduces the LCCCS; committed golden digest where one exists;
    # the commitment is linear (A (f + f) = 2 A f on the device: a size-independent property at the full 2^20)
    from tests import helpers
    verify = dict()
    try:
        lc_v = lf.linearization_verify(prob, lf.Transcript(RING), out_pf)
        verify["product_verifier"] = "accept" if np.array_equal(lc_v, out_lc) else "accept, but LCCCS differs"
    except lf.LfError as e:
        verify["product_verifier"] = f"REJECT: {e}"
    try:
        from oracle.pyoracle import Oracle
        orc = Oracle()
        lc_o = orc.linearization_verify({k: v for k, v in prob.items() if k != "w_i_f"}, orc.transcript(RING), out_pf)
        verify["oracle_verifier"] = "accept" if np.array_equal(lc_o, out_lc) else "accept, but LCCCS differs"
    except Exception as e:      # noqa: BLE001
        verify["oracle_verifier"] = f"REJECT: {e}"
    dg = {"cm": helpers.limb_digest(out_cm), "lcccs": helpers.limb_digest(out_lc), "lin_proof": helpers.limb_digest(out_pf)}
    gold = helpers.bench_golden().get(helpers.bench_case_key("c3lin", args.log_w)) if os.path.exists(helpers.BENCH_GOLDEN_PATH) else None
    verify["golden"] = "none committed for this size" if gold is None else ("match" if all(gold[k] == dg[k] for k in dg) else "MISMATCH")
    verify.update(cm_sha256=dg["cm"], proof_sha256=dg["lin_proof"], lcccs_sha256=dg["lcccs"])
    verify["verified"] = bool(verify["product_verifier"] == "accept" and verify["oracle_verifier"] == "accept" and verify["golden"] != "MISMATCH")
    ctx.profile(True); step_resident(); prof = ctx.profile_report(); ctx.profile(False)
    total_kernel_ms = sum(v[1] for v in prof.values())
    hbm, peak_src = peaks()
    ccs = prob["ccs"]; E = R["d"] * 4; n, m, kappa, t = wl["W"] * wl["L"], 1 << ccs["s"], wl["kappa"], ccs["t"]
    rows = wl["W"] + 2
    alg = {"k_dot": (kappa * n + n + kappa) * E, "k_sc_generic": (t + 1) * (m + 3 * m // 2) * E, "k_sc_wide": (t + 1) * 2 * m * E, "k_fold": (t + 1) * 3 * m * E,
           "k_spmv": t * (rows * (E + 8) + 2 * rows * E), "k_eq_combine": 2 * m * E, "k_coeff_eval": 4 * n * E, "k_dot_eval": (t + 1) * rows * E}
    kernels = []
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        b = alg.get(name)
        if ms >= 0.5 or not kernels:
            kernels.append(dict(kernel=name, launches=cnt, total_ms=round(ms, 4), algorithmic_bytes=b, achieved_gbs=(b / 1e9) / (ms / 1e3) if b else None,
                                frac=((b / 1e9) / (ms / 1e3) / hbm) if b else None))
    top = kernels[0]
    # the sumcheck kernel is bound by the FMA-heavy (integer multiplier) pipe, not by HBM (profiles/r02s_k_sc_wide_bb_full.md): report the
    # wide multiply-accumulates it issues against the measured IMAD.WIDE rate (profiles/r01d_imad_peak.md)
    int_pipe = None
    if "k_sc_wide" in prof:
        npts = ccs["d"] + 2; muls_per_point = sum(max(len(S_i) - 1, 0) for S_i in ccs["S"]) + 1      # +-1 coefficients are signs; + the eq factor
        wide = 2 * (m // 2) * R["S"] * npts * muls_per_point * 106                                   # all rounds ~ 2 x round 1; 89 + 17 wide multiplies per Fq9 product
        int_pipe = dict(kernel="k_sc_wide", wide_multiplies_per_step=int(wide), achieved_per_s=wide / (prof["k_sc_wide"][1] / 1e3), peak_per_s=6.97e12,
                        frac=wide / (prof["k_sc_wide"][1] / 1e3) / 6.97e12, fmaheavy_pipe_busy_ncu=0.648)
    line = dict(metric=METRIC, value=constraints / (ms_res / 1e3), unit="constraints/s", n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms_res,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype=DTYPE["c3"], data="synthetic",
                config=dict(config_of(wl, args, "gpu"), step="Witness::commit (A f) + LFLinearizationProver::prove (BASELINE.md C3: commit + linearization sumcheck)",
                            limbs="packed u32 planes: 288 B per ring element on the device (576 B in reference memory)", setup_s=round(t_setup, 1)),
                clocks=clk.summary(), gpu_launches=int(launches),
                e2e=dict(value=constraints / (ms_e2e / 1e3), unit="constraints/s", ms_per_step=ms_e2e, h2d_bytes_per_step=int(f_pin.nbytes),
                         d2h_bytes_per_step=int(out_lc.nbytes + out_pf.nbytes + out_cm.nbytes)),
                roofline=dict(bound="hbm", kernel=top["kernel"], launches_per_step=top["launches"], avg_launch_ms=top["total_ms"] / top["launches"],
                              share_of_kernel_time=top["total_ms"] / total_kernel_ms, achieved=top["achieved_gbs"], peak=hbm, unit="GB/s", frac=top["frac"], traffic=None,
                              peak_source=peak_src, algorithmic_bytes_per_step=top["algorithmic_bytes"], kernels=kernels,
                              note="algorithmic bytes per SURVEY 8(d): 288 B per BabyBear ring element (packed 4-byte limbs, the layout the planes have on the device; "
                                   "the reference's own memory image is 576 B per element)", int_pipe=int_pipe),
                kernels_ms={k: dict(launches=v[0], total_ms=round(v[1], 4)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                verified=verify["verified"], proof_sha256=verify["proof_sha256"], verify=verify,
                host=dict(poseidon=lf.Transcript(RING).backend(), cpus=os.cpu_count()))
    if not args.no_cpu_baseline:
        from tools.make_bench_golden import oracle_c3lin
        orc = Oracle(); cores = os.cpu_count() or 1; orc.set_threads(cores)
        sl = min(args.log_w, 14)
        _, sprob, _, _, ms_cpu = oracle_c3lin(orc, sl)
        line["cpu_baseline"] = dict(value=sprob["constraints"] / (ms_cpu / 1e3), unit="constraints/s", cores=cores, kind="port",
                                    sample=f"commit + linearization of the same workload at W=2^{sl} on {cores} host threads: {ms_cpu:.0f} ms")
    print(json.dumps(line), flush=True)
    pr.free_witness(w_i); pr.close(); ctx.close()


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the round's `ncu --set full` captures (profiles/r02*_full.md)
NCU_TRAFFIC = {"k_fold_sc_round": 2.578e9, "k_dot_commit": 2.086e9}
# what bounds each kernel (ncu evidence in profiles/): "hbm" unless stated
KERNEL_BOUND = {"k_fold_sc_round": "int-pipe (IMAD.WIDE / ALU)", "k_commit_mma": "hbm + L2 (tensor pipe idle-waiting on operand fill)", "k_dot_commit": "int-pipe (IMAD.WIDE)",
                "k_sc_generic": "latency (host-paced rounds)", "k_fold_sc_round1": "int-pipe"}


def bind_to_gpu_numa_node(local):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off (first-touch then places the pinned host buffers there too):
    with 8 ranks on a two-socket host, half of the host<->device copies otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit(); bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else str(bus)).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-"); cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def _commit_with_prover(ctx, pr, lf, prob, f):
    """cm_i.cm = A f using a temporary scheme object (setup only; all-reduced over the ranks when sharded)."""
    sch = lf.AjtaiCommitmentScheme(ctx, prob["A"])
    cm = sch.commit(ctx.upload(f))
    del sch
    return cm


if __name__ == "__main__":
    main()
