#!/usr/bin/env python3
"""bench.py -- LatticeFold prover step throughput (R1CS constraints / second) on B200.

  python bench.py --gpus N --steps K --warmup W            product arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the C++ restatement of the reference algorithm
                                                           (oracle/, kind "port": the reference itself is Rust with an
                                                           un-vendored dependency and cannot be built in this image)

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

RING = synth.RING_GOLDILOCKS
E_BYTES = 24 * 8     # bytes per ring element (reference layout: 24 u64 limbs)


def workload(log_w):
    """configs[1] of BASELINE.json = SURVEY 8 row C2: W = 2^16, DP (B, L, b, K) = (65536, 4, 2, 16), kappa = 26
    (crates/latticefold/benches/config.toml goldilocks row n=32768 extrapolated, as SURVEY.md 8 does)."""
    return dict(W=1 << log_w, B=1 << 16, L=4, b=2, K=16, kappa=26, kind="non_scalar")


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
    _, _, _, ms = orc.nifs_prove(prob, orc.transcript(RING), want_f=True)
    return ms


def run_reference(args, rank, world):
    """CPU arm: C++ restatement of the reference algorithm on the host cores, on a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle.pyoracle import Oracle
    from tests.helpers import OracleOps
    orc = Oracle()
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    import math
    wl = workload(args.log_w + int(math.log2(max(args.gpus, 1))))
    sample_log_w = min(args.log_w, args.cpu_sample_log_w)
    swl = dict(wl, W=1 << sample_log_w)
    prob = synth.make_instance(RING, swl["W"], swl["B"], swl["L"], swl["b"], swl["K"], swl["kappa"], kind=swl["kind"], config_id=2, ops=OracleOps(orc))
    for _ in range(args.warmup):
        cpu_step(orc, prob, cores)
    t = [cpu_step(orc, prob, cores) for _ in range(args.steps)]
    ms = float(np.mean(t))
    value = prob["constraints"] / (ms / 1e3)
    sample = f"W=2^{sample_log_w} slice of the W={wl['W']} workload (same ring, DP, kappa={wl['kappa']}); constraints/s = (W+2)/step time"
    line = dict(metric="prover constraints/sec (commit+decomp+sumcheck)", value=value, unit="constraints/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64 (mod 2^64-2^32+1)", data="synthetic", impl="reference",
                config=config_of(wl, args, "cpu"), cpu_baseline=dict(value=value, unit="constraints/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="constraints/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def config_of(wl, args, where):
    return dict(workload=f"Goldilocks ring X^24-X^12+1, dummy R1CS ({wl['kind']} witness) with {wl['W']}+2 constraints, one NIFSProver::prove step "
                         f"(BASELINE.json configs[1]; weak-scaled to W = gpus * 2^{args.log_w} when gpus > 1, as configs[3])", W=wl["W"], B=wl["B"], L=wl["L"], b=wl["b"], K=wl["K"], kappa=wl["kappa"], n=wl["W"] * wl["L"],
                parallelism=("cpu threads" if where == "cpu" else ("1 GPU" if args.gpus == 1 else
                             f"{args.gpus} GPUs: witness columns / hypercube sharded, one all-reduce per commit batch, sumcheck round and evaluation")),
                l2="inputs larger than L2: Ajtai matrix 1.3 GB + 2K witness pieces 1.6 GB per step vs 126 MB L2")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-w", type=int, default=16, dest="log_w")
    ap.add_argument("--cpu-sample-log-w", type=int, default=11, dest="cpu_sample_log_w")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import latticefold_b200 as lf
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Weak scaling: N GPUs prove ONE instance of N * 2^log_w constraints, witness columns / hypercube sharded over the
    # ranks with one small all-reduce per commit batch, sumcheck round and evaluation (BASELINE configs[3]).
    import math
    wl = workload(args.log_w + int(math.log2(world)))
    assert world & (world - 1) == 0, "rank count must be a power of two"
    ctx = lf.Context(RING, local)
    if world > 1:
        ctx.set_shard(rank, world)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    prob = make_sharded_instance(wl, rank, world)
    pr = lf.NIFSProver(ctx, prob)                 # static inputs (Ajtai matrix slice, CCS) go to HBM once, outside the timed region
    W_loc = wl["W"] // world
    f = ctx.witness_f_from_w_ccs(RING, prob["w_ccs"][rank * W_loc:(rank + 1) * W_loc], wl["B"], wl["L"])     # elementwise: local slice
    # pinned host copies of the per-step inputs / outputs for the end-to-end leg
    def pinned_like(a):
        t_ = torch.empty(a.shape, dtype=torch.int64, pin_memory=True)
        arr = t_.numpy().view(np.uint64); arr[...] = a
        return arr, t_
    keep = []
    f_pin, k_ = pinned_like(f); keep.append(k_)
    prob["w_i_f"], prob["w_acc_f"] = f_pin, f_pin
    prob["cm_i_cm"] = np.ascontiguousarray(_commit_with_prover(ctx, pr, lf, prob, f_pin))
    lc, _ = pr.linearize(prob, lf.Transcript(RING))
    prob["acc"] = synth.split_lcccs(RING, prob, lc)
    ccs = prob["ccs"]
    w_acc, w_i = pr.upload_witness(f_pin), pr.upload_witness(f_pin)
    out_proof, k1 = pinned_like(np.zeros(pr.proof_words, dtype=np.uint64)); out_lc, k2 = pinned_like(np.zeros(pr.lcccs_words, dtype=np.uint64))
    out_f, k3 = pinned_like(np.zeros((pr.n, 24), dtype=np.uint64)); keep += [k1, k2, k3]

    def step_resident():
        return pr.prove_resident(prob, w_acc, w_i, lf.Transcript(RING))

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

    # per-kernel device time of one extra step (events around every launch; not part of the timed region above)
    ctx.profile(True)
    step_resident()
    prof = ctx.profile_report()
    ctx.profile(False)
    total_kernel_ms = sum(v[1] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][1])
    hbm, peak_src = peaks()
    s, t = ccs["s"], ccs["t"]
    n, K, kappa = wl["W"] * wl["L"] // world, wl["K"], wl["kappa"]          # per rank
    M_fold = 5 + 2 * K * 3
    alg = {
        # SURVEY 8(d) per-unit figures in the reference layout (E = 192 B), summed over that kernel's launches in one step
        "k_dot_commit": 2 * (kappa * n + (K - 1) * n + (K - 1) * kappa) * E_BYTES,
        "k_fold_sc_round": M_fold * (((1 << s) // world) - 2 + 2 * max(world - 1, 0)) * E_BYTES,              # rounds 2..s read M tables of length 2^(s-r+1)
        "k_fold_sc_round1": M_fold * ((1 << s) // world) * E_BYTES,
        "k_fold": None, "k_matrix_apply": 2 * (2 * K + 3) * n * E_BYTES,
    }
    # 64-bit multiply-accumulates per step of the two integer-bound kernels against the measured
    # IMAD.WIDE-bound ceiling of tools/microbench/imad_peak.cu on B200: 1.785e12 lazily reduced MACs / s
    MAC_PEAK = 1.785e12
    S_slots, tau = 8, 3
    pairs_r2 = max(((1 << s) // world) // 2 - 1, 0) + max(world - 1, 0)                # sum over rounds >= 2 of the pair count
    macs = {"k_dot_commit": 2 * kappa * (K - 1) * n * S_slots * 9,
            "k_fold_sc_round": pairs_r2 * S_slots * (2 * K * tau) * 66}       # two lanes x (mu*t 9 + t^2 6 + two MACs 18)
    top_name, (top_cnt, top_ms) = top
    a_bytes = alg.get(top_name)
    # DRAM bytes of this kernel from the committed `ncu --set full` capture (profiles/): largest launch (round 2 at C2) moved
    # 2.545 GB read + 5 MB written for 2.54 GB algorithmic; the batched commit 2.10 GB for 2.06 GB algorithmic
    ncu_traffic = {"k_fold_sc_round": 2.578e9, "k_dot_commit": 2.086e9}      # profiles/r01d_*: 2.544 GB + 34 MB; 2.073 GB + 13.5 MB
    roofline = dict(bound="hbm", kernel=top_name, launches_per_step=top_cnt, avg_launch_ms=top_ms / top_cnt, share_of_kernel_time=top_ms / total_kernel_ms,
                    achieved=(a_bytes / 1e9) / (top_ms / 1e3) if a_bytes else None, peak=hbm, unit="GB/s",
                    frac=((a_bytes / 1e9) / (top_ms / 1e3) / hbm) if a_bytes else None,
                    traffic=ncu_traffic.get(top_name) if (world == 1 and args.log_w == 16) else None,
                    traffic_note="dram read+write of the largest launch of this kernel in profiles/ (ncu --set full); its algorithmic bytes are the same to 1%",
                    peak_source=peak_src, algorithmic_bytes_per_step=a_bytes,
                    int_pipe={k: dict(macs_per_step=v, achieved_mac_per_s=v / (prof[k][1] / 1e3), peak_mac_per_s=MAC_PEAK,
                                      frac=v / (prof[k][1] / 1e3) / MAC_PEAK) for k, v in macs.items() if k in prof and prof[k][1] > 0},
                    note="this kernel is bound by the IMAD.WIDE issue rate (64-bit modular multiply-accumulates on 32-bit pipes), not by HBM: "
                         "frac is the HBM fraction the contract asks for, int_pipe.frac the fraction of the measured multiply-accumulate ceiling")
    line = dict(metric="prover constraints/sec (commit+decomp+sumcheck)", value=value, unit="constraints/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_res, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64 (mod 2^64-2^32+1)", data="synthetic",
                config=config_of(wl, args, "gpu"), clocks=clk.summary(), gpu_launches=int(launches), collectives_per_step=collectives,
                e2e=dict(value=e2e_value, unit="constraints/s", ms_per_step=ms_e2e, h2d_bytes_per_step=int(2 * f_pin.nbytes),
                         d2h_bytes_per_step=int(out_proof.nbytes + out_lc.nbytes + out_f.nbytes)),
                roofline=roofline, phases_ms=phases,
                kernels_ms={k: dict(launches=v[0], total_ms=round(v[1], 4)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})
    line["host"] = dict(poseidon=lf.Transcript(RING).backend(), cpus=os.cpu_count())   # dense-layer code path of the host transcript
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import Oracle
        from tests.helpers import OracleOps
        orc = Oracle(); cores = os.cpu_count() or 1
        sl = min(args.log_w, args.cpu_sample_log_w)
        sprob = synth.make_instance(RING, 1 << sl, wl["B"], wl["L"], wl["b"], wl["K"], wl["kappa"], kind=wl["kind"], config_id=2, ops=OracleOps(orc))
        ms_cpu = cpu_step(orc, sprob, cores)
        line["cpu_baseline"] = dict(value=sprob["constraints"] / (ms_cpu / 1e3), unit="constraints/s", cores=cores, kind="port",
                                    sample=f"one step at W=2^{sl} (same ring, DP, kappa) on {cores} host threads: {ms_cpu:.0f} ms")
    if rank == 0:
        print(json.dumps(line), flush=True)
    pr.free_witness(w_acc); pr.free_witness(w_i); pr.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def make_sharded_instance(wl, rank, world):
    """Synthetic inputs of one step, holding only this rank's column slice of the Ajtai matrix (kappa x n/world independent
    uniform ring elements, SplitMix64 seeded per rank) -- the full 8-GPU matrix would be 10.5 GB per process."""
    R = synth.RINGS[RING]
    seed = (synth.SEED_BASE + 100) & synth.MASK
    n = wl["W"] * wl["L"]
    w_ccs = synth.make_witness(RING, wl["W"], wl["kind"], seed)
    ccs = synth.make_ccs(RING, wl["W"], wl["L"], wl["kind"], w_ccs, 1)
    A = synth.uniform_field(R["p"], wl["kappa"] * (n // world) * R["d"], seed + 7919 * (rank + 1)).reshape(wl["kappa"], n // world, R["d"])
    return dict(ring=RING, B=wl["B"], L=wl["L"], b=wl["b"], K=wl["K"], kappa=wl["kappa"], n=n, W=wl["W"], A=A, ccs=ccs, w_ccs=w_ccs,
                cm_i_x_ccs=synth.one(RING, 1), constraints=1 + wl["W"] + 1, kind=wl["kind"])


def _commit_with_prover(ctx, pr, lf, prob, f):
    """cm_i.cm = A f using a temporary scheme object (setup only; all-reduced over the ranks when sharded)."""
    sch = lf.AjtaiCommitmentScheme(ctx, prob["A"])
    cm = sch.commit(ctx.upload(f))
    del sch
    return cm


if __name__ == "__main__":
    main()
