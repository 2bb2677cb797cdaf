"""bench.py --config ntt: BASELINE configs[4] -- batched negacyclic NTT / INTT throughput sweep, degree 2^10 .. 2^16,
batch 2^8 .. 2^20 ring elements (capped so that one buffer stays below 4 GiB), Goldilocks (u64) and BabyBear (u32), on 1 .. 8 GPUs.

Transforms are independent, so N GPUs split the batch evenly with no data-path collective (weak scaling: every rank runs the full
per-GPU batch; the aggregate is the sum).  Algorithmic bytes per transform = 2 * N * batch * word (read once + write once); time = CUDA
events on the library's stream, max over ranks; the headline `value` is the aggregate forward GB/s at N = 2^16 with the largest batch.
Buffers below the 126 MB L2 are flagged (`fits_l2`): their rate is an L2 number, not an HBM one.  Correctness inside the run: every
timed configuration round-trips (INTT(NTT(a)) == a) and rank 0 checks one small batch per size against the CPU oracle's transform."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = {0: 0xFFFFFFFF00000001, 1: 2013265921}
NAME = {0: "goldilocks", 1: "babybear"}


def main(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import latticefold_b200 as lf
    from latticefold_b200 import synth
    if args.impl == "reference":
        return reference(args, rank)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --config ntt: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ctx = lf.Context(synth.RING_GOLDILOCKS, local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    iters, warm = max(args.steps, 1), max(args.warmup, 3)
    orc = None
    if rank == 0 and not args.no_verify:
        from oracle.pyoracle import Oracle
        orc = Oracle()
    rows, ok = [], True
    for field in (0, 1):
        word = 8 if field == 0 else 4; tdt = torch.int64 if field == 0 else torch.int32
        for lg in range(10, 17):
            n = 1 << lg
            plan = lf.NttPlan(ctx, field, lg)
            if orc is not None:      # one small batch against the oracle's transform (definition in include/lf_b200.h)
                a = synth.uniform_field(P[field], 3 * n, 1000 + lg).reshape(3, n)
                got = plan.forward(a.astype(np.uint64 if field == 0 else np.uint32))
                ok = ok and bool(np.array_equal(got.astype(np.uint64), orc.ntt(field, lg, a)))
            for lb in range(8, 21, 2):
                batch = 1 << lb
                if n * batch * word > (4 << 30):
                    continue
                g = torch.Generator(device="cuda"); g.manual_seed(lg * 64 + lb + 7919 * rank)
                a = torch.randint(0, P[field] if field == 1 else (1 << 62), (batch, n), dtype=tdt, device="cuda", generator=g)
                f = torch.empty_like(a); b = torch.empty_like(a)
                torch.cuda.synchronize()
                for _ in range(warm):
                    plan.forward_device(a.data_ptr(), f.data_ptr(), batch); plan.inverse_device(f.data_ptr(), b.data_ptr(), batch)
                ctx.sync()
                ok = ok and bool(torch.equal(a, b))
                res = {}
                for name, fn, src, dst in (("fwd", plan.forward_device, a, f), ("inv", plan.inverse_device, f, b)):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize(); e0.record(stream)
                    for _ in range(iters):
                        fn(src.data_ptr(), dst.data_ptr(), batch)
                    e1.record(stream); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / iters
                    if world > 1:
                        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
                    gbs = world * 2 * n * batch * word / 1e9 / (ms / 1e3)
                    res[name] = dict(ms=round(ms, 4), GBps=round(gbs, 1), frac_of_hbm=round(gbs / (world * peak), 4))
                rows.append(dict(field=NAME[field], log_n=lg, log_batch_per_gpu=lb, fits_l2=bool(n * batch * word < (126 << 20)), fwd=res["fwd"], inv=res["inv"]))
                del a, f, b
            plan.close()
    if world > 1:
        t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); ok = bool(t.item())
    if rank == 0:
        head = max((r for r in rows if r["field"] == "goldilocks" and r["log_n"] == 16), key=lambda r: r["log_batch_per_gpu"])
        best = {}
        for r in rows:
            k = (r["field"], r["log_n"])
            if not r["fits_l2"] and (k not in best or r["fwd"]["GBps"] > best[k]["fwd"]["GBps"]):
                best[k] = r
        line = dict(metric="negacyclic NTT throughput (forward, N = 2^16, Goldilocks)", value=head["fwd"]["GBps"], unit="GB/s", n_gpus=world, steps=iters, warmup=warm,
                    ms_per_step=head["fwd"]["ms"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64 / u32", data="synthetic",
                    config=dict(workload="BASELINE.json configs[4]: batched negacyclic NTT/INTT over Z_p[X]/(X^N+1), N = 2^10..2^16, batch 2^8..2^20 per GPU (buffers <= 4 GiB), "
                                         "Goldilocks u64 and BabyBear u32; transforms split evenly over the GPUs, no collective", l2="rows with fits_l2 = true are L2-resident; all others stream from HBM"),
                    roofline=dict(bound="hbm", kernel="k_ntt_cta", achieved=head["fwd"]["GBps"], peak=world * peak, unit="GB/s", frac=head["fwd"]["frac_of_hbm"], traffic=None,
                                  peak_source=peak_src, note="algorithmic bytes = 2 * N * batch * word; the transform is ALU-pipe bound on B200 (profiles/r01c_k_ntt*_f*.md)"),
                    verified=ok, best_per_size=[dict(field=k[0], log_n=k[1], log_batch_per_gpu=v["log_batch_per_gpu"], fwd_GBps=v["fwd"]["GBps"], inv_GBps=v["inv"]["GBps"],
                                                     fwd_frac=v["fwd"]["frac_of_hbm"]) for k, v in sorted(best.items())],
                    rows=rows)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def reference(args, rank):
    """CPU arm: the oracle's textbook radix-2 transform (OpenMP over the batch) at N = 2^16, bounded batch"""
    if rank != 0:
        return
    import time
    from oracle.pyoracle import Oracle
    from latticefold_b200 import synth
    orc = Oracle(); cores = os.cpu_count() or 1; orc.set_threads(cores)
    n, batch = 1 << 16, 256
    a = synth.uniform_field(P[0], n * batch, 5).reshape(batch, n)
    orc.ntt(0, 16, a)
    t0 = time.time(); k = 0
    while k < max(args.steps, 1) and time.time() - t0 < 60:
        orc.ntt(0, 16, a); k += 1
    ms = (time.time() - t0) * 1e3 / k
    gbs = 2 * n * batch * 8 / 1e9 / (ms / 1e3)
    print(json.dumps(dict(metric="negacyclic NTT throughput (forward, N = 2^16, Goldilocks)", value=gbs, unit="GB/s", n_gpus=args.gpus, steps=k, warmup=1, ms_per_step=ms,
                          higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64", data="synthetic", impl="reference",
                          config=dict(workload="BASELINE.json configs[4], CPU: radix-2 negacyclic NTT, N = 2^16, batch 256"),
                          cpu_baseline=dict(value=gbs, unit="GB/s", cores=cores, kind="port", sample=f"{k} transforms of a 256 x 2^16 batch"),
                          e2e=dict(value=gbs, unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)
