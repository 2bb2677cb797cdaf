"""Worker for the sharded-prover tests: run under torch.distributed.run with WORLD_SIZE ranks.  Every rank proves its
shard of the same instance; the proof must be byte-identical to the CPU oracle's (and hence to the 1-GPU proof).
Ranks share cuda:0 through gloo when the box has fewer GPUs than ranks, and use NCCL otherwise."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import latticefold_b200 as lf
    from latticefold_b200 import parallel, synth
    from oracle.pyoracle import Oracle
    from tests.helpers import OracleOps
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    ndev = torch.cuda.device_count()
    backend = "nccl" if ndev >= world else "gloo"
    dev = local % ndev
    torch.cuda.set_device(dev)
    dist.init_process_group(backend)
    G = synth.RING_GOLDILOCKS
    orc = Oracle(); orc.set_threads(2)
    cases = [(64, 1 << 16, 4, 2, 16, 5, "non_scalar"), (256, 1 << 16, 4, 2, 16, 3, "uniform"), (16, 1 << 16, 4, 2, 16, 4, "scalar")]
    ctx = lf.Context(G, dev)
    ctx.set_shard(rank, world)
    for (W, B, L, b, K, kappa, kind) in cases:
        if W % world or (W * L) // world < 2:
            continue
        prob = synth.make_instance(G, W, B, L, b, K, kappa, kind=kind, config_id=3, ops=OracleOps(orc))
        eproof, elc, ef, _ = orc.nifs_prove(prob, orc.transcript(G))
        mine = parallel.shard_instance(prob, rank, world)
        pr = lf.NIFSProver(ctx, mine)
        # sharded commit of the witness reproduces cm_i.cm on every rank
        lc_lin, pf_lin = pr.linearize(mine, lf.Transcript(G))
        elc_lin, epf_lin = orc.linearize(prob, orc.transcript(G))
        assert np.array_equal(pf_lin, epf_lin) and np.array_equal(lc_lin, elc_lin), "sharded linearization differs"
        proof, lc, f = pr.prove(mine, lf.Transcript(G))
        n_loc = prob["n"] // world
        assert np.array_equal(proof, eproof), f"rank {rank}: sharded proof differs from the oracle (W={W})"
        assert np.array_equal(lc, elc), "folded LCCCS differs"
        assert np.array_equal(f, ef[rank * n_loc:(rank + 1) * n_loc]), "folded witness slice differs"
        pr.close()
        print(f"rank {rank}/{world} [{backend}] W={W} {kind}: sharded step bit-exact, {ctx.collectives()} collectives so far, p2p={getattr(ctx, "p2p", False)}", flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"SHARDED_OK rank {rank}", flush=True)


if __name__ == "__main__":
    main()
