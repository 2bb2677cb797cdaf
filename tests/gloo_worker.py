"""CPU worker (gloo, no GPU): the host-side half of the N>1 path -- column-sharded Ajtai commit and hypercube-sharded MLE
evaluation / sumcheck round messages computed per rank with the oracle, combined with the split-limb all-reduce, must
equal the unsharded result.  Also checks shard_instance's slicing."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    from latticefold_b200 import parallel, synth
    from oracle.pyoracle import Oracle
    from tests.helpers import OracleOps, rand_elems, rand_sf_broadcast
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    G = synth.RING_GOLDILOCKS; p = synth.RINGS[G]["p"]
    orc = Oracle(); orc.set_threads(1)
    # 1. commit: rank g holds A[:, cols_g] and f[cols_g]
    kappa, n = 3, 64
    A = rand_elems(G, kappa * n, 1).reshape(kappa, n, 24); f = rand_elems(G, n, 2)
    lo, hi = rank * n // world, (rank + 1) * n // world
    part = orc.commit(G, np.ascontiguousarray(A[:, lo:hi]), np.ascontiguousarray(f[lo:hi]))
    assert np.array_equal(parallel.allreduce_field(part, p), orc.commit(G, A, f)), "sharded commit"
    # 2. MLE evaluation: slab of the table times slab of eq(., r)
    nv = 6; r = rand_sf_broadcast(G, nv, 3); tab = rand_elems(G, 1 << nv, 4)
    eq = orc.eq_table(G, r); m = 1 << nv; lo, hi = rank * m // world, (rank + 1) * m // world
    acc = np.zeros(24, dtype=object)
    for x in range(lo, hi):
        acc = (acc + orc.ntt_mul(G, eq[x:x + 1], tab[x:x + 1])[0].astype(object)) % p
    full = orc.evaluate_mles(G, tab[None], nv, r)[0]
    assert np.array_equal(parallel.allreduce_field(acc.astype(np.uint64), p), full), "sharded evaluation"
    # 3. worst case for the limb split: every rank contributes p - 1
    worst = np.full(5, p - 1, dtype=np.uint64)
    assert np.array_equal(parallel.allreduce_field(worst, p), np.full(5, (world * (p - 1)) % p, dtype=np.uint64))
    # 4. slicing of a full instance
    prob = synth.make_instance(G, 8, 1 << 16, 4, 2, 16, 2, kind="uniform", config_id=5, ops=OracleOps(orc))
    mine = parallel.shard_instance(prob, rank, world)
    nl = prob["n"] // world
    assert mine["A"].shape == (2, nl, 24) and np.array_equal(mine["w_i_f"], prob["w_i_f"][rank * nl:(rank + 1) * nl])
    parts = orc.commit(G, mine["A"], mine["w_i_f"])
    assert np.array_equal(parallel.allreduce_field(parts, p), prob["cm_i_cm"]), "sharded cm_i"
    dist.barrier(); dist.destroy_process_group()
    print(f"GLOO_OK rank {rank}", flush=True)


if __name__ == "__main__":
    main()
