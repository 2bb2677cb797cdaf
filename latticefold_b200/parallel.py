"""Multi-GPU plumbing: the collective the C library calls back into, on torch.distributed (NCCL over NVLink between the
GPUs of one box; gloo for CPU-side tests), and the slicing of a prover-step input set across ranks.

The library shards the Ajtai column axis and the hypercube's high bits; each batched dot product / sumcheck round /
MLE evaluation ends in ONE all-reduce of a few ring elements.  NCCL has no "sum mod p", so field elements travel as two
32-bit halves in u64 lanes summed with SUM (world * 2^32 cannot wrap) and are folded mod p afterwards -- on the device in
the library (k_split_limbs / k_combine_limbs), and in `allreduce_field` below for host-side arrays.
"""
import numpy as np

from . import api, synth


class _DevPtr:
    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def make_collective(ctx, group=None):
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", ctx.device)
    backend = dist.get_backend(group)

    def cb(user, op, ptr, words):
        try:
            with torch.cuda.device(dev):
                world = dist.get_world_size(group)
                t = torch.as_tensor(_DevPtr(ptr, words * (world if op == 1 else 1)), device=dev)
                if backend == "nccl":
                    if op == 0:
                        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                    else:
                        dist.all_gather_into_tensor(t, t[dist.get_rank(group) * words:(dist.get_rank(group) + 1) * words].clone(), group=group)
                    torch.cuda.current_stream(dev).synchronize()
                else:       # gloo: stage through host memory
                    h = t.cpu()
                    if op == 0:
                        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
                    else:
                        r = dist.get_rank(group)
                        parts = [torch.empty(words, dtype=torch.int64) for _ in range(world)]
                        dist.all_gather(parts, h[r * words:(r + 1) * words].contiguous(), group=group)
                        h = torch.cat(parts)
                    t.copy_(h)
                    torch.cuda.current_stream(dev).synchronize()
            return 0
        except Exception as e:      # never unwind through C
            import sys
            print(f"[latticefold_b200] collective failed: {e!r}", file=sys.stderr, flush=True)
            return 1

    return api.COLLECTIVE_FN(cb)


def allreduce_field(values, p, group=None):
    """Sum mod p over ranks of a numpy uint64 array of canonical field elements (host-side twin of the device path)."""
    import torch
    import torch.distributed as dist
    v = np.ascontiguousarray(values, dtype=np.uint64)
    halves = np.stack([v & np.uint64(0xFFFFFFFF), v >> np.uint64(32)], axis=-1).astype(np.int64)
    t = torch.from_numpy(halves)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    h = t.numpy().astype(object)
    return ((h[..., 0] + (h[..., 1] << 32)) % p).astype(np.uint64)


def shard_instance(prob, rank, world):
    """This rank's view of a prover-step input set: its column slice of the Ajtai matrix and its element slice of both
    witnesses.  Everything small (CCS, accumulator LCCCS, commitments, statements) is replicated."""
    n = prob["n"]
    assert n % world == 0
    lo, hi = rank * (n // world), (rank + 1) * (n // world)
    out = dict(prob)
    out["A"] = np.ascontiguousarray(prob["A"][:, lo:hi])
    for k in ("w_i_f", "w_acc_f"):
        if prob.get(k) is not None:
            out[k] = np.ascontiguousarray(prob[k][lo:hi])
    return out
