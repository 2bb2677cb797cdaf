"""Host Poseidon (width 24, the product's Fiat-Shamir transcript) permutation time through the C ABI: min / median over 300 batches
of 200 permutations per ring.  No GPU needed.  python tools/poseidon_bench.py [path/to/liblf_b200.so]"""
import ctypes, os, sys, time
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
L = ctypes.CDLL(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "latticefold_b200", "_lib", "liblf_b200.so"))
L.lf_transcript_create.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]
L.lf_transcript_absorb_base.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
L.lf_transcript_permutations.restype = ctypes.c_uint64
L.lf_transcript_permutations.argtypes = [ctypes.c_void_p]
L.lf_host_poseidon_backend.restype = ctypes.c_char_p
print("goldilocks dense layers:", L.lf_host_poseidon_backend().decode())
for ring, name in ((0, "goldilocks"), (1, "babybear"), (2, "frog")):
    t = ctypes.c_void_p()
    if L.lf_transcript_create(ring, ctypes.byref(t)):
        continue
    a = np.arange(1, 20 * 200 + 1, dtype=np.uint64)
    L.lf_transcript_absorb_base(t, a.ctypes.data, 20 * 100)
    ts = []
    for _ in range(300):
        p0 = L.lf_transcript_permutations(t)
        t0 = time.perf_counter(); L.lf_transcript_absorb_base(t, a.ctypes.data, a.size); dt = time.perf_counter() - t0
        ts.append(dt / (L.lf_transcript_permutations(t) - p0) * 1e6)
    ts.sort()
    print("poseidon %s: min %.2f med %.2f us/permutation" % (name, ts[0], ts[len(ts) // 2]))
