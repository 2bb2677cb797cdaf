"""CPU-side checks of the product library: it loads, exports every symbol include/lf_b200.h declares, its host-side
transcript / RotSum reproduce the reference KATs, and compute entry points fail loudly without a GPU (no fallback)."""
import json
import os
import re

import numpy as np
import pytest

import latticefold_b200 as lf
from latticefold_b200 import api, plus, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
G = synth.RING_GOLDILOCKS


def test_library_exports_every_declared_symbol():
    L = lf.lib()
    hdr = open(os.path.join(ROOT, "include", "lf_b200.h")).read()
    declared = set(re.findall(r"\b(lf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lf_status"}
    assert declared == set(api.SYMBOLS) | set(plus.SYMBOLS), declared ^ (set(api.SYMBOLS) | set(plus.SYMBOLS))
    for s in declared:
        assert hasattr(L, s), s


def test_product_transcript_kats():
    # crates/latticefold/src/transcript/poseidon.rs:86-142, through the product's own Poseidon
    g = json.load(open(os.path.join(GOLD, "transcript_goldilocks.json")))
    t = lf.Transcript(G); t.absorb_base(np.array(g["absorbed"], dtype=np.uint64))
    assert list(map(int, t.get_challenge())) == g["big_challenge"]
    t = lf.Transcript(G); t.absorb_base(np.array(g["absorbed"], dtype=np.uint64))
    assert list(map(int, t.get_short_challenge())) == g["small_challenge_coeffs"]


def test_product_transcript_matches_oracle_on_long_schedule(oracle):
    # same absorb / challenge schedule on both implementations, crossing many rate boundaries
    a, b = lf.Transcript(G), oracle.transcript(G)
    els = synth.uniform_field(synth.RINGS[G]["p"], 37 * 24, 5).reshape(37, 24)
    for step in range(6):
        a.absorb(els[step * 5:(step + 1) * 5 + step]); b.absorb(els[step * 5:(step + 1) * 5 + step])
        a.absorb_tag("beta_s"); b.absorb_tag("beta_s")
        for _ in range(step + 1):
            assert np.array_equal(a.get_challenge(), b.get_challenge())
        assert np.array_equal(a.get_short_challenge(), b.get_short_challenge())


@pytest.mark.parametrize("ring", [synth.RING_BABYBEAR, synth.RING_FROG])
def test_product_transcript_other_rings_match_oracle(oracle, ring):
    """the BabyBear sponge (one-word remainders, split-multiplier dot products) and the Frog sponge (Montgomery form, branch-free reductions)
    against the oracle's canonical arithmetic: random and edge lanes, many rate boundaries"""
    R = synth.RINGS[ring]; p, d = R["p"], R["d"]
    a, b = lf.Transcript(ring), oracle.transcript(ring)
    els = synth.uniform_field(p, 23 * d, 11).reshape(23, d)
    edge = np.array([0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 1 << 16, (1 << 16) - 1, min(p - 1, (1 << 31) - 1)], dtype=np.uint64)
    for step in range(5):
        a.absorb(els[step * 4:(step + 1) * 4 + step]); b.absorb(els[step * 4:(step + 1) * 4 + step])
        v = np.resize(np.roll(edge, step), 20 * 2 + step + 1); a.absorb_base(v); b.absorb_base(v)
        for _ in range(step + 2):
            assert np.array_equal(a.get_challenge(), b.get_challenge())
        assert np.array_equal(a.get_short_challenge(), b.get_short_challenge())
    v = np.full(20 * 4 + 3, p - 1, dtype=np.uint64); a.absorb_base(v); b.absorb_base(v)
    assert np.array_equal(a.get_challenge(), b.get_challenge())


def test_product_transcript_extreme_lanes_match_oracle(oracle):
    """lanes at the edges of the field (0, 1, 2^32 +- 1, p - 2^32, p - 1): the product's Poseidon keeps lazily reduced
    representatives inside a round, the oracle reduces canonically after every operation"""
    p = synth.RINGS[G]["p"]
    edge = np.array([0, 1, 2, (1 << 32) - 1, 1 << 32, (1 << 32) + 1, p - (1 << 32), p - (1 << 32) - 1, p - 2, p - 1], dtype=np.uint64)
    a, b = lf.Transcript(G), oracle.transcript(G)
    for rep in range(4):
        v = np.resize(np.roll(edge, rep), 20 * 3 + rep)
        a.absorb_base(v); b.absorb_base(v)
        assert np.array_equal(a.get_challenge(), b.get_challenge())
    v = np.full(20 * 5, p - 1, dtype=np.uint64)
    a.absorb_base(v); b.absorb_base(v)
    assert np.array_equal(a.get_challenge(), b.get_challenge())
    assert np.array_equal(a.get_short_challenge(), b.get_short_challenge())


_DIGEST_SNIPPET = """
import hashlib, numpy as np, latticefold_b200 as lf
from latticefold_b200 import synth
G = synth.RING_GOLDILOCKS; p = synth.RINGS[G]["p"]
t = lf.Transcript(G); h = hashlib.sha256()
edge = np.array([0, 1, (1 << 32) - 1, 1 << 32, p - (1 << 32), p - 2, p - 1], dtype=np.uint64)
for step in range(12):
    t.absorb_base(synth.uniform_field(p, 20 * 7 + step, 100 + step)); t.absorb_base(np.roll(edge, step)); t.absorb_tag("rho_s")
    for _ in range(3):
        h.update(t.get_challenge().tobytes())
    h.update(t.get_short_challenge().tobytes())
print(t.backend(), t.permutations(), h.hexdigest())
"""


def test_product_poseidon_backends_agree():
    """the AVX-512 IFMA dense layer (csrc/poseidon_ifma.cpp, picked at run time) and the scalar one give the same transcript; on a
    host without IFMA both runs are the scalar path and the test only checks that the override is harmless"""
    import subprocess, sys
    outs = {}
    for force_scalar in ("0", "1"):
        env = dict(os.environ, LF_POSEIDON_SCALAR=force_scalar, PYTHONPATH=ROOT)
        outs[force_scalar] = subprocess.run([sys.executable, "-c", _DIGEST_SNIPPET], env=env, capture_output=True, text=True, check=True, cwd=ROOT).stdout.split()
    assert outs["1"][0] == "scalar" and outs["0"][0] in ("scalar", "avx512-ifma")
    assert outs["0"][1:] == outs["1"][1:]
    assert lf.Transcript(G).backend() in ("scalar", "avx512-ifma")


def test_poseidon_ifma_dense_layer_edges(tmp_path):
    """csrc/poseidon_ifma.cpp against 128-bit `%` arithmetic on 2.8 million lanes, random and at the edges of every limb / carry
    decision (tests/native/poseidon_ifma_edge.cpp); skipped where the host has no AVX-512 IFMA or no g++"""
    import shutil, subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "ifma_edge")
    subprocess.run([gxx, "-O2", "-march=x86-64-v3", "-std=c++17", os.path.join(ROOT, "tests", "native", "poseidon_ifma_edge.cpp"),
                    os.path.join(ROOT, "latticefold_b200", "csrc", "poseidon_ifma.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    if out[:1] == ["no-ifma"]:
        pytest.skip("host without AVX-512 IFMA")
    assert out[-4] == "lanes" and int(out[-3]) > 2_000_000 and out[-2:] == ["bad", "0"], out


def test_babybear_balanced_slot_field_arithmetic(tmp_path):
    """csrc/field.cuh BbBal -- the balanced-representative Fq9 arithmetic of the wide-slot-field sumcheck kernels -- against 128-bit
    arithmetic in Fq[Y]/(Y^9 - nu): reductions over the whole signed 64-bit range, products with every operand at +-(p-1)/2 (the
    accumulator bound), fixed-operand products with an addend, and agreement with the canonical SlotField multiplication
    (tests/native/bb_balanced.cpp; the same code runs on the device, where tests/test_gpu_parity.py holds it to the oracle)"""
    import shutil, subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "bb_balanced")
    subprocess.run([gxx, "-O2", "-std=c++17", os.path.join(ROOT, "tests", "native", "bb_balanced.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-2000:]


def test_product_rot_lin_combination_kat():
    # crates/cyclotomic-rings/src/rotation.rs:174-776
    g = json.load(open(os.path.join(GOLD, "rotsum_goldilocks.json")))
    rho = np.array(g["rho"], dtype=np.uint64).reshape(3, 24); theta = np.array(g["theta"], dtype=np.uint64).reshape(3, 3, 24)
    out = np.empty((3, 24), dtype=np.uint64)
    rc = lf.lib().lf_rot_lin_combination(G, api.ptr(rho), api.ptr(theta), 3, api.ptr(out))
    assert rc == 0 and np.array_equal(out, np.array(g["expected"], dtype=np.uint64).reshape(3, 24))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lf.LfError) as e:
        lf.Context(G, 0)
    assert e.value.code == -20   # LF_ERR_CUDA


def test_unsupported_ring_is_reported():
    h = api.vp()
    rc = lf.lib().lf_transcript_create(3, api.C.byref(h))       # the Stark ring (256-bit field) is not implemented
    assert rc == -8


@pytest.mark.parametrize("ring", [synth.RING_GOLDILOCKS, synth.RING_BABYBEAR, synth.RING_FROG])
def test_product_host_side_matches_oracle_all_rings(oracle, ring):
    """transcript (incl. the ring's short-challenge set) and RotSum of the product's host code vs the oracle"""
    R = synth.RINGS[ring]; d, tau = R["d"], R["tau"]
    a, b = lf.Transcript(ring), oracle.transcript(ring)
    els = synth.uniform_field(R["p"], 9 * d, 5).reshape(9, d)
    for step in range(3):
        a.absorb(els[step * 3:(step + 1) * 3]); b.absorb(els[step * 3:(step + 1) * 3])
        a.absorb_tag("rho_s"); b.absorb_tag("rho_s")
        assert np.array_equal(a.get_challenge(), b.get_challenge())
        assert np.array_equal(a.get_short_challenge(), b.get_short_challenge())
    rho = synth.uniform_field(R["p"], 2 * d, 6).reshape(2, d); theta = synth.uniform_field(R["p"], 2 * tau * d, 7).reshape(2, tau, d)
    out = np.empty((tau, d), dtype=np.uint64)
    assert lf.lib().lf_rot_lin_combination(ring, api.ptr(rho), api.ptr(theta), 2, api.ptr(out)) == 0
    assert np.array_equal(out, oracle.rot_lin_combination(ring, rho, theta))


VERIFY_CASES = [  # ring, W, B, L, b, K, kappa, kind, CCS degree
    (synth.RING_GOLDILOCKS, 4, 1 << 15, 5, 2, 15, 4, "non_scalar", 2),
    (synth.RING_GOLDILOCKS, 8, 1 << 16, 4, 2, 16, 3, "uniform", 3),
    (synth.RING_BABYBEAR, 4, 1 << 8, 4, 2, 8, 4, "non_scalar", 3),
    (synth.RING_FROG, 4, 1 << 8, 8, 2, 10, 4, "uniform", 2),
]


@pytest.mark.parametrize("ring,W,B,L,b,K,kappa,kind,degree", VERIFY_CASES)
def test_product_verifier_accepts_oracle_proofs_and_rejects_tampering(oracle, oracle_ops, ring, W, B, L, b, K, kappa, kind, degree):
    """NIFSVerifier::verify of the product library (host code, csrc/verifier_host.hpp; nifs.rs:117-162) against the oracle's
    prover: same folded instance, and each sub-proof region is protected (nifs/tests.rs:58-117 and the per-protocol tamper tests)."""
    prob = synth.make_instance(ring, W, B, L, b, K, kappa, kind=kind, config_id=21, ops=oracle_ops, degree=degree)
    proof, lc, _, _ = oracle.nifs_prove(prob, oracle.transcript(ring))
    assert np.array_equal(lf.nifs_verify(prob, lf.Transcript(ring), proof), lc)
    assert np.array_equal(lc, oracle.nifs_verify(prob, oracle.transcript(ring), proof))
    p = synth.RINGS[ring]["p"]
    codes = set()
    for pos in (0, proof.size // 5, proof.size // 3, proof.size // 2, (2 * proof.size) // 3, proof.size - 1):
        bad = proof.copy(); bad[pos] = (int(bad[pos]) + 1) % p
        with pytest.raises(lf.LfError) as e:
            lf.nifs_verify(prob, lf.Transcript(ring), bad)
        codes.add(e.value.code)
    assert codes <= {-10, -11} and -10 in codes      # LF_ERR_SUMCHECK_FAILED / LF_ERR_RECOMPOSED
    # a wrong accumulator is rejected as well
    prob2 = dict(prob); acc = dict(prob["acc"]); acc["v"] = acc["v"].copy(); acc["v"][0, 0] = (int(acc["v"][0, 0]) + 1) % p; prob2["acc"] = acc
    with pytest.raises(lf.LfError):
        lf.nifs_verify(prob2, lf.Transcript(ring), proof)
    # non-canonical encodings (limb + k p, congruent to the honest value) are rejected like arkworks' deserialisation does:
    # accepting them would make proofs malleable
    for pos in (0, 7, proof.size // 4, proof.size // 2, proof.size - 1):
        for k in (1, 2, 1024):
            if int(proof[pos]) + k * p >= 1 << 64:
                continue
            bad = proof.copy(); bad[pos] = int(bad[pos]) + k * p
            with pytest.raises(lf.LfError) as e:
                lf.nifs_verify(prob, lf.Transcript(ring), bad)
            assert e.value.code == -21, (pos, k)        # LF_ERR_INVALID_ARG
    if int(prob["acc"]["u"][0, 0]) + p < 1 << 64:
        prob3 = dict(prob); acc = dict(prob["acc"]); acc["u"] = acc["u"].copy(); acc["u"][0, 0] = int(acc["u"][0, 0]) + p; prob3["acc"] = acc
        with pytest.raises(lf.LfError) as e:
            lf.nifs_verify(prob3, lf.Transcript(ring), proof)
        assert e.value.code == -21


@pytest.mark.parametrize("ring", [synth.RING_GOLDILOCKS, synth.RING_BABYBEAR, synth.RING_FROG])
def test_ring_describe_and_slot_product_agree_with_oracle(oracle, ring):
    """lf_ring_describe (host, no GPU) reports the ring shape the generator assumes, and the slot-field product built from its nu
    (synth.sf_mul, used for the degree-three CCS's diag(z^2)) equals the oracle's NTT-form product"""
    class Info(api.C.Structure):
        _fields_ = [("p", api.C.c_uint64), ("d", api.C.c_int32), ("n_slots", api.C.c_int32), ("tau", api.C.c_int32), ("nu", api.C.c_uint64)]
    info = Info(); assert lf.lib().lf_ring_describe(ring, api.C.byref(info)) == 0
    R = synth.RINGS[ring]
    assert (info.p, info.d, info.n_slots, info.tau) == (R["p"], R["d"], R["S"], R["tau"]) and info.nu == oracle.info(ring)["nu"]
    a = synth.uniform_field(R["p"], 5 * R["d"], 31).reshape(5, R["d"]); b = synth.uniform_field(R["p"], 5 * R["d"], 32).reshape(5, R["d"])
    assert np.array_equal(synth.sf_mul(ring, a, b, int(info.nu)), oracle.ntt_mul(ring, a, b))


@pytest.mark.parametrize("ring,W,B,L,b,K,kappa,kind,degree", VERIFY_CASES[:3])
def test_proof_wire_format_roundtrip_and_validation(oracle, oracle_ops, ring, W, B, L, b, K, kappa, kind, degree):
    """LFProof::serialize_with_mode(Compress::Yes) restated (nifs.rs:28-34, examples/e2e.rs:126-146; csrc/wire_host.hpp): size formula,
    round trip, field order (the v of the linearization proof sits right behind its sumcheck), and the deserialiser's validation"""
    prob = synth.make_instance(ring, W, B, L, b, K, kappa, kind=kind, config_id=23, ops=oracle_ops, degree=degree)
    proof, lc, _, _ = oracle.nifs_prove(prob, oracle.transcript(ring))
    R = synth.RINGS[ring]; d, tau, fb = R["d"], R["tau"], (8 if R["p"] >> 32 else 4)
    ccs = prob["ccs"]; s, t, l = ccs["s"], ccs["t"], ccs["l"]; rb = d * fb
    vec = lambda n: 8 + n * rb
    want = (8 + s * vec(ccs["d"] + 2) + vec(tau) + vec(t)) + 2 * (4 * 8 + K * (vec(t) + vec(tau) + vec(l + 1) + vec(kappa))) + (8 + s * vec(2 * b + 1) + 2 * 8 + 2 * K * (vec(tau) + vec(t)))
    data = lf.proof_to_bytes(prob, proof)
    assert len(data) == want
    assert np.array_equal(lf.proof_from_bytes(prob, data), proof)
    assert int.from_bytes(data[:8], "little") == s and int.from_bytes(data[8:16], "little") == ccs["d"] + 2
    first = np.frombuffer(data[16:16 + rb], dtype="<u8" if fb == 8 else "<u4").astype(np.uint64)
    assert np.array_equal(first, proof[:d])                                       # first evaluation of the first round message
    off_v = 8 + s * vec(ccs["d"] + 2)
    assert int.from_bytes(data[off_v:off_v + 8], "little") == tau
    # the verifier accepts what comes back from the wire
    assert np.array_equal(lf.nifs_verify(prob, lf.Transcript(ring), lf.proof_from_bytes(prob, data)), lc)
    bad = bytearray(data); bad[0] ^= 1                                            # wrong length prefix
    with pytest.raises(lf.LfError) as e:
        lf.proof_from_bytes(prob, bytes(bad))
    assert e.value.code == -5
    bad = bytearray(data); bad[16:16 + fb] = (R["p"]).to_bytes(fb, "little")      # non-canonical field element (= p)
    with pytest.raises(lf.LfError) as e:
        lf.proof_from_bytes(prob, bytes(bad))
    assert e.value.code == -21
    with pytest.raises(lf.LfError):
        lf.proof_from_bytes(prob, data[:-1])
    with pytest.raises(lf.LfError):
        lf.proof_from_bytes(prob, data + b"\\0")
