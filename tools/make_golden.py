#!/usr/bin/env python3
"""Extract the reference's own known-answer tests for the hot path into tests/golden/*.json.

Only NUMBERS are taken (the literals inside the reference's #[test] functions); nothing executable.
Sources ("Golden vectors / KATs"):
  rotsum_goldilocks.json        crates/cyclotomic-rings/src/rotation.rs:174-776  (test_rot_lin_combination)
  transcript_goldilocks.json    crates/latticefold/src/transcript/poseidon.rs:86-142
  challenge_sets.json           crates/cyclotomic-rings/src/rings/{goldilocks,babybear,frog}.rs test_small_challenge_from_random_bytes
Closed-form KATs (test_commit_ntt, test_get_fhat, arith/utils small vectors) need no data file; the tests
restate their formulae with the file:line.
Run in the build container only (the reference tree is absent on the GPU box).
"""
import json, re, pathlib
REF = pathlib.Path("/root/reference/crates")
OUT = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"

def test_body(src, name):
    i = src.index("fn " + name)
    # crude brace matching from the first '{' after the fn name
    j = src.index("{", i); depth = 0
    for k in range(j, len(src)):
        if src[k] == "{": depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0: return src[j:k + 1]
    raise ValueError(name)

def rotsum():
    src = (REF / "cyclotomic-rings/src/rotation.rs").read_text()
    body = test_body(src, "test_rot_lin_combination")
    i_rho, i_theta, i_exp = body.index("let rho_s"), body.index("let theta_s"), body.index("let expected")
    lit = re.compile(r"Fq::from\(\s*(\d+)u64\s*\)")
    order = sorted([(i_rho, "rho"), (i_theta, "theta"), (i_exp, "expected")])
    segs = {}
    for n, (pos, name) in enumerate(order):
        end = order[n + 1][0] if n + 1 < len(order) else len(body)
        segs[name] = [int(x) for x in lit.findall(body[pos:end])]
    assert len(segs["rho"]) == 72 and len(segs["theta"]) == 216 and len(segs["expected"]) == 72, {k: len(v) for k, v in segs.items()}
    return {"source": "crates/cyclotomic-rings/src/rotation.rs:174-776",
            "layout": {"rho": "3 polys x 24 coeffs", "theta": "3 vectors x 3 ring elems x 8 slots x 3 limbs",
                       "expected": "3 ring elems x 8 slots x 3 limbs"}, **segs}

def transcript():
    src = (REF / "latticefold/src/transcript/poseidon.rs").read_text()
    lit = re.compile(r"BigInt\(\[(\d+)\]\)")
    big = [int(x) for x in lit.findall(test_body(src, "test_get_big_challenge"))]
    small = [int(x) for x in lit.findall(test_body(src, "test_get_small_challenge"))]
    assert len(big) == 3 and len(small) == 24
    return {"source": "crates/latticefold/src/transcript/poseidon.rs:86-142", "absorbed": [255],
            "big_challenge": big, "small_challenge_coeffs": small}

def challenge_sets():
    out = {}
    for ring in ("goldilocks", "babybear", "frog"):
        src = (REF / f"cyclotomic-rings/src/rings/{ring}.rs").read_text()
        body = test_body(src, "test_small_challenge_from_random_bytes")
        bs = [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})\b", body[:body.index("res_coeffs")])]
        coeffs = [int(x) for x in re.findall(r"BigInt\(\[(\d+)\]\)", body)]
        out[ring] = {"source": f"crates/cyclotomic-rings/src/rings/{ring}.rs", "bytes": bs, "coeffs": coeffs}
    return out

def main():
    OUT.mkdir(parents=True, exist_ok=True)
    for name, fn in (("rotsum_goldilocks", rotsum), ("transcript_goldilocks", transcript), ("challenge_sets", challenge_sets)):
        (OUT / f"{name}.json").write_text(json.dumps(fn(), indent=0))
        print("wrote", name)

if __name__ == "__main__":
    main()
