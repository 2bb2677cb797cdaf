#!/usr/bin/env python3
"""profiles/r02_sass_excerpts.md: what the PTX of the Blackwell-specific kernels became in SASS (cuobjdump -sass of the built library;
runs without a GPU).  B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "latticefold_b200", "_lib", "liblf_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)[1:]
INS = re.compile(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);")
out = ["# r02 — SASS excerpts of the built library (`cuobjdump -sass latticefold_b200/_lib/liblf_b200.so`, sm_100a; `tools/sass_excerpts.py`)\n\n",
       "What the PTX became (B200_PROFILING.md: the PTX names never appear in SASS).  Offsets are instruction addresses inside each kernel.\n"]


def dump(fun_pat, title, pats, limit):
    for b in blocks:
        name = b.split("\n", 1)[0].strip()
        if not re.search(fun_pat, name):
            continue
        lines = [m for m in (INS.search(l) for l in b.split("\n") if re.search(pats, l)) if m]
        out.append(f"\n## {title}\n\n`{name}` — {len(lines)} matching instructions\n\n```\n")
        out.extend(f"/*{m.group(1)}*/  {m.group(2).strip()} ;\n" for m in lines[:limit])
        out.append("```\n")
        return


dump(r"k_commit_mmaINS_14GoldilocksRing", "tensor-core digit commit: tcgen05.mma.kind::i8 (UTCIMMA), TMEM allocation / loads (UTCATOMSWS, LDTM), bulk copies (UBLKCP), "
     "mbarriers (SYNCS), tcgen05.commit (UTCBAR)", r"UTCIMMA|UBLKCP|LDTM|UTCBAR|UTCATOMSWS|SYNCS|ELECT", 60)
dump(r"k_ntt_ctaIN2lf3ntt3GlFELi12ELb0ELb0", "negacyclic NTT (N = 2^12, Goldilocks, forward): bulk load of the polynomials (UBLKCP) completing on an mbarrier", r"UBLKCP|SYNCS", 20)
dump(r"k_ntt_ctaIN2lf3ntt3BbFELi12ELb0ELb0", "negacyclic NTT (N = 2^12, BabyBear, forward)", r"UBLKCP|SYNCS", 20)
dump(r"k_sc_wide_bbILi160ELi5", "wide-slot-field sumcheck round (BabyBear Fq9, csrc/sumcheck_wide.cuh): table tiles by bulk copies (UBLKCP) through a two-stage mbarrier ring (SYNCS)",
     r"UBLKCP|SYNCS|ELECT", 20)
out.append("\n## instruction mix of the integer kernels (counts of SASS mnemonics in the kernel body)\n\n| kernel | IMAD.WIDE | other IMAD | IADD3 / IADD3.X | SHFL | LDG | LDS | total |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
for pat, label in ((r"15k_fold_sc_roundINS_14GoldilocksRing", "k_fold_sc_round<Goldilocks>"), (r"k_fold_sc_round2INS_14GoldilocksRing", "k_fold_sc_round2<Goldilocks>"),
                   (r"k_fold_sc_round1INS_14GoldilocksRing", "k_fold_sc_round1<Goldilocks>"), (r"k_sc_pointsINS_14GoldilocksRingELi4", "k_sc_points<Goldilocks, 4>"),
                   (r"k_sc_pointsINS_12BabyBearRingELi5", "k_sc_points<BabyBear, 5>"),
                   (r"k_sc_wide_bbILi160ELi5", "k_sc_wide_bb<160, 5> (BabyBear, balanced Fq9)"), (r"k_fold_wide_bb", "k_fold_wide_bb (BabyBear)"), (r"5k_dotINS_14GoldilocksRingELi2ELi256", "k_dot<Goldilocks, 2>"),
                   (r"5k_dotINS_12BabyBearRingELi1ELi256", "k_dot<BabyBear, 1>"), (r"k_commit_mmaINS_14GoldilocksRing", "k_commit_mma<Goldilocks>")):
    for b in blocks:
        if re.search(pat, b.split("\n", 1)[0]):
            ins = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b)]
            c = lambda p: sum(1 for i in ins if re.match(p, i))
            out.append(f"| {label} | {c(r'IMAD\.WIDE')} | {c(r'IMAD(?!\.WIDE)')} | {c(r'IADD3')} | {c(r'SHFL')} | {c(r'LDG')} | {c(r'LDS')} | {len(ins)} |\n")
            break
open(os.path.join(ROOT, "profiles", "r02_sass_excerpts.md"), "w").write("".join(out))
print("".join(out)[-1500:])
