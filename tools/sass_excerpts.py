#!/usr/bin/env python
"""SASS evidence per kernel family: `cuobjdump -sass` of the built library, reduced to instruction counts, the counts of the
Blackwell-specific mnemonics and a few excerpt lines. Runs in the build container (no GPU needed).
Usage: python tools/sass_excerpts.py > profiles/r2_sass_excerpts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spectrograms_b200", "lib", "libsgx_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "FADD2", "FMUL2", "FFMA2", "DFMA", "DADD", "DMUL", "MUFU",
        "REDUX", "ELECT"]
SHOW = ["n400_tm", "n400_tc", "dct2_lifter_tc", "fused_pow2<float, 1024", "fused_pow2<double, 2048", "fused_n400<", "fused_mixed<float, 400",
        "istft_pow2<float, 256", "k_r2c_fused_generic<float"]


def demangle(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"sgx::\(anonymous namespace\)::", "", d)


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = f.split("\n", 1)
        funcs[demangle(name.strip())] = body
    print("# SASS evidence per kernel family (cuobjdump -sass spectrograms_b200/lib/libsgx_b200.so, sm_100a; tools/sass_excerpts.py)")
    print("# LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk")
    print("# (TMA 1-D), SYNCS = mbarrier ops, LDGSTS = cp.async, FADD2 / FMUL2 / FFMA2 = packed FP32x2, DFMA / DADD / DMUL = FP64.\n")
    for d in sorted(funcs):
        if not any(k in d for k in SHOW):
            continue
        ops = collections.Counter()
        for line in funcs[d].split("\n"):
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                ops[m.group(1)] += 1
        print(d[:150])
        print(f"    {sum(ops.values())} instructions; " + ", ".join(f"{k} {ops[k]}" for k in KEYS if ops[k]) + "\n")

    def excerpt(title, sub, pattern, maxn):
        print(title)
        for d, body in funcs.items():
            if sub in d:
                for line in [ln for ln in body.split("\n") if re.search(pattern, ln)][:maxn]:
                    print("    " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", line).strip())
                break

    print("# excerpts")
    excerpt("## k_r2c_fused_n400_tm<4, 1>: mbarrier arm + bulk copies of the signal tile", "k_r2c_fused_n400_tm<4, 1>", r"UBLKCP|SYNCS.ARRIVE", 5)
    excerpt("## k_r2c_fused_n400_tm<4, 1>: Y exchange through tensor memory", "k_r2c_fused_n400_tm<4, 1>", r"STTM|LDTM", 6)
    excerpt("## k_dct2_lifter_tc: MMA issue and commit", "k_dct2_lifter_tc", r"UTCHMMA|UTCBAR", 5)
    excerpt("## k_dct2_lifter_tc: A operand into / coefficients out of tensor memory", "k_dct2_lifter_tc", r"STTM|LDTM", 4)
    excerpt("## k_r2c_fused_n400_tc: MMA issue", "k_r2c_fused_n400_tc", r"UTCHMMA|UTCBAR", 4)
    excerpt("## k_r2c_fused_pow2<float, 1024, 4, false>: the measured-and-rejected bulk-staged path (run-time opt-in)", "k_r2c_fused_pow2<float, 1024, 4, false>",
            r"UBLKCP", 2)


if __name__ == "__main__":
    main()
