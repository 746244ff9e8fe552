#!/bin/bash
# SASS evidence (no GPU needed): per-kernel instruction histogram of the Blackwell-specific mnemonics in the built library
# usage: tools/sass_histogram.sh > profiles/rNN_sass_histogram.txt
LIB=timeviper_b200/libtimeviper_b200.so
echo "# cuobjdump -sass $LIB  (sm_100a) -- tcgen05 / TMEM / TMA / mbarrier mnemonics per kernel"
cuobjdump -sass "$LIB" | python3 -c '
import sys, re, collections
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTCCP|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|SYNCS|MUFU|F2FP|FMUL2|FFMA2|FADD2|HMUL2|LDGSTS|REDUX|FENCE|BAR)\S*")
name = None
hist = collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); hist[name] = collections.Counter(); continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        hist[name]["(instructions)"] += 1
        m = pat.search(line)
        if m: hist[name][m.group(0).rstrip(";")] += 1
tot = collections.Counter()
for k, c in hist.items():
    if any(x in k for x in ("ssd_fused", "ssd_state_kernel", "conv1d_fwd_kernelI13__nv_bfloat16Li4ELb1", "gated_rmsnormI", "gated_rmsnorm_kernelI13__nv_bfloat16Lb1ELb0ELb0", "ssd_dt_cumsum_bf16x2", "fold_boundary_p2p")):
        print("\n== " + k[:110])
        for m, v in sorted(c.items(), key=lambda kv: -kv[1])[:24]: print(f"   {v:7d}  {m}")
    for m, v in c.items(): tot[m.split(".")[0]] += v
print("\n== whole library, by base mnemonic")
for m, v in sorted(tot.items(), key=lambda kv: -kv[1]): print(f"   {v:7d}  {m}")
'
