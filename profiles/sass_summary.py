#!/usr/bin/env python
"""Static SASS instruction mix per kernel of libbabe_b200.so (run here, no GPU):
    python profiles/sass_summary.py > profiles/r02_sass_summary.md
UBLKCP = 1-D bulk TMA copy (cp.async.bulk), SYNCS = mbarrier operations, LDGSTS = cp.async."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "babe_b200/libbabe_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("babe::", "").replace("void ", "")
        rows[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        rows[kern][m.group(1)] += 1
        rows[kern]["_n"] += 1
cols = ["FFMA2", "FADD2", "FMUL2", "LDGSTS", "UBLKCP", "SYNCS", "BAR", "SHFL", "MUFU"]
fp64 = ("DADD", "DMUL", "DFMA", "DSETP")
print("# SASS summary of libbabe_b200.so (sm_100a), `cuobjdump -sass`, round 2\n")
print("Static instruction counts per kernel.  FFMA2 / FADD2 / FMUL2: Blackwell two-wide fp32 arithmetic; LDGSTS: `cp.async`;")
print("**UBLKCP: 1-D bulk TMA copy (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`), SYNCS: its mbarrier**")
print("(init / arrive.expect_tx / try_wait) -- the frame tiles of the fused NFFT-4096 operators (`k_filter_fused`, `k_stats_fused`,")
print("`k_fir_fused`, csrc/stft_fused.cu) are staged by the TMA engine; the round-1 kernels they replace (`k_apply_filter<Core3,1>`,")
print("`k_stft_stats<Core3,1>`, `k_fir_filter<Core3>`) remain as the path for rows that cannot be bulk-copied; `k_cqt_analysis<true>`")
print("stages each band's window slice of the spectrum the same way.  The prime-factor passes of the CQT (`pfa::k_pfa*`,")
print("csrc/cqt_pfa.cuh) and the band cores (csrc/bandfft_v.cuh) run on the two-wide instructions.  No tensor-core")
print("instructions (north star: not a dense contraction).\n")
print("| kernel | instructions | " + " | ".join(cols) + " | fp64 |")
print("|---|---:|" + "---:|" * (len(cols) + 1))
for k, c in sorted(rows.items(), key=lambda kv: -kv[1]["_n"]):
    print(f"| `{k}` | {c['_n']} | " + " | ".join(str(c[x]) for x in cols) + f" | {sum(c[x] for x in fp64)} |")
