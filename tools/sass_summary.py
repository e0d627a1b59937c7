#!/usr/bin/env python
"""SASS evidence for profiles/: instruction counts per kernel of the shipped library (no GPU needed).
    python tools/sass_summary.py r02      -> profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
so = os.path.join(ROOT, "dependence_free_rl_b200", "libdfrl_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
pat = collections.OrderedDict([
    ("UTCHMMA (tcgen05.mma)", r"\bUTCHMMA\b"), ("of which A operand from TMEM", r"UTCHMMA tmem\["), ("LDTM (tcgen05.ld)", r"\bLDTM\b"),
    ("STTM (tcgen05.st)", r"\bSTTM\b"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR\b"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP\b"),
    ("UTMALDG / UTMASTG (tensor-map TMA)", r"\bUTMA(LDG|STG)\b"), ("HMMA (mma.sync)", r"\bHMMA\b"), ("HGMMA (wgmma)", r"\bHGMMA\b")])
tot = collections.Counter()
out = [f"# SASS summary of `dependence_free_rl_b200/libdfrl_b200.so` ({tag})\n\n`python tools/sass_summary.py {tag}` = `cuobjdump -sass` "
       "instruction counts per kernel (names demangled by `c++filt`).\n\n", "| kernel | " + " | ".join(pat) + " |\n|---|" + "---|" * len(pat) + "\n"]
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", re.sub(r"\(anonymous namespace\)::", "", dem)).replace("void ", "")
    c = {k: len(re.findall(v, f)) for k, v in pat.items()}
    tot.update(c)
    if c["UTCHMMA (tcgen05.mma)"] or c["UBLKCP (cp.async.bulk)"]:
        out.append("| `" + dem + "` | " + " | ".join(str(c[k]) for k in pat) + " |\n")
out.append("| **whole library** | " + " | ".join(str(tot[k]) for k in pat) + " |\n")
out.append("\nArchitectures in the fat binary: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", txt)))) +
           ". Operands reach shared memory through `cp.async.bulk` (UBLKCP: the flat parameter vector, the pre-split B chunks of the "
           "layered GEMMs) and through the epilogues' own swizzled stores; no tensor-map TMA (`UTMALDG`) is used: the fused kernels' "
           "operands are PRODUCED on chip (activations) or converted fp32 -> 16-bit pairs on the way in, not copied from global tensors.\n")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md"), "w").write("".join(out))
print("".join(out)[-600:])
