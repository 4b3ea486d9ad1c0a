#!/usr/bin/env python3
"""SASS evidence for profiles/: instruction mix and a dense stretch of the mixed addition inside k_msm_accumulate (every
IMAD.WIDE.U32.X is one fused mad.lo.cc / madc.hi.cc pair), and the asynchronous-copy instructions of k_matvec_stream
(UBLKCP = TMA bulk copy, SYNCS = mbarrier, LDGSTS = cp.async).   usage: python tools/sass_excerpts.py > profiles/r2_sass_excerpts.txt"""
import collections
import re
import subprocess

sass = subprocess.run(["cuobjdump", "-sass", "build/obj/curve_pallas.o"], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        funcs[cur].append(line)


def ins(lines):
    out = []
    for l in lines:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((m.group(1), m.group(2).strip()))
    return out


def opcode(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]


print("# SASS excerpts of the round-2 build (cuobjdump -sass build/obj/curve_pallas.o, sm_100a, nvcc 12.9; tools/sass_excerpts.py)\n")
acc = [k for k in funcs if "k_msm_accumulate" in k][0]
I = ins(funcs[acc])
mix = collections.Counter(opcode(i[1]) for i in I)
print(f"## k_msm_accumulate<CurvePallas>: {len(I)} instructions ({len(I) * 16 // 1024} KB); instruction mix (top 14):")
print("   " + ", ".join(f"{k} {v}" for k, v in mix.most_common(14)))
best, bi = -1, 0
for i in range(0, len(I) - 48):
    c = sum(1 for x in I[i:i + 48] if "IMAD.WIDE.U32.X" in x[1])
    if c > best:
        best, bi = c, i
print(f"\n## inside the mixed addition: 48 consecutive instructions with {best} IMAD.WIDE.U32.X (one fused mad.lo.cc/madc.hi.cc pair each, carry in a predicate):")
for a, t in I[bi:bi + 48]:
    print(f"  /*{a}*/  {t}")
mv = [k for k in funcs if "k_matvec_stream" in k][0]
I = ins(funcs[mv])
print(f"\n## k_matvec_stream<FieldVestaP> ({len(I)} instructions): the asynchronous-copy instructions (TMA bulk copy of the index stream, mbarrier, cp.async z gathers)")
for a, t in I:
    if any(k in t for k in ("UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR")):
        print(f"  /*{a}*/  {t}")
allmix = collections.Counter()
for k, lines in funcs.items():
    for a, t in ins(lines):
        op = opcode(t)
        if any(x in op for x in ("UBLKCP", "LDGSTS", "UTMA", "SYNCS", "UTCMMA", "HMMA")):
            allmix[op] += 1
print("\n## whole translation unit: " + ", ".join(f"{k} {v}" for k, v in sorted(allmix.items())) +
      "   (no tensor-core instruction: nothing on this path is a contraction)")
