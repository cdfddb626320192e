#!/usr/bin/env python3
"""Tensor-core / TMEM / TMA opcodes per kernel of lib/libtome_b200.so (cuobjdump -sass), the listing committed under
profiles/ (python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "video-how-do-your-tokens-merge_b200", "lib", "libtome_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
want = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOM", "SYNCS", "ELECT", "BRA.U.ANY", "FENCE.VIEW.ASYNC")
print("# cuobjdump -sass lib/libtome_b200.so: tensor-core / TMEM / TMA opcodes per kernel (sm_100a)")
print("# UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,")
print("# SYNCS = mbarrier, UTCATOM = tcgen05.alloc / dealloc; BRA.U.ANY = the vote loop around a uniform-datapath issue from divergent code")
print("# (DESIGN.md, 'Issuing tcgen05.mma': absent where the issuing warp runs converged)\n")
idx = 0
for block in sass.split("Function : ")[1:]:
    name = names[idx] if idx < len(names) else block.split("\n")[0]
    idx += 1
    c = collections.Counter()
    for line in block.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for w in want:
                if op == w or op.startswith(w + ".") or (w == "BRA.U.ANY" and op == "BRA.U.ANY"):
                    c[w] += 1
    if any(c[w] for w in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG")):
        print(name[:150])
        print("    " + "  ".join(f"{w}={c[w]}" for w in want if c[w]))
