"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): the instructions that prove the
tcgen05 / TMEM / TMA / REDUX / atomics claims of DESIGN.md.  Writes profiles/sass_summary.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cytospace_b200", "libcytospace_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "REDUX", "CREDUX", "ATOMG", "ATOM.", "RED.", "ATOMS", "LDG", "LDS", "STS", "STG",
         "MATCH", "BAR.SYNC", "MEMBAR", "IMAD.WIDE", "CCTL", "FENCE", "ERRBAR", "VOTE", "SHFL", "LDL", "STL"]
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = kern.split("(")[0]
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
lines = ["# SASS mnemonic counts per kernel (cuobjdump -sass cytospace_b200/libcytospace_b200.so; sm_100a)",
         "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit,",
         "# SYNCS = mbarrier ops, REDUX = warp reduce, ATOMG/RED = global atomics, LDL/STL = local memory (spills)", ""]
for k, c in counts.items():
    parts = [f"{w}={c[w]}" for w in WATCH if c[w]]
    lines.append(f"{k}\n    instructions={c['_total']}  " + "  ".join(parts))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
