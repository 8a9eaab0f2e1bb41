"""Per-kernel SASS evidence of the built library: counts of the mnemonics that prove which
hardware paths the kernels use (UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, DMMA = fp64 tensor
core, LDGSTS = cp.async, REDG/ATOMG = global atomics, MEMBAR, BAR).  Writes profiles/sass_summary.txt.
usage: python tools/sass_summary.py   (needs cuobjdump and c++filt; no GPU)"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "krylov_b200", "csrc", "libkrylov_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UBLKCP", "SYNCS", "DMMA", "UTMALDG", "UTCMMA", "LDGSTS", "ATOMG", "REDG", "MEMBAR", "BAR.SYNC", "DADD", "DMUL", "DFMA"]
per = collections.OrderedDict()
cur = None
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
fam = collections.OrderedDict()
for mangled, name in zip(per, names):
    base = re.sub(r"<.*", "", name.replace("void ", "")).split("(")[0]
    f = fam.setdefault(base, {"n": 0, "c": collections.Counter()})
    f["n"] += 1
    f["c"].update(per[mangled])
out = [f"# SASS summary of krylov_b200/csrc/libkrylov_b200.so (arch {', '.join(sorted(arch))}); regenerate with tools/sass_summary.py",
       "# per kernel family (all template instantiations summed): instantiations, SASS instructions, and the counts of",
       "# " + " ".join(KEYS), ""]
out.append(f"{'kernel':42s} {'inst':>4s} {'sass':>8s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for base, f in fam.items():
    out.append(f"{base:42s} {f['n']:4d} {f['c']['_total']:8d} " + " ".join(f"{f['c'][k]:8d}" for k in KEYS))
    tot.update(f["c"])
out.append(f"{'TOTAL':42s} {sum(f['n'] for f in fam.values()):4d} {tot['_total']:8d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
path = os.path.join(root, "profiles", "sass_summary.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
print("wrote", path)
