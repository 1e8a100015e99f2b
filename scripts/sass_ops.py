"""Counts the Blackwell-specific SASS mnemonics per kernel of the shipped library -> profiles/sass_ops.txt
(UTCHMMA / UTCIMMA = tcgen05.mma f16 / i8, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTCCP = tcgen05.cp,
UBLKCP = cp.async.bulk, UTMALDG = tensor TMA, SYNCS = mbarrier, ATOMS/REDG = shared / global reductions)."""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "cellula_b200", "libcellula_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "ATOMS", "ATOMG", "REDG", "RED.", "DFMA", "DADD", "DMUL"]
per = collections.OrderedDict(); cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        per[cur] = collections.Counter(); continue
    if cur is None: continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        per[cur]["_total"] += 1
        for o in ops:
            if m.group(1).startswith(o): per[cur][o] += 1
out = os.path.join(root, "profiles", "sass_ops.txt")
with open(out, "w") as f:
    f.write("cuobjdump -sass cellula_b200/libcellula_b200.so (sm_100a), instruction counts per kernel (static)\n\n")
    f.write(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{o:>7s}" for o in ops) + "\n")
    tot = collections.Counter()
    for k, c in per.items():
        f.write(f"{k[:70]:70s} {c['_total']:7d} " + " ".join(f"{c[o]:7d}" for o in ops) + "\n")
        tot.update(c)
    f.write(f"{'TOTAL':70s} {tot['_total']:7d} " + " ".join(f"{tot[o]:7d}" for o in ops) + "\n")
print(open(out).read()[:6000])
