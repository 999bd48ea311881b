"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass`: per-opcode samples, per-phase stall
reasons (phases split at the first/last DMMA of the kernel) and the hottest instructions."""
import collections
import csv
import sys

path = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
section = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
secs, cur, hdr, names = [], None, None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        if cur is not None:
            secs.append(cur)
        cur = []
        names.append(r[1])
    elif r and r[0] == "Address":
        hdr = r
    elif cur is not None and len(r) > 10:
        cur.append(r)
secs.append(cur)
idx = {}
for i, n in enumerate(hdr):
    idx.setdefault(n, i)
print("sections:", [(n[:40], len(s)) for n, s in zip(names, secs)])
sass = secs[section]


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


S, E = idx["# Samples"], idx["Instructions Executed"]
tot = sum(I(r[S]) for r in sass)
print("instructions", len(sass), "samples", tot)
stall = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]


def opcode(src):
    t = src.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else "?"


byop, byop_e = collections.Counter(), collections.Counter()
for r in sass:
    byop[opcode(r[1])] += I(r[S])
    byop_e[opcode(r[1])] += I(r[E])
for op, n in byop.most_common(18):
    print("  %-10s samples %8d (%5.1f%%)  warp-instr executed %12d" % (op, n, 100.0 * n / tot, byop_e[op]))
dm = [i for i, r in enumerate(sass) if "DMMA" in r[1]]
if dm:
    lo, hi = dm[0] - 40, dm[-1] + 8
    for name, (x, y) in (("before first DMMA (setup + phase A)", (0, lo)), ("DMMA loop (phase B)", (lo, hi)),
                         ("after last DMMA (epilogue)", (hi, len(sass)))):
        n = sum(I(r[S]) for r in sass[x:y])
        st = collections.Counter()
        for r in sass[x:y]:
            for k in stall:
                st[k] += I(r[idx[k]])
        print("%-38s instr %5d samples %8d (%5.1f%%) %s" % (name, y - x, n, 100.0 * n / tot, st.most_common(5)))
for i, r in sorted(enumerate(sass), key=lambda x: -I(x[1][S]))[:ntop]:
    st = sorted({k: I(r[idx[k]]) for k in stall}.items(), key=lambda x: -x[1])[:3]
    print("#%5d %7d %11d %-58s %s" % (i, I(r[S]), I(r[E]), r[1].strip()[:58], st))
