"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv, sys
rows = list(csv.reader(sys.stdin)); h, u = rows[0], rows[1]
want = sys.argv[1:]
for r in rows[2:]:
    for k in want:
        for i, n in enumerate(h):
            if n == k:
                print("%-75s %s %s" % (k, r[i], u[i]))
