"""Top stalled SASS instructions of an `ncu --page source --csv` export: address, samples, dominant stall, instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = 0; items = []
for r in rows[2:]:
    if len(r) < len(h): continue
    try: n = int(r[ix["# Samples"]])
    except ValueError: continue
    tot += n
    best = max(stalls, key=lambda k: int(r[ix[k]] or 0))
    items.append((n, r[ix["Address"]], best, r[ix[best]], r[ix["Source"]][:90], r[ix["Instructions Executed"]]))
print("total samples", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for n, a, b, bv, src, ex in sorted(items, key=lambda t: -t[0])[:top]:
    print("%6d %5.1f%% %s %-18s %-6s x%-9s %s" % (n, 100.0 * n / tot, a[-5:], b, bv, ex, src))
