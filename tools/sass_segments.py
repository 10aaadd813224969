"""Split the SASS source page of an .ncu-rep kernel into runs of equal execution count (= basic-block groups) and print
each run's share of the executed warp instructions, its SIMT width and stall samples.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass -k regex:NAME > src.csv; python tools/sass_segments.py src.csv [block]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Address":
        cur = {"hdr": r, "data": []}
        blocks.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr, data = b["hdr"], b["data"]
isrc, ii, it, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
tot = sum(int(r[ii]) for r in data)
print("total warp instructions", tot, "| SASS lines", len(data), "| kernels in file", len(blocks))
runs, cur = [], None
for k, r in enumerate(data):
    n = int(r[ii])
    if cur and abs(n - cur["n"]) <= 0.02 * max(n, cur["n"], 1):
        cur["len"] += 1; cur["sum"] += n; cur["thr"] += float(r[it]) * n; cur["smp"] += int(r[iss])
    else:
        if cur: runs.append(cur)
        cur = {"start": k, "n": n, "len": 1, "sum": n, "thr": float(r[it]) * n, "smp": int(r[iss]), "first": r[isrc].strip()}
runs.append(cur)
for r in runs:
    if r["sum"] > 0.004 * tot:
        print(f"{r['start']:5d} len {r['len']:4d} exec {r['n']:9d} share {100*r['sum']/tot:5.1f}% lanes {r['thr']/max(r['sum'],1):5.1f} samples {r['smp']:6d}  {r['first'][:48]}")
