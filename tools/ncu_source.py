"""Per-instruction stall summary of one kernel in an .ncu-rep.  Usage: ncu_source.py file.ncu-rep kernel_regex"""
import csv, collections, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
h = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(h) and r[0].startswith("0x")]
ix = {n: i for i, n in enumerate(h)}
S = lambda r: int(r[ix["# Samples"]]); X = lambda r: int(r[ix["Instructions Executed"]])
tot = sum(S(r) for r in data); te = sum(X(r) for r in data)
print(rows[hi - 1][1][:100]); print("samples", tot, "static instr", len(data), "executed warp-instr", te)
byop = collections.Counter(); ex = collections.Counter()
for r in data:
    t = r[ix["Source"]].strip().split(); op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    byop[op] += S(r); ex[op] += X(r)
for op, c in byop.most_common(14): print(f"  {op:10s} samples {100 * c / tot:6.2f}%  exec {100 * ex[op] / te:6.2f}%")
stl = [k for k in h if k.startswith("stall_") and "(" not in k]
print("stall totals:", ", ".join(f"{k[6:]}={sum(int(r[ix[k]]) for r in data) * 100 // tot}%" for k in stl if sum(int(r[ix[k]]) for r in data) * 50 > tot))
for i in sorted(range(len(data)), key=lambda i: -S(data[i]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    r = data[i]; st = sorted(((int(r[ix[k]]), k[6:]) for k in stl), reverse=True)[:2]
    print(f"  [{i:5d}] {r[ix['Source']].strip()[:64]:64s} {S(r):7d} x{X(r):8d} {st}")
n = len(data)
print("samples / executed by tenth of the code:", [(sum(S(r) for r in data[d * n // 10:(d + 1) * n // 10]) * 100 // tot, sum(X(r) for r in data[d * n // 10:(d + 1) * n // 10]) * 100 // te) for d in range(10)])
