"""Development aid: static SASS opcode histogram of one kernel of a built library (cuobjdump -sass)."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist, n = None, collections.Counter(), 0
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            hist[m.group(2)] += 1; n += 1
print(pat, n, "instructions")
print(", ".join(f"{k} {v}" for k, v in hist.most_common(30)))
