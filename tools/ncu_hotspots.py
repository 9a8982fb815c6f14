"""Dev tool: `ncu -i <rep> --page source --csv` -> the source lines / SASS instructions that hold most warp-state samples of
one kernel (needs -lineinfo and --import-source on).  python tools/ncu_hotspots.py <rep> [kernel-regex] [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
if kern:
    cmd += ["-k", "regex:" + kern, "-c", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next((i for i, r in enumerate(rows) if any("Samples" in c for c in r)), None)
if hdr_i is None:
    print("no sampling columns found; first lines:\n" + "\n".join(out.splitlines()[:20]))
    sys.exit(0)
hdr = rows[hdr_i]
scol = next(i for i, c in enumerate(hdr) if "Samples" in c)
src = next((i for i, c in enumerate(hdr) if c.strip() in ("Source", "SASS")), 1)
data = []
tot = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= max(scol, src):
        continue
    try:
        v = float(r[scol].replace(",", ""))
    except ValueError:
        continue
    tot += v
    data.append((v, r[src].strip()))
data.sort(key=lambda t: -t[0])
print(f"# {rep} kernel~{kern or '*'}: {len(data)} instructions, {tot:.0f} samples ({hdr[scol]})")
opc = {}
for v, s in data:
    op = s.split()[0] if s else "?"
    if op.startswith("@"):
        op = s.split()[1] if len(s.split()) > 1 else op
    op = op.split(".")[0]
    opc[op] = opc.get(op, 0) + v
print("# samples by opcode: " + ", ".join(f"{k} {v / max(tot, 1):.1%}" for k, v in sorted(opc.items(), key=lambda t: -t[1])[:14]))
for v, s in data[:top]:
    print(f"{v:10.0f} {v / max(tot, 1):7.2%}  {s[:150]}")
