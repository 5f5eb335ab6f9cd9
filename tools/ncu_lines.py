"""Per-source-line stall samples of kernels in an .ncu-rep (needs --import-source on / -lineinfo):
   python tools/ncu_lines.py file.ncu-rep kernel-regex [launch-index] [top-N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}; blocks.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]) and r[2] == "-":
        cur["rows"].append(r)
b = blocks[which]
hdr = b["hdr"]
S = hdr.index("Warp Stall Sampling (All Samples)")
st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
lines = []
for r in b["rows"]:
    try:
        v = int(r[S])
    except ValueError:
        continue
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, h[6:]) for i, h in st), reverse=True)[:2]
    lines.append((v, int(r[0]), r[1], top))
tot = sum(x[0] for x in lines)
print("launch", which, "of", len(blocks), "total samples", tot)
for v, ln, text, top in sorted(lines, reverse=True)[:topn]:
    print("%7d %5.1f%%  L%-5d %-100s %s" % (v, 100.0 * v / max(tot, 1), ln, text.strip()[:100], " ".join("%s=%d" % (n, c) for c, n in top)))
