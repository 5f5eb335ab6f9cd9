"""Per-opcode and per-instruction stall summary of one kernel in an .ncu-rep (needs --import-source on / -lineinfo):
   python tools/ncu_source.py file.ncu-rep kernel-regex [top-N]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
def I(r, name):
    try: return int(r[hdr.index(name)])
    except ValueError: return 0
S = "Warp Stall Sampling (All Samples)"
tot = sum(I(r, S) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("kernel", rows[0][1][:100]); print("total samples", tot)
agg = collections.Counter()
for r in body:
    for h in stalls: agg[h] += I(r, h)
print("stall mix:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / max(tot, 1)) for h, v in agg.most_common(8)))
byop = collections.Counter(); ex = collections.Counter(); st = collections.Counter()
for r in body:
    t = r[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0] + ("." + op.split(".")[-1] if op.startswith(("LDS", "STS", "LDG", "STG")) and "." in op else "")
    byop[op] += I(r, S); ex[op] += I(r, "Instructions Executed"); st[op] += 1
print("%-14s %8s %6s %7s %14s" % ("opcode", "samples", "%", "static", "executed"))
for op, v in byop.most_common(14): print("%-14s %8d %5.1f%% %7d %14d" % (op, v, 100 * v / max(tot, 1), st[op], ex[op]))
print("top instructions:")
for r in sorted(body, key=lambda r: -I(r, S))[:topn]:
    top = sorted(((I(r, h), h[6:]) for h in stalls), reverse=True)[:2]
    print("%6d  %-60s %s" % (I(r, S), r[1][:60], " ".join("%s=%d" % (n, v) for v, n in top)))
