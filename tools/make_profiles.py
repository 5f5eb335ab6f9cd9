"""Copy the judged evidence from gpurun_out/ (scratch) into profiles/ (tracked): the ncu launch list of the bench's timed
region and text summaries (key metrics + per-opcode stall table) of the ncu --set full captures.
   python tools/make_profiles.py <round-tag> launches.csv name=file.ncu-rep:kernel-regex ..."""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1]
os.makedirs("profiles", exist_ok=True)
out_traffic = {}
for a in sys.argv[2:]:
    if a.endswith(".csv"):
        rows = list(csv.reader(open(a)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.OrderedDict(); lines = []
        for r in rows[hi + 1:]:
            if len(r) <= vi: continue
            v = float(r[vi].replace(",", "")); u = r[ui]
            v = v / 1000 if u == "us" else v / 1e6 if u == "ns" else v * 1000 if u == "s" else v
            n = r[ki].split("(")[0]
            e = agg.setdefault(n, [0, 0.0]); e[0] += 1; e[1] += v
            lines.append("%s,%s,%s,%.6f" % (r[0], n, r[hdr.index("Grid Size")].replace(",", " "), v))
        tot = sum(e[1] for e in agg.values())
        with open("profiles/%s_launches.csv" % tag, "w") as f:
            f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none: every launch of one timed bench step\n")
            f.write("id,kernel,grid,duration_ms\n" + "\n".join(lines) + "\n")
        with open("profiles/%s_launch_shares.txt" % tag, "w") as f:
            f.write("total %.3f ms over %d launches (one lens plane, cold-cache serialised ncu timings)\n" % (tot, sum(e[0] for e in agg.values())))
            for n, e in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write("%-60s %4d launches %10.3f ms %5.1f%%\n" % (n[:60], e[0], e[1], 100 * e[1] / tot))
    else:
        name, rest = a.split("=", 1)
        rep, kre = rest.split(":", 1)
        s1 = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
        s2 = subprocess.run([sys.executable, "tools/ncu_source.py", rep, kre, "20"], capture_output=True, text=True).stdout
        with open("profiles/%s_%s.txt" % (tag, name), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on, report %s\n" % os.path.basename(rep))
            f.write(s1 + "\n" + s2)
        # dram traffic per launch of the first kernel in the report
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        h, u = rr[0], rr[1]
        for row in rr[2:]:
            if kre.split("|")[0] in row[h.index("Kernel Name")]:
                def val(k):
                    x = float(row[h.index(k)]); un = u[h.index(k)]
                    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un]
                out_traffic[name] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                     "duration_ms": float(row[h.index("gpu__time_duration.sum")]), "report": os.path.basename(rep)}
                break
if out_traffic:
    p = "profiles/%s_traffic.json" % tag
    old = json.load(open(p)) if os.path.exists(p) else {}
    old.update(out_traffic)
    json.dump(old, open(p, "w"), indent=1)
