#!/bin/bash
# per-class ring FFT durations for a list of tuning settings: tools/fft_classes.sh "fg=1" "fg=3 dbg=1" ...
for cfg in "$@"; do
  tag=$(echo $cfg | tr ' =' '__')
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ring_synthesis -c 4 --csv --log-file gpurun_out/fftc_$tag.csv python tools/stage_bench.py 12 8192 1 $cfg > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/fftc_$tag.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]; h=rows[hi]
out=[]
for r in rows[hi+1:]:
    out.append("%s%s=%.2f" % (r[h.index("Kernel Name")][5:8], r[h.index("Grid Size")].replace(" ",""), float(r[h.index("Metric Value")])/1e6))
print("$cfg :", " ".join(out))
PY
done
