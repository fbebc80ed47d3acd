#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/retr_launches.csv \
    python tools/prof_retr.py > gpurun_out/retr_launches_run.log 2>&1
echo "launch list exit=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/retr_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:][-22:]:
    print(r[ki][:70], r[vi], r[ui])
PY
