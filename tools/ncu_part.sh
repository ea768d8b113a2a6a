#!/bin/bash
# usage: tools/ncu_part.sh tag [nshapes]   (env PWA_TMA_* select the variant)
python tools/part_probe.py ${2:-1} > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/part_$1.csv python tools/part_probe.py ${2:-1} > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/part_$1.csv") if not l.startswith("=="))]
agg=collections.OrderedDict()
for r in rows:
    n=r["Kernel Name"]
    if "partition" not in n and "reverse" not in n: continue
    k=n.split("(")[0].replace("void pwa::<unnamed>::","").replace("void pwa::","")
    agg.setdefault(k,[]).append(float(r["Metric Value"].replace(",",""))/1e3)
print("$1", {k: [round(min(v),1), len(v)] for k,v in agg.items()})
PY
