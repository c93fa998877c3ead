#!/bin/bash
# per-kernel warm durations of the fused step against the in-tree library and each experimental build build/var_*.so
for so in "" build/var_*.so; do
  echo "== ${so:-in-tree}"
  for B in 512 4096; do
  DSF_B200_LIB=${so:+$PWD/$so} ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/var_l.csv python tools/prof_fused.py $B 6 > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/var_l.csv')) if len(r)>5]
h=rows[0]; d=collections.OrderedDict()
for r in rows[1:]:
    try: d.setdefault(r[h.index('Kernel Name')][:44],[]).append(float(r[-1]))
    except: pass
print('B=$B', ' | '.join(f"{k.split('(')[0][-22:]} {sum(v[2:])/max(len(v[2:]),1)/1000:.1f}" for k,v in d.items() if 'at::' not in k and 'view' not in k))
PY
  done
done
