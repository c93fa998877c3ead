#!/bin/bash
# MANO parity tests + warm per-kernel durations of the fused step at 512 and 4096 hands
python -m pytest tests -x -q -m gpu -k "mano or fit_step or bench_size" 2>&1 | tail -3
for B in 512 4096; do
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/mano_l_$B.csv python tools/prof_fused.py $B 6 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/mano_l_$B.csv')) if len(r)>5]
h=rows[0]; d=collections.OrderedDict()
for r in rows[1:]:
    try: d.setdefault(r[h.index('Kernel Name')][:44],[]).append(float(r[-1]))
    except: pass
print('B=$B', ' | '.join(f"{k.split('(')[0][-28:]} {sum(v[2:])/max(len(v[2:]),1)/1000:.1f}" for k,v in d.items() if 'at::' not in k and 'view' not in k))
PY
done
python tools/time_steps.py 512 4096 2>&1 | grep -E "chunks=[12] "
