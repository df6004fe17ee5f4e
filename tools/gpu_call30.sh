#!/bin/bash
# round 2, call 30: ncu launch list of one evaluation at N = 20000 (final kernels), ncu --set full of the dominant GEMM launch,
# per-launch timelines at N = 5018 and 10570
mkdir -p gpurun_out
for n in 5018 10570; do
  PIGP_PROF_DUMP=gpurun_out/r02_c30_timeline_$n.csv timeout 120 python tools/one_step.py $n >> gpurun_out/r02_c30_onestep.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c30_launches.csv python tools/one_step.py 20000 > gpurun_out/r02_c30_ncu_list.log 2>&1
IDX=$(python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02_c30_launches.csv') if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
g=[float(r[iv].replace(",","")) for r in rows[1:] if "k_gemm" in r[ik]]
print(max(range(len(g)), key=lambda i:g[i]))
PY
)
echo "dominant k_gemm launch index among k_gemm* launches: $IDX" > gpurun_out/r02_c30_idx.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm -s $IDX -c 1 -o gpurun_out/r02_c30_gemm_full python tools/one_step.py 20000 > gpurun_out/r02_c30_ncu_full.log 2>&1
gzip -f gpurun_out/r02_c30_launches.csv
