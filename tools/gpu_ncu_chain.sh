#!/bin/bash
# quick ncu look at the chain kernels only: tensor-pipe activity, issue stalls (a handful of launches, a few metrics)
mkdir -p gpurun_out
timeout 300 python tools/prof_target.py > gpurun_out/prof_target.log 2>&1 || { echo "plain run failed"; tail gpurun_out/prof_target.log; exit 1; }
timeout 600 ncu --clock-control none -k regex:"k_chain" -s 6 -c 3 --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__cycles_active.avg,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --csv --log-file gpurun_out/ncu_chain.csv python tools/prof_target.py > gpurun_out/ncu_chain.log 2>&1
echo "ncu exit $?"; python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_chain.csv')) if len(r)>10]
hdr=rows[0]; 
for r in rows[1:]:
    d=dict(zip(hdr,r)); print(d.get('Kernel Name','')[:40], d.get('Metric Name'), d.get('Metric Value'))
PY
