# ncu --set full of every library kernel of one eager config-2 training step (scripts/ncu_step.py runs two identical
# steps; scripts/ncu_summary.py keeps the second).  The report is too large to travel: the raw CSV is exported here.
# A second, metrics-only pass gives the launch list (cold-cache, serialised durations).
set -e
K='regex:gemm_kernel|attn_|pack128|ln_|colsum|compat_|head_|csa_head|grad_unscale|block_add|sgemm'
python scripts/ncu_step.py
ncu --set full --clock-control none -k "$K" -o /tmp/full_step python scripts/ncu_step.py > gpurun_out/full_step_ncu.log 2>&1
ncu -i /tmp/full_step.ncu-rep --page raw --csv > gpurun_out/full_step_raw.csv
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/full_step_launches.csv python scripts/ncu_step.py > /dev/null 2>&1
ls -la /tmp/full_step.ncu-rep gpurun_out/full_step_raw.csv gpurun_out/full_step_launches.csv
