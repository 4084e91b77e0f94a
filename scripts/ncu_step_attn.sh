# ncu --set full (with SASS source view) of the three attention kernels of one config-2 training step.
set -e
python scripts/ncu_step.py
ncu --set full --clock-control none --import-source on -k regex:"attn_" -c 3 -o /tmp/step_attn python scripts/ncu_step.py > gpurun_out/step_attn_ncu.log 2>&1
cp /tmp/step_attn.ncu-rep gpurun_out/step_attn.ncu-rep; ls -la gpurun_out/step_attn.ncu-rep
