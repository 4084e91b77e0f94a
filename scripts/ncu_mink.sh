# ncu --set full of the d_head-64 attention kernels of one MinkowskiNet config-4 step; the report travels back
# (about 20 MB) and is read with `ncu -i ... --page source --csv --print-source sass`.
set -e
python scripts/mink_ncu.py
ncu --set full --clock-control none --import-source on -k regex:"${1:-attn_}" -c ${2:-3} -o /tmp/mink_attn python scripts/mink_ncu.py > gpurun_out/mink_ncu.log 2>&1
ncu -i /tmp/mink_attn.ncu-rep --page raw --csv > gpurun_out/mink_attn_raw.csv
cp /tmp/mink_attn.ncu-rep gpurun_out/mink_attn.ncu-rep; ls -la gpurun_out/mink_attn.ncu-rep
