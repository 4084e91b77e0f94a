set -e
python scripts/mink_ncu.py
ncu --set full --clock-control none -k regex:attn_ -c 3 -o /tmp/mink_attn python scripts/mink_ncu.py > gpurun_out/mink_ncu.log 2>&1
ncu -i /tmp/mink_attn.ncu-rep --page raw --csv > gpurun_out/mink_attn_raw.csv
cp /tmp/mink_attn.ncu-rep gpurun_out/mink_attn.ncu-rep; ls -la gpurun_out/mink_attn.ncu-rep
