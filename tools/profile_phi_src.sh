set -e
timeout 200 python tools/eval_bench.py 20 var 5 > gpurun_out/phi_src_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"blu_phi_partial" -s 4 -c 1 -o gpurun_out/phi_src python tools/eval_bench.py 20 var 5 > gpurun_out/ncu_phi_src.log 2>&1 || true
ncu -i gpurun_out/phi_src.ncu-rep --page raw --csv > gpurun_out/phi_src.raw.csv 2>/dev/null || true
ncu -i gpurun_out/phi_src.ncu-rep --page source --csv --print-source sass > gpurun_out/phi_src.sass.csv 2>/dev/null || true
rm -f gpurun_out/phi_src.ncu-rep
ls -la gpurun_out/phi_src*
