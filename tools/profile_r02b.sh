# final measurement pass of round 2 (after the run-length Phi kernel): plain bench, launch list, full capture of the Phi / gradient streams
set -e
python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
python bench.py --steps 2 --warmup 1 > gpurun_out/r02b_bench_short.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02b_ncu_launches.log 2>&1 || true
python tools/profile_targets.py n20 > gpurun_out/r02b_prof_plain.log 2>&1
ncu --set full --clock-control none -k regex:"blu_phi_partial|blu_grad_soa" -c 4 -o gpurun_out/r02b_n20 python tools/profile_targets.py n20 > gpurun_out/r02b_ncu_n20.log 2>&1 || true
ncu -i gpurun_out/r02b_n20.ncu-rep --page raw --csv > gpurun_out/r02b_n20.raw.csv 2>/dev/null || true
rm -f gpurun_out/r02b_n20.ncu-rep
ls -la gpurun_out/r02b_*
