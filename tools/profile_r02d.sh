# measurement pass after the second-generation operator kernels and the KKT kernels (round 2, final session)
set -e
python bench.py > gpurun_out/r02d_bench_default.json 2> gpurun_out/r02d_bench_default.err
python bench.py --impl reference > gpurun_out/r02d_bench_reference.json 2> /dev/null || true
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02d_ncu_launches.log 2>&1 || true
python tools/kkt_bench.py 15 > gpurun_out/r02d_kkt_bench15.txt 2>&1
ncu --set full --clock-control none -k regex:"blu_kkt_|blu_hv_" -s 12 -c 14 -o gpurun_out/r02d_kkt python tools/profile_targets.py kkt_hv > gpurun_out/r02d_ncu_kkt.log 2>&1 || true
ncu -i gpurun_out/r02d_kkt.ncu-rep --page raw --csv > gpurun_out/r02d_kkt.raw.csv 2>/dev/null || true
rm -f gpurun_out/r02d_kkt.ncu-rep
ls -la gpurun_out/r02d_*
