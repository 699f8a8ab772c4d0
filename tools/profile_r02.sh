set -e
python tools/profile_targets.py all > gpurun_out/prof_plain.log 2>&1
run() {  # name, section, kernel regex, count
  ncu --set full --clock-control none -k regex:"$3" -c $4 -o gpurun_out/r02_$1 python tools/profile_targets.py $2 > gpurun_out/ncu_r02_$1.log 2>&1 || true
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1.raw.csv 2>/dev/null || true
  rm -f gpurun_out/r02_$1.ncu-rep
}
run invert n15 "blu_invert_groups" 15
run kkt n15 "blu_kkt_syrk|blu_kkt_rows|blu_kkt_chol|blu_kkt_apply" 4
run gram gram "blu_gram_kernel" 4
run n20 n20 "blu_phi_partial|blu_grad_soa" 4
run batch batch "blu_batch_eval" 2
ls -la gpurun_out/r02_*.raw.csv
