# Short GPU-box pass: parity tests, two-level diagnostics, device-only bench, launch list.
mkdir -p gpurun_out
KREGEX='estep|sstat|nz_|mask_|gather_list|colsum|row_norm|build_items|convert_f32|absmax|labels_to_q'
timeout 300 python tests/two_level_diag.py --big > gpurun_out/diag.log 2>&1; echo "diag rc=$?"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tall.log 2>&1; echo "tall rc=$?"; tail -6 gpurun_out/tall.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; cat gpurun_out/bench_q.json; tail -3 gpurun_out/bench_q.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREGEX" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
for kn in $NCU_KERNELS; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$kn" -s 1 -c 1 -f -o gpurun_out/ncu_$kn python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
  ncu -i gpurun_out/ncu_$kn.ncu-rep --page raw --csv > gpurun_out/ncu_${kn}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$kn.ncu-rep --page source --csv > gpurun_out/ncu_${kn}_src.csv 2>/dev/null
done
echo ==== diag; grep "two_level=\|^True\|^False" gpurun_out/diag.log | tail -12
