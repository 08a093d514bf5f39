# One GPU-box pass over the two-level E step: diagnostics, parity tests, bench, launch list, ncu capture.
mkdir -p gpurun_out
KREGEX='estep|sstat|nz_|mask_|gather_list|colsum|row_norm|build_items|convert_f32|absmax|labels_to_q'
timeout 600 python tests/two_level_diag.py ${DIAG_ARGS:---big} > gpurun_out/diag.log 2>&1; echo "diag rc=$?"
timeout 600 python -m pytest tests/test_gpu_two_level.py -q -x > gpurun_out/t2.log 2>&1; echo "t2 rc=$?"; tail -5 gpurun_out/t2.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_two_level.py > gpurun_out/tall.log 2>&1; echo "tall rc=$?"; tail -4 gpurun_out/tall.log
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:---no-e2e --no-cpu-baseline} > gpurun_out/bench_two.json 2> gpurun_out/bench_two.err; echo "bench rc=$?"; cat gpurun_out/bench_two.json; tail -3 gpurun_out/bench_two.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREGEX" -c 60 --csv --log-file gpurun_out/launches_two.csv python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu1 rc=$?"
if [ -n "$NCU_FULL" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$NCU_FULL" -s 1 -c 1 -f -o gpurun_out/ncu_full python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1; echo "ncu2 rc=$?"
  ncu -i gpurun_out/ncu_full.ncu-rep --page raw --csv > gpurun_out/ncu_full_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_full.ncu-rep --page source --csv > gpurun_out/ncu_full_src.csv 2>/dev/null
fi
echo ==== diag; grep "two_level=\|^(\|^True\|^False" gpurun_out/diag.log | tail -24
