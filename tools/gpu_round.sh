# One GPU-box pass: parity tests, diagnostics, full bench (e2e + cpu baseline), launch list, ncu captures, host probe.
mkdir -p gpurun_out
KREGEX='estep|sstat|nz_|mask_|gather_list|colsum|row_norm|build_items|convert_f32|absmax|labels_to_q'
timeout 600 python tests/two_level_diag.py --big > gpurun_out/diag.log 2>&1; echo "diag rc=$?"
timeout 600 python -m pytest tests/test_gpu_two_level.py -q -x > gpurun_out/t2.log 2>&1; echo "t2 rc=$?"; tail -5 gpurun_out/t2.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_two_level.py > gpurun_out/tall.log 2>&1; echo "tall rc=$?"; tail -4 gpurun_out/tall.log
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREGEX" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
for kn in $NCU_KERNELS; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$kn" -s 1 -c 1 -f -o gpurun_out/ncu_$kn python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
  ncu -i gpurun_out/ncu_$kn.ncu-rep --page raw --csv > gpurun_out/ncu_${kn}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$kn.ncu-rep --page source --csv > gpurun_out/ncu_${kn}_src.csv 2>/dev/null
done
if [ -n "$HOST_PROBE" ]; then
  nproc; free -g | head -2; lscpu | grep -i "model name\|socket\|numa" | head -6
  gcc -O3 -march=native -fopenmp tools/host_probe.c -o /tmp/host_probe && /tmp/host_probe
fi
echo ==== diag; grep "two_level=\|^True\|^False" gpurun_out/diag.log | tail -12
