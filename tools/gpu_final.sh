# Round-end GPU pass: smoke, all GPU tests, the default bench line (e2e + cpu baseline), launch list, ncu of the top kernel.
mkdir -p gpurun_out
KREGEX='estep|sstat|nz_|mask_|gather_list|colsum|row_norm|build_items|convert_f32|absmax|labels_to_q'
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tall.log 2>&1; echo "tall rc=$?"; tail -4 gpurun_out/tall.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KREGEX" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
for kn in $NCU_KERNELS; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$kn" -s 1 -c 1 -f -o gpurun_out/ncu_$kn python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$kn.log 2>&1; echo "ncu $kn rc=$?"
  ncu -i gpurun_out/ncu_$kn.ncu-rep --page raw --csv > gpurun_out/ncu_${kn}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$kn.ncu-rep --page source --csv > gpurun_out/ncu_${kn}_src.csv 2>/dev/null
done
