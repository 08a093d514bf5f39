mkdir -p gpurun_out
timeout 600 python tests/two_level_diag.py --big > gpurun_out/diag.log 2>&1; echo "diag rc=$?"
timeout 600 python -m pytest tests/test_gpu_two_level.py -q -x > gpurun_out/t2.log 2>&1; echo "t2 rc=$?"; tail -15 gpurun_out/t2.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_two_level.py > gpurun_out/tall.log 2>&1; echo "tall rc=$?"; tail -8 gpurun_out/tall.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_two.json 2> gpurun_out/bench_two.err; echo "bench rc=$?"; cat gpurun_out/bench_two.json; tail -3 gpurun_out/bench_two.err
LCB_TC_TWO_LEVEL=0 timeout 400 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; cat gpurun_out/bench_dense.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_two.csv python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; echo "ncu1 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:estep_coarse -s 1 -c 1 -o gpurun_out/ncu_coarse python bench.py --n-points 4000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1; echo "ncu2 rc=$?"
echo ==== diag; tail -60 gpurun_out/diag.log
