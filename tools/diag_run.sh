#!/bin/bash
# run a python diagnostic with a host backtrace if it hangs: tools/diag_run.sh <seconds> <logname> <args...>
LIMIT=$1; LOG=$2; shift 2
export PYTHONUNBUFFERED=1
python -u "$@" > gpurun_out/$LOG.log 2>&1 &
PID=$!
for i in $(seq 1 $LIMIT); do sleep 1; kill -0 $PID 2>/dev/null || break; done
if kill -0 $PID 2>/dev/null; then
  echo "HUNG after $LIMIT s" >> gpurun_out/$LOG.log
  timeout 60 cuda-gdb -batch -p $PID -ex "thread apply all bt 14" > gpurun_out/$LOG.bt 2>&1
  kill -9 $PID
fi
wait $PID 2>/dev/null
tail -25 gpurun_out/$LOG.log
