#!/bin/bash
# ONE ncu use per gpurun call: `list` = launch list of one bench run, otherwise a --set full capture of the named kernel.
# The plain command runs first and must exit 0.  TAG names the output files (default r2).
set -u
TAG=${TAG:-r2}
WHAT=${1:-list}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --host-phantom"
timeout 900 $CMD > gpurun_out/plain_profile.log 2> gpurun_out/plain_profile.err || { echo "plain run failed"; tail -5 gpurun_out/plain_profile.err; exit 1; }
tail -1 gpurun_out/plain_profile.log | cut -c1-250
if [ "$WHAT" = list ]; then
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_512x720.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "launch list exit $?"
else
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$WHAT" -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$WHAT $CMD > gpurun_out/ncu_full_$WHAT.log 2>&1
  echo "ncu $WHAT exit $?"
  ls -la gpurun_out/prof_${TAG}_$WHAT.ncu-rep
fi
