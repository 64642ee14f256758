#!/bin/bash
# First GPU pass: parity tests, smoke, bench, launch list, one full ncu capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpu_count', os.cpu_count())" >> gpurun_out/gpu.txt
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
echo "== bench 256"; timeout 600 python bench.py --size 256 --views 360 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; echo "exit $?"; cat gpurun_out/bench_256.json; tail -5 gpurun_out/bench_256.err
echo "== bench 512"; timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "exit $?"; cat gpurun_out/bench_512.json; tail -5 gpurun_out/bench_512.err
echo "== ncu"
CMD="python bench.py --size 256 --views 48 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ray_kernel|adjoint_gather' -s 4 -c 3 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
