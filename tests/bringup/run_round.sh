#!/bin/bash
# One GPU call: parity tests, smoke, bench, then the ncu launch list of the same bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_launch.log
