#!/bin/bash
# One GPU-box pass: parity tests, bench (graph + eager), per-kernel micro-bench.  Usage: tools/gpu_check.sh [tag]
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_${TAG}_eager.json 2>> gpurun_out/bench_${TAG}.err; echo "eager bench exit $?"; cut -c1-400 gpurun_out/bench_${TAG}_eager.json
