#!/bin/bash
# full verification of a round: GPU tests, smoke, bench (ours + reference arm), launch list + ncu captures
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; $T 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; $T 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "=== bench ours"; $T 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; head -c 300 gpurun_out/bench_n1.json; echo
echo "=== bench reference"; $T 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 400 gpurun_out/bench_ref.json; echo
if [ "$1" == "ncu" ]; then echo "=== ncu round"; bash tools/ncu_round.sh 2>&1 | tail -16; fi
