#!/usr/bin/env bash
# The evidence set of a finished build in one gpurun call: GPU tests, smoke(), the bench line, the ncu launch list of the same
# command, one `ncu --set full` capture of the two main kernels, and the cfg5 verification microbench (reduced table).
#   gpurun --timeout 600 -- 'bash tools/final_round.sh TAG'
set -u
TAG=${1:-final}; O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/${TAG}_smoke.log
timeout 500 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log; echo "bench exit $?"; cut -c1-400 $O/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > $O/${TAG}_launches_bench.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^(seed_reads|verify_windows)' --launch-skip 8 --launch-count 2 -f -o $O/${TAG}_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > $O/${TAG}_full_bench.log 2>&1; echo "ncu full exit $?"
timeout 300 python tools/bench_verify.py --rates 0,0.04,0.08 --reads-log2 20 --reps 3 --cpu-sample-log2 12 > $O/${TAG}_verify_microbench.jsonl 2> $O/${TAG}_verify.log; echo "bench_verify exit $?"; cut -c1-160 $O/${TAG}_verify_microbench.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
