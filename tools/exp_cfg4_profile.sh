#!/usr/bin/env bash
# on-box: `ncu --set full` of the pair kernels on cfg4 (--pe --sensitive, 500 k pairs per step); the dataset is built first so that
# the capture starts at the step's kernels
O=gpurun_out; mkdir -p $O
python bench.py --workload cfg4 --reads 500000 --steps 2 --warmup 3 --no-cpu-baseline > $O/cfg4_bench.json 2> $O/cfg4_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^(finish_pe|finish_pe_long|sens_pair|seed_reseed|sens_reseed_filter|sens_reseed_finish|votes_big)" \
  --launch-skip 10 --launch-count 12 -f -o $O/cfg4_full python bench.py --workload cfg4 --reads 500000 --steps 2 --warmup 3 --no-cpu-baseline > $O/cfg4_full_bench.log 2>&1
echo "ncu exit $?"
ls -la $O | tail -5
