#!/usr/bin/env bash
# on-box experiment: cfg4 launch list (dataset built first so that ncu's launch budget goes to the step), then the thread / warp
# boundary of the single-end finishing (FIN_THREAD_MAX) rebuilt and benched at a few values
O=gpurun_out; mkdir -p $O
python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > $O/expF_cfg4_bench.json 2> $O/expF_cfg4_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/expF_cfg4_launches.csv python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python bench.py --no-cpu-baseline --no-cfg4 > $O/expF_fin64.json 2>/dev/null
for T in 16 32; do
  sed -i "s/constexpr u32 FIN_THREAD_MAX = [0-9]*;/constexpr u32 FIN_THREAD_MAX = $T;/" bitmapperbs_b200/csrc/bmbs_kernels.cuh
  python -m bitmapperbs_b200.build > /dev/null 2>&1
  python bench.py --no-cpu-baseline --no-cfg4 > $O/expF_fin$T.json 2>/dev/null
done
ls $O
