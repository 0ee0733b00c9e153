#!/usr/bin/env bash
# Where the command-line mapper's time goes on the bench workload: runs bench.py once for the dataset, then the mapper with its
# timing output on (BMBS_TIMING / BMBS_VERBOSE), device finishing and host finishing, 16 and 8 threads.
#   gpurun -- 'BMBS_BENCH_SCALE=0.1 bash tools/cli_timing.sh tag'
TAG=${1:-cli}; O=gpurun_out; mkdir -p $O
SCALE=${BMBS_BENCH_SCALE:-1}
python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log
D=$(ls -d /tmp/bmbs_bench/cfg3_s1003_x* | head -1)
FQ=$(ls $D/wp_cfg3_*.fq | head -1)
EXE=$PWD/bitmapperbs_b200/_build/bmbs
cd $D
for rep in 1 2; do
  for mode in dev host; do
    for t in 16; do
      echo "=== rep $rep finish=$mode -t $t" >> $OLDPWD/$O/${TAG}_cli.log
      if [ $mode = host ]; then export BMBS_HOST_FINISH=1; else unset BMBS_HOST_FINISH; fi
      BMBS_TIMING=1 BMBS_VERBOSE=1 $EXE --search g.fa --seq $(basename $FQ) -t $t -o /dev/null 2>> $OLDPWD/$O/${TAG}_cli.log
    done
  done
done
cd $OLDPWD
tail -80 $O/${TAG}_cli.log
