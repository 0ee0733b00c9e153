#!/usr/bin/env bash
# One gpurun call's worth of evidence: GPU parity tests, the bench line, the ncu launch list of the same command and
# one `ncu --set full` capture of the main kernels.  Everything lands in gpurun_out/ (copied into profiles/ by hand).
#   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [tag] [what...]'      what: tests bench cfg2 refarm launches traffic full cliscale verify fullverify scale2   (votes_big runs twice per round now: see KERNELS)
set -u
TAG=${1:-run}; shift || true
WHAT=${*:-tests bench launches full}
O=gpurun_out; mkdir -p $O
KERNELS='regex:^(pack_reads|seed_reads|expand_locate|votes_classify|votes_sort|votes_mid|votes_big1k|votes_big|gather_work|verify_windows|finish_se|finish_long|finish_huge|finish_sorted)'
BENCH_ARGS=${BENCH_ARGS:-}
for w in $WHAT; do
  case $w in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log ;;
    bench)
      timeout 900 python bench.py $BENCH_ARGS > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log; echo "bench exit $?"; cat $O/${TAG}_bench.json ;;
    cfg2)
      # round 1's headline workload (100 Mbp uniform genome, pairs, fast mode) with the current build, for comparison
      timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.log; echo "bench cfg2 exit $?"; cut -c1-600 $O/${TAG}_bench_cfg2.json ;;
    refarm)
      timeout 600 python bench.py $BENCH_ARGS --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.log; cat $O/${TAG}_bench_reference.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
        python bench.py $BENCH_ARGS --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > $O/${TAG}_launches_bench.log 2>&1; echo "launch list exit $?" ;;
    full)
      timeout 900 ncu --set full --clock-control none --import-source on -k "$KERNELS" --launch-skip ${NCU_SKIP:-80} --launch-count ${NCU_COUNT:-14} -f -o $O/${TAG}_full \
        python bench.py $BENCH_ARGS --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > $O/${TAG}_full_bench.log 2>&1; echo "ncu full exit $?" ;;
    traffic)
      # DRAM bytes and duration of one launch of every step kernel at the bench's own scale (one pass: nothing is replayed)
      timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$KERNELS" \
        --launch-skip ${NCU_SKIP:-80} --launch-count ${NCU_COUNT:-14} --csv --log-file $O/${TAG}_traffic.csv \
        python bench.py $BENCH_ARGS --steps 2 --warmup 3 --no-cpu-baseline --no-cfg4 > $O/${TAG}_traffic_bench.log 2>&1; echo "traffic exit $?" ;;
    cliscale)
      timeout 1500 python tools/cli_scale.py ${CLI_SCALE_ARGS:-} --out $O/${TAG}_cli_scale.json > $O/${TAG}_cli_scale.log 2>&1; echo "cli_scale exit $?"; tail -5 $O/${TAG}_cli_scale.log ;;
    verify)
      timeout 900 python tools/bench_verify.py > $O/${TAG}_verify.json 2> $O/${TAG}_verify.log; echo "bench_verify exit $?"; tail -c 1500 $O/${TAG}_verify.json ;;
    fullverify)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_windows --launch-skip 3 --launch-count 1 -f -o $O/${TAG}_fullverify150 \
        python tools/bench_verify.py --lengths 150 --rates 0.02 --reps 2 --no-oracle --cpu-sample-log2 10 > $O/${TAG}_fullverify.log 2>&1; echo "ncu verify150 exit $?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:verify_windows --launch-skip 3 --launch-count 1 -f -o $O/${TAG}_fullverify250 \
        python tools/bench_verify.py --lengths 250 --rates 0.02 --reps 2 --no-oracle --cpu-sample-log2 10 >> $O/${TAG}_fullverify.log 2>&1; echo "ncu verify250 exit $?" ;;
    scale2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py $BENCH_ARGS --gpus 2 --steps 10 --warmup 3 \
        > $O/${TAG}_scale2.json 2> $O/${TAG}_scale2.log; cat $O/${TAG}_scale2.json ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
ls -la $O | tail -20
