#!/usr/bin/env bash
# verify_windows variants (BMBS_VERIFY_VARIANT = 0 / 1 / 2: where the shifts of the 32-bit band run): parity tests of the
# verification kernel and the cfg5 microbench for each, then the GPU tests and the bench line with the default.
#   gpurun --timeout 800 -- 'bash tools/exp_verify_variants.sh TAG'
set -u
TAG=${1:-vv}; O=gpurun_out; mkdir -p $O
for V in 0 1 2 3; do
  BMBS_VERIFY_VARIANT=$V timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "verify and default" > $O/${TAG}_pytest_verify_v$V.log 2>&1
  echo "variant $V verify tests exit $?"; tail -2 $O/${TAG}_pytest_verify_v$V.log
  BMBS_VERIFY_VARIANT=$V timeout 300 python tools/bench_verify.py --lengths 100,150,250 --rates 0.02 --reps 3 --reads-log2 20 --cpu-sample-log2 10 \
    > $O/${TAG}_verify_v$V.jsonl 2> $O/${TAG}_verify_v$V.log
  echo "variant $V microbench exit $?"; cut -c1-200 $O/${TAG}_verify_v$V.jsonl
done
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 700 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log; echo "bench exit $?"; cut -c1-1500 $O/${TAG}_bench.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
