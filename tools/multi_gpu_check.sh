#!/usr/bin/env bash
# 2-GPU checks on one box: bench.py under torchrun (weak scaling line) and the mapper with --gpus 2 (same SAM as one GPU).
set -e
N=${1:-2}
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/scale_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/scale_n$N.log > gpurun_out/scale_n$N.json || tail -20 gpurun_out/scale_n$N.log
python - <<PY
import json
for n in (1, $N):
    j = json.loads([l for l in open(f"gpurun_out/scale_n{n}.json") if l.startswith("{")][-1])
    print(n, "GPUs: value", round(j["value"] / 1e6, 1), "M reads/s  e2e", round(j["e2e"]["value"] / 1e6, 1), "M reads/s  ms/step", round(j["ms_per_step"], 3))
PY
python tools/cli_compare.py cfg2 --pairs 1000000 > /dev/null 2>&1 || true
cd /tmp/bmbs_bench/cfg2_s1002_x1
B=/root/repo/bitmapperbs_b200/_build/bmbs
$B --search g.fa --seq1 cmp_1000000_1.fq --seq2 cmp_1000000_2.fq --pe -t 16 --gpus 1 -o g1.sam 2>&1 | grep Total: || true
$B --search g.fa --seq1 cmp_1000000_1.fq --seq2 cmp_1000000_2.fq --pe -t 16 --gpus $N -o g2.sam 2>&1 | grep Total: || true
cmp <(grep -v "^@" g1.sam) <(grep -v "^@" g2.sam) && echo "SAM identical with --gpus 1 and --gpus $N (input order preserved)"
