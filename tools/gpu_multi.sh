#!/bin/bash
# usage: bash tools/gpu_multi.sh <tag> <ngpus>
TAG=${1:-m}
N=${2:-2}
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -15
for cfg in "poses 1920 1080" "tiles 1920 1080" "tiles 3840 2160"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --shard $1 --width $2 --height $3 > $O/bench_${TAG}_n${N}_$1_$2.json 2> $O/bench_${TAG}_n${N}_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${TAG}_n${N}_$1_$2.json").read().strip().splitlines()[-1])
    print("N=$N $1 $2x$3", round(d["value"],1), "Mpx/s  e2e", round(d["e2e"]["value"],1), d["scaling"], "step_ms", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("N=$N $1 $2 FAILED", e); print(open("$O/bench_${TAG}_n${N}_$1_$2.err").read()[-3000:])
PY
done
python bench.py --steps 20 --warmup 5 --width 3840 --height 2160 --no-cpu-baseline --no-second-flavour > $O/bench_${TAG}_n1_4k.json 2>$O/bench_${TAG}_n1_4k.err; python -c "
import json; d=json.load(open('$O/bench_${TAG}_n1_4k.json')); print('N=1 4K', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4))"
