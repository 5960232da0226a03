#!/bin/bash
# usage: bash tools/gpu_multi_r2.sh <N> [tag] : on an N-GPU box - the multi-GPU tests (fused tile gather across processes, device
# groups in one process) and the driver's own command at N GPUs (poses weak scaling + BASELINE.json config 3 tiles in one line)
N=${1:-2}; TAG=${2:-r2}
O=gpurun_out; mkdir -p $O
( time python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py -m gpu -x -q ) > $O/pytest_${TAG}_${N}gpu.log 2>&1; tail -4 $O/pytest_${TAG}_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N > $O/bench_${TAG}_n$N.json 2> $O/bench_${TAG}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${TAG}_n$N.json").read().strip().splitlines()[-1]); c=d.get("config3_tiles",{})
    print("N=$N value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/frame", round(d["extra"]["ms_per_frame"],4), "| config3 tiles", round(c.get("value",0),1), "ms/frame", c.get("ms_per_frame"), "e2e", c.get("e2e",{}).get("value"), c.get("completion"), c.get("error"))
except Exception as e:
    print("bench FAILED", e); print(open("$O/bench_${TAG}_n$N.err").read()[-3000:])
PY
