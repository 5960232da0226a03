O=gpurun_out
N=${N:-2}
( time python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py -m gpu -x -q ) > $O/pytest_r2w_${N}gpu.log 2>&1; tail -4 $O/pytest_r2w_${N}gpu.log
for g in fused fused-nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --quick --shard tiles --gather $g --width 3840 --height 2160 --steps 6 --warmup 3 --frames-per-step 8 > $O/bench_r2w_n${N}_tiles_$g.json 2> $O/bench_r2w_n${N}_tiles_$g.err
python -c "
import json
d=json.loads(open('$O/bench_r2w_n${N}_tiles_$g.json').read().strip().splitlines()[-1])
print('N=$N tiles $g value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/frame', round(d['extra']['ms_per_frame'],4))" || tail -20 $O/bench_r2w_n${N}_tiles_$g.err
done
