O=gpurun_out
N=8
run() { g=$1; c=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $N --quick --shard tiles --gather $g --contexts $c --width 3840 --height 2160 --steps 6 --warmup 3 --frames-per-step 8 > $O/bench_r2x_n${N}_tiles_${g}_c$c.json 2> $O/bench_r2x_n${N}_tiles_${g}_c$c.err
python -c "
import json
d=json.loads(open('$O/bench_r2x_n${N}_tiles_${g}_c$c.json').read().strip().splitlines()[-1])
print('N=$N tiles $g contexts $c value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/frame', round(d['extra']['ms_per_frame'],4))" || tail -5 $O/bench_r2x_n${N}_tiles_${g}_c$c.err
}
run fused 3
run fused-nccl 3
run fused 4
run fused 6
