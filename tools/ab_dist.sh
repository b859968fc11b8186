#!/bin/bash
# A/B of the gather placement on 2 GPUs of one box: N=1, N=2 with the exchange on the side stream (default), N=2 with it
# on the context stream (TRT_DIST_MAIN_STREAM=1).  usage: gpurun --gpus 2 -- bash tools/ab_dist.sh
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
F="--statstr-only --no-cpu-baseline --steps 10 --warmup 3"
pr() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), d['timing'].get('per_rank_us_per_step_device_wall'), round(d['e2e']['value']))
" $1; }
timeout 150 $T --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 $F > gpurun_out/r2_ab_n2_side.json 2>>gpurun_out/r2_ab.err; pr gpurun_out/r2_ab_n2_side.json
TRT_DIST_MAIN_STREAM=1 timeout 150 $T --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 $F > gpurun_out/r2_ab_n2_main.json 2>>gpurun_out/r2_ab.err; pr gpurun_out/r2_ab_n2_main.json
