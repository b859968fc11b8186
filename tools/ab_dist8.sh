#!/bin/bash
# N = 8: exchange issued at once on the side stream (default) vs deferred behind the next scan (TRT_DIST_DEFER=1)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
F="--statstr-only --no-cpu-baseline --steps 10 --warmup 3"
pr() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), d['timing'].get('per_rank_us_per_step_device_wall'), round(d['e2e']['value']))
" $1; }
timeout 120 $T --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 $F > gpurun_out/r2_ab8_side.json 2>>gpurun_out/r2_ab8.err; pr gpurun_out/r2_ab8_side.json
TRT_DIST_DEFER=1 timeout 120 $T --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 $F > gpurun_out/r2_ab8_defer.json 2>>gpurun_out/r2_ab8.err; pr gpurun_out/r2_ab8_defer.json
