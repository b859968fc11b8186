#!/bin/bash
# End-of-round run on one GPU: the whole GPU suite, the default bench line, and the ncu launch list of the bench command.
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dist.py > gpurun_out/r2_pytest_final.log 2>&1; tail -2 gpurun_out/r2_pytest_final.log
timeout 500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], e["value"], e["packed_2_bytes"]["value"], e["cyvcf2_layout"]["value"], d["parity"]["ok"], d["parity"]["max_rel"], d["gpu_launches"])
print({k:(round(v["ms_per_step"],2), v.get("roofline_frac")) for k,v in d["tools"].items()})
print(d["cpu_baseline"]["value"], {k:v["value"] for k,v in d["cpu_baseline"]["tools"].items()}, d["ingest"]["value"])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_final_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bfinal_ncu.log 2>&1; echo ncu rc=$?
