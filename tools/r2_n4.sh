#!/bin/bash
N=${1:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r2g_n$N.json 2> gpurun_out/r2g_n$N.err; echo "bench rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "n", d["n_gpus"])
s=d.get("sharded_step") or {}
print("sharded", s.get("steps_per_s"), s.get("e2e_steps_per_s"), s.get("equals_one_gpu_fold"))
print("msm", [(m["log2_points"], m["mpts_per_s"], m.get("sharding"), m.get("result_equals_closed_form")) for m in d["msm"]])
c=d["configs"]
print("cfg sharded", [(x["circuit"], round(x["steps_per_s"])) for x in c.get("sharded_step",[])])
print("cfg msm", [(m["log2_points"], round(m["mpts_per_s"])) for m in c.get("msm",[])])
print("mixed", c.get("mixed_transformations"))
PY
tail -5 gpurun_out/r2g_n$N.err
