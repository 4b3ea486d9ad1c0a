
for c in grayscale brightness contrast resize crop blur sharpness hash; do
  python bench.py --circuit $c --steps 100 --no-cpu-baseline --msm-log2 > gpurun_out/s3_circuit_$c.json 2>> gpurun_out/s3_circuits.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s3_circuit_$c.json"))
print("$c", round(d["value"],1), round(d["e2e"]["value"],1), d["config"]["workload"][:90], d["roofline"]["frac"], d["clocks"]["sm_mhz"])
PY
done
