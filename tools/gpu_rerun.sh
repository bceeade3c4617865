N=$1; W=$2; shift 2
timeout 900 python bench.py --gpus $N --workload $W "$@" > gpurun_out/r02_bench_$(echo $W | tr -d -)_n${N}.json 2> gpurun_out/r02_bench_$(echo $W | tr -d -)_n${N}.err
echo rc=$?; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r02_bench_$(echo $W | tr -d -)_n${N}.err | tail -3
python - <<PY
import json
l = json.loads(open("gpurun_out/r02_bench_$(echo $W | tr -d -)_n${N}.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "steps_per_s", "ms_per_step")}, "roofline", round(l["step_roofline"]["frac"], 3), l["run"]["exchange"])
PY
