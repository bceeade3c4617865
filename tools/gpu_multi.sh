# multi-GPU bench lines of a round: bash tools/gpu_multi.sh <N> <round tag>   (run under `gpurun --gpus N`)
N=${1:-8}; R=${2:-r02}; OUT=gpurun_out
mkdir -p $OUT
run() {  # name, extra args
  name=$1; shift
  timeout 900 python bench.py --gpus $N "$@" > $OUT/${R}_bench_${name}_n${N}.json 2> $OUT/${R}_bench_${name}_n${N}.err
  echo "== $name N=$N rc=$?"; tail -c 300 $OUT/${R}_bench_${name}_n${N}.err | grep -v OMP_NUM | tail -3
  python - <<PY
import json
try:
    l = json.loads(open("$OUT/${R}_bench_${name}_n${N}.json").read().strip().splitlines()[-1])
    print({k: l[k] for k in ("value", "steps_per_s", "ms_per_step", "gpu_launches")}, "roofline", round(l["step_roofline"]["frac"], 3), "nonoverlapped", round(l["step_roofline"]["frac_non_overlapped"], 3))
    print("parity", l.get("parity")); print("weak_ref", l.get("weak_ref")); print("fft", l.get("fft")); print("exchange", l["run"]["exchange"])
    for k in l["kernels"][:8]: print("  ", k)
except Exception as e:
    print("no line:", e)
PY
}
run c5 --steps 6 --warmup 2
run c5lsrk54 --workload c5-lsrk54 --exchange peer-store --steps 3 --warmup 1
run c4 --workload c4 --steps 4 --warmup 2
