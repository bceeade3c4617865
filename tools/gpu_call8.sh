echo "== tests"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== sweep"; timeout 300 python tools/l2four_sweep.py quick 2>&1 | cut -c1-150
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_v2.json 2> gpurun_out/r02_bench_n1_v2.err; tail -c 1000 gpurun_out/r02_bench_n1_v2.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02_bench_n1_v2.json").read().strip().splitlines()[-1])
print(l["ms_per_step"], l["value"], l["step_roofline"]["frac"], l["fft"], l["parity"])
for k in l["kernels"]: print(k)
PY
