#!/bin/bash
# A/B timing on one box: the in-tree library against an alternative build (FFB_LIB_PATH) and against env switches.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_ab.sh'
OUT=gpurun_out
mkdir -p $OUT
ALT=$PWD/fourierflows_jl_b200/lib_ab/libfourierflows_b200.so
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee $OUT/ab_pytest_gpu.log
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "launches", d["gpu_launches"], "fft", d.get("fft", {}).get("rfft_ms"), d.get("fft", {}).get("irfft_ms"),
          "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
    for k in d.get("kernels", [])[:6]:
        print("    ", k["name"], k["us_per_launch"], k.get("frac"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
B="python bench.py --no-cpu-baseline --no-weak-ref --steps 20 --warmup 3"
echo "== bench: in-tree, multi-A on";   timeout 200 $B > $OUT/ab_main_multi1.json 2> $OUT/ab_main_multi1.err; line $OUT/ab_main_multi1.json
echo "== bench: in-tree, multi-A off";  FFB_MULTI_A=0 timeout 200 $B > $OUT/ab_main_multi0.json 2> $OUT/ab_main_multi0.err; line $OUT/ab_main_multi0.json
if [ -f "$ALT" ]; then
echo "== bench: two-phase rows build, multi-A on"; FFB_LIB_PATH=$ALT timeout 200 $B > $OUT/ab_alt_multi1.json 2> $OUT/ab_alt_multi1.err; line $OUT/ab_alt_multi1.json
fi
echo "== bench: in-tree, multi-A on (again)"; timeout 200 $B > $OUT/ab_main_multi1b.json 2> $OUT/ab_main_multi1b.err; line $OUT/ab_main_multi1b.json
echo "== bench c2"; timeout 200 python bench.py --workload c2 > $OUT/r02_bench_c2_n1.json 2> $OUT/r02_bench_c2_n1.err; cut -c1-600 $OUT/r02_bench_c2_n1.json
