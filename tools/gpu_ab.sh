#!/bin/bash
# A/B timing on one box: the in-tree library against an alternative build (FFB_LIB_PATH) and against env switches.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_ab.sh'
OUT=gpurun_out
mkdir -p $OUT
ALT=$PWD/fourierflows_jl_b200/lib_ab/libfourierflows_b200.so
echo "== pytest -m gpu (${AB_TESTS:-all})"; timeout 600 python -m pytest ${AB_TESTS:-tests} -q -m gpu -x 2>&1 | tail -6 | tee $OUT/ab_pytest_gpu.log
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
for rep in a b; do
echo "== bench: in-tree build ($rep)"; timeout 200 $B > $OUT/ab_main_$rep.json 2> $OUT/ab_main_$rep.err; line $OUT/ab_main_$rep.json
if [ -f "$ALT" ]; then
echo "== bench: alternative build fourierflows_jl_b200/lib_ab ($rep)"; FFB_LIB_PATH=$ALT timeout 200 $B > $OUT/ab_alt_$rep.json 2> $OUT/ab_alt_$rep.err; line $OUT/ab_alt_$rep.json
fi
done
if [ -n "$AB_EXTRA" ]; then echo "== $AB_EXTRA"; bash -c "$AB_EXTRA"; fi
