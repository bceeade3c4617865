#!/bin/bash
# One gpurun call that produces the evidence of a round: GPU test-suite, the N = 1 bench line, the ncu launch list of the
# bench command and `ncu --set full` summaries of one step and of the Float32 3-D transform passes.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_checkpoint.sh r02'
# Everything lands in gpurun_out/<round>_*; copy what should be judged into profiles/.
# (Numbers printed by the runs under ncu are never bench values.)
# Second argument `lite`: skip the parts that do not depend on the latest kernel changes (Float32 3-D ncu capture, reference arm,
# cuFFT times).
R=${1:-rXX}
LITE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $OUT/${R}_pytest_gpu.log
echo "== bench N=1"; timeout 300 python bench.py > $OUT/${R}_bench_n1.json 2> $OUT/${R}_bench_n1.err
python - <<PY
import json
d = json.loads(open("$OUT/${R}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "steps_per_s", "ms_per_step", "gpu_launches")}, "roofline", round(d["step_roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 3), d["clocks"])
for k in d["kernels"][:8]:
    print("  ", k)
PY
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${R}_launches_raw.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summarize.py launches $OUT/${R}_launches_raw.csv > $OUT/${R}_ncu_launches_n1.csv && head -12 $OUT/${R}_ncu_launches_n1.csv
echo "== ncu --set full: one ETDRK4 step at 8192^2 F64 (fused calcN)"
# skip the problem set-up (3 transform kernels) and the first step (56 launches, cold tables); capture the second step's 56 kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fft_pow2_kernel|stage_kernel|fs_pass' --launch-skip 59 -c 56 \
  -o $OUT/${R}_step python tools/run_step_once.py 8192 2 > /dev/null 2>&1
ncu -i $OUT/${R}_step.ncu-rep --page raw --csv > $OUT/${R}_step_raw.csv 2>/dev/null
python tools/ncu_summarize.py full $OUT/${R}_step_raw.csv > $OUT/${R}_ncu_full_step_kernels.csv && head -8 $OUT/${R}_ncu_full_step_kernels.csv
rm -f $OUT/${R}_step.ncu-rep          # > 64 MiB reports are not copied back; the CSV exports are
if [ "$LITE" != "lite" ]; then
echo "== ncu --set full: Float32 3-D r2c passes (per-GPU share of C5)"
timeout 400 ncu --set full --clock-control none -k regex:'fft_pow2_kernel|fs_pass' -c 6 -o $OUT/${R}_fft3d python tools/run_fft_once.py 2048x2048x256 f32 1 > /dev/null 2>&1
ncu -i $OUT/${R}_fft3d.ncu-rep --page raw --csv > $OUT/${R}_fft3d_raw.csv 2>/dev/null
python tools/ncu_summarize.py full $OUT/${R}_fft3d_raw.csv > $OUT/${R}_ncu_full_fft3d_f32_kernels.csv && cat $OUT/${R}_ncu_full_fft3d_f32_kernels.csv
rm -f $OUT/${R}_fft3d.ncu-rep
echo "== reference arm (driver flags)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${R}_bench_reference_n1.json 2> $OUT/${R}_bench_reference_n1.err; cut -c1-400 $OUT/${R}_bench_reference_n1.json
echo "== cuFFT reference times (torch.fft on the same box)"; timeout 200 python tools/cufft_times.py | tee $OUT/${R}_cufft_times.log
fi
echo "== transform sweep"; timeout 300 python tools/gpu_sweep.py > $OUT/${R}_fft_sweep.log 2>&1; grep -E "FAILURES|^time" $OUT/${R}_fft_sweep.log
