# quick GPU check: parity suite + a short bench line (usage: bash tools/gpu_quick.sh <tag>)
tag=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_suite_${tag}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_suite_${tag}.txt; tail -15 gpurun_out/gpu_suite_${tag}.txt
python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -3 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}.json"))
print("value %.4g px/s  ms/step %.3f  e2e %.4g  kernels %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
PY
