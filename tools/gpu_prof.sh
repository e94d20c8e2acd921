# parity suite + short bench + one full ncu capture of a kernel (usage: bash tools/gpu_prof.sh <tag> <kernel-regex> [pytest-args])
tag=${1:-p}; kern=${2:-k1_glyph_tiles}; shift; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" > gpurun_out/gpu_suite_${tag}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_suite_${tag}.txt; tail -6 gpurun_out/gpu_suite_${tag}.txt
python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -3 gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}.json"))
print("value %.4g px/s  ms/step %.3f  e2e %.4g  kernels %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"]))
PY
for k in ${kern//,/ }; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${k}_${tag} \
    python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_${k}_${tag}.log 2>&1
ls -la gpurun_out/${k}_${tag}.ncu-rep
done
