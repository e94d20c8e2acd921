# parity suite on the main build, then the bench kernel times of the main build and of every variant in msdfgen_b200/variants/
# usage: bash tools/gpu_ab.sh <tag> [ncu-kernel-regex]
tag=${1:-ab}; kern=$2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_suite_${tag}.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_suite_${tag}.txt; tail -4 gpurun_out/gpu_suite_${tag}.txt
shopt -s nullglob
for lib in main msdfgen_b200/variants/*.so; do
  name=$(basename $lib .so)
  if [ "$lib" = main ]; then unset B200MSDF_LIB; else export B200MSDF_LIB=$PWD/$lib; fi
  python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_${name}.json 2> gpurun_out/bench_${tag}_${name}.err || tail -3 gpurun_out/bench_${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_${name}.json"))
    print("%-28s value %.4g px/s  ms/step %.3f  e2e %.4g  kernels %s" % ("${name}", d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}))
except Exception as e: print("${name}", "failed", e)
PY
done
unset B200MSDF_LIB
for v in $AB_ENV; do          # e.g. AB_ENV="B200MSDF_K1_ROUND=512 B200MSDF_K1_ROUND=100000"
  env $v python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_${tag}_env.json 2> gpurun_out/bench_${tag}_env.err || tail -3 gpurun_out/bench_${tag}_env.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_env.json"))
print("%-28s value %.4g px/s  ms/step %.3f  e2e %.4g  kernels %s" % ("$v", d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}))
PY
done
if [ -n "$kern" ]; then
for k in ${kern//,/ }; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${k}_${tag} \
    python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 > gpurun_out/ncu_${k}_${tag}.log 2>&1
ls -la gpurun_out/${k}_${tag}.ncu-rep
done
fi
