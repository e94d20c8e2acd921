# compute-sanitizer logs for profiles/ (usage: bash tools/gpu_sanitize.sh <tag>)
tag=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/_gpu_sanitize.py > gpurun_out/sanitizer_${tool}_${tag}.txt 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_${tool}_${tag}.txt) summary line(s)"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|done " gpurun_out/sanitizer_${tool}_${tag}.txt | head -5
done
