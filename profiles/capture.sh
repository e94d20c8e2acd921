#!/bin/bash
# Run on the GPU box (gpurun): ncu launch list of bench.py + one full capture per hot kernel. Outputs under gpurun_out/.
# Usage: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/capture.sh r2'
tag=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    $B --steps 2 --warmup 3 > gpurun_out/bench_under_ncu_${tag}.log 2>&1
for k in k1a_filter k1b_exact k2_classify k3_exact_distance; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${k}_${tag} \
      $B --steps 1 --warmup 3 > gpurun_out/ncu_${k}_${tag}.log 2>&1
done
# big shapes (BASELINE configs[3]): one launch of the block-list kernel, sdf 2048x2048 of the 10k-segment shape, overlapping combiner
SIZE=2048 MODE=1 ncu --set full --clock-control none --import-source on -k regex:k1_big_blocks -s 1 -c 1 -o gpurun_out/k1_big_blocks_${tag} \
    python tools/_gpu_big_one.py > gpurun_out/ncu_k1_big_blocks_${tag}.log 2>&1
# summarise on the box (the .ncu-rep files are ~25 MB each; gpurun copies back at most 64 MiB) and keep only the texts
SUMMARY_OUT=gpurun_out/summary_${tag} python profiles/summarize.py ${tag}
python tools/ncu_lines.py gpurun_out/k1b_exact_${tag}.ncu-rep 40 > gpurun_out/summary_${tag}/k1b_exact_${tag}_lines.txt 2>&1
python tools/ncu_lines.py gpurun_out/k2_classify_${tag}.ncu-rep 40 > gpurun_out/summary_${tag}/k2_classify_${tag}_lines.txt 2>&1
rm -f gpurun_out/*_${tag}.ncu-rep
ls -la gpurun_out/summary_${tag}
