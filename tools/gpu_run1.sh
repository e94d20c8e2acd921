mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_suite_r2a.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_suite_r2a.txt; tail -5 gpurun_out/gpu_suite_r2a.txt
python bench.py --impl reference > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err; cat gpurun_out/bench_r2a_ref.json | cut -c1-600
python bench.py > gpurun_out/bench_r2a_n1.json 2> gpurun_out/bench_r2a_n1.err; tail -3 gpurun_out/bench_r2a_n1.err; cat gpurun_out/bench_r2a_n1.json
bash profiles/capture.sh r2a > gpurun_out/capture_r2a.log 2>&1; tail -8 gpurun_out/capture_r2a.log
