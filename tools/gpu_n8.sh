# N-GPU bench: default line (all BASELINE configs), with the NUMA binding and (COMPARE=1) without it / with staged D2H
mkdir -p gpurun_out
N=${1:-8}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n${N}_${tag}.json 2> gpurun_out/bench_r2_n${N}_${tag}.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2_n${N}_${tag}.json").read().splitlines() if l.startswith("{")][0])
print("${tag}: value %.4g (%.3f ms) e2e %.4g (%.3f ms)  binding=%s clocks=%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("host_binding"), d["clocks"]))
print("   per-rank k1/k2 ms:", d["roofline"].get("kernel_ms_per_rank"))
for k,v in d.get("configs",{}).items(): print("   %-10s value %.4g  ms %.3f  e2e %.4g  %s" % (k, v["value"], v["ms_per_step"], v["e2e"]["value"], v.get("broadcast_ms","")))
PY
}
run numa A=1
if [ -n "$COMPARE" ]; then run nonuma B200MSDF_NO_NUMA_BIND=1; fi
