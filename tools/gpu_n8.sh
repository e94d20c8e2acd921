# 8-GPU bench: default line (all BASELINE configs) with and without NUMA binding, and with staged D2H instead of zero-copy stores
mkdir -p gpurun_out
N=${1:-8}
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n${N}_${tag}.json 2> gpurun_out/bench_r2_n${N}_${tag}.err; tail -2 gpurun_out/bench_r2_n${N}_${tag}.err | cut -c1-300; python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2_n${N}_${tag}.json"))
print("${tag}: value %.4g e2e %.4g (%.3f ms)  binding=%s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("host_binding")))
for k,v in d.get("configs",{}).items(): print("   %-10s value %.4g  ms %.3f  e2e %.4g  %s" % (k, v["value"], v["ms_per_step"], v["e2e"]["value"], v.get("broadcast_ms","")))
PY
}
run numa A=1
run nonuma B200MSDF_NO_NUMA_BIND=1
run staged B200MSDF_NO_ZEROCOPY=1
nvidia-smi topo -m > gpurun_out/topo_n${N}.txt 2>&1
