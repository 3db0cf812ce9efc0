#!/bin/bash
# Multi-GPU call: gpurun --gpus N --timeout 1500 -- bash tools/gpu_multi.sh N tag [pytest]
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/${TAG}_topo.txt 2>&1
if [ "$3" = "pytest" ]; then
  timeout 900 python -m pytest tests/test_dist_nccl.py tests/test_gpu_parity.py -x -q -m gpu -k "nccl or another_device" > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
fi
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/h2d_bw.py > gpurun_out/${TAG}_h2d.json 2> gpurun_out/${TAG}_h2d.err; echo "h2d rc=$?"
timeout 300 $RUN tools/h2d_bw.py --bind > gpurun_out/${TAG}_h2d_bind.json 2> gpurun_out/${TAG}_h2d_bind.err; echo "h2d bind rc=$?"
FSEM_BIND_NUMA=0 timeout 900 $RUN bench.py --gpus $N --no-cpu --steps 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
FSEM_BIND_NUMA=1 timeout 900 $RUN bench.py --gpus $N --no-cpu > gpurun_out/${TAG}_bench_bind.json 2> gpurun_out/${TAG}_bench_bind.err; echo "bench bind rc=$?"
python - <<PY
import json
for f in ("h2d", "h2d_bind"):
    try:
        j = json.loads([l for l in open("gpurun_out/${TAG}_%s.json" % f) if l.startswith("H2D_JSON ")][0][9:]); print(f, j["aggregate_concurrent_gbs"], [ (r["solo_gbs"], r["concurrent_gbs"], r["gpu_numa_node"]) for r in j["ranks"]])
    except Exception as e: print(f, "failed", e)
for f in ("bench", "bench_bind"):
    try:
        j = json.load(open("gpurun_out/${TAG}_%s.json" % f)); print(f, "value %.0f ms %.3f e2e %.0f sep %.0f i16 %.0f parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["separate_calls"]["value"], j["e2e"]["int16_ingest"]["value"], j["parity"]["ok"]))
    except Exception as e: print(f, "failed", e)
PY
