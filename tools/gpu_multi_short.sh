#!/bin/bash
# Short multi-GPU call (charged N x): NCCL tests + one bench line.  gpurun --gpus N -- bash tools/gpu_multi_short.sh N tag
N=${1:-8}; TAG=${2:-m8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_nccl.py -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
j = json.load(open("gpurun_out/${TAG}_bench.json")); print("value %.0f ms %.3f e2e %.0f sep %.0f i16 %.0f parity %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["separate_calls"]["value"], j["e2e"]["int16_ingest"]["value"], j["parity"]))
print({k: round(v["ms_per_step"], 3) for k, v in j["kernels"].items() if v["ms_per_step"] > 0})
PY
