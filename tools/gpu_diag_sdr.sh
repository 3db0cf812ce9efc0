python tools/diag_sdr.py
bash tools/gpu_sdr.sh 2>&1 | grep -v "^shard"
