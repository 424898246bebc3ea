# 2-GPU comparison of the gradient-exchange schedules through bench.py (CUDA-graph replay), gpurun --gpus 2
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 15 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', {k:d[k] for k in ('value','ms_per_step')}, d.get('clocks'), flush=True)"; }
LGB200_EXCHANGE=end run end_sync
LGB200_EXCHANGE=none run none
LGB200_EXCHANGE=chunked run chunked
LGB200_EXCHANGE=end run end_sync_again
