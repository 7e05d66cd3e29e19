export MVB200_NO_BUILD=1
NG=4
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 400 $L --master-port 29602 bench.py --gpus $NG --steps 10 --warmup 3 --no-table-bw > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench rc=$?"; python - <<PY
import json
for l in open('gpurun_out/bench_n4.json'):
    if l.startswith('{'):
        d=json.loads(l); print(' device ms %.2f  e2e ms %.2f'%(d['ms_per_step'], d['e2e']['ms_per_step'])); print(d['extra']['e2e_trace_rank0']); print(d['extra']['monitors_device_arm'])
PY
tail -3 gpurun_out/bench_n4.err
