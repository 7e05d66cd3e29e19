export MVB200_NO_BUILD=1
NG=8
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 400 $L --master-port 29601 tests/mp_device_check.py > gpurun_out/mp_check_n8.log 2>&1; echo "mp rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/mp_check_n8.log | cut -c1-60 | head -10
timeout 400 $L --master-port 29602 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
bash tools/gpu_suite.sh 8 allreduce replica
