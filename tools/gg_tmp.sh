export MVB200_NO_BUILD=1
NG=8
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 400 $L --master-port 29601 tests/mp_device_check.py > gpurun_out/mp_check_n8.log 2>&1; echo "mp rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/mp_check_n8.log | cut -c1-200 | head -10
timeout 400 $L --master-port 29602 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench pipelined rc=$?"; cut -c1-2600 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
timeout 400 $L --master-port 29603 bench.py --gpus $NG --steps 10 --warmup 3 --no-pipeline --no-table-bw > gpurun_out/bench_n8_nopipe.json 2> gpurun_out/bench_n8_nopipe.err; echo "bench nopipe rc=$?"; cut -c1-1800 gpurun_out/bench_n8_nopipe.json
timeout 300 $L --master-port 29604 bench/get_gemm.py > gpurun_out/get_gemm_n8.log 2>&1; echo "gemm rc=$?"; grep '^\[' gpurun_out/get_gemm_n8.log | python -c "
import sys,json
for l in sys.stdin:
    for r in json.loads(l): print(r['M'],r['N'],r['K'],'fused %.3f ms %.0f TF unfused %.3f get %.3f'%(r['fused_ms'],r['fused_tflops'],r['unfused_get_plus_cublas_ms'],r['get_only_ms']))
"
