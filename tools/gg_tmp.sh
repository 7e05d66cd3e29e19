export MVB200_NO_BUILD=1
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $L --master-port 29601 tests/mp_device_check.py > gpurun_out/mp_check.log 2>&1; echo "mp rc=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/mp_check.log | cut -c1-100 | head -4; grep -o '"aggregate_[a-z_]*": [a-z]*' gpurun_out/mp_check.log | sort | uniq -c
bash tools/gpu_suite.sh 2 allreduce
