mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s3_multi.log 2>&1; echo "multi rc=$?"
tail -n 5 gpurun_out/r2s3_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_probe.py 28 24 > gpurun_out/r2s3_allreduce_probe_n2.txt 2>&1
grep -v "^\*\|OMP" gpurun_out/r2s3_allreduce_probe_n2.txt | tail -n 40
