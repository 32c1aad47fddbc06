mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2final_multi.log 2>&1; echo "multi rc=$?"
tail -n 3 gpurun_out/r2final_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2final_bench_n2.json 2> gpurun_out/r2final_bench_n2.err; echo "bench rc=$?"
