mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/r2s3_fused.log 2>&1; echo "fused rc=$?"; tail -n 3 gpurun_out/r2s3_fused.log
python tools/sumbench.py > gpurun_out/r2s3_sumbench_v3.txt 2>&1; cat gpurun_out/r2s3_sumbench_v3.txt
