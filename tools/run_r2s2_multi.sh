set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -x -q > gpurun_out/r2s2_fused.log 2>&1; echo "fused rc=$?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2s2_multi.log 2>&1; echo "multi rc=$?"
tail -30 gpurun_out/r2s2_fused.log gpurun_out/r2s2_multi.log
