mkdir -p gpurun_out
# What the round-end evidence under profiles/ was produced with on a 1-GPU box: gpurun -- bash tools/gpu_validation.sh
python -m pytest tests -m gpu -x -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2final_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2final_smoke.log
python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo "bench rc=$?"
