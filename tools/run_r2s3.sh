set -x
mkdir -p gpurun_out tools/bin
nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include tools/call_overhead.cu -L pi-quant_b200/piquant -lpiquant -Xlinker -rpath=$PWD/pi-quant_b200/piquant -o tools/bin/call_overhead && tools/bin/call_overhead > gpurun_out/r2s3_call_overhead.txt 2>&1
cat gpurun_out/r2s3_call_overhead.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2s3_pytest.log
python bench.py > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2s3_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['pageable_host'])
print(json.dumps(d['extra']['small_tensor_regime'], indent=1))
"
