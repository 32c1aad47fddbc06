set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2s2_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2s2_bench_ref.json 2> gpurun_out/r2s2_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2s2_launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r2s2_bench_under_ncu.log 2>&1
for v in 1 2; do
 ncu --set full --clock-control none --import-source on -k regex:'quant' -o gpurun_out/r2_bf16u4_v${v}_1e9 -f python tools/ncu_cells.py --numel 1000000000 --launches 2 --cells bf16u4n --variant $v > gpurun_out/r2s2_ncu_v${v}.log 2>&1
 ncu --set full --clock-control none --import-source on -k regex:'quant' -o gpurun_out/r2_bf16u4_v${v}_27M -f python tools/ncu_cells.py --numel 27264000 --launches 2 --cells bf16u4n --variant $v >> gpurun_out/r2s2_ncu_v${v}.log 2>&1
done
ls -la gpurun_out | tail -20
