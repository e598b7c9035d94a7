timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t21_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t21_pytest.log
for v in 1 0; do
VIT_PDL=$v timeout 300 python tools/latency_breakdown.py 1 2>&1 | tail -3
VIT_PDL=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t21_bench_$v.json 2> gpurun_out/t21_bench.err; echo "bench pdl=$v rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t21_bench_$v.json'));print(d['value'],d['ms_per_step'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['e2e']['value'],d['batch1_latency'])"
done
