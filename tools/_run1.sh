timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -x -q -k "model or head" > gpurun_out/t22_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t22_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t22_bench.json 2> gpurun_out/t22_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t22_bench.json'));print(d['value'],d['ms_per_step'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['e2e']['value'],d['batch1_latency'])"
python tools/latency_breakdown.py 1 2>&1 | tail -12
