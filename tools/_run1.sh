python -m pytest tests -m gpu -x -q > gpurun_out/t29_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t29_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/t29_bench.json 2> gpurun_out/t29_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t29_bench.json'));print(d['value'],d['ms_per_step'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['e2e']['value'],d['batch1_latency'],d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/t29_ncu.log 2>&1; echo "ncu rc=$?"
