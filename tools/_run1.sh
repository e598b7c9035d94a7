python -m pytest tests -m gpu -x -q > gpurun_out/t17_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t17_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/t17_bench.json 2> gpurun_out/t17_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t17_bench.json'));print(d['value'],d['ms_per_step'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['e2e'],d['clocks'],d['batch1_latency'],d['roofline'],d['cpu_baseline'])"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/t17_bench_ref.json 2>gpurun_out/t17_ref.err; cat gpurun_out/t17_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t17_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/t17_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_b1024_final.csv python tools/profile_one.py 1024 > gpurun_out/t17_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_sm100_staged|attention_sm100" -s 5 -c 5 -o gpurun_out/t17_layer python tools/profile_one.py 1024 > gpurun_out/t17_ncu2.log 2>&1; tail -1 gpurun_out/t17_ncu2.log
