timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" > gpurun_out/t16_attn.log 2>&1; echo "attn rc=$?"; tail -2 gpurun_out/t16_attn.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/t16_bench.json 2> gpurun_out/t16_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t16_bench.json'));print(d['value'],d['ms_per_step'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['clocks']['sm_mhz'])"
python tools/attn_trace.py 160 197 > gpurun_out/t16_trace_stream.log 2>&1; tail -1 gpurun_out/t16_trace_stream.log; sed -n 40,48p gpurun_out/t16_trace_stream.log
