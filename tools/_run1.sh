python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" > gpurun_out/t4_pytest_attn.log 2>&1; echo "attn rc=$?"; tail -2 gpurun_out/t4_pytest_attn.log
VIT_ATTN_NO_PINGPONG=1 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" > gpurun_out/t4_pytest_attn_np.log 2>&1; echo "attn(no pingpong) rc=$?"; tail -2 gpurun_out/t4_pytest_attn_np.log
for v in 0 1; do
VIT_ATTN_NO_PINGPONG=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/t4_bench_np$v.json 2> gpurun_out/t4_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t4_bench_np$v.json'));print(d['value'],d['ms_per_step'],d['step_breakdown_ms']['attention'],d['clocks'])"
done
VIT_ATTN_NO_PINGPONG=1 python tools/attn_trace.py 160 197 > gpurun_out/t4_trace_np.log 2>&1
