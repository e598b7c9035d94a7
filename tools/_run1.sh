python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "linear or folded or residual" > gpurun_out/t10_ops.log 2>&1; echo "ops rc=$?"; tail -2 gpurun_out/t10_ops.log
for cfg in "1 0" "1 4" "0 0" "0 5"; do set -- $cfg
VIT_LN_FUSED=$1 VIT_RES_CFG=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/t10_bench_$1_$2.json 2> gpurun_out/t10_bench.err; echo "bench fused=$1 cfg=$2 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/t10_bench_$1_$2.json'));print(d['value'],d['ms_per_step'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['clocks']['sm_mhz'])"
done
