timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" > gpurun_out/t24_attn.log 2>&1; echo "attn rc=$?"; tail -3 gpurun_out/t24_attn.log
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "384" > gpurun_out/t24_model.log 2>&1; echo "model384 rc=$?"; tail -3 gpurun_out/t24_model.log
for v in 3 2; do
VIT_ATTN_IMPL=$v timeout 600 python bench.py --img-size 384 --batch 512 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/t24_bench384_$v.json 2> gpurun_out/t24_bench.err; echo "bench384 impl=$v rc=$?"; tail -1 gpurun_out/t24_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t24_bench384_$v.json'));print(d['value'],d['ms_per_step'],d['model_frac_of_peak'],{k:round(v,2) for k,v in d['step_breakdown_ms'].items()},d['e2e']['value'])"
done
