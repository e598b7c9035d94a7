L=vision-transformer-opencl_b200/lib
cp $L/libvit_b200.so /tmp/new.so
for i in 1 2 3; do
for v in new tanh; do
if [ $v = tanh ]; then cp $L/libvit_b200_tanh.so $L/libvit_b200.so; else cp /tmp/new.so $L/libvit_b200.so; fi
echo $v $(timeout 300 python tools/ab_step.py 20 3)
done; done
cp $L/libvit_b200_tanh.so $L/libvit_b200.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-variants --no-inproc > gpurun_out/ab_tanh.json 2> gpurun_out/ab_tanh.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ab_tanh.json').read().strip().splitlines()[-1])
print('tanh', round(d['value'],1), round(d['ms_per_step'],3), d['parity']['max_abs_dlogit'], d['parity']['mean_abs_dlogit'], d['parity']['top1_equal'], {k:round(v,2) for k,v in d['step_breakdown_ms'].items()})
PY
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k gelu 2>&1 | tail -5
cp /tmp/new.so $L/libvit_b200.so
