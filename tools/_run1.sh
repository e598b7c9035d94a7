timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/t19_model.log 2>&1; echo "model rc=$?"; tail -3 gpurun_out/t19_model.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/t19_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t19_bench.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['batch1_latency'])"
VIT_GRAPHS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t19_bench_nog.json 2> gpurun_out/t19_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t19_bench_nog.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['batch1_latency'])"
