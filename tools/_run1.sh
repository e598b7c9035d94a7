timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/t31_model.log 2>&1; echo "model rc=$?"; tail -4 gpurun_out/t31_model.log; grep "pruned vs full" gpurun_out/t31_model.log | head
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/t31_bench.json 2> gpurun_out/t31_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/t31_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t31_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['class_row_pruning'],d['batch1_latency'])"
