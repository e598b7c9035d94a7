python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t35_g2.json 2> gpurun_out/t35_g2.err; echo rc=$?; wc -l gpurun_out/t35_g2.json; python -c "
import json;d=json.load(open('gpurun_out/t35_g2.json'));print(d['n_gpus'],d['value'],d['e2e']['value'],d.get('nccl_logit_allgather_ms'),d['class_row_pruning']['value'])"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/t35_g1.json 2>/dev/null; wc -l gpurun_out/t35_g1.json
python tools/step_trend.py 100 > gpurun_out/r1_step_trend.txt 2>&1; tail -1 gpurun_out/r1_step_trend.txt
