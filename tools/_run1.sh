python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
python - <<'PY'
import sys
sys.path.insert(0, "vision-transformer-opencl_b200")
import numpy as np, vit_b200 as V
w = V.synth_weights(224, 42)
imgs = V.synth_images(203, 224, 7)          # ragged over 4 GPUs: 51 + 51 + 51 + 50
with V.Engine(w, 224, max_batch=64, n_gpus=1) as eng:
    one = eng.forward(imgs)
for g in (2, 4):
    with V.Engine(w, 224, max_batch=64, n_gpus=g) as eng:
        out = eng.forward(imgs)
    print(g, "GPUs in one process: bit-identical to 1 GPU:", np.array_equal(one, out))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/t34_bench_g4.json 2> gpurun_out/t34_g4.err; echo rc=$?; wc -l gpurun_out/t34_bench_g4.json; head -c 400 gpurun_out/t34_bench_g4.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 4 --steps 1 --warmup 0 2>/dev/null | head -c 300; echo
