timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "golden or full_size" > gpurun_out/t30_model.log 2>&1; echo "model rc=$?"; tail -12 gpurun_out/t30_model.log
