python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "embed" > gpurun_out/t12_ops.log 2>&1; echo "ops rc=$?"; tail -2 gpurun_out/t12_ops.log
python -m pytest tests/test_gpu_model.py -m gpu -x -q > gpurun_out/t12_model.log 2>&1; echo "model rc=$?"; tail -2 gpurun_out/t12_model.log
