timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "odd_image or empty" > gpurun_out/t28_model.log 2>&1; echo "model rc=$?"; tail -12 gpurun_out/t28_model.log
