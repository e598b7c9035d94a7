L=vision-transformer-opencl_b200/lib
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== GELU probe, product build (tanh.approx)"; python tools/gelu_probe.py
cp $L/libvit_b200.so /tmp/new.so; cp $L/libvit_b200_ex2rcp.so $L/libvit_b200.so
echo "== GELU probe, -DVIT_GELU_EX2RCP build (ex2 + rcp)"; python tools/gelu_probe.py
echo ex2rcp $(python tools/ab_step.py 20 3)
cp /tmp/new.so $L/libvit_b200.so
echo product $(python tools/ab_step.py 20 3)
cp $L/libvit_b200_ex2rcp.so $L/libvit_b200.so; echo ex2rcp $(python tools/ab_step.py 20 3)
cp /tmp/new.so $L/libvit_b200.so; echo product $(python tools/ab_step.py 20 3)
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
python tools/profile_one.py 1024 2 && ncu --set full --clock-control none --import-source on -k 'regex:gemm_sm100_staged|attention_sm100' -s 62 -c 5 -f -o gpurun_out/r2_layer_final python tools/profile_one.py 1024 2 > gpurun_out/ncu_layer.log 2>&1
python bench.py --steps 2 --warmup 3 --no-variants --no-inproc > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-variants --no-inproc > gpurun_out/ncu_launch.log 2>&1
echo done
