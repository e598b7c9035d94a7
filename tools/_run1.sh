python tools/determinism_probe.py 150 1024 bf16 | tail -3
python tools/determinism_probe.py 100 777 bf16 | tail -3
python tools/determinism_probe.py 100 1024 fp16 | tail -3
VIT_LN_FUSED=0 python tools/determinism_probe.py 40 1024 bf16 | tail -3
VIT_PRUNE_LAST=0 python tools/determinism_probe.py 60 1024 bf16 | tail -3
python tools/determinism_probe.py 200 64 bf16 | tail -3
