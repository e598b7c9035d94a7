echo default; python tools/determinism_probe.py 100 | grep -v ": 0 images" | head -20; echo "--"
echo default again; python tools/determinism_probe.py 100 | grep -v ": 0 images" | head -20; echo "--"
