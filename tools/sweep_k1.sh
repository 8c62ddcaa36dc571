SWEEP_MORPH=3 python tools/sweep_k1.py
SWEEP_MORPH=7 python tools/sweep_k1.py
