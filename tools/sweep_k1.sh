SWEEP_MORPH=3 python tools/sweep_k1.py
SWEEP_MORPH=3 HV_MORPH_TILES_PER_SM=2 python tools/sweep_k1.py
SWEEP_MORPH=3 HV_MORPH_TILES_PER_SM=3 python tools/sweep_k1.py
SWEEP_MORPH=7 HV_MORPH_TILES_PER_SM=2 python tools/sweep_k1.py
