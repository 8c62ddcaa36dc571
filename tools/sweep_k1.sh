SWEEP_MORPH=3 python tools/sweep_k1.py
SWEEP_MORPH=3 HV_MORPH_CTA_TILES=1 python tools/sweep_k1.py
SWEEP_MORPH=7 python tools/sweep_k1.py
SWEEP_MORPH=3 HV_MORPH_WARP_CTAS_PER_SM=2 python tools/sweep_k1.py
