python tools/sweep_k1.py
HV_K1_CTAS_PER_SM=4 python tools/sweep_k1.py
HV_CCL_BIG=1 python tools/sweep_k1.py
