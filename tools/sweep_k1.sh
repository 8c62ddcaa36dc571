HV_K1_CTAS_PER_SM=5 python tools/sweep_k1.py
HV_K1_CTAS_PER_SM=4 python tools/sweep_k1.py
HV_K1_CTAS_PER_SM=5 python tools/sweep_k1.py
HV_K1_CTAS_PER_SM=4 python tools/sweep_k1.py
HV_K1_CTAS_PER_SM=5 HV_K1_STATIC=1 python tools/sweep_k1.py
HV_K1_CTAS_PER_SM=3 python tools/sweep_k1.py
