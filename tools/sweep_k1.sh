python tools/sweep_k1.py
HV_EXP_K1_ONLY=1 HV_K1_CTAS_PER_SM=4 python tools/sweep_k1.py
HV_EXP_K1_ONLY=1 HV_K1_CTAS_PER_SM=5 python tools/sweep_k1.py
HV_K1_TAIL_ROUNDS=0 python tools/sweep_k1.py
HV_K1_TAIL_ROUNDS=2 python tools/sweep_k1.py
SWEEP_COMPRESS=0 python tools/sweep_k1.py
