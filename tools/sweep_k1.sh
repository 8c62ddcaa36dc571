SWEEP_COMPRESS=0 python tools/sweep_k1.py
SWEEP_COMPRESS=1 python tools/sweep_k1.py
SWEEP_COMPRESS=1 HV_K1_LOOKAHEAD=3 HV_K1_TAIL_ROUNDS=0 python tools/sweep_k1.py
SWEEP_COMPRESS=1 HV_K1_CTAS_PER_SM=3 python tools/sweep_k1.py
