SWEEP_COMPRESS=0 python tools/sweep_k1.py
SWEEP_COMPRESS=0 HV_EXP_CCL_NOOP=1 python tools/sweep_k1.py
SWEEP_COMPRESS=0 HV_EXP_K1_ONLY=1 HV_K1_CTAS_PER_SM=4 python tools/sweep_k1.py
