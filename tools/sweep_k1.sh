python tools/sweep_k1.py
HV_PIPELINE_DEPTH=8 python tools/sweep_k1.py
HV_PIPELINE_DEPTH=4 python tools/sweep_k1.py
HV_PIPELINE_DEPTH=3 python tools/sweep_k1.py
HV_PIPELINE_DEPTH=8 HV_K1_CTAS_PER_SM=3 python tools/sweep_k1.py
