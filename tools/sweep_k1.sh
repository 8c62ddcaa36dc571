SWEEP_OUTS=5 python tools/sweep_k1.py
SWEEP_OUTS=5 HV_CCL_BIG=1 python tools/sweep_k1.py
SWEEP_OUTS=5 python tools/sweep_k1.py
