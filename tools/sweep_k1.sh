python tools/sweep_k1.py
HV_NO_CCL_QUEUE=1 python tools/sweep_k1.py
python tools/sweep_k1.py
HV_NO_CCL_QUEUE=1 python tools/sweep_k1.py
