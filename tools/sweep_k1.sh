python tools/sweep_k1.py
HV_NO_EARLY_K1=1 python tools/sweep_k1.py
python tools/sweep_k1.py
HV_NO_EARLY_K1=1 python tools/sweep_k1.py
