python tools/sweep_k1.py
HV_CCL_THREADS=1024 python tools/sweep_k1.py
python tools/sweep_k1.py
HV_CCL_THREADS=1024 python tools/sweep_k1.py
