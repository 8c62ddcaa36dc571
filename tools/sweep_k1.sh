# One gpurun call, one process per variant (the library reads its HV_* switches once per process; boxes differ by ~4 %,
# so only numbers from the same call are comparable).  Edit to taste.
python tools/sweep_k1.py
HV_EXP_CCL_NOOP=1 python tools/sweep_k1.py
HV_EXP_K1_ONLY=1 HV_K1_CTAS_PER_SM=4 python tools/sweep_k1.py
HV_CCL_BIG=1 python tools/sweep_k1.py
SWEEP_MORPH=3 python tools/sweep_k1.py
SWEEP_MORPH=3 HV_NO_MORPH_CHAIN=1 python tools/sweep_k1.py
