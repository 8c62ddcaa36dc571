HV_K1_CLAIM_AHEAD=1 python tools/sweep_k1.py
HV_K1_CLAIM_AHEAD=0 python tools/sweep_k1.py
HV_K1_CLAIM_AHEAD=1 python tools/sweep_k1.py
HV_K1_CLAIM_AHEAD=0 python tools/sweep_k1.py
