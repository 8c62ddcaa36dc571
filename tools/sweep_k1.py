"""Headline batch (25 x 1280x1024), steady state as in bench.py (8 rotating input batches, 2 output sets): step time with the
K1/CCL overlap, and K1 alone (events around every launch).  Environment switches (HV_K1_*) are read by the library once per
process, so run one process per variant:  HV_K1_LOOKAHEAD=2 python tools/sweep_k1.py
SWEEP_N/H/W (batch shape), SWEEP_POOL, SWEEP_OUTS, SWEEP_COMPRESS, SWEEP_MORPH=k, SWEEP_GAUSS=k,sigma, SWEEP_DEFER=1
(HV_FLAG_DEFER_TAIL); HEIMDALL_CUDA_LIB=<path> runs another build of the library (A/B inside one gpurun call)."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
n, h, w = int(os.environ.get('SWEEP_N', 25)), int(os.environ.get('SWEEP_H', 1024)), int(os.environ.get('SWEEP_W', 1280))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
st = torch.cuda.current_stream().cuda_stream
npool = int(os.environ.get('SWEEP_POOL', 8))
_distinct = min(n, 8)
_base = [synth.bottle_batch(_distinct, h, w, start_index=100 * i) for i in range(npool)]
pool = [torch.from_numpy(np.ascontiguousarray(np.resize(b, (n, h, w)))).cuda() for b in _base]
det = hc.Detector(0, defer_tail=os.environ.get('SWEEP_DEFER', '0') == '1'); det.set_stream(st)
comp = os.environ.get('SWEEP_COMPRESS', '1') == '1'
nout = int(os.environ.get('SWEEP_OUTS', str(det.pipeline_depth())))
outs = [(det.device_alloc((n, h, w), np.uint8, comp), det.device_alloc((n, h, w), np.int32, comp)) for _ in range(nout)]
print('outputs compressed:', outs[0][0].compressed, outs[0][1].compressed)
mk = int(os.environ.get('SWEEP_MORPH', '0'))
params = hc.make_params(morph_open_k=mk, morph_close_k=mk) if mk else None
if os.environ.get('SWEEP_GAUSS'):
    gk, gs = os.environ['SWEEP_GAUSS'].split(',')
    params = hc.make_params(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=int(gk), gauss_sigma=float(gs))
def step(i): det.enqueue_device(pool[i % npool].data_ptr(), n, h, w, 1, params, outs[i % nout][0].data_ptr(), outs[i % nout][1].data_ptr())
for i in range(10): step(i)
torch.cuda.synchronize()
best = 1e9; tot = 0
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / K * 1e3; best = min(best, us); tot += us
r = det.fetch_results(n)
det.profile_enable(None)
for i in range(K): step(i)
torch.cuda.synchronize()
prof = {k: round(v['ms'] / K * 1e3, 2) for k, v in det.profile().items() if v['launches']}
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith(('HV_', 'SWEEP_')))
print(f'[{tag}] step us mean {tot / 3:.2f} best {best:.2f}  frames/s {n / (tot / 3) * 1e6:,.0f}  pipeline frac {n * h * w * 6 / (tot / 3) / 1e3 / 6546.6:.3f} | alone us {prof} | defects {int(r.frames["n_defects"].sum())}')
