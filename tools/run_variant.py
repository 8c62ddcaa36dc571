"""Run a few device-resident steps of one pipeline variant at 5 MP (for ncu launch lists).
usage: run_variant.py [morph K | gauss K SIGMA | box] [frames]"""
import sys
import numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
mode = sys.argv[1] if len(sys.argv) > 1 else 'box'
n = 16
h, w = 2048, 2448
base = [synth.bottle_frame(h, w, 500 + i, contaminants=i % 4) for i in range(4)]
batch = np.stack([base[i % 4] for i in range(n)])
if mode == 'morph':
    k = int(sys.argv[2]); p = hc.make_params(morph_open_k=k, morph_close_k=k)
elif mode == 'gauss':
    k = int(sys.argv[2]); s = float(sys.argv[3]); p = hc.make_params(blur_mode=hc._abi.HV_BLUR_GAUSSIAN, blur_ksize=k, gauss_sigma=s)
else:
    p = hc.make_params()
d_in = torch.from_numpy(batch).cuda()
det = hc.Detector(0)
det.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(3):
    det.enqueue_device(d_in.data_ptr(), n, h, w, 1, p)
r = det.fetch_results(n)
print('ok', int(r.frames['n_defects'].sum()))
