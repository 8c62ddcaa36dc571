"""Run a few device-resident steps of the hot path (for ncu captures). usage: run_steps.py [steps] [bottle|textured]"""
import sys
import numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kind = sys.argv[2] if len(sys.argv) > 2 else 'bottle'
n, h, w = 25, 1024, 1280
if kind == 'uniform':
    batch = np.full((n, h, w), 220, np.uint8)
elif kind == 'bottle':
    batch = synth.bottle_batch(n, h, w, start_index=0)
else:
    batch = np.random.default_rng(0).integers(0, 256, (n, h, w), dtype=np.uint8)
d_in = torch.from_numpy(batch).cuda()
det = hc.Detector(0)
det.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(steps):
    det.enqueue_device(d_in.data_ptr(), n, h, w)
r = det.fetch_results(n)
print('ok', int(r.frames['n_defects'].sum()))
