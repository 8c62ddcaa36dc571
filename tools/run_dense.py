"""A few device-resident steps of the 12 MP high-contamination config (global-memory CCL path), for ncu.
usage: run_dense.py [frames]"""
import sys
import numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h, w = 3000, 4096
base = [synth.high_contamination_frame(h, w, i) for i in range(2)]
batch = np.stack([base[i % 2] for i in range(n)])
d_in = torch.from_numpy(batch).cuda()
det = hc.Detector(0, max_defects_per_frame=32768)
det.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(4):
    det.enqueue_device(d_in.data_ptr(), n, h, w)
r = det.fetch_results(n)
print('ok', int(r.frames['n_components'].sum()), int(r.frames['n_defects'].sum()))
