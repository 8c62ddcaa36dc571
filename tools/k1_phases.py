"""K1 on the headline batch with phase timing: per-tile TMA wait / processing time of flat and non-flat tiles, CTA lifetimes."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
n, h, w = 25, 1024, 1280
pool = [torch.from_numpy(synth.bottle_batch(n, h, w, start_index=1000 * i)).cuda() for i in range(8)]
st = torch.cuda.current_stream().cuda_stream
grid = 148 * int(os.environ.get('HV_K1_CTAS_PER_SM', '5'))
det = hc.Detector(0, profile=True, phase_timing=True)
det.set_stream(st)
mk = int(os.environ.get('SWEEP_MORPH', '0'))
params = hc.make_params(morph_open_k=mk, morph_close_k=mk) if mk else None
for it in range(4): det.detect_device(pool[it % 8].data_ptr(), n, h, w, 1, params)
det.profile()
res = []
for it in range(10):
    det.detect_device(pool[it % 8].data_ptr(), n, h, w, 1, params)
    k1 = det.phase_times()[200:206]; q = det.phase_times()[248:256]
    res.append((k1, q))
pr = det.profile()
print('K1 us %.1f  CCL us %.1f' % (pr['preprocess_mask']['ms'] / pr['preprocess_mask']['launches'] * 1e3,
      pr['ccl_frame_fused']['ms'] / max(pr['ccl_frame_fused']['launches'], 1) * 1e3))
for k1, q in res[-3:]:
    print('flat wait %.2f proc %.2f n=%d | nonflat wait %.2f proc %.2f n=%d' % (k1[0]/max(k1[2],1)/1e3, k1[1]/max(k1[2],1)/1e3, k1[2], k1[3]/max(k1[5],1)/1e3, k1[4]/max(k1[5],1)/1e3, k1[5]),
          '| span %.1f us, mean CTA life %.1f, max %.1f, earliest end %.1f, latest start %.1f' % ((q[1] - q[0]) / 1e3, q[2] / grid / 1e3, q[3] / 1e3, (q[4] - q[0]) / 1e3, (q[5] - q[0]) / 1e3),
          '| sum tile time per CTA %.1f us' % ((k1[0] + k1[1] + k1[3] + k1[4]) / grid / 1e3))
