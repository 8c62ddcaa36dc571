"""Phase stamps of the per-frame CCL kernel (frame 0 of the headline batch), isolated launch."""
import os, sys, numpy as np, torch
sys.path.insert(0, 'heimdall-vision_b200'); sys.path.insert(0, '.')
import heimdall_core as hc, synth
n, h, w = 25, 1024, 1280
batch = torch.from_numpy(synth.bottle_batch(n, h, w, start_index=0)).cuda()
det = hc.Detector(0, profile=True, phase_timing=True)
det.set_stream(torch.cuda.current_stream().cuda_stream)
for it in range(4): det.detect_device(batch.data_ptr(), n, h, w)
det.profile()
det.detect_device(batch.data_ptr(), n, h, w)
pr = det.profile()
print('tag', os.environ.get('HV_CCL_BIG', 'small'), {k: round(v['ms'] * 1e3, 1) for k, v in pr.items() if v['launches']})
pt = np.array(det.phase_times()[:192], dtype=np.int64).reshape(12, 16); t0 = pt[0][pt[0] > 0].min()
names = ['start', 'p0 compaction', 'p1 fetch+nodes', 'p2a hook', 'p2b jump', 'p2c edges', 'p3 rank', 'p4 labels+stats', 'p5 score']
for i in range(9):
    v = pt[i][pt[i] > 0]
    if len(v): print('  %-16s warps reach it at %.1f .. %.1f us' % (names[i], (v.min() - t0) / 1e3, (v.max() - t0) / 1e3), ' per warp:', ' '.join('%.1f' % ((x - t0) / 1e3) for x in pt[i] if x > 0))
