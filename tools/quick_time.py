import sys, time, numpy as np, torch
sys.path.insert(0,'heimdall-vision_b200'); sys.path.insert(0,'.')
import heimdall_core as hc, synth
det = hc.Detector(0, profile=True, phase_timing=True)
n,h,w=25,1024,1280
batch = synth.bottle_batch(n,h,w,start_index=0)
d_in = torch.from_numpy(batch).cuda()
det.set_stream(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    r = det.detect_device(d_in.data_ptr(), n,h,w)
det.profile()
r = det.detect_device(d_in.data_ptr(), n,h,w)
print('profile ms', {k:round(v['ms'],4) for k,v in det.profile().items()}, 'defects', int(r.frames['n_defects'].sum()))
pt=np.array(det.phase_times(),dtype=np.int64).reshape(16,16); t0=pt[0].min()
for i in range(8): print('stamp',i,'warps min/max us', (pt[i].min()-t0)/1e3, (pt[i].max()-t0)/1e3, 'argmax', int(pt[i].argmax()))
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
det2 = hc.Detector(0)
det2.set_stream(torch.cuda.current_stream().cuda_stream)
for it in range(3): det2.enqueue_device(d_in.data_ptr(), n,h,w)
torch.cuda.synchronize()
e0.record()
K=20
for it in range(K): det2.enqueue_device(d_in.data_ptr(), n,h,w)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/K
print('ms/step', ms, 'frames/s', n/ms*1e3, 'GB/s alg', n*h*w*6/ms/1e6)
rng=np.random.default_rng(0)
tex = torch.from_numpy(rng.integers(0,256,(n,h,w),dtype=np.uint8)).cuda()
for it in range(2): r = det.detect_device(tex.data_ptr(), n,h,w)
det.profile()
r = det.detect_device(tex.data_ptr(), n,h,w)
print('textured profile ms', {k:round(v['ms'],4) for k,v in det.profile().items()}, int(r.frames['n_components'].sum()))
