import sys, os, numpy as np, torch
sys.path.insert(0, '.')
from metro_pose3d_b200.inference import MetroModel
n = 160
torch.manual_seed(0)
img = torch.rand((n, 256, 256, 3), dtype=torch.float32)
m = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=n)
a1 = m.infer(img.cuda()).cpu().numpy()
a2 = m.infer(img.cuda()).cpu().numpy()
b1 = m.infer_host(img.numpy()).copy()
b2 = m.infer_host(img.numpy()).copy()
a3 = m.infer(img.cuda()).cpu().numpy()
k = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=n, keep_activations=True)
c = k.infer(img.cuda()).cpu().numpy()
def d(x, y, name):
    bad = np.where(np.abs(x - y).reshape(n, -1).max(1) > 0)[0]
    print(name, 'crops differing:', len(bad), bad[:20], 'max', np.abs(x - y).max())
d(a1, a2, 'dev vs dev'); d(a1, a3, 'dev vs dev(after host)'); d(b1, b2, 'host vs host'); d(a1, b1, 'dev vs host'); d(a1, c, 'dev vs keep'); d(b1, c, 'host vs keep')
