import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import decode as D
from oriented_object_detection_b200 import ops
dev=torch.device('cuda:0')
tiles=[(300, 40, 121, 128)]
head = D.synthetic_head(4, 12, 128, seed=128, density=0.05)[3:4]
plan = ops.plan_from_tiles(2000, 2000, tiles, device=dev)
boxes, cls, conf, count = ops.decode_tiles(torch.from_numpy(head.copy()).to(dev), plan, 128, 0.25, 0.7, 300)
torch.cuda.synchronize()
A=head.shape[2]
ws=ops._workspaces[(str(dev),'decode')].cpu().numpy().view(np.float32)
np.set_printoptions(linewidth=200, precision=4, suppress=True)
print('sorted conf field', ws[8*A:8*A+20])
print('sorted cls field', ws[9*A:9*A+20].view(np.int32))
print('sorted cx', ws[0:20])
sc=head[0,4:16]; cf=sc.max(0); cand=np.nonzero(cf>0.25)[0]; print('K',len(cand)); o=cand[np.lexsort((cand,-cf[cand].astype(np.float64)))]
print('ref sorted conf', cf[o]); print('ref cx', head[0,0,o])
k=int(count[0]); print('out', conf.cpu().numpy()[:k])
