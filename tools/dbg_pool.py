import sys, os, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools'); sys.path.insert(0,'/root/repo/tests')
import cv2, tempfile
from conftest import make_model
from oracle import segnet_oracle as S
from sivo_b200 import BayesianSegNet, BayesianSegNetParams
img0=cv2.imread('/root/repo/tests/golden/kitti_000000_1242x375.png')
img=np.ascontiguousarray(img0[100:164,300:428])
tmp=tempfile.mkdtemp()
net,w,proto,model=make_model(tmp,'basic',T=3,H=64,W=128,seed=0)
for engine in ('simt','auto'):
    seg=BayesianSegNet(BayesianSegNetParams(proto,model),seed=1234,precision='fp16',engine=engine,keep_blobs=True)
    seg.set_frame(0); seg.segmentImage(img)
    masks={ly.tops[1]: seg.blob(ly.tops[1]).astype(np.int64) for ly in net.layers if ly.type=='Pooling'}
    prob,blobs=S.forward(net,w,img,seed=1234,frame=0,precision='fp16',return_blobs=True,masks=masks)
    for name in ('norm','conv1','pool1','conv2','pool2'):
        ref=blobs[name].numpy(); got=seg.blob(name)
        d=np.abs(ref-got); print(engine,name,'max err',d.max(),'at',np.unravel_index(d.argmax(),d.shape),'ref',ref.flat[d.argmax()],'got',got.flat[d.argmax()],'scale',np.abs(ref).max())
    x=blobs['conv2']; _,own=S.pool_with_mask(x); m=torch.from_numpy(masks['pool2_mask'])
    diff=(m!=own); print('pool2 differing',int(diff.sum()))
    idx=diff.nonzero()[:5]
    for n,c,h,ww in idx.tolist():
        win=x[n,c,2*h:2*h+2,2*ww:2*ww+2]; dwin=seg.blob('conv2')[n,c,2*h:2*h+2,2*ww:2*ww+2]
        print(' at',(n,c,h,ww),'oracle win',win.flatten().tolist(),'dev win',dwin.flatten().tolist(),'own',int(own[n,c,h,ww]),'dev',int(m[n,c,h,ww]), 'expected idx base',(2*h)*64+2*ww)
