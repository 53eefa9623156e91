import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, cv2
import test_gpu_config2 as t
from deepemia_b200 import engine, synthetic as syn
dev = torch.device("cuda:0")
probs, boxes, scores = t._torchvision_heads(dev)
n = probs.shape[0]
iset = engine.paste(probs, boxes, 1024, 1024)
engine.measure(iset, um_pix=0.5)
rec = iset.records.cpu().numpy(); co = iset.cont_off.cpu().numpy()
cont = engine.contours_to_host(iset)
for i in range(n):
    for j, c in enumerate(cont[i]):
        if len(c) >= 5 and cv2.contourArea(c.reshape(-1,1,2)) >= 5:
            r = rec[co[i] + j]
            e1, e2 = cv2.fitEllipse(c.reshape(-1,1,2)), cv2.fitEllipse(c.reshape(-1,1,2))
            if not np.allclose([r[0], r[1]], [e1[1][0]*0.5, e1[1][1]*0.5], rtol=1e-5):
                print(i, j, c.tolist(), "mine", r[0]/0.5, r[1]/0.5, "cv", e1, "cv again", e2)
