import numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import maskrcnn_b200 as m
cfg = m.MaskRCNNConfig(); cfg.maxBatch = 4
_, blobs = m.weights.synthetic_blobs(101)
model = m.MaskRCNN(cfg, blobs=blobs, anchors=m.synth.generate_anchors(1024,1024))
rng = np.random.default_rng(20260)
img = rng.integers(0,256,(4,1024,1024,3),dtype=np.uint8)
det, masks = model.prediction_batch(img)
print('detections per image', [(d[:,5]>0).sum() for d in det], 'classes', [np.unique(d[d[:,5]>0,4]).size for d in det])
print('score range', det[...,5].max())
