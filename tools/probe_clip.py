#!/usr/bin/env python
"""Conditioning encoders on the device (surfd_b200/clip_encoder.py, seeded ViT-B/32-shaped checkpoint): CUDA-event time of one
image / one prompt / a batch of 8, best of 5 after warm-up, next to the same code on the host cores."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.clip_encoder import ClipEncoder

sd = synth.synth_clip(77)
dev = ClipEncoder(sd, "cuda")
img = torch.randn(8, 3, 224, 224, device="cuda")
tok = torch.zeros(8, 77, dtype=torch.int64, device="cuda"); tok[:, 0] = 49406; tok[:, 1:9] = 320; tok[:, 9] = 49407
out = {}
for name, fn in (("image_b1", lambda: dev.encode_image(img[:1])), ("image_b8", lambda: dev.encode_image(img)),
                 ("text_b1", lambda: dev.encode_text(tok[:1])), ("text_b8", lambda: dev.encode_text(tok))):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_ms"] = round(best, 3)
host = ClipEncoder(sd, "cpu")
t = time.perf_counter(); host.encode_image(img[:1].cpu()); out["host_image_b1_ms"] = round((time.perf_counter() - t) * 1e3, 1)
t = time.perf_counter(); host.encode_text(tok[:1].cpu()); out["host_text_b1_ms"] = round((time.perf_counter() - t) * 1e3, 1)
out["host_threads"] = torch.get_num_threads()
print(json.dumps(out))
