"""One eager forward of workload B inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_case
from egtr_b200.model.egtr import DetrForSceneGraphGeneration

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg, sd, px, mask, _ = build_case("B", batch)
model = DetrForSceneGraphGeneration(cfg)
model.load_state_dict(sd)
model.cuda().eval()
px, mask = px.cuda(), mask.cuda()
for _ in range(2):
    model(pixel_values=px, pixel_mask=mask)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(pixel_values=px, pixel_mask=mask)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one forward, batch", batch)
