"""Phase timeline of the fused decoder kernel (decoder.cu) at workload B: %globaltimer stamps of CTA 0 after every phase."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import build_case
from egtr_b200._lib import call
from egtr_b200.model.egtr import DetrForSceneGraphGeneration

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg, sd, px, mask, _ = build_case("B", batch)
model = DetrForSceneGraphGeneration(cfg)
model.load_state_dict(sd)
model.cuda().eval()
px, mask = px.cuda(), mask.cuda()
eng = model.engine()
for _ in range(2):
    eng.forward(px, mask)
torch.cuda.synchronize()
L = cfg.decoder_layers
if len(sys.argv) > 2:
    call("egtr_set_debug_flags", int(sys.argv[2]))
    print("debug flags", sys.argv[2])
buf = torch.zeros(4 + 12 * L, dtype=torch.int64, device="cuda")
call("egtr_decoder_debug_profile", buf.data_ptr())
NAMES = ["init", "qkv", "mha", "oproj", "ln1", "offaw", "msda", "outproj", "ln2", "fc1", "fc2", "ln3"]
tot = {n: 0.0 for n in NAMES}
REP = 5
span = 0.0
for _ in range(REP):
    buf.zero_()
    eng.forward(px, mask)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    prev = t[0]
    for l in range(L):
        for p, n in enumerate(NAMES):
            v = t[1 + l * 12 + p]
            if v == 0:
                continue
            tot[n] += (v - prev) / 1e3
            prev = v
    span += (prev - t[0]) / 1e3
    clk = (t[2 + 12 * L] - t[1 + 12 * L]) / max(1.0, (prev - t[0]) / 1e3)  # SM cycles per us
    bare = (t[3 + 12 * L] - t[0]) / 1e3 / 64 if t[3 + 12 * L] else 0.0
call("egtr_decoder_debug_profile", None)
print(f"decoder kernel, batch {batch}: {span / REP:.1f} us from first phase to last ({L} layers); SM clock {clk:.0f} MHz; bare phase end {bare:.2f} us")
for n in NAMES:
    print(f"  {n:<8} {tot[n] / REP / L:7.2f} us per layer   {tot[n] / REP:8.1f} us total")
