"""Dev experiment: how much end-to-end error do candidate tensor-core operand formats add?
Emulates operand rounding inside every linear/conv of the oracle and compares with plain fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from egtr_b200.config import WORKLOADS, workload_config
from egtr_b200.synth import synth_images, synth_state_dict
from oracle import egtr_oracle as orc

def rn_tf32(x):
    i = x.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF   # round-half-up on magnitude to 10 mantissa bits
    return i.view(torch.float32)
def tr_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
def bf(x): return x.bfloat16().float()
def split(x):
    h = bf(x); return h, bf(x - h)

MODES = {
 "bf16x1": lambda a, w, op: op(bf(a), bf(w)),
 "tf32_rn": lambda a, w, op: op(rn_tf32(a), rn_tf32(w)),
 "tf32_trunc_act": lambda a, w, op: op(tr_tf32(a), rn_tf32(w)),
 "bf16x3": lambda a, w, op: (lambda ah, al, wh, wl: op(ah, wh) + op(ah, wl) + op(al, wh))(*split(a), *split(w)),
 "bf16x2_wsingle": lambda a, w, op: (lambda ah, al: op(ah, bf(w)) + op(al, bf(w)))(*split(a)),
}

def main(wl="A"):
    cfg = workload_config(wl); H, W = WORKLOADS[wl]["image"]
    sd = synth_state_dict(cfg, 30); px, m = synth_images(1, H, W, 31)
    base = orc.forward(sd, cfg, px, m)
    lin0, conv0 = F.linear, F.conv2d
    for name, f in MODES.items():
        def lin(x, w, b=None): 
            y = f(x, w, lambda a, ww: lin0(a, ww))
            return y if b is None else y + b
        def conv(x, w, b=None, **kw):
            y = f(x, w, lambda a, ww: conv0(a, ww, **kw))
            return y if b is None else y + b.view(1, -1, 1, 1)
        orc.F.linear, orc.F.conv2d = lin, conv
        try:
            out = orc.forward(sd, cfg, px, m)
        finally:
            orc.F.linear, orc.F.conv2d = lin0, conv0
        flips = (out["logits"].argmax(-1) != base["logits"].argmax(-1)).sum().item()
        s = " ".join(f"{k}={float((out[k]-base[k]).abs().max()/base[k].abs().max()):.2e}" for k in
                     ("encoder_last_hidden_state","last_hidden_state","logits","pred_boxes","pred_rel","pred_connectivity"))
        print(f"{name:16s} argmax_flips={flips} {s}", flush=True)
main(*sys.argv[1:])
