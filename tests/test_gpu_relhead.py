"""GPU: the fused relation-head pair kernel (relhead.cu) through the C ABI against an fp64 torch restatement of
`model/egtr.py:366-418, 507-516` on the factorised inputs (U_l(i), V_l(j): layer 1 applied per query)."""
import ctypes as C

import pytest
import torch

from tests.util import relerr

pytestmark = pytest.mark.gpu


def w3_perm(dev):
    r = torch.arange(64, device=dev)
    return (32 * ((r % 32) // 16) + 16 * (r // 32) + r % 16).to(torch.int32).contiguous()


def pack_w3(w3, P, dev):
    """Layer-3 weights as the fused kernel wants them (include/egtr_b200.h): P <= 64 -> 64 permuted rows, else 64*ceil(P/64) rows."""
    from egtr_b200 import _lib
    from egtr_b200.engine import _ptr, _stream
    rows = 64 if P <= 64 else 64 * ((P + 63) // 64)
    out = torch.empty(rows, 512, dtype=torch.bfloat16, device=dev)
    perm = w3_perm(dev) if P <= 64 else None
    _lib.call("egtr_pack_weight_p32g", _ptr(w3), P, 256, rows, _ptr(perm), _ptr(out), _stream())
    return out, perm


def make_case(dev, B, N, Lr, P, K1, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rnd = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
    U = rnd(B * N, Lr, 516, sc=0.5)
    V = rnd(B * N, Lr, 516, sc=0.5)
    U[..., 513:] = float("nan")  # padding columns are never read
    V[..., 513:] = float("nan")
    w = dict(b1=rnd(512, sc=0.1), w2=rnd(512, 256, sc=1 / 16), b2=rnd(512, sc=0.1), w3=rnd(P, 256, sc=1 / 16), b3=rnd(P, sc=0.1),
             w3c=rnd(256, sc=1 / 16), b3c=0.05)
    cls = torch.randint(0, K1, (B * N,), generator=g).to(dev).to(torch.int32)
    triplet = rnd(K1, K1, P)
    rel_dist = (torch.rand(P, generator=g) * 0.9 + 0.05).to(dev)
    return U, V, w, cls, triplet, rel_dist


def reference(U, V, w, cls, triplet, rel_dist, tau, B, N, Lr, P):
    U, V = U.double().view(B, N, Lr, 516), V.double().view(B, N, Lr, 516)
    gate = torch.sigmoid(U[:, :, None, :, 512] + V[:, None, :, :, 512])  # [B,N,N,Lr]
    h1 = w["b1"].double() + torch.einsum("bijl,bijlc->bijc", gate, U[:, :, None, :, :512] + V[:, None, :, :, :512])
    h1 = torch.relu(h1)
    w2, b2 = w["w2"].double(), w["b2"].double()
    hr = torch.relu(h1[..., :256] @ w2[:256].T + b2[:256])
    hc = torch.relu(h1[..., 256:] @ w2[256:].T + b2[256:])
    rel = hr @ w["w3"].double().T + w["b3"].double()
    if cls is not None:
        c = cls.long().view(B, N)
        rel = rel + torch.stack([triplet.double()[c[b]][:, c[b]] for b in range(B)], 0)
    if rel_dist is not None:
        rel = rel - tau * rel_dist.double().log()
    conn = hc @ w["w3c"].double() + w["b3c"]
    return rel.sigmoid(), conn.sigmoid()


@pytest.mark.parametrize("B,N,Lr,P,freq,adj", [
    (1, 200, 7, 50, True, False),   # VG shape (workloads B / C)
    (2, 37, 3, 30, True, True),     # ragged tiles, few layers, logit adjustment
    (1, 16, 7, 64, False, False),   # exactly one tile, all 64 predicate columns
    (3, 100, 7, 51, True, True),    # odd P: scalar store path
    (1, 1, 1, 1, False, True),      # degenerate
    (1, 300, 7, 200, True, False),  # stress config E: N_q = 300, 200 predicates (one N = 256 layer-3 MMA, own accumulator)
    (2, 50, 7, 65, True, True),     # just past the small-P variant, odd P
    (1, 33, 4, 128, False, True),   # P3 = 128
    (1, 20, 7, 256, True, False),   # every layer-3 column in use
])
def test_relation_pairs_fused_matches_fp64(cuda, B, N, Lr, P, freq, adj):
    from egtr_b200 import _lib
    from egtr_b200.engine import _ptr, _stream
    K1, tau = 11, 0.3
    U, V, w, cls, triplet, rel_dist = make_case(cuda, B, N, Lr, P, K1, seed=B * 1000 + N)
    bf16 = dict(dtype=torch.bfloat16, device=cuda)
    w2g = torch.empty(512, 512, **bf16)
    _lib.call("egtr_pack_weight_p32g", _ptr(w["w2"]), 512, 256, 512, None, _ptr(w2g), _stream())
    w3g, perm = pack_w3(w["w3"], P, cuda)
    hw = _lib.RelheadWeights()
    hw.layers = Lr
    hw.b1, hw.w2g, hw.b2, hw.w3g, hw.b3, hw.w3c, hw.b3c = _ptr(w["b1"]), _ptr(w2g), _ptr(w["b2"]), _ptr(w3g), _ptr(w["b3"]), _ptr(w["w3c"]), w["b3c"]
    pred_rel = torch.full((B, N, N, P), float("nan"), device=cuda)
    pred_conn = torch.full((B, N, N), float("nan"), device=cuda)
    for div in (1, 2):  # full grid / half of the SMs (forwards in flight)
        _lib.call("egtr_set_grid_div", div)
        pred_rel.fill_(float("nan"))
        pred_conn.fill_(float("nan"))
        _lib.call("egtr_relation_pairs_fused_f32", _ptr(U), _ptr(V), 516, Lr, C.byref(hw), _ptr(cls) if freq else None,
                  _ptr(triplet) if freq else None, K1, _ptr(rel_dist) if adj else None, tau, B, N, P, _ptr(pred_rel), _ptr(pred_conn),
                  _stream())
        torch.cuda.synchronize()
        want_rel, want_conn = reference(U, V, w, cls if freq else None, triplet, rel_dist if adj else None, tau, B, N, Lr, P)
        assert torch.isfinite(pred_rel).all() and torch.isfinite(pred_conn).all()
        e_rel, e_conn = relerr(pred_rel, want_rel), relerr(pred_conn, want_conn)
        # pre-sigmoid agreement too: sigmoid squashes errors of saturated logits
        lg = lambda p: torch.log(p.double().clamp(1e-12, 1 - 1e-12)) - torch.log1p(-p.double().clamp(1e-12, 1 - 1e-12))  # noqa: E731
        sel = (want_rel > 1e-4) & (want_rel < 1 - 1e-4)
        e_logit = float((lg(pred_rel)[sel] - lg(want_rel)[sel]).abs().max()) if sel.any() else 0.0
        print(f"B={B} N={N} Lr={Lr} P={P} div={div}: pred_rel {e_rel:.2e} pred_conn {e_conn:.2e} logit abs {e_logit:.2e}")
        assert e_rel < 1e-4 and e_conn < 1e-4 and e_logit < 2e-3
    _lib.call("egtr_set_grid_div", 1)


@pytest.mark.parametrize("N,P", [(200, 50), (150, 200)])
def test_relation_pairs_fused_is_deterministic(cuda, N, P):
    """Repeated launches are bit-identical (no atomics, fixed summation order) — full and partial grids alike."""
    from egtr_b200 import _lib
    from egtr_b200.engine import _ptr, _stream
    B, Lr, K1 = 2, 7, 151
    U, V, w, cls, triplet, rel_dist = make_case(cuda, B, N, Lr, P, K1, seed=5)
    bf16 = dict(dtype=torch.bfloat16, device=cuda)
    w2g = torch.empty(512, 512, **bf16)
    _lib.call("egtr_pack_weight_p32g", _ptr(w["w2"]), 512, 256, 512, None, _ptr(w2g), _stream())
    w3g, perm = pack_w3(w["w3"], P, cuda)
    hw = _lib.RelheadWeights()
    hw.layers = Lr
    hw.b1, hw.w2g, hw.b2, hw.w3g, hw.b3, hw.w3c, hw.b3c = _ptr(w["b1"]), _ptr(w2g), _ptr(w["b2"]), _ptr(w3g), _ptr(w["b3"]), _ptr(w["w3c"]), w["b3c"]
    first = None
    for it in range(30):
        _lib.call("egtr_set_grid_div", 1 + it % 3)
        pred_rel = torch.empty(B, N, N, P, device=cuda)
        pred_conn = torch.empty(B, N, N, device=cuda)
        _lib.call("egtr_relation_pairs_fused_f32", _ptr(U), _ptr(V), 516, Lr, C.byref(hw), _ptr(cls), _ptr(triplet), K1, _ptr(rel_dist), 0.3,
                  B, N, P, _ptr(pred_rel), _ptr(pred_conn), _stream())
        torch.cuda.synchronize()
        if first is None:
            first = (pred_rel, pred_conn)
        else:
            assert torch.equal(first[0], pred_rel) and torch.equal(first[1], pred_conn), it
    _lib.call("egtr_set_grid_div", 1)


def test_relation_pairs_fused_rejects_unsupported(cuda):
    from egtr_b200 import _lib
    from egtr_b200.engine import _ptr, _stream
    U = torch.zeros(4, 7, 516, device=cuda)
    hw = _lib.RelheadWeights()
    out = torch.zeros(4, 4, 300, device=cuda)
    with pytest.raises(_lib.EgtrError):  # null weights
        _lib.call("egtr_relation_pairs_fused_f32", _ptr(U), _ptr(U), 516, 7, C.byref(hw), None, None, 0, None, 0.3, 1, 4, 80, _ptr(out), _ptr(out), _stream())
    z = torch.zeros(512, 512, device=cuda)
    hw.b1 = hw.w2g = hw.b2 = hw.w3g = hw.b3 = hw.w3c = _ptr(z)
    with pytest.raises(_lib.EgtrError):  # more predicates than layer-3 columns
        _lib.call("egtr_relation_pairs_fused_f32", _ptr(U), _ptr(U), 516, 7, C.byref(hw), None, None, 0, None, 0.3, 1, 4, 300, _ptr(out), _ptr(out), _stream())


@pytest.mark.parametrize("wl,overrides", [("small", {}), ("small", dict(logit_adjustment=True, use_freq_bias=False)), ("A", {})])
def test_relation_head_operator_matches_oracle_relation_head(cuda, wl, overrides):
    """`RelationHead` (egtr_relation_head_fwd_f32) fed with the ORACLE's captured queries / keys / hidden state / logits against
    the oracle's own relation head: the operator boundary of SURVEY §8b on its own, no other kernel of the library involved."""
    from egtr_b200.config import WORKLOADS, workload_config
    from egtr_b200.relation_head import RelationHead
    from egtr_b200.synth import synth_images, synth_state_dict
    from oracle import egtr_oracle as orc
    from tests.util import class_flips, pred_rel_err
    cfg = workload_config(wl, **overrides)
    H, W = WORKLOADS[wl]["image"]
    sd = synth_state_dict(cfg, 91)
    px, mask = synth_images(2, H, W, seed=92)
    ref = orc.forward(sd, cfg, px, mask)
    head = RelationHead(cfg, sd, cuda)
    pred_rel, pred_con = head(ref["decoder_attention_queries"], ref["decoder_attention_keys"], ref["last_hidden_state"], ref["logits"])
    torch.cuda.synchronize()
    flipped = class_flips(ref["logits"], ref["logits"])  # same logits in and out: no flips by construction
    assert not flipped.any()
    e_rel, e_con = pred_rel_err(pred_rel, ref["pred_rel"], flipped), relerr(pred_con, ref["pred_connectivity"])
    print(wl, overrides, f"pred_rel {e_rel:.2e} pred_connectivity {e_con:.2e}")
    assert e_rel < 1e-4 and e_con < 1e-4
