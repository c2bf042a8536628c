"""GPU: the serving runner's device-side boundary formats: uint8 images in (staging captured in the
graph) and triplet records out (extraction captured in the graph), against the plain model API + the CPU oracles."""
import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as post_orc
from oracle import preprocess_oracle as pre_orc

pytestmark = pytest.mark.gpu


def _model(cuda, wl="tiny", seed=3):
    from egtr_b200.config import workload_config
    from egtr_b200.model.egtr import DetrForSceneGraphGeneration
    from egtr_b200.synth import synth_state_dict
    cfg = workload_config(wl)
    model = DetrForSceneGraphGeneration(cfg)
    model.load_state_dict(synth_state_dict(cfg, seed))
    model.cuda().eval()
    return cfg, model


@pytest.mark.parametrize("raw_hw,hw", [((96, 128), (96, 128)), ((60, 80), (96, 128))])  # images already at model size / resized on the device
def test_u8_in_triplets_out_matches_model_api_and_oracles(cuda, raw_hw, hw):
    from egtr_b200.serving import PipelinedRunner
    cfg, model = _model(cuda)
    B, (H, W) = 2, hw
    rng = np.random.default_rng(5)
    batches = [rng.integers(0, 256, (B,) + raw_hw + (3,), dtype=np.uint8) for _ in range(5)]
    topk = 50
    # concurrency 1 = the latency configuration, the same launches as the eager model API below: results are comparable bit for bit
    pipe = PipelinedRunner(model, B, H, W, depth=3, concurrency=1, input_format="u8", output="triplets", topk=topk,
                           raw_hw=None if raw_hw == hw else raw_hw, resize=None if raw_hw == hw else (H, W))
    got = [{k: v.clone() for k, v in r.items()} for r in pipe.run([(torch.from_numpy(b).pin_memory(), None) for b in batches])]
    assert pipe.h2d_bytes == B * raw_hw[0] * raw_hw[1] * 3
    assert pipe.d2h_bytes < 64 * 1024
    for imgs, g in zip(batches, got):
        # host-side staging oracle (Pillow-pinned) -> plain model API -> CPU triplet oracle
        px, mask, _ = pre_orc.stage_batch(list(imgs), size=H, max_size=W)
        assert px.shape[-2:] == (H, W)
        o = model(torch.from_numpy(px).to(cuda), torch.from_numpy(mask).to(cuda))
        want = post_orc.extract(o.logits.cpu(), o.pred_rel.cpu(), o.pred_connectivity.cpu(), cfg.num_labels, single=False, topk=topk)
        assert torch.equal(g["pred_boxes"], o.pred_boxes.cpu())
        for j in range(B):
            assert np.array_equal(g["pred_classes"][j].numpy(), want[j]["pred_classes"])
            assert np.allclose(g["obj_scores"][j].numpy(), want[j]["obj_scores"], rtol=2e-6)
            same = (g["pred_rel_inds"][j].numpy() == want[j]["pred_rel_inds"]).all(1)
            assert same.mean() > 0.95, same.mean()  # rank swaps only between scores equal to fp32 rounding
            assert np.allclose(g["rel_scores"][j].numpy()[same], want[j]["rel_scores"][same], rtol=2e-6, atol=1e-7)


def test_u8_triplets_runner_in_flight_agrees_with_lone_runner(cuda):
    """Several forwards in flight (throughput configuration, other tilings) give the same triplets as the lone runner up to the
    parity bar: same boxes / scores to 1e-3, and top-k sets that differ only where scores tie to that precision."""
    from egtr_b200.serving import PipelinedRunner
    cfg, model = _model(cuda)
    B, H, W, topk = 2, 96, 128, 50
    rng = np.random.default_rng(8)
    batches = [torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory() for _ in range(6)]
    res = []
    for conc in (1, 3):
        pipe = PipelinedRunner(model, B, H, W, depth=2 * conc, concurrency=conc, input_format="u8", output="triplets", topk=topk)
        res.append([{k: v.clone() for k, v in r.items()} for r in pipe.run([(b, None) for b in batches])])
    for a, b in zip(*res):
        assert torch.allclose(a["pred_boxes"], b["pred_boxes"], rtol=0, atol=1e-3)
        assert torch.allclose(a["obj_scores"], b["obj_scores"], rtol=2e-3, atol=1e-6)
        for j in range(B):
            sa = {tuple(t) for t in a["pred_rel_inds"][j].tolist()}
            sb = {tuple(t) for t in b["pred_rel_inds"][j].tolist()}
            assert len(sa & sb) >= 0.9 * topk


def test_static_stager_matches_oracle_bit_exact(cuda):
    from egtr_b200.preprocess import StaticStager
    rng = np.random.default_rng(6)
    ims = rng.integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
    want_px, want_mask, sizes = pre_orc.stage_batch(list(ims), size=192, max_size=320)
    st = StaticStager(3, (120, 160), (200, 260), cuda, size=192, max_size=320)  # batch tensor larger than the resized image: padding
    assert st.out_hw == sizes[0]
    px = torch.full((3, 3, 200, 260), float("nan"), device=cuda)
    pm = torch.full((3, 200, 260), 7, dtype=torch.int64, device=cuda)
    st.u8.copy_(torch.from_numpy(ims))
    st.enqueue(px, pm)
    torch.cuda.synchronize()
    oh, ow = sizes[0]
    assert torch.equal(px[:, :, :oh, :ow].cpu(), torch.from_numpy(want_px))
    assert float(px[:, :, oh:].abs().sum()) == 0 and float(px[:, :, :, ow:].abs().sum()) == 0
    assert bool(pm[:, :oh, :ow].all()) and int(pm[:, oh:].sum()) == 0 and int(pm[:, :, ow:].sum()) == 0


def test_triplet_records_match_extract_triplets(cuda):
    from egtr_b200.postprocess import TripletRecords, extract_triplets
    g = torch.Generator().manual_seed(11)
    B, N, K, P = 2, 40, 30, 16
    outs = dict(logits=(torch.randn(B, N, K + 1, generator=g) * 2).to(cuda), pred_rel=(torch.rand(B, N, N, P, generator=g) ** 3).to(cuda),
                pred_connectivity=torch.rand(B, N, N, 1, generator=g).to(cuda), pred_boxes=torch.rand(B, N, 4, generator=g).to(cuda))
    for single in (False, True):
        rec = TripletRecords(B, N, K + 1, P, K, cuda, topk=64, single=single)
        got = rec.decode(rec.enqueue(outs))
        want = extract_triplets(outs, K, single=single, topk=64)
        torch.cuda.synchronize()
        assert torch.equal(got["pred_boxes"], outs["pred_boxes"])
        for k in want:
            assert torch.equal(got[k], want[k]), k
