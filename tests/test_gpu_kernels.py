"""GPU: every kernel of libegtr_b200.so, called through the C ABI, against the CPU oracle / torch fp64."""
import ctypes as C

import pytest
import torch

from tests.util import load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _splitk_enabled():
    """Kernel tests exercise the split-K paths too (a forward in throughput mode leaves the process-wide cap at 1)."""
    from egtr_b200 import _lib
    _lib.call("egtr_set_splitk_max", 64)
    yield


def _eng_helpers():
    from egtr_b200 import _lib
    from egtr_b200.engine import Lin, _ptr, _stream
    return _lib, Lin, _ptr, _stream


def _gemm(lib, Lin, _ptr, _stream, backend, a, w, bias, *, relu=False, res=None, a2=None, conv=None, M=None, remap=None, out=None, ldo=None):
    from egtr_b200._lib import ASrc, Epilogue
    dev = w.device
    lin = Lin(w, bias, dev)
    src, ep = ASrc(), Epilogue()
    if conv is None:
        src.a, src.a2, src.mode, src.lda = _ptr(a), _ptr(a2), 0, a.shape[1]
        M = a.shape[0]
    else:
        src.a, src.a2, src.mode, src.lda = _ptr(conv["x"]), None, conv.get("mode", 1), 0
        for k in ("H", "W", "C", "OH", "OW", "KH", "KW", "stride", "pad"):
            setattr(src, k, conv[k])
    if out is None:
        out = torch.full((M, lin.N), float("nan"), device=dev)
    ep.bias, ep.res, ep.out = _ptr(lin.b), _ptr(res), _ptr(out)
    ep.ldo = ldo or lin.N
    ep.ldr = ep.ldo
    ep.relu = int(relu)
    ep.rows_per_b, ep.bstride, ep.off = remap or (0, 0, 0)
    if backend == "simt":
        lib.call("egtr_gemm_f32", C.byref(src), _ptr(lin.w), M, lin.N, lin.K, C.byref(ep), _stream())
    else:
        lib.call("egtr_gemm_sbf16", C.byref(src), _ptr(lin.planes), M, lin.N, lin.Npad, lin.K, C.byref(ep), _stream())
    torch.cuda.synchronize()
    return out


GEMM_SHAPES = [(128, 64, 64), (200, 256, 256), (1000, 384, 256), (333, 150, 256), (4100, 1024, 256),
               (2500, 256, 1024), (77, 513, 256), (129, 50, 256), (40000, 256, 256), (5, 64, 128),
               (200, 256, 1024), (273, 256, 2048), (130, 1024, 512), (600, 150, 4096)]  # the last four take the split-K path


@pytest.mark.parametrize("backend", ["simt", "tc"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(cuda, backend, M, N, K):
    lib, Lin, _ptr, _stream = _eng_helpers()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    a2 = torch.randn(M, K, generator=g).to(cuda)
    ref = ((a.double() + a2.double()) @ w.double().t() + b.double() + res.double()).relu()
    out = _gemm(lib, Lin, _ptr, _stream, backend, a, w, b, relu=True, res=res, a2=a2)
    tol = (2e-6 if backend == "simt" else 2e-5) * max(1.0, (K / 1024) ** 0.5)
    assert relerr(out, ref) < tol
    ref2 = a.double() @ w.double().t()
    out2 = _gemm(lib, Lin, _ptr, _stream, backend, a, w, None)
    assert relerr(out2, ref2) < tol


@pytest.mark.parametrize("M,N,K", [(200, 256, 256), (200, 1024, 256), (200, 256, 1024), (300, 150, 256), (7, 513, 256), (512, 601, 256)])
def test_gemm_skinny_f32(cuda, M, N, K):
    from egtr_b200 import _lib
    from egtr_b200._lib import Epilogue
    g = torch.Generator().manual_seed(M + N + K)
    a, a2 = torch.randn(M, K, generator=g).to(cuda), torch.randn(M, K, generator=g).to(cuda)
    w, b = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda), torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    out = torch.full((M, N), float("nan"), device=cuda)
    ep = Epilogue()
    ep.bias, ep.res, ep.out, ep.ldo, ep.ldr, ep.relu = b.data_ptr(), res.data_ptr(), out.data_ptr(), N, N, 1
    one = (C.c_void_p * 1)
    _lib.call("egtr_gemm_f32_grouped", one(a.data_ptr()), one(a2.data_ptr()), one(out.data_ptr()), (C.c_int * 1)(0), 1, (C.c_int * 1)(K),
              w.data_ptr(), M, N, K, C.byref(ep), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = ((a.double() + a2.double()) @ w.double().t() + b.double() + res.double()).relu()
    assert relerr(out, want) < 2e-6 * max(1.0, (K / 256) ** 0.5)


@pytest.mark.parametrize("M", [200, 700])  # 200 rows -> skinny fp32 kernel, 700 rows -> tcgen05 grouped launch
def test_gemm_grouped_launch(cuda, M):
    """Relation-style groups (N=513, mixed row strides) and the decoder's q|k|v triple in one launch each."""
    from egtr_b200.config import workload_config
    from egtr_b200.engine import Engine, LinStack
    import types
    g = torch.Generator().manual_seed(21)
    eng = types.SimpleNamespace()
    bufs = [torch.randn(M, 768, generator=g).to(cuda) for _ in range(3)] + [torch.randn(M, 256, generator=g).to(cuda)]
    ws = [torch.randn(513, 256, generator=g) / 16 for _ in range(6)]
    bs = [torch.randn(513, generator=g) for _ in range(6)]
    st = LinStack(ws, bs, cuda)
    a = [bufs[0], bufs[1], bufs[3], bufs[0], bufs[2], bufs[3]]
    cols = [0, 256, 0, 256, 0, 0]
    ldas = [768, 768, 256, 768, 768, 256]
    out = torch.full((M, 6 * 516), float("nan"), device=cuda)
    Engine.gemm_grouped(eng, st, M, a=a, a_col=cols, lda=ldas, out=[(out, i * 516) for i in range(6)], ldo=6 * 516)
    torch.cuda.synchronize()
    for i in range(6):
        x = a[i][:, cols[i]:cols[i] + 256].double().cpu()
        want = x @ ws[i].double().t() + bs[i].double()
        assert relerr(out[:, i * 516:i * 516 + 513], want) < 2e-5, i
    assert torch.isnan(out[:, 513:516]).all()  # padding columns are never written
    # q|k|v: a2 (position embedding) only on the first two groups
    h, pos = torch.randn(M, 256, generator=g).to(cuda), torch.randn(M, 256, generator=g).to(cuda)
    w3 = [torch.randn(256, 256, generator=g) / 16 for _ in range(3)]
    b3 = [torch.randn(256, generator=g) for _ in range(3)]
    st3 = LinStack(w3, b3, cuda)
    qkv = torch.empty(M, 768, device=cuda)
    Engine.gemm_grouped(eng, st3, M, a=[h, h, h], a2=[pos, pos, None], lda=[256] * 3, out=[(qkv, 0), (qkv, 256), (qkv, 512)], ldo=768)
    torch.cuda.synchronize()
    for i in range(3):
        x = (h + pos if i < 2 else h).double().cpu()
        assert relerr(qkv[:, i * 256:(i + 1) * 256], x @ w3[i].double().t() + b3[i].double()) < 2e-5


@pytest.mark.parametrize("backend", ["simt", "tc"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", [(2, 13, 17, 64, 64, 3, 1, 1), (1, 20, 31, 128, 256, 3, 2, 1),
                                                  (2, 9, 11, 256, 512, 1, 2, 0), (1, 7, 5, 2048, 256, 3, 2, 1)])
def test_gemm_conv_nhwc(cuda, backend, B, H, W, Cin, Cout, k, s, p):
    lib, Lin, _ptr, _stream = _eng_helpers()
    from egtr_b200.engine import _conv_mat, conv_out
    g = torch.Generator().manual_seed(H * W + Cin)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p).permute(0, 2, 3, 1).reshape(-1, Cout)
    OH, OW = conv_out(H, k, s, p), conv_out(W, k, s, p)
    xn = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    out = _gemm(lib, Lin, _ptr, _stream, backend, None, _conv_mat(w).to(cuda), b.to(cuda), M=B * OH * OW,
                conv=dict(x=xn, H=H, W=W, C=Cin, OH=OH, OW=OW, KH=k, KW=k, stride=s, pad=p))
    assert relerr(out, ref) < (5e-6 if backend == "simt" else 1e-4)  # K up to 18432


@pytest.mark.parametrize("backend", ["simt"])  # the NCHW element gather (mode 2) only exists on the CUDA-core kernel; the
def test_gemm_stem_nchw_gather_and_remap(cuda, backend):  # tensor-core stem uses the padded NHWC4 form (next test)
    lib, Lin, _ptr, _stream = _eng_helpers()
    from egtr_b200.engine import _conv_mat, conv_out
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 37, 45
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
    b = torch.randn(64, generator=g)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=2, padding=3).relu().permute(0, 2, 3, 1).reshape(B, -1, 64)
    OH, OW = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
    # write each image's rows at offset 5 of a [B, OH*OW + 9, 64] buffer (the level-slice remap)
    S = OH * OW + 9
    out = torch.zeros(B * S, 64, device=cuda)
    _gemm(lib, Lin, _ptr, _stream, backend, None, _conv_mat(w, 192).to(cuda), b.to(cuda), relu=True, M=B * OH * OW, out=out,
          remap=(OH * OW, S, 5), conv=dict(x=x.to(cuda), mode=2, H=H, W=W, C=3, OH=OH, OW=OW, KH=7, KW=7, stride=2, pad=3))
    got = out.view(B, S, 64)[:, 5:5 + OH * OW]
    assert relerr(got, ref) < (2e-6 if backend == "simt" else 2e-5)
    assert float(out.view(B, S, 64)[:, :5].abs().max()) == 0.0 and float(out.view(B, S, 64)[:, 5 + OH * OW:].abs().max()) == 0.0


def test_gemm_tc_row_remap_into_level_slice(cuda):
    lib, Lin, _ptr, _stream = _eng_helpers()
    g = torch.Generator().manual_seed(8)
    B, rows, S, off = 2, 300, 700, 130
    a = torch.randn(B * rows, 128, generator=g).to(cuda)
    w, b = (torch.randn(256, 128, generator=g) / 11).to(cuda), torch.randn(256, generator=g).to(cuda)
    out = torch.zeros(B * S, 256, device=cuda)
    _gemm(lib, Lin, _ptr, _stream, "tc", a, w, b, out=out, remap=(rows, S, off))
    want = (a.double() @ w.double().t() + b.double()).view(B, rows, 256)
    got = out.view(B, S, 256)
    assert relerr(got[:, off:off + rows], want) < 2e-5
    assert float(got[:, :off].abs().max()) == 0.0 and float(got[:, off + rows:].abs().max()) == 0.0


@pytest.mark.parametrize("backend", ["simt", "tc"])
def test_gemm_stem_padded_nhwc4(cuda, backend):
    lib, Lin, _ptr, _stream = _eng_helpers()
    from egtr_b200.engine import _conv_mat, conv_out
    g = torch.Generator().manual_seed(6)
    B, H, W = 2, 41, 53
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
    b = torch.randn(64, generator=g)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=2, padding=3).relu().permute(0, 2, 3, 1).reshape(-1, 64)
    OH, OW = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
    xd = x.to(cuda)
    x4 = torch.full((B, H + 6, W + 6, 4), float("nan"), device=cuda)
    lib.call("egtr_pad_nchw3_to_nhwc4_f32", xd.data_ptr(), B, H, W, 3, x4.data_ptr(), _stream())
    assert torch.equal(x4[:, 3:-3, 3:-3, :3].permute(0, 3, 1, 2).cpu(), x) and float(x4[..., 3].abs().max()) == 0 and float(x4[:, :3].abs().max()) == 0
    w4 = torch.cat([w, torch.zeros(64, 1, 7, 7)], 1)
    out = _gemm(lib, Lin, _ptr, _stream, backend, None, _conv_mat(w4, 256).to(cuda), b.to(cuda), relu=True, M=B * OH * OW,
                conv=dict(x=x4, mode=3, H=H + 6, W=W + 6, C=4, OH=OH, OW=OW, KH=7, KW=7, stride=2, pad=0))
    assert relerr(out, ref) < (2e-6 if backend == "simt" else 2e-5)


@pytest.mark.parametrize("name", ["msda_enc_small", "msda_dec_small", "msda_edge"])
def test_msda_dropin_matches_reference_golden(cuda, name):
    from egtr_b200.model.deformable_detr import MultiScaleDeformableAttentionFunction
    from egtr_b200.synth import synth_msda_inputs
    ref, meta = load_golden(name)
    value, spatial, start, loc, w = synth_msda_inputs(meta["batch"], [tuple(s) for s in meta["shapes"]], meta["n_query"], seed=meta["seed"])
    if meta["edge"]:
        loc = ref["loc"]
    out = MultiScaleDeformableAttentionFunction.apply(value.to(cuda), spatial.to(cuda), start.to(cuda), loc.to(cuda), w.to(cuda), 64)
    assert out.shape == ref["out"].shape
    assert relerr(out, ref["out"]) < 1e-5


def test_msda_dropin_errors_like_reference(cuda):
    from egtr_b200.model.deformable_detr import ms_deform_attn_forward
    from egtr_b200.synth import synth_msda_inputs
    v, sp, st, loc, w = synth_msda_inputs(1, [(4, 4), (2, 2), (1, 1), (1, 1)], 3)
    with pytest.raises(RuntimeError):
        ms_deform_attn_forward(v, sp, st, loc, w, 64)  # CPU tensors: "Not implemented on the CPU"
    with pytest.raises(RuntimeError):
        ms_deform_attn_forward(v.to(cuda).transpose(1, 2), sp.to(cuda), st.to(cuda), loc.to(cuda), w.to(cuda), 64)  # non-contiguous


@pytest.mark.parametrize("enc", [True, False])
def test_msda_fused_matches_oracle_module_arithmetic(cuda, enc):
    """softmax + sampling locations + gather fused, vs the oracle's step-by-step msda_module."""
    from egtr_b200 import _lib
    from oracle import egtr_oracle as orc
    g = torch.Generator().manual_seed(11)
    B, shapes = 2, [(9, 13), (5, 7), (3, 4), (2, 2)]
    S = sum(h * w for h, w in shapes)
    Lq = S if enc else 21
    value = torch.randn(B, S, 8, 32, generator=g)
    offaw = torch.randn(B, Lq, 384, generator=g) * torch.cat([torch.full((256,), 2.0), torch.ones(128)])
    vr = torch.rand(B, 4, 2, generator=g) * 0.4 + 0.6
    if enc:
        ref_pts = orc.encoder_reference_points(shapes, vr)
        refarg = None
    else:
        pts = torch.rand(Lq, 2, generator=g)
        ref_pts = pts[None, :, None, :] * vr[:, None]
        refarg = pts.to(cuda)
    off = offaw[..., :256].view(B, Lq, 8, 4, 4, 2)
    aw = torch.softmax(offaw[..., 256:].view(B, Lq, 8, 16), -1).view(B, Lq, 8, 4, 4)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    loc = ref_pts[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    want = orc.msda_core(value, shapes, loc, aw)
    out = torch.full((B, Lq, 256), float("nan"), device=cuda)
    shp = (C.c_int * 8)(*[v for hw in shapes for v in hw])
    vd, od, vrd = value.to(cuda), offaw.to(cuda), vr.to(cuda)
    _lib.call("egtr_msda_fused_fwd_f32", vd.data_ptr(), 256, shp, od.data_ptr(), 384, refarg.data_ptr() if refarg is not None else None,
              vrd.data_ptr(), int(enc), B, S, 8, 32, 4, Lq, 4, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert relerr(out, want) < 1e-5


def test_layernorm_maxpool_groupnorm(cuda):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(3)
    st = torch.cuda.current_stream().cuda_stream
    x, r = torch.randn(1001, 256, generator=g) * 3 + 1, torch.randn(1001, 256, generator=g)
    gam, bet = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    out = torch.empty(1001, 256, device=cuda)
    xd, rd, gd, bd = x.to(cuda), r.to(cuda), gam.to(cuda), bet.to(cuda)
    _lib.call("egtr_add_layernorm_f32", xd.data_ptr(), rd.data_ptr(), gd.data_ptr(), bd.data_ptr(), 1001, 256, out.data_ptr(), st)
    assert relerr(out, torch.nn.functional.layer_norm((x + r).double(), (256,), gam.double(), bet.double())) < 1e-5
    # max-pool
    xi = torch.randn(2, 64, 23, 31, generator=g)
    want = torch.nn.functional.max_pool2d(xi, 3, 2, 1).permute(0, 2, 3, 1)
    xn = xi.permute(0, 2, 3, 1).contiguous().to(cuda)
    o = torch.empty(2, 12, 16, 64, device=cuda)
    _lib.call("egtr_maxpool3x3s2_nhwc_f32", xn.data_ptr(), 2, 23, 31, 64, o.data_ptr(), st)
    assert torch.equal(o.cpu(), want.contiguous())
    # GroupNorm on a level slice of a [B,S,256] buffer
    B, S, off, rows = 2, 700, 130, 431
    buf = torch.randn(B, S, 256, generator=g) * 2 + 0.5
    want = buf.clone()
    sl = buf[:, off:off + rows].permute(0, 2, 1).reshape(B, 256, rows, 1)
    want[:, off:off + rows] = torch.nn.functional.group_norm(sl.double(), 32, gam.double(), bet.double(), 1e-5).float().reshape(B, 256, rows).permute(0, 2, 1)
    bd2 = buf.to(cuda)
    scratch = torch.empty(int(_lib.call("egtr_groupnorm_scratch_doubles", B, rows)), dtype=torch.float64, device=cuda)
    _lib.call("egtr_groupnorm_f32", bd2.data_ptr(), B, rows, S, off, 256, 32, gd.data_ptr(), bd.data_ptr(), scratch.data_ptr(), st)
    assert relerr(bd2, want) < 1e-5


def test_levels_geometry_matches_oracle(cuda):
    from egtr_b200 import _lib
    from egtr_b200.engine import level_shapes
    from egtr_b200.synth import synth_images
    from oracle import egtr_oracle as orc
    H, W, B = 150, 203, 3
    _, mask = synth_images(B, H, W, pad_to=[(150, 203), (97, 203), (150, 121)])
    shapes = level_shapes(H, W)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator().manual_seed(9)
    lvl = torch.randn(4, 256, generator=g)
    masks = [orc.nearest_mask(mask, s) for s in shapes]
    want_mask = torch.cat([m.flatten(1) for m in masks], 1)
    want_pos = torch.cat([orc.sine_position_embedding(m).flatten(2).transpose(1, 2) + lvl[l] for l, m in enumerate(masks)], 1)
    want_vr = torch.stack([orc.valid_ratio(m) for m in masks], 1)
    md, ld = mask.to(cuda), lvl.to(cuda)
    i = torch.arange(128, dtype=torch.float32)
    dim_t = (10000.0 ** (2 * torch.div(i, 2, rounding_mode="trunc") / 128)).to(cuda)
    mf = torch.empty(B, S, dtype=torch.uint8, device=cuda)
    pos = torch.empty(B, S, 256, device=cuda)
    vr = torch.empty(B, 4, 2, device=cuda)
    scr = torch.empty(2 * B * S, device=cuda)
    shp = (C.c_int * 8)(*[v for hw in shapes for v in hw])
    _lib.call("egtr_levels_geometry_f32", md.data_ptr(), B, H, W, shp, 4, ld.data_ptr(), dim_t.data_ptr(), 256, mf.data_ptr(), pos.data_ptr(), vr.data_ptr(),
              scr.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(mf.cpu().bool(), want_mask)
    assert torch.equal(vr.cpu(), want_vr)
    valid = want_mask[..., None].expand_as(want_pos)
    assert float((pos.cpu() - want_pos)[valid].abs().max()) < 2e-5
    # fully padded columns normalise by (0 + 1e-6): arguments of ~3e6 rad, where a 1-ulp difference between
    # CUDA's and the host's sinf/cosf argument reduction is visible; they belong to masked tokens only.
    assert float((pos.cpu() - want_pos).abs().max()) < 5e-2


def test_mha_core_and_small_linear(cuda):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(4)
    st = torch.cuda.current_stream().cuda_stream
    B, N = 2, 77
    qkv = torch.randn(B * N, 768, generator=g)
    q, k, v = [qkv[:, i * 256:(i + 1) * 256].view(B, N, 8, 32).transpose(1, 2).double() for i in range(3)]
    want = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(B * N, 256)
    qd = qkv.to(cuda)
    out = torch.empty(B * N, 256, device=cuda)
    _lib.call("egtr_mha_core_f32", qd.data_ptr(), 768, B, N, 8, 32, out.data_ptr(), st)
    assert relerr(out, want) < 1e-5
    # bbox-style small linear with inverse-sigmoid reference add
    x = torch.randn(B * N, 256, generator=g)
    w, b = torch.randn(4, 256, generator=g) / 16, torch.randn(4, generator=g)
    ref = torch.rand(N, 2, generator=g)
    ref[0, 0], ref[1, 1] = 0.0, 1.0
    y = x.double() @ w.double().t() + b.double()
    r = ref.double().repeat(B, 1).clamp(0, 1)
    y[:, :2] += torch.log(r.clamp(min=1e-5) / (1 - r).clamp(min=1e-5))
    xd, wd, bd, rd = x.to(cuda), w.to(cuda), b.to(cuda), ref.to(cuda)
    o = torch.empty(B * N, 4, device=cuda)
    _lib.call("egtr_small_linear_f32", xd.data_ptr(), 256, wd.data_ptr(), bd.data_ptr(), B * N, 256, 4, 2, rd.data_ptr(), 2, N, o.data_ptr(), 4, st)
    assert relerr(o, y.sigmoid()) < 1e-5


@pytest.mark.parametrize("M,K,splits", [(200, 256, 2), (200, 1024, 8), (300, 1024, 4), (7, 256, 2), (800, 1024, 8)])
def test_gemm_splitk_sum_layernorm(cuda, M, K, splits):
    """Decoder sub-layer tail: split-K partial sums of a Linear, then bias + residual + LayerNorm (+ strided second copy)."""
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(M + K + splits)
    a = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(256, K, generator=g) / K ** 0.5).to(cuda)
    b, res = torch.randn(256, generator=g).to(cuda), torch.randn(M, 256, generator=g).to(cuda)
    gamma, beta = torch.randn(256, generator=g).to(cuda), torch.randn(256, generator=g).to(cuda)
    part = torch.full((splits, M, 256), float("nan"), device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("egtr_gemm_f32_splitk", a.data_ptr(), None, K, w.data_ptr(), M, 256, K, splits, part.data_ptr(), st)
    torch.cuda.synchronize()
    lin = a.double() @ w.double().t()
    assert relerr(part.double().sum(0), lin) < 2e-6 * max(1.0, (K / 256) ** 0.5)
    out = torch.empty(M, 256, device=cuda)
    nb = 1 if M % 100 else M // 100  # second copy: [nb, 3, rows_per_b, 256] slot 1
    rpb = M // nb
    out2 = torch.full((nb, 3, rpb, 256), 7.0, device=cuda)
    _lib.call("egtr_sum_layernorm_f32", part.data_ptr(), splits, M * 256, b.data_ptr(), res.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
              M, 256, out.data_ptr(), out2.data_ptr() + rpb * 256 * 4, rpb, 3 * rpb * 256, st)
    torch.cuda.synchronize()
    want = torch.nn.functional.layer_norm(lin + b.double() + res.double(), (256,), gamma.double(), beta.double(), 1e-5)
    assert relerr(out, want) < 5e-6
    assert torch.equal(out2[:, 1].reshape(M, 256), out)
    assert (out2[:, 0] == 7.0).all() and (out2[:, 2] == 7.0).all()
