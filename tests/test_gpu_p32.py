"""GPU: the P32 (split-bf16 rows) data path — conversions, the TMA-fed tcgen05 GEMM, LayerNorm and MSDA writers."""
import ctypes as C

import pytest
import torch

from tests.util import p32_decode, p32_encode, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _splitk_enabled():
    """Kernel tests exercise the split-K paths too (a forward in throughput mode leaves the process-wide cap at 1)."""
    from egtr_b200 import _lib
    _lib.call("egtr_set_splitk_max", 64)
    yield


def _st():
    return torch.cuda.current_stream().cuda_stream


def test_p32_encode_decode_host():
    x = torch.randn(7, 96)
    assert relerr(p32_decode(p32_encode(x)), x) < 2 ** -16


@pytest.mark.parametrize("rows,Cn", [(1, 32), (333, 256), (1000, 1024)])
def test_rows_to_p32_roundtrip(cuda, rows, Cn):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(rows + Cn)
    x, add = torch.randn(rows, Cn, generator=g).to(cuda), torch.randn(rows, Cn, generator=g).to(cuda)
    out = torch.empty(rows, Cn, device=cuda)
    _lib.call("egtr_rows_to_p32", x.data_ptr(), None, rows, Cn, Cn, out.data_ptr(), _st())
    assert torch.equal(out.view(torch.int32), p32_encode(x).view(torch.int32))  # bit-exact split
    _lib.call("egtr_rows_to_p32", x.data_ptr(), add.data_ptr(), rows, Cn, Cn, out.data_ptr(), _st())
    assert torch.equal(out.view(torch.int32), p32_encode(x + add).view(torch.int32))
    back = torch.empty(rows, Cn, device=cuda)
    _lib.call("egtr_p32_to_rows", out.data_ptr(), rows, Cn, back.data_ptr(), Cn, _st())
    assert torch.equal(back, p32_decode(out))


def _gemm_p32(a_p32, w, bias, *, relu=False, res=None, res_fmt=0, out_fmt=0, keep=None, remap=None, out=None, ldo=None):
    from egtr_b200 import _lib
    from egtr_b200._lib import ASrc, Epilogue
    from egtr_b200.engine import Lin
    dev = w.device
    lin = Lin(w, bias, dev)
    M, K = a_p32.shape
    src, ep = ASrc(), Epilogue()
    src.a, src.mode, src.lda, src.fmt = a_p32.data_ptr(), 0, K, 1
    if out is None:
        out = torch.full((M, lin.N), float("nan"), device=dev)
    ep.bias, ep.res, ep.out = (lin.b.data_ptr() if bias is not None else None), (res.data_ptr() if res is not None else None), out.data_ptr()
    ep.ldo = ldo or lin.N
    ep.ldr = ep.ldo
    ep.relu, ep.out_fmt, ep.res_fmt = int(relu), out_fmt, res_fmt
    ep.rows_per_b, ep.bstride, ep.off = remap or (0, 0, 0)
    ep.row_keep = keep.data_ptr() if keep is not None else None
    _lib.call("egtr_gemm_sbf16", C.byref(src), lin.planes.data_ptr(), M, lin.N, lin.Npad, lin.K, C.byref(ep), _st())
    torch.cuda.synchronize()
    return out


P32_SHAPES = [(128, 64, 64), (200, 256, 256), (1000, 384, 256), (4100, 1024, 256), (2500, 256, 1024), (77, 512, 256),
              (129, 64, 256), (22223, 256, 256), (5, 64, 128), (333, 1536, 256), (66800, 64, 256), (1030, 96, 128), (500, 32, 64),
              (200, 256, 1024), (273, 256, 2048), (130, 1024, 512), (1050, 512, 2048)]  # the last four take the split-K path


@pytest.mark.parametrize("M,N,K", P32_SHAPES)
def test_gemm_p32_plain(cuda, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = p32_encode(torch.randn(M, K, generator=g)).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    ad = p32_decode(a).double()
    tol = 2e-5 * max(1.0, (K / 1024) ** 0.5)
    out = _gemm_p32(a, w, None)
    assert relerr(out, ad @ w.double().t()) < tol
    out = _gemm_p32(a, w, b, relu=True)
    assert relerr(out, (ad @ w.double().t() + b.double()).relu()) < tol


@pytest.mark.parametrize("M,N,K", [(200, 256, 256), (1000, 384, 256), (4100, 1024, 256), (2500, 256, 1024), (333, 64, 64), (273, 256, 2048)])
@pytest.mark.parametrize("res_fmt", [0, 1])
@pytest.mark.parametrize("out_fmt", [0, 1])
def test_gemm_p32_formats(cuda, M, N, K, res_fmt, out_fmt):
    g = torch.Generator().manual_seed(M + N + K + res_fmt * 2 + out_fmt)
    a = p32_encode(torch.randn(M, K, generator=g)).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g)
    res_dev = (p32_encode(res) if res_fmt else res).to(cuda)
    res_val = p32_decode(res_dev) if res_fmt else res_dev
    keep = (torch.rand(M, generator=g) > 0.3).to(torch.uint8).to(cuda)
    ref = (p32_decode(a).double() @ w.double().t() + b.double() + res_val.double()).relu() * keep.double().unsqueeze(1)
    out = _gemm_p32(a, w, b, relu=True, res=res_dev, res_fmt=res_fmt, out_fmt=out_fmt, keep=keep)
    got = p32_decode(out) if out_fmt else out
    assert relerr(got, ref) < 2e-5 * max(1.0, (K / 1024) ** 0.5)


def test_gemm_p32_remap_levels(cuda):
    """input_proj style: per-image level slices of a [B, S, 256] buffer (rows_per_b / bstride / off)."""
    g = torch.Generator().manual_seed(5)
    B, hw, S, off, K = 3, 300, 1000, 450, 512
    a = p32_encode(torch.randn(B * hw, K, generator=g)).to(cuda)
    w = (torch.randn(256, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(256, generator=g).to(cuda)
    out = torch.full((B * S, 256), 7.0, device=cuda)
    _gemm_p32(a, w, b, remap=(hw, S, off), out=out, ldo=256)
    ref = (p32_decode(a).double() @ w.double().t() + b.double()).view(B, hw, 256)
    o3 = out.view(B, S, 256)
    assert relerr(o3[:, off:off + hw], ref) < 2e-5
    assert (o3[:, :off] == 7.0).all() and (o3[:, off + hw:] == 7.0).all()  # nothing outside the slice is touched


@pytest.mark.parametrize("res_fmt", [None, 0, 1])
def test_add_layernorm_p32(cuda, res_fmt):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(3)
    rows = 1237
    x, res, pos = (torch.randn(rows, 256, generator=g) for _ in range(3))
    gamma, beta = torch.randn(256, generator=g).to(cuda), torch.randn(256, generator=g).to(cuda)
    res_dev = None if res_fmt is None else (p32_encode(res) if res_fmt else res).to(cuda)
    res_val = 0 if res_fmt is None else (p32_decode(res_dev) if res_fmt else res_dev).double().cpu()
    want = torch.nn.functional.layer_norm(x.double() + res_val, (256,), gamma.double().cpu(), beta.double().cpu(), 1e-5)
    o_p, o_f, o_plus = (torch.empty(rows, 256, device=cuda) for _ in range(3))
    xd, pd = x.to(cuda), pos.to(cuda)
    _lib.call("egtr_add_layernorm_p32", xd.data_ptr(), res_dev.data_ptr() if res_dev is not None else None, res_fmt or 0,
              gamma.data_ptr(), beta.data_ptr(), rows, 256, o_p.data_ptr(), o_f.data_ptr(), pd.data_ptr(), o_plus.data_ptr(), _st())
    torch.cuda.synchronize()
    assert relerr(o_f, want) < 5e-6
    assert torch.equal(o_p.view(torch.int32), p32_encode(o_f.cpu()).to(cuda).view(torch.int32))
    assert torch.equal(o_plus.view(torch.int32), p32_encode((o_f + pd).cpu()).to(cuda).view(torch.int32))


def test_msda_fused_p32_out(cuda):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(11)
    shapes = [(12, 17), (6, 9), (3, 5), (2, 3)]
    S = sum(h * w for h, w in shapes)
    B = 2
    value = torch.randn(B * S, 256, generator=g).to(cuda)
    offaw = torch.randn(B * S, 384, generator=g).to(cuda)
    vr = (0.5 + 0.5 * torch.rand(B, 4, 2, generator=g)).to(cuda)
    sh = (C.c_int * 8)(*[v for hw in shapes for v in hw])
    o_f, o_p = torch.empty(B * S, 256, device=cuda), torch.empty(B * S, 256, device=cuda)
    for out, fmt in ((o_f, 0), (o_p, 1)):
        _lib.call("egtr_msda_fused_fwd_ex", value.data_ptr(), 256, sh, offaw.data_ptr(), 384, None, vr.data_ptr(), 1,
                  B, S, 8, 32, 4, S, 4, out.data_ptr(), fmt, _st())
    torch.cuda.synchronize()
    assert torch.equal(o_p.view(torch.int32), p32_encode(o_f.cpu()).to(cuda).view(torch.int32))


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", [(2, 13, 17, 64, 64, 3, 1, 1), (1, 20, 31, 128, 256, 3, 2, 1), (2, 9, 11, 256, 512, 1, 2, 0),
                                                  (1, 7, 5, 2048, 256, 3, 2, 1), (1, 50, 84, 256, 256, 3, 1, 1), (1, 100, 167, 128, 128, 3, 1, 1),
                                                  (1, 101, 167, 128, 128, 3, 2, 1), (3, 25, 42, 512, 512, 3, 1, 1)])
@pytest.mark.parametrize("out_fmt", [0, 1])
def test_gemm_p32_conv_nhwc(cuda, B, H, W, Cin, Cout, k, s, p, out_fmt):
    """Convolution as patch-tiled TMA boxes over the P32 NHWC tensor (padding = OOB fill, stride = element stride)."""
    from egtr_b200 import _lib
    from egtr_b200._lib import ASrc, Epilogue
    from egtr_b200.engine import Lin, _conv_mat, conv_out
    g = torch.Generator().manual_seed(H * W + Cin + k)
    x = torch.randn(B * H * W, Cin, generator=g)
    xp = p32_encode(x).to(cuda)
    xd = p32_decode(xp).cpu().view(B, H, W, Cin).permute(0, 3, 1, 2).double()
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    want = torch.nn.functional.conv2d(xd, w.double(), b.double(), stride=s, padding=p).relu()
    OH, OW = conv_out(H, k, s, p), conv_out(W, k, s, p)
    lin = Lin(_conv_mat(w), b, cuda)
    src, ep = ASrc(), Epilogue()
    src.a, src.mode, src.fmt = xp.data_ptr(), 1, 1
    src.H, src.W, src.C, src.OH, src.OW, src.KH, src.KW, src.stride, src.pad = H, W, Cin, OH, OW, k, k, s, p
    M = B * OH * OW
    out = torch.full((M, Cout), float("nan"), device=cuda)
    ep.bias, ep.out, ep.ldo, ep.ldr, ep.relu, ep.out_fmt = lin.b.data_ptr(), out.data_ptr(), Cout, Cout, 1, out_fmt
    _lib.call("egtr_gemm_sbf16", C.byref(src), lin.planes.data_ptr(), M, lin.N, lin.Npad, lin.K, C.byref(ep), _st())
    torch.cuda.synchronize()
    got = (p32_decode(out) if out_fmt else out).cpu().view(B, OH, OW, Cout).permute(0, 3, 1, 2)
    assert relerr(got, want) < 2e-5 * max(1.0, (lin.K / 1024) ** 0.5)


def test_gemm_p32_conv_remap(cuda):
    """The extra feature level: 3x3/2 conv on C5 written into its slice of source_flatten [B, S, 256]."""
    from egtr_b200 import _lib
    from egtr_b200._lib import ASrc, Epilogue
    from egtr_b200.engine import Lin, _conv_mat, conv_out
    g = torch.Generator().manual_seed(9)
    B, H, W, Cin, S, off = 2, 13, 21, 256, 400, 250
    xp = p32_encode(torch.randn(B * H * W, Cin, generator=g)).to(cuda)
    xd = p32_decode(xp).cpu().view(B, H, W, Cin).permute(0, 3, 1, 2).double()
    w = torch.randn(256, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(256, generator=g)
    want = torch.nn.functional.conv2d(xd, w.double(), b.double(), stride=2, padding=1)
    OH, OW = conv_out(H, 3, 2, 1), conv_out(W, 3, 2, 1)
    lin = Lin(_conv_mat(w), b, cuda)
    src, ep = ASrc(), Epilogue()
    src.a, src.mode, src.fmt = xp.data_ptr(), 1, 1
    src.H, src.W, src.C, src.OH, src.OW, src.KH, src.KW, src.stride, src.pad = H, W, Cin, OH, OW, 3, 3, 2, 1
    out = torch.full((B * S, 256), 7.0, device=cuda)
    ep.bias, ep.out, ep.ldo, ep.ldr = lin.b.data_ptr(), out.data_ptr(), 256, 256
    ep.rows_per_b, ep.bstride, ep.off = OH * OW, S, off
    _lib.call("egtr_gemm_sbf16", C.byref(src), lin.planes.data_ptr(), B * OH * OW, 256, 256, lin.K, C.byref(ep), _st())
    torch.cuda.synchronize()
    o3 = out.view(B, S, 256)
    got = o3[:, off:off + OH * OW].cpu().view(B, OH, OW, 256).permute(0, 3, 1, 2)
    assert relerr(got, want) < 2e-5
    assert (o3[:, :off] == 7.0).all() and (o3[:, off + OH * OW:] == 7.0).all()


def test_maxpool_p32_out(cuda):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(2)
    B, H, W, Cn = 2, 21, 34, 64
    x = torch.randn(B, H, W, Cn, generator=g).to(cuda)
    OH, OW = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    o_f, o_p = torch.empty(B * OH * OW, Cn, device=cuda), torch.empty(B * OH * OW, Cn, device=cuda)
    _lib.call("egtr_maxpool3x3s2_nhwc_ex", x.data_ptr(), B, H, W, Cn, o_f.data_ptr(), 0, _st())
    _lib.call("egtr_maxpool3x3s2_nhwc_ex", x.data_ptr(), B, H, W, Cn, o_p.data_ptr(), 1, _st())
    torch.cuda.synchronize()
    want = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).reshape(B * OH * OW, Cn)
    assert torch.equal(o_f, want)
    assert torch.equal(o_p.view(torch.int32), p32_encode(o_f.cpu()).to(cuda).view(torch.int32))


@pytest.mark.parametrize("M,K", [(200, 256), (22223, 256), (5000, 1024), (129, 64)])
@pytest.mark.parametrize("res_fmt", [None, 0, 1])
@pytest.mark.parametrize("with_out2", [False, True])
def test_gemm_p32_layernorm_epilogue(cuda, M, K, res_fmt, with_out2):
    """Linear + residual + LayerNorm in the GEMM epilogue (encoder sub-layer tails) vs fp64 torch."""
    from egtr_b200 import _lib
    from egtr_b200._lib import ASrc, Epilogue
    from egtr_b200.engine import Lin
    g = torch.Generator().manual_seed(M + K + (res_fmt or 0) * 3 + int(with_out2))
    a = p32_encode(torch.randn(M, K, generator=g)).to(cuda)
    w = (torch.randn(256, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(256, generator=g).to(cuda)
    gamma, beta = (1 + 0.2 * torch.randn(256, generator=g)).to(cuda), torch.randn(256, generator=g).to(cuda)
    res = torch.randn(M, 256, generator=g)
    res_dev = None if res_fmt is None else (p32_encode(res) if res_fmt else res).to(cuda)
    res_val = 0 if res_fmt is None else (p32_decode(res_dev) if res_fmt else res_dev).double()
    addend = torch.randn(M, 256, generator=g).to(cuda)
    lin = Lin(w, b, cuda)
    src, ep = ASrc(), Epilogue()
    src.a, src.mode, src.lda, src.fmt = a.data_ptr(), 0, K, 1
    out = torch.full((M, 256), float("nan"), device=cuda)
    out2 = torch.full((M, 256), float("nan"), device=cuda)
    ep.bias, ep.out, ep.ldo, ep.ldr, ep.out_fmt = lin.b.data_ptr(), out.data_ptr(), 256, 256, 1
    if res_dev is not None:
        ep.res, ep.res_fmt = res_dev.data_ptr(), res_fmt
    ep.ln_gamma, ep.ln_beta = gamma.data_ptr(), beta.data_ptr()
    if with_out2:
        ep.ln_out2, ep.ln_addend = out2.data_ptr(), addend.data_ptr()
    _lib.call("egtr_gemm_sbf16", C.byref(src), lin.planes.data_ptr(), M, 256, lin.Npad, K, C.byref(ep), _st())
    torch.cuda.synchronize()
    y = p32_decode(a).double() @ w.double().t() + b.double() + res_val
    want = torch.nn.functional.layer_norm(y, (256,), gamma.double(), beta.double(), 1e-5)
    assert relerr(p32_decode(out), want) < 3e-5
    if with_out2:
        assert relerr(p32_decode(out2), want + addend.double()) < 3e-5
    else:
        assert torch.isnan(out2).all()


def test_groupnorm_p32_outputs(cuda):
    """GroupNorm of a level slice, with the encoder's P32 operands (x and x + pos) written by the same pass."""
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(13)
    B, S, off, hw = 2, 500, 120, 301
    x = torch.randn(B, S, 256, generator=g).to(cuda)
    pos = torch.randn(B, S, 256, generator=g).to(cuda)
    gamma, beta = torch.randn(256, generator=g).to(cuda), torch.randn(256, generator=g).to(cuda)
    ref = x.clone()
    n = int(_lib.call("egtr_groupnorm_scratch_doubles", B, hw))
    scratch = torch.empty(n, dtype=torch.float64, device=cuda)
    _lib.call("egtr_groupnorm_f32", ref.data_ptr(), B, hw, S, off, 256, 32, gamma.data_ptr(), beta.data_ptr(), scratch.data_ptr(), _st())
    o_p = torch.full((B, S, 256), 3.0, device=cuda)
    o_q = torch.full((B, S, 256), 3.0, device=cuda)
    _lib.call("egtr_groupnorm_ex", x.data_ptr(), B, hw, S, off, 256, 32, gamma.data_ptr(), beta.data_ptr(), scratch.data_ptr(),
              o_p.data_ptr(), pos.data_ptr(), o_q.data_ptr(), _st())
    torch.cuda.synchronize()
    assert torch.equal(x, ref)
    sl = slice(off, off + hw)
    assert torch.equal(o_p[:, sl].reshape(-1, 256).contiguous().view(torch.int32), p32_encode(ref[:, sl].reshape(-1, 256).cpu()).to(cuda).view(torch.int32))
    assert torch.equal(o_q[:, sl].reshape(-1, 256).contiguous().view(torch.int32),
                       p32_encode((ref[:, sl] + pos[:, sl]).reshape(-1, 256).cpu()).to(cuda).view(torch.int32))
    assert (o_p[:, :off] == 3.0).all() and (o_p[:, off + hw:] == 3.0).all()
