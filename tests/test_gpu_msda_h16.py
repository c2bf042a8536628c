"""GPU: MSDeformAttn on the H16 pair-record `value` layout (include/egtr_b200.h, EGTR_FMT_H16PAIR): the GEMM epilogue that
writes it, and the gather kernel that reads it — against the fp32-row kernels of the same library and the CPU oracle."""
import ctypes as C

import pytest
import torch

from tests.util import p32_encode, relerr

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


def pair_records(value_rows: torch.Tensor, heads: int) -> torch.Tensor:
    """fp32 [rows, heads*32] -> the H16 pair-record tensor [heads, rows + 1, 2, 32] fp16 (padding slots zero)."""
    rows = value_rows.shape[0]
    v = value_rows.view(rows, heads, 32).permute(1, 0, 2).half()  # [heads, rows, 32]
    rec = torch.zeros(heads, rows + 1, 2, 32, dtype=torch.float16, device=value_rows.device)
    rec[:, 1:, 0] = v   # record r slot 0 = token r - 1
    rec[:, :-1, 1] = v  # record r slot 1 = token r
    return rec.contiguous()


@pytest.mark.parametrize("M,N,K,keep", [(200, 256, 256, False), (22223, 256, 256, True), (1000, 1536, 256, True), (129, 256, 64, False), (5, 256, 256, False)])
def test_gemm_p32_writes_h16_pair_records(cuda, M, N, K, keep):
    from egtr_b200 import _lib
    from egtr_b200._lib import ASrc, Epilogue
    from egtr_b200.engine import Lin
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    mask = (torch.rand(M, generator=g) > 0.2).to(torch.uint8).to(cuda) if keep else None
    lin = Lin(w, bias, cuda)
    a_p32 = p32_encode(a.cpu()).to(cuda)
    heads = N // 32
    out = torch.zeros(heads, M + 1, 2, 32, dtype=torch.float16, device=cuda)
    src, ep = ASrc(), Epilogue()
    src.a, src.mode, src.lda, src.fmt = a_p32.data_ptr(), 0, K, 1
    ep.bias, ep.out, ep.ldo, ep.ldr, ep.out_fmt = lin.b.data_ptr(), out.data_ptr(), N, N, 2
    ep.row_keep = mask.data_ptr() if keep else None
    for div in (1, 2):
        _lib.call("egtr_set_grid_div", div)
        out.zero_()
        _lib.call("egtr_gemm_sbf16", C.byref(src), lin.planes.data_ptr(), M, N, lin.Npad, K, C.byref(ep), _st())
        torch.cuda.synchronize()
        want_rows = a.double() @ w.double().T + bias.double()
        if keep:
            want_rows = want_rows * mask.double()[:, None]
        want = pair_records(want_rows.float(), heads)
        # fp16 rounding of an fp32-accurate product: half an ulp of fp16 relative to each element, plus the GEMM's own 1e-5
        err = (out.double() - want.double()).abs()
        tol = want.double().abs() * 2.0 ** -10 + 1e-4
        err[:, M, 1] = 0  # the trailing padding slot receives the tile's first out-of-range row (bias only): finite, never weighted
        assert bool((err <= tol).all()), float((err - tol).max())
        assert bool(torch.isfinite(out.float()).all())
        assert float(out[:, 0, 0].abs().max()) == 0.0  # the leading padding slot is never written
        assert torch.equal(out[:, 1:, 0], out[:, :-1, 1])  # both copies of every token are the same bits
    _lib.call("egtr_set_grid_div", 1)


@pytest.mark.parametrize("enc", [True, False])
def test_msda_h16_matches_f32_kernel_on_fp16_values(cuda, enc):
    """Same sampling arithmetic, same fp32 accumulation: with `value` pre-rounded to fp16 both layouts must agree to fp32 rounding."""
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(17)
    shapes = [(12, 17), (6, 9), (3, 5), (2, 3)]
    S = sum(h * w for h, w in shapes)
    B, layers = 2, 3
    heads_total = 8 * layers
    Lq = S if enc else 37
    value = torch.randn(B * S, 32 * heads_total, generator=g).half().float().to(cuda)  # exactly representable in fp16
    offaw = (torch.randn(B * Lq, 384, generator=g) * torch.cat([torch.full((256,), 3.0), torch.ones(128)])).to(cuda)
    vr = (0.5 + 0.5 * torch.rand(B, 4, 2, generator=g)).to(cuda)
    ref = torch.rand(Lq, 2, generator=g).to(cuda)
    sh = (C.c_int * 8)(*[v for hw in shapes for v in hw])
    rec = pair_records(value, heads_total)
    for layer in range(layers):
        for fmt in (0, 1):
            want = torch.empty(B * Lq, 256, device=cuda)
            got = torch.full((B * Lq, 256), float("nan"), device=cuda)
            _lib.call("egtr_msda_fused_fwd_ex", value.data_ptr() + layer * 256 * 4, 32 * heads_total, sh, offaw.data_ptr(), 384,
                      None if enc else ref.data_ptr(), vr.data_ptr(), int(enc), B, S, 8, 32, 4, Lq, 4, want.data_ptr(), fmt, _st())
            _lib.call("egtr_msda_fused_fwd_h16", rec.data_ptr(), B * S + 1, layer * 8, heads_total, sh, offaw.data_ptr(), 384,
                      None if enc else ref.data_ptr(), vr.data_ptr(), int(enc), B, S, 8, 32, 4, Lq, 4, got.data_ptr(), fmt, _st())
            torch.cuda.synchronize()
            # same products, different fp32 summation order (left and right corners accumulate in different lanes)
            if fmt == 0:
                assert relerr(got, want) < 2e-5, (layer, relerr(got, want))
            else:  # P32 rows: compare the decoded values (hi + lo)
                from tests.util import p32_decode
                assert relerr(p32_decode(got.cpu()), p32_decode(want.cpu())) < 2e-5


def test_msda_h16_large_decoder_query_set_and_errors(cuda):
    from egtr_b200 import _lib
    g = torch.Generator().manual_seed(3)
    shapes = [(20, 30), (10, 15), (5, 8), (3, 4)]
    S = sum(h * w for h, w in shapes)
    B, Lq = 3, 1500  # Lq * B > 4096: the 32-query decoder form
    value = torch.randn(B * S, 256, generator=g).half().float().to(cuda)
    offaw = torch.randn(B * Lq, 384, generator=g).to(cuda)
    vr = (0.5 + 0.5 * torch.rand(B, 4, 2, generator=g)).to(cuda)
    ref = torch.rand(Lq, 2, generator=g).to(cuda)
    sh = (C.c_int * 8)(*[v for hw in shapes for v in hw])
    rec = pair_records(value, 8)
    want, got = torch.empty(B * Lq, 256, device=cuda), torch.empty(B * Lq, 256, device=cuda)
    _lib.call("egtr_msda_fused_fwd_ex", value.data_ptr(), 256, sh, offaw.data_ptr(), 384, ref.data_ptr(), vr.data_ptr(), 0, B, S, 8, 32, 4, Lq, 4,
              want.data_ptr(), 0, _st())
    _lib.call("egtr_msda_fused_fwd_h16", rec.data_ptr(), B * S + 1, 0, 8, sh, offaw.data_ptr(), 384, ref.data_ptr(), vr.data_ptr(), 0, B, S, 8, 32,
              4, Lq, 4, got.data_ptr(), 0, _st())
    torch.cuda.synchronize()
    assert relerr(got, want) < 2e-5
    with pytest.raises(_lib.EgtrError):  # wrong record count
        _lib.call("egtr_msda_fused_fwd_h16", rec.data_ptr(), B * S, 0, 8, sh, offaw.data_ptr(), 384, ref.data_ptr(), vr.data_ptr(), 0, B, S, 8, 32,
                  4, Lq, 4, got.data_ptr(), 0, _st())
    with pytest.raises(_lib.EgtrError):  # heads out of range
        _lib.call("egtr_msda_fused_fwd_h16", rec.data_ptr(), B * S + 1, 4, 8, sh, offaw.data_ptr(), 384, ref.data_ptr(), vr.data_ptr(), 0, B, S, 8, 32,
                  4, Lq, 4, got.data_ptr(), 0, _st())
