"""GPU parity tests: the CUDA kernels, called through the Python shim and the C
ABI, against (1) the golden fixtures produced by the upstream functions, (2) the
C oracle on seeded inputs, (3) a torch restatement at BASELINE.json's full sizes.

Tolerances (north_star): 1e-5 relative (max-abs / max-abs) in fp32, 1e-2 in bf16.
"""
import numpy as np
import pytest
import torch

from helpers import TOL_BF16, TOL_F32, load_cases, rel_err
import ref_torch

pytestmark = pytest.mark.gpu

CORR = load_cases("corr")
WARP = load_cases("warp")
LRW = load_cases("lrwarp")


@pytest.fixture(scope="module")
def ops():
    import dssmd_b200
    from dssmd_b200 import _lib
    _lib.lib()  # fail loudly if the extension is missing
    return dssmd_b200


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    return O


def cuda(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def feat(shape, seed, relu=True, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(shape, generator=g, device="cuda")
    return (t.relu() if relu else t).to(dtype)


# ------------------------------------------------------------------ correlation
@pytest.mark.parametrize("name", sorted(CORR))
def test_corr_golden(ops, name):
    c = CORR[name]
    bf16 = bool(c["bf16"])
    dt = torch.bfloat16 if bf16 else torch.float32
    tol = TOL_BF16 if bf16 else TOL_F32
    L = cuda(c["L"], dt).requires_grad_(True)
    R = cuda(c["R"], dt).requires_grad_(True)
    out = ops.build_corr(L, R, int(c["D"]))
    assert out.shape == c["out"].shape and out.is_contiguous() and out.dtype == dt
    out.backward(cuda(c["g"], dt))
    assert rel_err(out.detach().float().cpu().numpy(), c["out"]) <= tol
    assert rel_err(L.grad.float().cpu().numpy(), c["gL"]) <= tol
    assert rel_err(R.grad.float().cpu().numpy(), c["gR"]) <= tol


CORR_SHAPES = [
    # B, C, H, W, D, relu          what it covers
    (1, 128, 40, 80, 24, True),    # cfg1 / model shape
    (2, 256, 6, 160, 41, True),    # cfg2 rows (C=256, W=160, D=41)
    (1, 128, 5, 120, 24, False),   # cfg5 row shape, cancellation inputs
    (2, 128, 3, 156, 24, True),    # 384x1248 / 8: W*4 bytes not 16-B aligned for bf16
    (1, 20, 3, 37, 11, False),     # ragged: W % 4 != 0, C not a power of two
    (1, 16, 3, 6, 9, False),       # D > W
    (1, 3, 2, 1, 2, True),         # W = 1
    (1, 1, 1, 8, 1, True),         # C = 1, D = 1
    (1, 64, 2, 300, 192, True),    # large D, x-tiled
    (3, 8, 33, 16, 5, True),       # many rows per block
    (1, 32, 4, 520, 40, False),    # W > 256: several x tiles
    (2, 70, 3, 244, 100, True),    # backward fast path: 16 lanes per pixel group, x-tiled, ragged channel blocks
    (1, 24, 2, 1000, 12, False),   # backward fast path: one lane per group, D = 12 exactly, 5 x tiles
    (1, 8, 2, 64, 192, True),      # backward fast path: largest supported D
]


@pytest.mark.parametrize("shape", CORR_SHAPES, ids=lambda s: "x".join(map(str, s[:5])))
def test_corr_vs_oracle_f32(ops, oracle, shape):
    B, C, H, W, D, relu = shape
    L = feat((B, C, H, W), 1, relu).requires_grad_(True)
    R = feat((B, C, H, W), 2, relu).requires_grad_(True)
    g = feat((B, D, H, W), 3, False)
    out = ops.build_corr(L, R, D)
    out.backward(g)
    Ln, Rn, gn = (t.detach().cpu().numpy() for t in (L, R, g))
    ref = oracle.corr_fwd(Ln, Rn, D)
    ref64 = oracle.corr_fwd(Ln.astype(np.float64), Rn.astype(np.float64), D)
    assert rel_err(out.detach().cpu().numpy(), ref) <= TOL_F32
    assert rel_err(out.detach().cpu().numpy(), ref64) <= TOL_F32
    gL, gR = oracle.corr_bwd(gn, Ln, Rn)
    assert rel_err(L.grad.cpu().numpy(), gL) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), gR) <= TOL_F32
    # exact structural zeros (Appendix B.2)
    o = out.detach()
    for d in range(min(D, W)):
        assert torch.count_nonzero(o[:, d, :, :d]) == 0
    if D > W:
        assert torch.count_nonzero(o[:, W:]) == 0


CORR_SHAPES_16 = CORR_SHAPES[:6] + [
    # 16-bit rows that are a multiple of 16 bytes take the TMA-staged register-blocked kernels
    (1, 64, 2, 304, 192, True),    # 16 lanes per pixel group, x-tiled
    (1, 32, 4, 520, 40, False),    # several x tiles, 4 lanes per group
    (2, 72, 3, 248, 100, True),    # 8 lanes per group, ragged channel blocks
    (1, 24, 2, 1000, 12, False),   # one lane per group: halo rounded up to 16 elements
    (1, 8, 2, 64, 192, True),      # largest supported D, D > W
]


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", CORR_SHAPES_16, ids=lambda s: "x".join(map(str, s[:5])))
def test_corr_vs_oracle_16bit(ops, oracle, shape, dt):
    B, C, H, W, D, relu = shape
    L = feat((B, C, H, W), 1, relu, dt).requires_grad_(True)
    R = feat((B, C, H, W), 2, relu, dt).requires_grad_(True)
    g = feat((B, D, H, W), 3, False, dt)
    out = ops.build_corr(L, R, D)
    assert out.dtype == dt
    out.backward(g)
    Ln, Rn, gn = (t.detach().float().cpu().numpy() for t in (L, R, g))
    assert rel_err(out.detach().float().cpu().numpy(), oracle.corr_fwd(Ln, Rn, D)) <= TOL_BF16
    gL, gR = oracle.corr_bwd(gn, Ln, Rn)
    assert rel_err(L.grad.float().cpu().numpy(), gL) <= TOL_BF16
    assert rel_err(R.grad.float().cpu().numpy(), gR) <= TOL_BF16


@pytest.mark.parametrize("shape", CORR_SHAPES, ids=lambda s: "x".join(map(str, s[:5])))
def test_corr_fwd_register_blocked_variant(ops, oracle, shape, monkeypatch):
    """The TMA-staged register-blocked forward forced on every shape it accepts (by default short fp32 rows and
    large problems go to other kernels), incl. the cross-warp partial-sum path and ragged channel blocks."""
    B, C, H, W, D, relu = shape
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "1")
    monkeypatch.setenv("DSSMD_CORR_FWD_REG", "2")
    L, R = feat((B, C, H, W), 1, relu), feat((B, C, H, W), 2, relu)
    ref = oracle.corr_fwd(L.cpu().numpy(), R.cpu().numpy(), D)
    for cgw, cb in (("0", "0"), ("1", "0"), ("3", "9")):
        monkeypatch.setenv("DSSMD_CORR_FWD_CGW", cgw)
        monkeypatch.setenv("DSSMD_CORR_FWD_CB", cb)
        out = ops.build_corr(L, R, D)
        assert rel_err(out.cpu().numpy(), ref) <= TOL_F32, (cgw, cb)
        for d in range(min(D, W)):
            assert torch.count_nonzero(out[:, d, :, :d]) == 0


UMMA_SHAPES = [
    # B, C, H, W, D, relu            tcgen05 kernel: 128-pixel tiles over the flattened H*W of one image
    (1, 16, 2, 64, 5, True),         # one tile = two whole rows, one pipeline stage
    (1, 32, 3, 80, 24, True),        # tiles cross rows mid-row; halo 23 (TMA box start needs flooring to 16 B)
    (2, 128, 4, 120, 24, False),     # cfg5 row shape, cancellation inputs, two images
    (1, 256, 4, 160, 41, True),      # cfg2 row shape: widest band (N = 256)
    (1, 32, 3, 44, 7, False),        # narrowest supported width: four segments per tile
    (3, 64, 7, 160, 58, True),       # largest supported D, ragged last tile per image
    (1, 48, 5, 300, 33, False),      # wide rows: tiles inside one row
]


@pytest.mark.parametrize("shape", UMMA_SHAPES, ids=lambda s: "x".join(map(str, s[:5])))
def test_corr_fwd_tensor_core_variant(ops, oracle, shape, monkeypatch, capfd):
    """DSSMD_CORR_FWD_VARIANT=2 forces the TMA + tcgen05 (3xTF32) forward that the
    library otherwise selects by size; it must meet the same fp32 bar as the FFMA kernel."""
    B, C, H, W, D, relu = shape
    L = feat((B, C, H, W), 21, relu)
    R = feat((B, C, H, W), 22, relu)
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "2")
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    tc = ops.build_corr(L, R, D)
    torch.cuda.synchronize()
    assert "corr_fwd umma cfg" in capfd.readouterr().err, "shape fell back to the FFMA kernel"
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "1")
    ff = ops.build_corr(L, R, D)
    assert "corr_fwd umma cfg" not in capfd.readouterr().err
    ref64 = oracle.corr_fwd(L.cpu().numpy().astype(np.float64), R.cpu().numpy().astype(np.float64), D)
    assert rel_err(tc.cpu().numpy(), ref64) <= TOL_F32
    assert rel_err(tc.cpu().numpy(), ff.cpu().numpy()) <= TOL_F32
    for d in range(min(D, W)):                       # exact structural zeros
        assert torch.count_nonzero(tc[:, d, :, :d]) == 0


UMMA2_SHAPES = [
    # B, C, H, W, D, relu
    (2, 64, 6, 120, 24, True),      # model row (cfg5): one tile, 120 valid pixels, N = 144
    (2, 64, 5, 160, 41, True),      # cfg2 row: a full tile (N = 176) and a 32-pixel tail tile (N = 80)
    (1, 64, 3, 160, 24, False),     # KITTI row, zero-mean features (cancellation)
    (1, 32, 4, 320, 58, True),      # three tiles per row, the largest D the epilogue window takes
    (1, 48, 7, 36, 2, True),        # a single short tile, D = 2
    (3, 128, 9, 132, 33, False),    # a 4-pixel tail tile; more tiles than one wave of stages
]


@pytest.mark.parametrize("n32", ["0", "1"])
@pytest.mark.parametrize("shape", UMMA2_SHAPES, ids=lambda s: "x".join(map(str, s[:5])))
def test_corr_fwd_tcgen05_row_tiles(ops, oracle, shape, n32, monkeypatch, capfd):
    """DSSMD_CORR_FWD_VARIANT=4: the second tcgen05 design (row-aligned tiles, raw ring + lo ring, truncation split)."""
    B, C, H, W, D, relu = shape
    L = feat((B, C, H, W), 23, relu)
    R = feat((B, C, H, W), 24, relu)
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "4")
    monkeypatch.setenv("DSSMD_UMMA2_N32", n32)
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    tc = ops.build_corr(L, R, D)
    torch.cuda.synchronize()
    assert "corr_fwd umma2 cfg" in capfd.readouterr().err, "shape fell back"
    ref64 = oracle.corr_fwd(L.cpu().numpy().astype(np.float64), R.cpu().numpy().astype(np.float64), D)
    assert rel_err(tc.cpu().numpy(), ref64) <= TOL_F32
    for d in range(min(D, W)):                       # exact structural zeros
        assert torch.count_nonzero(tc[:, d, :, :d]) == 0


def test_corr_f64_and_gradcheck(ops):
    L = feat((1, 5, 3, 9), 1, False, torch.float64).requires_grad_(True)
    R = feat((1, 5, 3, 9), 2, False, torch.float64).requires_grad_(True)
    out = ops.build_corr(L, R, 4)
    assert rel_err(out.detach().cpu().numpy(), ref_torch.corr_ref(L.detach(), R.detach(), 4).cpu().numpy()) <= 1e-12
    assert torch.autograd.gradcheck(lambda a, b: ops.build_corr(a, b, 4), (L, R), eps=1e-6, atol=1e-6)


@pytest.mark.parametrize("cfg", [
    ("cfg2", 8, 256, 80, 160, 41),
    ("cfg3", 8, 128, 40, 80, 24),
    ("cfg4", 16, 128, 48, 160, 24),
    ("cfg5", 4, 128, 72, 120, 24),
    ("cfg5_d192", 4, 128, 72, 120, 192),
], ids=lambda c: c[0])
def test_corr_full_size_vs_torch(ops, cfg):
    """BASELINE.json sizes: compare with the torch restatement on the same GPU
    (fp32 and an fp64 arbiter) plus size-independent properties."""
    _, B, C, H, W, D = cfg
    L = feat((B, C, H, W), 11).requires_grad_(True)
    R = feat((B, C, H, W), 12).requires_grad_(True)
    g = feat((B, D, H, W), 13, False)
    out = ops.build_corr(L, R, D)
    out.backward(g)
    L2, R2 = L.detach().clone().requires_grad_(True), R.detach().clone().requires_grad_(True)
    ref = ref_torch.corr_ref(L2, R2, D)
    ref.backward(g)
    ref64 = ref_torch.corr_ref(L.detach().double(), R.detach().double(), D)
    assert rel_err(out.detach().cpu().numpy(), ref64.cpu().numpy()) <= TOL_F32
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= TOL_F32
    assert rel_err(L.grad.cpu().numpy(), L2.grad.cpu().numpy()) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), R2.grad.cpu().numpy()) <= TOL_F32
    # properties: channel 0 is the plain mean of products; linearity in L
    assert rel_err(out.detach()[:, 0].cpu().numpy(), (L.detach() * R.detach()).mean(1).cpu().numpy()) <= TOL_F32
    out2 = ops.build_corr(2.0 * L.detach(), R.detach(), D)
    assert rel_err(out2.cpu().numpy(), (2.0 * out.detach()).cpu().numpy()) <= 1e-6
    # <g, corr(L,R)> == <gL, L> == <gR, R>  (bilinearity => Euler identity)
    s0 = float((g.double() * out.detach().double()).sum())
    s1 = float((L.grad.double() * L.detach().double()).sum())
    s2 = float((R.grad.double() * R.detach().double()).sum())
    assert abs(s0 - s1) <= 1e-4 * abs(s0) and abs(s0 - s2) <= 1e-4 * abs(s0)


def test_corr_noncontiguous_and_cost_volume(ops):
    L = feat((2, 8, 6, 16), 1).to(memory_format=torch.channels_last)
    R = feat((2, 8, 12, 16), 2)[:, :, ::2]
    out = ops.build_corr(L, R, 5)
    ref = ref_torch.corr_ref(L, R, 5)
    assert out.is_contiguous() and rel_err(out.cpu().numpy(), ref.cpu().numpy()) <= TOL_F32
    cv = ops.CostVolume(5)(L, R)
    assert torch.equal(cv, out)
    with pytest.raises(NotImplementedError):
        ops.CostVolume(5, "cosine")(L, R)


# ------------------------------------------------------------------------ warp
@pytest.mark.parametrize("name", sorted(WARP))
def test_warp_golden(ops, name):
    c = WARP[name]
    pad = str(c["pad"])
    img = cuda(c["img"]).requires_grad_(True)
    disp = cuda(c["disp"]).requires_grad_(True)
    warped, mask = ops.disp_warp(img, disp, padding_mode=pad)
    warped.backward(cuda(c["g"]))
    assert rel_err(warped.detach().cpu().numpy(), c["warped"]) <= TOL_F32
    assert np.array_equal(mask.cpu().numpy(), c["mask"])
    assert rel_err(img.grad.cpu().numpy(), c["g_img"]) <= TOL_F32
    assert rel_err(disp.grad.cpu().numpy(), c["g_disp"]) <= TOL_F32


def images(B, C, H, W, seed, smooth=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    img = torch.rand((B, C, H, W), generator=g, device="cuda")
    if smooth:
        img = torch.nn.functional.avg_pool2d(img, 9, 1, 4)
    base = torch.rand((B, 1, H // 8 + 1, W // 8 + 1), generator=g, device="cuda")
    disp = torch.nn.functional.interpolate(base, size=(H, W), mode="bilinear", align_corners=True)
    disp = 0.5 + disp * (192.0 * W / 960.0 - 0.5)
    return img, disp.contiguous()


@pytest.mark.parametrize("pad", ["border", "zeros"])
@pytest.mark.parametrize("shape", [(2, 3, 40, 80), (1, 3, 33, 157), (1, 1, 2, 9), (1, 7, 16, 64), (2, 3, 64, 1280),
                                   (1, 4, 5, 256),    # staged kernels: 4 channels, widest row of the first plane class
                                   (1, 2, 3, 2064),   # ... and of the last one (W/4 + 2 > 514)
                                   (1, 3, 2, 12)],    # one warp per row, 3 of 32 lanes active
                         ids=lambda s: "x".join(map(str, s)))
def test_warp_vs_oracle(ops, oracle, shape, pad):
    B, C, H, W = shape
    img, disp = images(B, C, H, W, 5, smooth=False)
    img.requires_grad_(True); disp.requires_grad_(True)
    g = feat((B, C, H, W), 6, False)
    warped, mask = ops.disp_warp(img, disp, padding_mode=pad)
    warped.backward(g)
    In, Dn, gn = (t.detach().cpu().numpy() for t in (img, disp, g))
    w_ref, m_ref, bad = oracle.warp_fwd(In, Dn, pad)
    assert not bad
    assert rel_err(warped.detach().cpu().numpy(), w_ref) <= TOL_F32
    assert np.array_equal(mask.cpu().numpy(), m_ref)
    gi_ref, gd_ref = oracle.warp_bwd(gn, In, Dn, pad)
    assert rel_err(img.grad.cpu().numpy(), gi_ref) <= TOL_F32
    assert rel_err(disp.grad.cpu().numpy(), gd_ref) <= TOL_F32
    m = mask.cpu().numpy()
    assert np.all(m[:, :, 0] == 0) and np.all(m[:, :, -1] == 0)
    assert set(np.unique(m)) <= {0.0, 1.0}


@pytest.mark.parametrize("cfg", [("cfg3", 8, 320, 640), ("cfg4", 2, 384, 1280), ("cfg4_1248", 1, 384, 1248), ("cfg5", 4, 576, 960)], ids=lambda c: c[0])
def test_warp_full_size_vs_torch(ops, cfg):
    """Full BASELINE sizes.  The parity reference is the torch restatement run on
    the CPU (that is what upstream's functions do there, and what the golden
    fixtures and the oracle are pinned to): 1e-5.  torch's *CUDA* elementwise
    kernels evaluate `t / (W-1)` as `t * (1/(W-1))`, so upstream itself differs
    between CPU and CUDA by ~1e-5 at these widths; against that variant the bound
    is 5e-5 and the mask may flip only within rounding of the 0.9999 threshold."""
    _, B, H, W = cfg
    img, disp = images(B, 3, H, W, 9)
    img.requires_grad_(True); disp.requires_grad_(True)
    g = feat((B, 3, H, W), 10, False)
    warped, mask = ops.disp_warp(img, disp)
    warped.backward(g)
    for dev, tol, mask_tol in (("cpu", TOL_F32, 0.0), ("cuda", 5e-5, 1e-5)):
        i2 = img.detach().to(dev).clone().requires_grad_(True)
        d2 = disp.detach().to(dev).clone().requires_grad_(True)
        w_ref, m_ref = ref_torch.warp_ref(i2, d2)
        w_ref.backward(g.to(dev))
        assert rel_err(warped.detach().cpu().numpy(), w_ref.detach().cpu().numpy()) <= tol
        assert float((mask.cpu() != m_ref.cpu()).float().mean()) <= mask_tol
        if dev == "cpu":
            assert rel_err(img.grad.cpu().numpy(), i2.grad.cpu().numpy()) <= tol
            assert rel_err(disp.grad.cpu().numpy(), d2.grad.cpu().numpy()) <= tol
        else:
            # a 1-ulp coordinate difference can move floor(ix) across an integer: the sample
            # is continuous there but its derivative is not, so isolated pixels may differ
            for ours, ref in ((img.grad, i2.grad), (disp.grad, d2.grad)):
                diff = (ours.cpu() - ref.cpu()).abs()
                scale = float(ref.abs().max())
                assert float((diff > 1e-4 * scale).float().mean()) <= 1e-4
                assert float(diff.mean()) <= 1e-6 * scale
    frac = float(mask.mean())
    assert 0.7 < frac < 1.0


def test_warp_f64_gradcheck(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand((1, 2, 5, 8), generator=g, device="cuda", dtype=torch.float64).requires_grad_(True)
    disp = (torch.rand((1, 1, 5, 8), generator=g, device="cuda", dtype=torch.float64) * 3 + 0.3).requires_grad_(True)
    for pad in ("border", "zeros"):
        assert torch.autograd.gradcheck(lambda a, b: ops.disp_warp(a, b, pad)[0], (img, disp), eps=1e-6, atol=1e-5)


def test_warp_bwd_many_taps_in_one_cell(ops, oracle):
    """Border mode clamps every pixel whose disparity exceeds x to column 0: one cell of
    g_img collects a whole row of gradients.  The integer accumulators must neither
    overflow nor lose the small terms next to the large ones."""
    B, C, H, W = 1, 3, 5, 512
    g0 = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand((B, C, H, W), generator=g0, device="cuda").requires_grad_(True)
    disp = torch.full((B, 1, H, W), 600.0, device="cuda")
    disp[:, :, :, W // 2:] = 3.25                     # right half: ordinary sub-pixel shift
    disp.requires_grad_(True)
    g = torch.randn((B, C, H, W), generator=g0, device="cuda")
    g[:, 0] = g[:, 0].abs() * 1e3 + 1.0               # same-sign, large: the sum is ~256x a single term
    g[:, 1, :, 7] = 1e6                               # one outlier next to unit-sized terms
    warped, _ = ops.disp_warp(img, disp)
    warped.backward(g)
    gi_ref, gd_ref = oracle.warp_bwd(g.cpu().numpy(), img.detach().cpu().numpy(), disp.detach().cpu().numpy(), "border")
    assert rel_err(img.grad.cpu().numpy(), gi_ref) <= TOL_F32
    assert rel_err(disp.grad.cpu().numpy(), gd_ref) <= TOL_F32
    # per channel too: the outlier of channel 1 must not wash out channels 0 and 2
    for c in range(C):
        assert rel_err(img.grad[:, c].cpu().numpy(), gi_ref[:, c]) <= TOL_F32
    # bit-for-bit repeatable
    img2 = img.detach().clone().requires_grad_(True)
    w2, _ = ops.disp_warp(img2, disp.detach())
    w2.backward(g)
    assert torch.equal(img2.grad, img.grad)


def test_warp_bwd_nonfinite_gradient_poisons_only_its_rows(ops):
    B, C, H, W = 1, 3, 8, 64
    img, disp = images(B, C, H, W, 9)
    img.requires_grad_(True)
    g = torch.ones((B, C, H, W), device="cuda")
    g[0, 1, 4, 10] = float("inf")
    warped, _ = ops.disp_warp(img, disp)
    warped.backward(g)
    gi = img.grad
    bad_rows = (~torch.isfinite(gi)).any(dim=3).any(dim=1)[0].nonzero().flatten().tolist()
    assert bad_rows and set(bad_rows) <= {3, 4, 5}   # output row 4 blends into image rows 3..5 at most
    assert torch.isfinite(gi[:, :, :3]).all() and torch.isfinite(gi[:, :, 6:]).all()


def test_warp_error_conventions(ops):
    img = torch.rand(1, 3, 6, 8, device="cuda")
    disp = torch.rand(1, 1, 6, 8, device="cuda")
    bad = disp.clone(); bad[0, 0, 2, 2] = -1.0
    with pytest.raises(AssertionError):
        ops.disp_warp(img, bad)
    nan = disp.clone(); nan[0, 0, 1, 1] = float("nan")
    with pytest.raises(AssertionError):
        ops.disp_warp(img, nan)
    with pytest.raises(RuntimeError):
        ops.disp_warp(img, disp.double())
    with pytest.raises(RuntimeError):
        ops.disp_warp(img.cpu(), disp.cpu())
    with pytest.raises(ValueError):
        ops.disp_warp(img, disp, padding_mode="nope")
    with pytest.raises(AssertionError):
        ops.normalize_coords(torch.zeros(1, 3, 4, 4, device="cuda"))
    w, m = ops.disp_warp(img, disp)
    assert not m.requires_grad and w.shape == img.shape and m.shape == img.shape


def test_negative_disparity_check_modes(ops):
    """True: upstream's behaviour (the call raises).  "deferred": no host wait in the call, a LATER call (or
    check_negative_disparity_now) raises.  False: no check.  While a CUDA graph is being captured no read-back is
    attempted (a replay could not be interrupted by it)."""
    img = torch.rand(1, 3, 16, 64, device="cuda")
    good = torch.rand(1, 1, 16, 64, device="cuda") * 5
    bad = good.clone(); bad[0, 0, 3, 7] = -0.5
    try:
        ops.set_check_negative_disparity("deferred")
        ops.disp_warp(img, good)
        ops.disp_warp(img, bad)                      # does not raise here
        torch.cuda.synchronize()
        with pytest.raises(AssertionError, match="deferred"):
            ops.disp_warp(img, good)                 # ... but at the next call
        ops.disp_warp(img, bad)
        with pytest.raises(AssertionError):
            ops.check_negative_disparity_now()
        ops.check_negative_disparity_now()           # nothing pending: silent
        for _ in range(200):                         # the flag ring wraps without losing or inventing verdicts
            ops.disp_warp(img, good)
        ops.check_negative_disparity_now()
        ops.set_check_negative_disparity(False)
        ops.disp_warp(img, bad)
        ops.set_check_negative_disparity(True)
        with pytest.raises(AssertionError):
            ops.disp_warp(img, bad)
        # graph capture of the public API: no flag, no host sync inside the captured region
        w_ref, m_ref = ops.disp_warp(img, good)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ops.disp_warp(img, good)
            with torch.cuda.graph(g):
                w, m = ops.disp_warp(img, good)
        torch.cuda.current_stream().wait_stream(s)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(w, w_ref) and torch.equal(m, m_ref)
        with pytest.raises(ValueError):
            ops.set_check_negative_disparity("later")
    finally:
        ops.set_check_negative_disparity(True)


def test_lrwarp_error_resizes_each_image_on_its_own(ops):
    """Upstream resizes imgL and imgR independently against disp (utils/disparity_warper.py:110-113): the two images need
    not have the same size."""
    _, disp = images(2, 3, 40, 80, 27)
    imgL, _ = images(2, 3, 80, 160, 25)      # wider than disp: resized
    imgR, _ = images(2, 3, 40, 80, 26)       # already at disp's size: used as it is
    L = imgL.clone().requires_grad_(True); R = imgR.clone().requires_grad_(True); d = disp.clone().requires_grad_(True)
    err = ops.LRwarp_error(L, d, R)
    Lr = imgL.clone().requires_grad_(True); Rr = imgR.clone().requires_grad_(True); dr = disp.clone().requires_grad_(True)
    ref = ref_torch.lrwarp_ref(Lr, dr, Rr)
    assert err.shape == ref.shape == (2, 3, 40, 80)
    assert rel_err(err.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= TOL_F32
    g = feat((2, 3, 40, 80), 28, False)
    err.backward(g); ref.backward(g)
    assert rel_err(L.grad.cpu().numpy(), Lr.grad.cpu().numpy()) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), Rr.grad.cpu().numpy()) <= TOL_F32
    assert rel_err(d.grad.cpu().numpy(), dr.grad.cpu().numpy()) <= TOL_F32


# ---------------------------------------------------------------- LRwarp_error
@pytest.mark.parametrize("name", sorted(LRW))
def test_lrwarp_golden(ops, name):
    c = LRW[name]
    L = cuda(c["imgL"]).requires_grad_(True)
    R = cuda(c["imgR"]).requires_grad_(True)
    d = cuda(c["disp"]).requires_grad_(True)
    err = ops.LRwarp_error(L, d, R)
    err.backward(cuda(c["g"]))
    assert rel_err(err.detach().cpu().numpy(), c["err"]) <= TOL_F32
    assert rel_err(L.grad.cpu().numpy(), c["g_imgL"]) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), c["g_imgR"]) <= TOL_F32
    assert rel_err(d.grad.cpu().numpy(), c["g_disp"]) <= TOL_F32


@pytest.mark.parametrize("scale", [1, 2, 4, 8])
def test_lrwarp_pyramid_vs_torch(ops, scale):
    B, H0, W0 = 2, 320, 640
    imgL, _ = images(B, 3, H0, W0, 21)
    imgR, _ = images(B, 3, H0, W0, 22)
    _, disp = images(B, 3, H0 // scale, W0 // scale, 23)
    err = ops.LRwarp_error(imgL, disp, imgR)
    ref = ref_torch.lrwarp_ref(imgL, disp, imgR)
    assert err.shape == ref.shape
    assert rel_err(err.cpu().numpy(), ref.cpu().numpy()) <= TOL_F32
    rs = ops.resize_bilinear(imgL, disp.shape[-2:])
    rr = torch.nn.functional.interpolate(imgL, size=disp.shape[-2:], mode="bilinear", align_corners=False)
    assert rel_err(rs.cpu().numpy(), rr.cpu().numpy()) <= TOL_F32


# ---------------------------------------------------------------- photometric L1
PHOTO = load_cases("photo")


@pytest.mark.parametrize("name", sorted(PHOTO))
def test_photometric_golden(ops, name):
    """sum_s w_s * LRwarp_error(imgL, disp_s, imgR).abs().mean() and d/d(disp_s), produced by upstream's functions."""
    c = PHOTO[name]
    L, R = cuda(c["imgL"]), cuda(c["imgR"])
    disps = [cuda(c[f"disp{k}"]).requires_grad_(True) for k in range(int(c["n"]))]
    total = ops.photometric_loss(L, R, disps, weights=[float(w) for w in c["weights"]])
    total.backward()
    assert abs(float(total.detach()) - float(c["total"])) <= TOL_F32 * abs(float(c["total"]))
    for k, d in enumerate(disps):
        assert rel_err(d.grad.cpu().numpy(), c[f"g_disp{k}"]) <= TOL_F32


@pytest.mark.parametrize("shape", [(2, 3, 40, 80, 1), (1, 3, 32, 64, 4), (1, 1, 2, 8, 1), (2, 4, 9, 260, 2), (1, 2, 3, 2064, 1),
                                   (1, 3, 6, 30, 1),      # W % 4 != 0: composed from LRwarp_error
                                   (1, 5, 6, 32, 1)],     # C > 4: composed
                         ids=lambda s: "x".join(map(str, s)))
def test_photometric_vs_oracle(ops, oracle, shape):
    B, C, H, W, scale = shape
    imgL, _ = images(B, C, H * scale, W * scale, 31, smooth=False)
    imgR, _ = images(B, C, H * scale, W * scale, 32, smooth=False)
    _, disp = images(B, C, H, W, 33)
    disp.requires_grad_(True)
    loss = ops.photometric_l1(imgL, disp, imgR)
    (loss * 3.0).backward()
    l_ref, g_ref = oracle.photometric_l1(imgL.cpu().numpy(), disp.detach().cpu().numpy(), imgR.cpu().numpy())
    assert abs(float(loss) - l_ref) <= TOL_F32 * abs(l_ref)
    assert rel_err(disp.grad.cpu().numpy(), 3.0 * g_ref) <= TOL_F32


def test_photometric_pyramid_full_size_vs_composed(ops):
    """cfg3's loss: 320x640 images, B8, the models' [1/8, 1/4, 1/2, 1] disparity pyramid.  The fused kernel against the
    same loss composed from LRwarp_error (whose pieces are pinned above), value and gradient; run-to-run bit equality."""
    B, H0, W0 = 8, 320, 640
    imgL, _ = images(B, 3, H0, W0, 41)
    imgR, _ = images(B, 3, H0, W0, 42)
    pyr = [images(B, 3, H0 // s, W0 // s, 43 + s)[1].requires_grad_(True) for s in (8, 4, 2, 1)]
    total = ops.photometric_loss(imgL, imgR, pyr)
    total.backward()
    got = [d.grad.clone() for d in pyr]
    for d in pyr:
        d.grad = None
    ref = sum(w * ops.LRwarp_error(imgL, d, imgR).abs().mean() for w, d in zip(ops.losses.PYRAMID_WEIGHTS, pyr))
    ref.backward()
    assert abs(float(total) - float(ref)) <= TOL_F32 * abs(float(ref))
    for g, d in zip(got, pyr):
        # |err| is not differentiable at err = 0: where the two paths' samples differ in the last ulp and the residual
        # is ~1e-8, sign(err) flips.  Isolated pixels may differ; everything else must agree to 1e-5.
        diff = (g - d.grad).abs()
        scale = float(d.grad.abs().max())
        assert float((diff > TOL_F32 * scale).float().mean()) <= 2e-5
        assert float(diff.mean()) <= 1e-6 * scale
    again = ops.photometric_loss(imgL, imgR, [d.detach().requires_grad_(True) for d in pyr])
    assert float(again) == float(total)


def test_photometric_error_conventions(ops):
    imgL, disp = images(1, 3, 8, 16, 51)
    imgR, _ = images(1, 3, 8, 16, 52)
    bad = disp.clone(); bad[0, 0, 3, 5] = -1.0
    with pytest.raises(AssertionError):
        ops.photometric_l1(imgL, bad, imgR)
    with pytest.raises(RuntimeError):
        ops.photometric_l1(imgL.cpu(), disp.cpu(), imgR.cpu())      # no CPU fallback
    with pytest.raises(RuntimeError):
        ops.photometric_l1(imgL, disp, imgR[:, :2])
    with pytest.raises(ValueError):
        ops.photometric_loss(imgL, imgR, [disp, disp], weights=[1.0])
    # an image that needs a gradient takes upstream's composed graph
    Lg = imgL.clone().requires_grad_(True)
    ops.photometric_l1(Lg, disp, imgR).backward()
    assert Lg.grad is not None and float(Lg.grad.abs().sum()) > 0


# ---------------------------------------------------------------- warp backward, one-pass band kernel
@pytest.mark.parametrize("pad", ["border", "zeros"])
@pytest.mark.parametrize("rows", [1, 3, 4, 32])
@pytest.mark.parametrize("shape", [(2, 3, 37, 96), (1, 4, 7, 520), (1, 1, 3, 8), (1, 2, 66, 1028), (2, 3, 6, 512)], ids=lambda s: "x".join(map(str, s)))
def test_warp_bwd_band_rows(ops, oracle, shape, rows, pad, monkeypatch):
    """The fused backward with forced band heights: bands of one row, bands that do not divide H, a band taller than
    the image; every band boundary is a halo row scattered twice."""
    monkeypatch.setenv("DSSMD_WARP_BWD_BAND_ROWS", str(rows))
    B, C, H, W = shape
    img, disp = images(B, C, H, W, 61, smooth=False)
    img.requires_grad_(True); disp.requires_grad_(True)
    g = feat((B, C, H, W), 62, False)
    warped, _ = ops.disp_warp(img, disp, padding_mode=pad)
    warped.backward(g)
    gi_ref, gd_ref = oracle.warp_bwd(g.cpu().numpy(), img.detach().cpu().numpy(), disp.detach().cpu().numpy(), pad)
    assert rel_err(img.grad.cpu().numpy(), gi_ref) <= TOL_F32
    assert rel_err(disp.grad.cpu().numpy(), gd_ref) <= TOL_F32
    # only the image needs a gradient: same g_img bits, no g_disp
    img2 = img.detach().clone().requires_grad_(True)
    w2, _ = ops.disp_warp(img2, disp.detach(), padding_mode=pad)
    w2.backward(g)
    assert torch.equal(img2.grad, img.grad)


def test_warp_bwd_band_matches_two_kernel_path(ops, monkeypatch):
    """The one-pass kernels (7: one fixed-point scale per row, 6: one per row and channel) and the scatter + vertical-mix
    pair (5) must agree to the fixed-point quantum, at a full BASELINE size: 2^-30 of a row's sum of |g| over one channel
    (6, 5: <= 1e-6 of the largest gradient) or over all three (7: <= 2e-6; the north_star bar is 1e-5)."""
    img, disp = images(4, 3, 576, 960, 71)
    g = feat((4, 3, 576, 960), 72, False)
    outs = []
    for variant in ("7", "6", "5"):
        monkeypatch.setenv("DSSMD_WARP_BWD_VARIANT", variant)
        i2 = img.clone().requires_grad_(True); d2 = disp.clone().requires_grad_(True)
        w, _ = ops.disp_warp(i2, d2)
        w.backward(g)
        outs.append((i2.grad.clone(), d2.grad.clone()))
    assert rel_err(outs[0][0].cpu().numpy(), outs[2][0].cpu().numpy()) <= 2e-6
    assert rel_err(outs[1][0].cpu().numpy(), outs[2][0].cpu().numpy()) <= 1e-6
    assert torch.equal(outs[1][1], outs[2][1])
    # variant 7 differentiates the blended rows (B[k+1] - B[k]) instead of blending the differences: same value to rounding
    assert rel_err(outs[0][1].cpu().numpy(), outs[2][1].cpu().numpy()) <= 2e-6
    # three resident blocks per SM instead of two (spilling build): same bits as the default build
    monkeypatch.setenv("DSSMD_WARP_BWD_VARIANT", "7")
    monkeypatch.setenv("DSSMD_WARP_BWD_BAND_OCC", "3")
    i3 = img.clone().requires_grad_(True); d3 = disp.clone().requires_grad_(True)
    w, _ = ops.disp_warp(i3, d3)
    w.backward(g)
    for name, a, b in (("g_img", i3.grad, outs[0][0]), ("g_disp", d3.grad, outs[0][1])):
        bad = (a != b)
        assert not bool(bad.any()), (name, int(bad.sum()), float((a - b).abs().max()), bad.nonzero()[:8].tolist())


@pytest.mark.parametrize("pad", ["border", "zeros"])
def test_warp_bwd_heavy_tailed_gradient_is_elementwise_accurate(ops, oracle, pad):
    """the fixed-point scale of a row comes from the row's sum of |g|, so ONE outlier used to set the quantum for
    every other element of its row (relative error of the small elements >> 1e-5, invisible to a max-normalised metric).
    Rows in which one thread's pixels hold more than half of the row's |g| now take the exact (hi, lo) path: every element
    of g_img must be close to the oracle (same fp32 coordinates, sequential fp32 sums) RELATIVE TO THE MAGNITUDES THAT MEET
    IN ITS OWN CELL, outlier rows included; and the result stays bit-reproducible."""
    B, C, H, W = 2, 3, 24, 512
    img, disp = images(B, C, H, W, 91)
    g = feat((B, C, H, W), 92, False)
    g[0, 1, 7, 300] = 3.0e6          # one outlier: row 7 of image 0
    g[1, :, 12, :] = 0.0             # a masked row with a single non-zero pixel
    g[1, 2, 12, 40] = 1.0
    g[1, :, 13, 64:] = 0.0           # a sparse row: 64 live pixels (fast path: its quantum follows its own, small sum)
    i32 = img.clone().requires_grad_(True)
    w, _ = ops.disp_warp(i32, disp, padding_mode=pad)
    w.backward(g)
    In, Dn = img.cpu().numpy(), disp.cpu().numpy()
    ref, _ = oracle.warp_bwd(g.cpu().numpy(), In, Dn, pad)
    mag, _ = oracle.warp_bwd(g.abs().cpu().numpy(), In, Dn, pad)   # sum of |contributions| per cell: what fp32 summation is relative to
    err = np.abs(i32.grad.cpu().numpy().astype(np.float64) - ref)
    # fixed-point rows: n_taps * 2^-30 * sum|g_row| -- with g ~ N(0, 1) on 512 x 3 pixels ~2e-6 absolute, cell magnitudes of order 1;
    # exact rows and the oracle's own fp32 sums: a few 1e-7 of the cell's magnitude
    assert float((err / (mag + 0.1)).max()) <= 2e-5
    assert rel_err(i32.grad[0, 1].cpu().numpy(), ref[0, 1]) <= 1e-6      # the outlier itself
    i2 = img.clone().requires_grad_(True)
    ops.disp_warp(i2, disp, padding_mode=pad)[0].backward(g)
    assert torch.equal(i2.grad, i32.grad)


def test_warp_bwd_unbalanced_channels_keep_their_own_precision(ops, oracle):
    """One fixed-point scale per row serves all channels: a channel 1000x smaller than the others would be quantised
    against THEIR sum.  Rows in which most threads see such a channel take the exact path: each channel, normalised by its
    own maximum, meets the bar."""
    B, C, H, W = 1, 3, 12, 256
    img, disp = images(B, C, H, W, 95)
    g = feat((B, C, H, W), 96, False)
    g[:, 0] *= 1.0e4
    g[:, 2] *= 1.0e-3
    i32 = img.clone().requires_grad_(True)
    ops.disp_warp(i32, disp)[0].backward(g)
    ref, _ = oracle.warp_bwd(g.cpu().numpy(), img.cpu().numpy(), disp.cpu().numpy(), "border")
    for c in range(C):
        assert rel_err(i32.grad[:, c].cpu().numpy(), ref[:, c]) <= TOL_F32
    i2 = img.clone().requires_grad_(True)
    ops.disp_warp(i2, disp)[0].backward(g)
    assert torch.equal(i2.grad, i32.grad)


def test_warp_bwd_nonfinite_gradient_touches_only_its_taps(ops):
    """A single inf in g reaches exactly the cells ATen's scatter would reach (its row takes the fp32 path): the rest of
    the row, the other channels and all other rows stay finite and correct."""
    B, C, H, W = 1, 3, 16, 256
    img, disp = images(B, C, H, W, 93)
    g = feat((B, C, H, W), 94, False)
    g_bad = g.clone()
    g_bad[0, 1, 8, 100] = float("inf")
    i_ok = img.clone().requires_grad_(True)
    ops.disp_warp(i_ok, disp)[0].backward(g)
    i_bad = img.clone().requires_grad_(True)
    ops.disp_warp(i_bad, disp)[0].backward(g_bad)
    bad = ~torch.isfinite(i_bad.grad)
    assert 1 <= int(bad.sum()) <= 4                      # the 2 x 2 bilinear taps of one pixel, one channel
    assert not bad[0, 0].any() and not bad[0, 2].any()
    ys = bad[0, 1].any(dim=1).nonzero().flatten().tolist()
    assert set(ys) <= {7, 8, 9}
    ok = ~bad
    # every other element equals the clean run up to the different accumulation of row 8 (fp32 instead of fixed point)
    assert float((i_bad.grad[ok] - i_ok.grad[ok]).abs().max()) <= 1e-5 * float(i_ok.grad.abs().max())


def test_hoisted_reciprocal_division_is_ieee(ops):
    """The warp kernels' x / (W - 1) with a hoisted reciprocal must equal IEEE division bit for bit (the validity mask
    is compared exactly): checked on the device for every W in [4, 4096] over 1/64-pixel numerators +- 1 ulp."""
    from dssmd_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().dssmd_selftest_division(4, 4096, _lib.ptr(bad), _lib.stream_ptr(bad.device)))
    assert int(bad.item()) == 0


def test_back_to_back_dependent_launches(ops, oracle):
    """The hot-path kernels are launched with programmatic stream serialization (a kernel may become resident while
    its predecessor drains).  Stream order must still hold: chains in which each launch consumes the previous
    launch's output directly, with no other kernel in between, repeated to give a race a chance to show."""
    img, disp = images(2, 3, 96, 320, 81, smooth=False)
    w_ref = img.cpu().numpy()
    for _ in range(3):
        w_ref, _, _ = oracle.warp_fwd(w_ref, disp.cpu().numpy(), "border")
    L, R = feat((2, 64, 24, 80), 82), feat((2, 64, 24, 80), 83)
    c_ref = oracle.corr_fwd(L.cpu().numpy(), R.cpu().numpy(), 64)
    c_ref = oracle.corr_fwd(c_ref, c_ref, 16)
    for it in range(25):
        # a different power-of-two scale per iteration (exact in both ops, so the references just scale): the buffers of one
        # iteration are recycled by the next with DIFFERENT contents, and a consumer that read a stale cache line of its
        # predecessor's output would show
        sc = float(2 ** (it % 5 - 2))
        w = img * sc
        for _ in range(3):
            w, _ = ops.disp_warp(w, disp)
        Ls = L * sc
        c = ops.build_corr(ops.build_corr(Ls, R, 64), ops.build_corr(Ls, R, 64), 16)
        assert rel_err(w.cpu().numpy(), w_ref * sc) <= TOL_F32
        assert rel_err(c.cpu().numpy(), c_ref * sc * sc) <= TOL_F32
    # backward chain: g_img of one warp is the upstream gradient of the next
    g = feat((2, 3, 96, 320), 84, False)
    gi_ref = g.cpu().numpy()
    for _ in range(2):
        gi_ref, _ = oracle.warp_bwd(gi_ref, img.cpu().numpy(), disp.cpu().numpy(), "border")
    for it in range(10):
        sc = float(2 ** (it % 5 - 2))
        x = img.clone().requires_grad_(True)
        y, _ = ops.disp_warp(x, disp)
        (gi1,) = torch.autograd.grad(y, x, g * sc)
        x2 = img.clone().requires_grad_(True)
        y2, _ = ops.disp_warp(x2, disp)
        (gi2,) = torch.autograd.grad(y2, x2, gi1)
        assert rel_err(gi2.cpu().numpy(), gi_ref * sc) <= TOL_F32


# ------------------------------------------------------------------ haze synthesis
HAZE = load_cases("haze")


def _haze_tol(a):
    return TOL_F32 if a.dtype == np.float32 else 1e-13


@pytest.mark.parametrize("name", sorted(HAZE))
def test_haze_golden(ops, name):
    """Upstream's convert_disp_to_depth / depth2trans / recover_haze_images on the loaders' dtypes (float64 scalars with
    float32 images promote to float64), one by one and as the single fused launch."""
    c = HAZE[name]
    t = {k: torch.from_numpy(c[k]).cuda() for k in ("clear", "disp", "baseline", "focal", "beta", "airlight")}
    depth = ops.convert_disp_to_depth(baseline=t["baseline"], focal_length=t["focal"], disp=t["disp"])
    trans = ops.depth2trans(depth, beta=t["beta"])
    haze = ops.recover_haze_images(clean_images=t["clear"], beta=t["beta"], A=t["airlight"], depth=depth)
    fh, ft, fd = ops.synthesize_haze(t["clear"], t["disp"], t["baseline"], t["focal"], t["beta"], t["airlight"])
    for got, fused, key in ((depth, fd, "depth"), (trans, ft, "trans"), (haze, fh, "haze")):
        ref = c[key]
        assert got.shape == ref.shape and got.is_contiguous() and str(got.dtype).endswith(str(ref.dtype)), key
        assert rel_err(got.cpu().numpy(), ref) <= _haze_tol(ref), key
        assert torch.equal(got, fused), key                     # same arithmetic, same bits


@pytest.mark.parametrize("shape,idt,sdt", [((4, 576, 960), torch.float32, torch.float64),
                                           ((8, 320, 640), torch.float32, torch.float32),
                                           ((2, 37, 53), torch.float32, torch.float64),
                                           ((1, 5, 7), torch.float64, torch.float64)])
def test_haze_vs_oracle(ops, oracle, shape, idt, sdt):
    B, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(1024 + H)
    clear = torch.rand(B, 3, H, W, generator=g, device="cuda").to(idt)
    disp = (0.5 + torch.rand(B, 1, H, W, generator=g, device="cuda") * 191.5).to(idt)
    baseline = torch.full((B,), 0.1, dtype=sdt, device="cuda")
    focal = torch.full((B,), 1050.0, dtype=sdt, device="cuda")
    beta = (0.02 + 0.28 * torch.rand(B, generator=g, device="cuda")).to(sdt)
    A = (0.8 + 0.2 * torch.rand(B, generator=g, device="cuda")).to(sdt)
    haze, trans, depth = ops.synthesize_haze(clear, disp, baseline, focal, beta, A)
    n = lambda x: x.cpu().numpy()  # noqa: E731
    ref = oracle.haze(n(clear), n(disp), True, n(baseline), n(focal), n(beta), n(A))
    for got, key in ((haze, "haze"), (trans, "trans"), (depth, "depth")):
        assert n(got).dtype == ref[key].dtype, key
        assert rel_err(n(got), ref[key]) <= _haze_tol(ref[key]), key
    # size-independent properties: transmission in (0, 1], haze between the clear image and the airlight,
    # depth * disp == f * b, and the ASM inverse recovers the clear image where the transmission is not tiny
    assert bool(((trans > 0) & (trans <= 1)).all())
    Ab = A.view(B, 1, 1, 1).to(haze.dtype)
    assert bool((haze >= torch.minimum(clear.to(haze.dtype), Ab) - 1e-6).all())
    assert bool((haze <= torch.maximum(clear.to(haze.dtype), Ab) + 1e-6).all())
    assert rel_err(n(depth * disp), np.full(n(depth).shape, 105.0)) <= TOL_F32
    ok = trans > 0.05
    back = (haze - Ab * (1 - trans)) / trans
    assert float(((back - clear).abs() * ok).max()) <= 1e-4


def test_haze_needs_grad_raises_and_guarded_wrapper_hands_it_to_upstream(ops):
    """The kernels are forward-only: no eager-torch branch inside the package.  What install() binds in upstream's
    modules (haze.guarded) sends such a call to upstream's own function instead."""
    from dssmd_b200 import haze
    disp = (torch.rand(2, 1, 8, 16, device="cuda") + 1).requires_grad_(True)
    one = torch.ones(2, device="cuda")
    with pytest.raises(NotImplementedError, match="forward-only"):
        ops.convert_disp_to_depth(one * 0.1, one * 1050, disp)

    def upstream(baseline, focal_length, disp):       # DeBug/inference.py:57-64
        b = disp.shape[0]
        return focal_length.view(b, 1, 1, 1) * baseline.view(b, 1, 1, 1) / disp
    g = haze.guarded(haze.convert_disp_to_depth, upstream)
    depth = g(baseline=one * 0.1, focal_length=one * 1050, disp=disp)   # the trainers call with keywords
    depth.sum().backward()
    assert rel_err(disp.grad.cpu().numpy(), (-105.0 / disp.detach() ** 2).cpu().numpy()) <= TOL_F32
    with torch.no_grad():                                                # no gradient wanted: the kernel
        k = g(one * 0.1, one * 1050, disp)
    assert rel_err(k.cpu().numpy(), depth.detach().cpu().numpy()) <= TOL_F32


# ------------------------------------------------------------------ correlation into the concat buffer
CAT_SHAPES = [
    # B, C, H, W, D, head channels, dtype, env
    (4, 128, 72, 120, 24, 32, torch.float32, {}),                                   # model shape (cfg5), register-blocked TMA kernels
    (2, 128, 40, 80, 24, 32, torch.float32, {}),                                    # cfg1 / cfg3
    (2, 64, 12, 64, 41, 7, torch.float32, {"DSSMD_CORR_FWD_VARIANT": "2"}),         # tcgen05 forward forced
    (2, 64, 12, 64, 24, 5, torch.float32, {"DSSMD_CORR_FWD_VARIANT": "1", "DSSMD_CORR_FWD_REG": "0", "DSSMD_CORR_BWD_VARIANT": "0"}),  # cp.async kernels
    (2, 64, 12, 64, 24, 5, torch.float32, {"DSSMD_CORR_FWD_VARIANT": "0", "DSSMD_CORR_BWD_VARIANT": "0"}),  # wide-tile forward
    (2, 20, 5, 37, 11, 3, torch.float32, {}),                                       # ragged W: scalar paths
    (2, 64, 8, 48, 12, 16, torch.bfloat16, {}),
    (1, 64, 8, 48, 12, 1, torch.float16, {}),
    (1, 16, 4, 12, 5, 2, torch.float64, {}),
    (1, 16, 3, 6, 9, 4, torch.float32, {}),                                         # D > W
]


@pytest.mark.parametrize("B,C,H,W,D,Ch,dt,env", CAT_SHAPES)
def test_corr_cat_equals_cat_of_corr(ops, monkeypatch, B, C, H, W, D, Ch, dt, env):
    """build_corr_cat(head, L, R, D) == torch.cat((head, build_corr(L, R, D)), 1), forward and all three gradients, to
    the bit: same kernels, only the batch stride of the volume differs (upstream models/DSSMD.py:129-131)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    L0, R0 = feat((B, C, H, W), 31, dtype=dt), feat((B, C, H, W), 32, dtype=dt)
    h0 = feat((B, Ch, H, W), 33, relu=False, dtype=dt)
    g = feat((B, Ch + D, H, W), 34, relu=False, dtype=dt)
    outs = []
    for fused in (False, True):
        L, R, h = (t.clone().requires_grad_(True) for t in (L0, R0, h0))
        out = ops.build_corr_cat(h, L, R, D) if fused else torch.cat((h, ops.build_corr(L, R, D)), dim=1)
        assert out.shape == (B, Ch + D, H, W) and out.is_contiguous() and out.dtype == dt
        out.backward(g)
        outs.append((out.detach(), h.grad, L.grad, R.grad))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert torch.equal(outs[1][1], g[:, :Ch])


def test_corr_cat_vs_oracle(ops, oracle):
    B, C, H, W, D, Ch = 2, 128, 40, 80, 24, 32
    L, R = feat((B, C, H, W), 41), feat((B, C, H, W), 42)
    h = feat((B, Ch, H, W), 43)
    out = ops.build_corr_cat(h, L, R, D)
    ref = oracle.corr_fwd(L.cpu().numpy(), R.cpu().numpy(), D)
    assert torch.equal(out[:, :Ch], h)
    assert rel_err(out[:, Ch:].cpu().numpy(), ref) <= TOL_F32


def test_corr_cat_errors(ops):
    L = feat((1, 8, 4, 8), 1)
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        ops.build_corr_cat(feat((1, 4, 5, 8), 2), L, L, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.build_corr_cat(torch.zeros(1, 4, 4, 8), L.cpu(), L.cpu(), 3)
    from dssmd_b200 import _lib
    vp = _lib.ptr
    out = torch.empty(1, 6, 4, 8, device="cuda")
    rc = _lib.lib().dssmd_corr_fwd_cat(vp(L), vp(L), vp(out), 1, 8, 4, 8, 3, 5, 3, 0, None)
    assert rc == 1 and b"do not fit" in _lib.lib().dssmd_last_error()


# ------------------------------------------------------------------ ASM reconstruction loss
ASM = load_cases("asm")


@pytest.mark.parametrize("name", sorted(ASM))
def test_recovered_clean_loss_golden(ops, name):
    c = ASM[name]
    v2 = bool(c["v2"])
    B = c["airlight"].shape[0]
    trans = cuda(c["trans"]).requires_grad_(True)
    A = cuda(c["airlight"]).requires_grad_(True)
    mod = (ops.RecoveredCleanImagesLossV2 if v2 else ops.RecoveredCleanImagesLoss)(type="normal")
    loss = mod(trans, A.view(B, 1, 1, 1) if v2 else A, cuda(c["haze"]), cuda(c["clean"]))
    loss.backward()
    assert abs(float(loss.detach()) - float(c["loss"])) <= TOL_F32 * abs(float(c["loss"]))
    assert rel_err(trans.grad.cpu().numpy(), c["g_trans"]) <= TOL_F32
    assert A.grad.shape == A.shape
    assert rel_err(A.grad.cpu().numpy(), c["g_airlight"]) <= TOL_F32


def _asm_inputs(B, H0, W0, scale, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g, device="cuda")  # noqa: E731
    h, w = H0 // scale, W0 // scale
    clean = r(B, 3, H0, W0)
    t_true = 0.15 + 0.8 * r(B, 1, H0, W0)
    A_true = 0.8 + 0.2 * r(B, 1)
    haze = clean * t_true + A_true.view(B, 1, 1, 1) * (1 - t_true)
    t_small = torch.nn.functional.avg_pool2d(t_true, scale) if scale > 1 else t_true
    trans = (t_small * (0.7 + 0.6 * r(B, 1, h, w))).clamp(0.05, 1.0)
    A = A_true + 0.2 * (r(B, 1) - 0.5)
    return trans, A, haze, clean


@pytest.mark.parametrize("B,H0,W0,scale", [(4, 576, 960, 1), (8, 320, 640, 1), (8, 320, 640, 2), (8, 320, 640, 4),
                                            (8, 320, 640, 8), (2, 37, 53, 1), (1, 30, 42, 3)])
def test_recovered_clean_loss_vs_oracle(ops, oracle, B, H0, W0, scale):
    trans, A, haze, clean = _asm_inputs(B, H0, W0, scale, 1024 + scale)
    tr, Ar = trans.clone().requires_grad_(True), A.clone().requires_grad_(True)
    loss = ops.recovered_clean_l1(tr, Ar, haze, clean)
    loss.backward()
    n = lambda x: x.detach().cpu().numpy()  # noqa: E731
    l_ref, gt_ref, ga_ref = oracle.recovered_clean_l1(n(trans), n(A), n(haze), n(clean))
    assert abs(float(loss.detach()) - l_ref) <= TOL_F32 * abs(l_ref)
    assert rel_err(n(tr.grad), gt_ref) <= TOL_F32
    assert rel_err(n(Ar.grad).reshape(-1), ga_ref) <= 5 * TOL_F32       # a sum of ~1e5..1e6 signed fp32 terms per sample
    # run-to-run bit equality (fixed reduction order)
    tr2, Ar2 = trans.clone().requires_grad_(True), A.clone().requires_grad_(True)
    loss2 = ops.recovered_clean_l1(tr2, Ar2, haze, clean)
    loss2.backward()
    assert torch.equal(loss2.detach(), loss.detach()) and torch.equal(tr2.grad, tr.grad) and torch.equal(Ar2.grad, Ar.grad)


def test_recovered_clean_loss_pyramid_matches_upstream_expression(ops):
    """The trainer's loop (trainfiles/dssmd_traner_new.py:298-302) over the 4-level transmission pyramid against upstream's
    expression evaluated with torch ops on the same GPU, value and gradients; zero loss for the true parameters."""
    B, H0, W0 = 4, 320, 640
    weights = [0.6, 0.8, 1.0, 1.0]
    pyr = [_asm_inputs(B, H0, W0, s, 77 + s) for s in (8, 4, 2, 1)]
    haze, clean, A0 = pyr[-1][2], pyr[-1][3], pyr[-1][1]

    def upstream(t, A):
        s = haze.shape[-1] // t.shape[-1]
        hz, cl = haze, clean
        if s != 1:
            hz = torch.nn.functional.interpolate(hz, scale_factor=1. / s, mode='bilinear', align_corners=False)
            cl = torch.nn.functional.interpolate(cl, scale_factor=1. / s, mode='bilinear', align_corners=False)
        a = A.unsqueeze(-1).unsqueeze(-1)
        rec = torch.clamp((hz - a * (1 - t)) / t, min=0, max=1.0)
        return torch.nn.functional.l1_loss(rec, cl)

    res = []
    for fn in (lambda t, A: ops.RecoveredCleanImagesLoss()(t, A, haze, clean), upstream):
        ts = [p[0].clone().requires_grad_(True) for p in pyr]
        A = A0.clone().requires_grad_(True)
        total = sum(fn(t, A) * wt for t, wt in zip(ts, weights))
        total.backward()
        res.append((total.detach(), [t.grad for t in ts], A.grad))
    assert abs(float(res[0][0]) - float(res[1][0])) <= TOL_F32 * abs(float(res[1][0]))
    for a, b in zip(res[0][1], res[1][1]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= TOL_F32
    assert rel_err(res[0][2].cpu().numpy(), res[1][2].cpu().numpy()) <= 5 * TOL_F32
    # the true transmission and airlight recover the clean image: zero loss
    t_true = 0.2 + 0.7 * torch.rand(B, 1, H0, W0, device="cuda")
    A_true = torch.full((B, 1), 0.9, device="cuda")
    hz = clean * t_true + 0.9 * (1 - t_true)
    assert float(ops.recovered_clean_l1(t_true, A_true, hz, clean)) <= 1e-6


def test_recovered_clean_loss_errors(ops):
    t = torch.rand(2, 1, 8, 16, device="cuda") + 0.1
    A = torch.rand(2, 1, device="cuda")
    img = torch.rand(2, 3, 16, 32, device="cuda")
    with pytest.raises(NotImplementedError, match="integer multiple"):
        ops.recovered_clean_l1(t, A, torch.rand(2, 3, 17, 32, device="cuda"), torch.rand(2, 3, 17, 32, device="cuda"))
    with pytest.raises(RuntimeError, match="one value per sample"):
        ops.recovered_clean_l1(t, torch.rand(3, 1, device="cuda"), img, img)
    with pytest.raises(NotImplementedError, match="float32"):
        ops.recovered_clean_l1(t.double(), A.double(), img.double(), img.double())
    with pytest.raises(UnboundLocalError):
        ops.RecoveredCleanImagesLoss(type="other")(t, A, img, img)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_data_parallel_two_gpus_matches_one(ops):
    """Upstream's live parallel mode is nn.DataParallel (trainfiles/dssmd_traner_new.py:131-135): one Python thread per GPU
    calls the ops concurrently, each on its own device and stream.  Forward and backward of a module built on build_corr +
    disp_warp must give the same numbers split over two GPUs as on one."""
    import torch.nn as nn

    class Path(nn.Module):
        def __init__(self):
            super().__init__()
            self.scale = nn.Parameter(torch.tensor(1.5))

        def forward(self, L, R, img, disp):
            vol = ops.build_corr(L * self.scale, R, 24)
            warped, mask = ops.disp_warp(img, disp)
            return vol.mean((1, 2, 3)) + (warped * mask).mean((1, 2, 3))

    B = 4
    L, R = feat((B, 128, 16, 120), 101), feat((B, 128, 16, 120), 102)
    img, disp = images(B, 3, 64, 256, 103)
    outs = []
    for ids in ([0], [0, 1]):
        m = nn.DataParallel(Path().cuda(0), device_ids=ids)
        Lg, Rg = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
        ig, dg = img.clone().requires_grad_(True), disp.clone().requires_grad_(True)
        y = m(Lg, Rg, ig, dg)
        y.sum().backward()
        outs.append([t.detach().cpu().numpy() for t in (y, Lg.grad, Rg.grad, ig.grad, dg.grad, m.module.scale.grad)])
    for a, b in zip(*outs):
        assert rel_err(a, b) <= TOL_F32


def test_default_dispatch_of_the_baseline_configs():
    """Which kernel each BASELINE.json config takes BY DEFAULT is part of the measured result (the dispatcher's rules were set
    by measurement): a heuristic regression must fail here, not only show up as a slower bench.  DSSMD_CORR_DEBUG prints the
    configuration of every launch; run in a subprocess with no variant override."""
    import os
    import subprocess
    import sys
    from helpers import ROOT
    code = r'''
import sys
sys.path.insert(0, %r)
import torch, dssmd_b200
def corr(tag, B, C, H, W, D):
    print("##", tag, file=sys.stderr, flush=True)
    L = torch.randn(B, C, H, W, device="cuda").relu().requires_grad_(True)
    R = torch.randn(B, C, H, W, device="cuda").relu().requires_grad_(True)
    dssmd_b200.build_corr(L, R, D).sum().backward()
    torch.cuda.synchronize()
def warp(tag, B, H, W):
    print("##", tag, file=sys.stderr, flush=True)
    img = torch.rand(B, 3, H, W, device="cuda", requires_grad=True)
    disp = (torch.rand(B, 1, H, W, device="cuda") * 20).requires_grad_(True)
    dssmd_b200.disp_warp(img, disp)[0].sum().backward()
    torch.cuda.synchronize()
corr("cfg5", 4, 128, 72, 120, 24); warp("cfg5w", 4, 576, 960)
corr("cfg5_d12", 4, 128, 72, 120, 12)
corr("cfg2", 8, 256, 80, 160, 41)
corr("cfg3", 8, 128, 40, 80, 24); warp("cfg3w", 8, 320, 640)
corr("cfg4", 16, 128, 48, 160, 24); warp("cfg4w", 2, 384, 1280)
print("done")
''' % ROOT
    env = {k: v for k, v in os.environ.items() if not k.startswith("DSSMD_")}
    env.update(DSSMD_TUNING="1", DSSMD_CORR_DEBUG="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "done" in r.stdout, r.stderr[-3000:]
    sec, cur = {}, None
    for line in r.stderr.splitlines():
        if line.startswith("## "):
            cur = line[3:].strip(); sec[cur] = []
        elif cur and " cfg:" in line:
            sec[cur].append(line.split(" cfg:")[0].strip())
    assert "not eligible" not in r.stderr, r.stderr[-3000:]
    # the model shape of the headline config: the mma.sync forward (D = 24; the register-blocked FFMA kernel keeps D < 20), ONE
    # backward launch for both gradients (the D <= 24 kernel with a lane owning all disparities of its pixels); 320x640 feature
    # maps (W = 80, 62 % of a 128-pixel lane tile): the mma.sync kernels, both gradients in ONE launch
    assert sec["cfg5"] == ["corr_fwd mma32", "corr_bwd wide"], sec
    assert sec["cfg5_d12"][0] == "corr_fwd reg", sec
    assert sec["cfg3"] == ["corr_fwd mma32", "corr_bwd mma32", "corr_bwd mma32", "corr_bwd mma32 both"], sec
    # C256 80x160 D41 and KITTI feature maps: the fp32 mma.sync (3xTF32) kernels, both gradients planned and issued as ONE launch
    for k in ("cfg2", "cfg4"):
        assert sec[k] == ["corr_fwd mma32", "corr_bwd mma32", "corr_bwd mma32", "corr_bwd mma32 both"], (k, sec)
    # image-sized warps: lean forward, one-pass band backward (rewritten kernel)
    for k in ("cfg5w", "cfg3w", "cfg4w"):
        assert sec[k] == ["warp_fwd lean", "warp_bwd band2"], (k, sec)


def test_fast_paths_are_taken_from_fresh_threads():
    """The TMA-staged kernels need tensor maps, and the driver's encoder needs a current context in the calling thread.
    An autograd worker on its first backward and every nn.DataParallel replica thread (trainfiles/dssmd_traner_new.py:131-135)
    start without one: the first call from such a thread must still get the fast kernels (DSSMD_CORR_DEBUG prints the chosen
    configuration), not the generic fallback.  Runs in a subprocess so that the threads are really fresh."""
    import os
    import subprocess
    import sys
    from helpers import ROOT
    code = r'''
import sys, threading
sys.path.insert(0, %r)
import torch, dssmd_b200
L = torch.randn(2, 128, 36, 120, device="cuda").relu().requires_grad_(True)
R = torch.randn(2, 128, 36, 120, device="cuda").relu().requires_grad_(True)
torch.cuda.synchronize()
def worker():
    with torch.cuda.device(0):
        out = dssmd_b200.build_corr(L, R, 24)      # forward from a fresh replica thread
    box.append(out)
box = []
t = threading.Thread(target=worker); t.start(); t.join()
box[0].sum().backward()                           # backward on the autograd worker's first call
torch.cuda.synchronize()
print("done")
''' % ROOT
    env = dict(os.environ, DSSMD_CORR_DEBUG="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "done" in r.stdout, r.stderr[-2000:]
    assert "corr_fwd reg cfg" in r.stderr or "corr_fwd mma32 cfg" in r.stderr, r.stderr[-2000:]
    # one fused TMA-staged launch for both gradients, or a TMA-staged kernel each
    assert (r.stderr.count("corr_bwd fused cfg") + r.stderr.count("corr_bwd wide cfg") == 1
            or r.stderr.count("corr_bwd reg cfg") + r.stderr.count("corr_bwd mma32 cfg") == 2), r.stderr[-2000:]
    assert "not eligible" not in r.stderr, r.stderr[-2000:]


# ------------------------------------------------------------------ attention along the epipolar line
ATTN = load_cases("attn")


@pytest.mark.parametrize("name", sorted(ATTN))
def test_epipolar_attention_golden(ops, name):
    c = ATTN[name]
    q, k, v = (cuda(c[n]).requires_grad_(True) for n in ("q", "k", "v"))
    out = ops.epipolar_attention(q, k, v, float(c["scale"]))
    assert out.shape == q.shape and out.is_contiguous()
    out.backward(cuda(c["g_out"]))
    assert rel_err(out.detach().cpu().numpy(), c["out"]) <= TOL_F32
    assert rel_err(q.grad.cpu().numpy(), c["g_q"]) <= TOL_F32
    assert rel_err(k.grad.cpu().numpy(), c["g_k"]) <= TOL_F32
    assert rel_err(v.grad.cpu().numpy(), c["g_v"]) <= TOL_F32


@pytest.mark.parametrize("name", ["bidnet_like", "short_row"])
def test_sttm_single_module_drop_in(ops, name):
    """dssmd_b200.STTM_Single loads upstream's state_dict unchanged and reproduces the module's output and input gradients
    (train mode: BatchNorm uses batch statistics, as in the fixture)."""
    c = ATTN[name]
    dim, heads, dh, out_dim = (int(t) for t in c["cfg"])
    mod = ops.STTM_Single(dim=dim, heads=heads, dim_head=dh, out_dim=out_dim).cuda().train()
    sd = {k_[2:]: torch.from_numpy(np.ascontiguousarray(c[k_])) for k_ in c if k_.startswith("w:")}
    mod.load_state_dict(sd)                       # strict: same parameter and buffer names as upstream
    left, right = cuda(c["left"]).requires_grad_(True), cuda(c["right"]).requires_grad_(True)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False       # the 1x1 convolutions around the kernel: fp32 like the CPU fixture
    try:
        y = mod(left, right)
        y.backward(cuda(c["g_y"]))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel_err(y.detach().cpu().numpy(), c["y"]) <= 2 * TOL_F32
    assert rel_err(left.grad.cpu().numpy(), c["g_left"]) <= 1e-4     # through cuDNN convolutions and BatchNorm on either side
    assert rel_err(right.grad.cpu().numpy(), c["g_right"]) <= 1e-4


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 6, 640), (1, 32, 4, 1280), (2, 128, 5, 80), (1, 64, 3, 160), (2, 16, 3, 44), (1, 32, 2, 4)])
def test_epipolar_attention_vs_torch_and_oracle(ops, oracle, B, C, H, W):
    """BidNet's full-resolution widths against upstream's expression with torch ops on the same GPU (fp32 matmul) and, for the
    small cases, the C oracle; rows independent; deterministic."""
    g = torch.Generator(device="cuda").manual_seed(100 + W)
    q, k, v, go = (torch.randn(B, C, H, W, generator=g, device="cuda") for _ in range(4))
    scale = 32 ** -0.5
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        res = []
        for fused in (True, False):
            qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
            if fused:
                out = ops.epipolar_attention(qq, kk, vv, scale)
            else:                                   # models/BidNet/STTM_Simple.py:43-57
                b, c, h, w = qq.shape
                q2, k2, v2 = (t.permute(0, 2, 3, 1).reshape(b * h, w, c).contiguous() for t in (qq, kk, vv))
                attn = torch.softmax(torch.matmul(q2, k2.transpose(-1, -2)) * scale, dim=-1)
                out = torch.matmul(attn, v2).view(b, h, w, c).contiguous().permute(0, 3, 1, 2)
            out.backward(go)
            res.append([t.detach().cpu().numpy() for t in (out, qq.grad, kk.grad, vv.grad)])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    for a, b_ in zip(*res):
        assert rel_err(a, b_) <= TOL_F32
    if B * H * W * W * C <= 2e8:
        ref = oracle.epipolar_attention(*(t.cpu().numpy() for t in (q, k, v)), scale, go.cpu().numpy())
        for a, b_ in zip(res[0], ref):
            assert rel_err(a, b_) <= TOL_F32
    out2 = ops.epipolar_attention(q, k, v, scale)
    assert torch.equal(out2, torch.from_numpy(res[0][0]).cuda())
    k3 = k.clone(); k3[:, :, 0] += 1.0               # row 0 of k only touches row 0 of the output
    out3 = ops.epipolar_attention(q, k3, v, scale)
    assert torch.equal(out3[:, :, 1:], out2[:, :, 1:]) and not torch.equal(out3[:, :, 0], out2[:, :, 0])


@pytest.mark.parametrize("case", ["large", "small", "mixed_rows", "outlier", "zero_v", "c16"])
def test_epipolar_attention_mma_forward_scaling(ops, case, monkeypatch, capfd):
    """The tensor-core forward splits every staged tile into fp16 hi / lo after dividing it by a power of two >= its largest
    magnitude (epipolar_attn_mma.cu): inputs far outside fp16's range, rows of very different magnitude inside one image, a
    single outlier key and an all-zero value tensor must stay at the fp32 bar against upstream's expression in float64."""
    B, C, H, W = (2, 16, 3, 200) if case == "c16" else (2, 32, 3, 328)
    g = torch.Generator(device="cuda").manual_seed(7)
    q, k, v = (torch.randn(B, C, H, W, generator=g, device="cuda") for _ in range(3))
    scale = C ** -0.5
    if case == "large":
        q, k, v, scale = q * 3.0e4, k * 2.0e3, v * 1.0e6, scale / 6.0e7
    elif case == "small":
        q, k, v, scale = q * 1.0e-5, k * 3.0e-6, v * 1.0e-7, scale * 3.3e10
    elif case == "mixed_rows":
        k = k * torch.logspace(-6, 6, W, device="cuda").roll(37)       # key tiles of very different scale
        v = v * torch.logspace(5, -5, W, device="cuda")
        scale = scale * 1e-6
    elif case == "outlier":
        k[0, :, 1, 77] *= 1.0e4
        v[1, 3, 2, 5] = 3.0e5
    elif case == "zero_v":
        v = torch.zeros_like(v)
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    out = ops.epipolar_attention(q, k, v, scale)
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    assert "epipolar_attn_fwd mma cfg" in err or "epipolar_attn_fwd umma cfg" in err
    q2, k2, v2 = (t.double().permute(0, 2, 3, 1).reshape(B * H, W, C) for t in (q, k, v))
    ref = torch.matmul(torch.softmax(torch.matmul(q2, k2.transpose(-1, -2)) * scale, dim=-1), v2).view(B, H, W, C).permute(0, 3, 1, 2)
    if case == "zero_v":
        assert torch.count_nonzero(out) == 0
    else:
        assert rel_err(out.cpu().numpy(), ref.float().cpu().numpy()) <= TOL_F32
        if case == "mixed_rows":   # element-wise where the reference is not tiny: small-magnitude outputs keep their own precision
            big = ref.abs() > 1e-3 * ref.abs().amax(dim=(1, 3), keepdim=True)
            assert ((out.double() - ref).abs()[big] <= 1e-4 * ref.abs()[big] + 1e-6 * ref.abs().amax()).all()
    monkeypatch.setenv("DSSMD_ATTN_FWD_MMA", "0")       # the FFMA kernel it replaces agrees
    out2 = ops.epipolar_attention(q, k, v, scale)
    assert rel_err(out.cpu().numpy(), out2.cpu().numpy()) <= TOL_F32


@pytest.mark.parametrize("case", ["plain", "large", "small", "mixed_rows", "outlier", "c16"])
def test_epipolar_attention_mma_backward_scaling(ops, case, monkeypatch, capfd):
    """The tensor-core backward (one power-of-two scale per image row and tensor, fp16 hi / lo split operands, seven GEMMs) against
    float64 autograd of upstream's expression: gradients at the fp32 bar for inputs far outside fp16's range, image rows of very
    different magnitude, single outliers; bit-identical from run to run; the FFMA kernels it replaces agree."""
    B, C, H, W = (2, 16, 3, 200) if case == "c16" else (2, 32, 3, 328)
    g = torch.Generator(device="cuda").manual_seed(11)
    q, k, v, go = (torch.randn(B, C, H, W, generator=g, device="cuda") for _ in range(4))
    scale = C ** -0.5
    if case == "large":
        q, k, v, go, scale = q * 3.0e4, k * 2.0e3, v * 1.0e6, go * 5.0e3, scale / 6.0e7
    elif case == "small":
        q, k, v, go, scale = q * 1.0e-5, k * 3.0e-6, v * 1.0e-7, go * 1.0e-4, scale * 3.3e10
    elif case == "mixed_rows":     # every image row (b, y) at its own magnitude
        m = torch.logspace(-4, 4, B * H, device="cuda").view(B, 1, H, 1)
        q, v, go = q * m, v * m.flip(2), go / m
        scale_rows = None
        k = k / m
    elif case == "outlier":
        k[0, :, 1, 77] *= 50.0
        v[1, 3, 2, 5] = 3.0e3
        go[0, 7, 0, 100] = -2.0e3
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")

    def run():
        qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
        ops.epipolar_attention(qq, kk, vv, scale).backward(go)
        return qq.grad, kk.grad, vv.grad
    got = run()
    torch.cuda.synchronize()
    assert "epipolar_attn_bwd mma cfg" in capfd.readouterr().err
    q2, k2, v2 = (t.double().requires_grad_(True) for t in (q, k, v))
    qr_, kr_, vr_ = (t.permute(0, 2, 3, 1).reshape(B * H, W, C) for t in (q2, k2, v2))
    ref = torch.matmul(torch.softmax(torch.matmul(qr_, kr_.transpose(-1, -2)) * scale, dim=-1), vr_).view(B, H, W, C).permute(0, 3, 1, 2)
    ref.backward(go.double())
    for a, b_, name in zip(got, (q2.grad, k2.grad, v2.grad), "qkv"):
        if case == "mixed_rows":   # rows differ by 8 decades: compare every image row with its own maximum
            for bb in range(B):
                for yy in range(H):
                    assert rel_err(a[bb, :, yy].cpu().numpy(), b_[bb, :, yy].float().cpu().numpy()) <= TOL_F32, (name, bb, yy)
        else:
            assert rel_err(a.cpu().numpy(), b_.float().cpu().numpy()) <= TOL_F32, name
    again = run()
    assert all(torch.equal(x, y) for x, y in zip(got, again))
    monkeypatch.setenv("DSSMD_ATTN_BWD_MMA", "0")
    monkeypatch.setenv("DSSMD_ATTN_FWD_MMA", "0")
    old = run()
    for a, b_ in zip(got, old):
        if case != "mixed_rows":
            assert rel_err(a.cpu().numpy(), b_.cpu().numpy()) <= 2 * TOL_F32


def test_epipolar_attention_errors(ops):
    x = torch.randn(1, 32, 2, 8, device="cuda")
    with pytest.raises(NotImplementedError, match="needs C in"):
        ops.epipolar_attention(x[:, :24], x[:, :24], x[:, :24], 1.0)
    with pytest.raises(NotImplementedError, match="needs C in"):
        ops.epipolar_attention(x[..., :6], x[..., :6], x[..., :6], 1.0)
    with pytest.raises(NotImplementedError, match="float32"):
        ops.epipolar_attention(x.half(), x.half(), x.half(), 1.0)
    with pytest.raises(RuntimeError, match="differ"):
        ops.epipolar_attention(x, x[:, :16], x, 1.0)


def test_epipolar_attention_repeated_with_changing_data(ops, oracle):
    """The backward is three dependent launches (delta -> dQ -> dK, dV) whose buffers the allocator recycles from call to call:
    new data every iteration, every result checked (a consumer reading a stale line of its predecessor's output would show)."""
    B, C, H, W = 1, 16, 2, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    for it in range(12):
        q, k, v, go = (torch.randn(B, C, H, W, generator=g, device="cuda").requires_grad_(True) for _ in range(4))
        out = ops.epipolar_attention(q, k, v, 0.25)
        out.backward(go.detach())
        ref = oracle.epipolar_attention(*(t.detach().cpu().numpy() for t in (q, k, v)), 0.25, go.detach().cpu().numpy())
        for a, b_ in zip((out, q.grad, k.grad, v.grad), ref):
            assert rel_err(a.detach().cpu().numpy(), b_) <= TOL_F32, it


# ------------------------------------------------------------------ 16-bit correlation forward on the tensor cores (mma.sync)
MMA16_SHAPES = [
    # B, C, H, W, D
    (2, 128, 6, 120, 24),      # model shape (cfg5): 8 warps, the last one half empty
    (1, 256, 4, 160, 41),      # cfg2: 10 warps, 4 column-tile pairs
    (2, 128, 5, 80, 24),       # cfg1 / cfg3
    (1, 64, 3, 80, 5),
    (1, 40, 2, 48, 12),        # channels not a multiple of the 32-channel stage
    (1, 64, 2, 264, 64),       # two x tiles per row
    (1, 32, 2, 64, 81),        # largest D it takes, D > W
    (1, 16, 2, 16, 1),         # D = 1: no halo
    (1, 48, 3, 1000, 40),      # four x tiles, ragged last tile
]


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", MMA16_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_corr_fwd_mma16(ops, oracle, shape, dt, monkeypatch, capfd):
    """bf16 / fp16 forward: the banded row GEMM on mma.sync.m16n8k16 (fp32 accumulation) against the oracle on the rounded inputs,
    and against the FFMA kernels it replaces (DSSMD_CORR_FWD_MMA16=0); also through the concat buffer."""
    B, C, H, W, D = shape
    L, R = feat((B, C, H, W), 11, True, dt), feat((B, C, H, W), 12, False, dt)
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    out = ops.build_corr(L, R, D)
    torch.cuda.synchronize()
    assert "corr_fwd mma16 cfg" in capfd.readouterr().err
    monkeypatch.delenv("DSSMD_CORR_DEBUG")
    assert out.dtype == dt and out.is_contiguous()
    ref = oracle.corr_fwd(L.float().cpu().numpy(), R.float().cpu().numpy(), D)
    assert rel_err(out.float().cpu().numpy(), ref) <= TOL_BF16
    for d in range(min(D, W)):
        assert bool((out[:, d, :, :d] == 0).all())            # the zero left border is exact
    assert bool((out[:, W:] == 0).all())                       # channels d >= W
    monkeypatch.setenv("DSSMD_CORR_FWD_MMA16", "0")
    old = ops.build_corr(L, R, D)
    monkeypatch.delenv("DSSMD_CORR_FWD_MMA16")
    # both accumulate exact products in fp32 and round once: they agree to an ulp of the 16-bit result
    assert rel_err(out.float().cpu().numpy(), old.float().cpu().numpy()) <= (2 ** -7 if dt == torch.bfloat16 else 2 ** -10)
    head = feat((B, 3, H, W), 13, False, dt)
    assert torch.equal(ops.build_corr_cat(head, L, R, D), torch.cat((head, out), dim=1))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", MMA16_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_corr_bwd_mma16(ops, oracle, shape, dt, monkeypatch, capfd):
    """bf16 / fp16 backward: both gradients as banded row GEMMs on mma.sync.m16n8k16 (the band operand built once per warp from
    the staged g tile) against the oracle on the rounded inputs and against the FFMA kernels (DSSMD_CORR_BWD_MMA16=0), and
    through the concat gradient."""
    B, C, H, W, D = shape
    L0, R0 = feat((B, C, H, W), 21, True, dt), feat((B, C, H, W), 22, False, dt)
    g = feat((B, D, H, W), 23, False, dt)
    res = []
    for new in (True, False):
        if new:
            monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
        else:
            monkeypatch.setenv("DSSMD_CORR_BWD_MMA16", "0")
        L, R = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
        ops.build_corr(L, R, D).backward(g)
        torch.cuda.synchronize()
        if new:
            err = capfd.readouterr().err
            assert err.count("corr_bwd mma16 cfg") == 2, err
            monkeypatch.delenv("DSSMD_CORR_DEBUG")
        else:
            monkeypatch.delenv("DSSMD_CORR_BWD_MMA16")
        res.append((L.grad, R.grad))
    gL, gR = oracle.corr_bwd(g.float().cpu().numpy(), L0.float().cpu().numpy(), R0.float().cpu().numpy())
    assert rel_err(res[0][0].float().cpu().numpy(), gL) <= TOL_BF16
    assert rel_err(res[0][1].float().cpu().numpy(), gR) <= TOL_BF16
    ulp = 2 ** -7 if dt == torch.bfloat16 else 2 ** -10
    assert rel_err(res[0][0].float().cpu().numpy(), res[1][0].float().cpu().numpy()) <= ulp
    assert rel_err(res[0][1].float().cpu().numpy(), res[1][1].float().cpu().numpy()) <= ulp
    # the concat gradient read in place
    head = feat((B, 3, H, W), 24, False, dt).requires_grad_(True)
    L, R = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
    gcat = torch.cat((feat((B, 3, H, W), 25, False, dt), g), dim=1)
    ops.build_corr_cat(head, L, R, D).backward(gcat)
    assert torch.equal(L.grad, res[0][0]) and torch.equal(R.grad, res[0][1])


# ------------------------------------------------------------------ fp32 forward on mma.sync with 3xTF32 splitting
MMA32_SHAPES = [
    (2, 128, 6, 120, 24),      # cfg5
    (1, 256, 4, 160, 41),      # cfg2
    (2, 128, 5, 80, 24),       # cfg1 / cfg3
    (1, 64, 3, 156, 24),       # 384x1248 / 8: W % 8 != 0
    (1, 20, 2, 44, 12),        # channels not a multiple of the 16-channel stage
    (1, 64, 2, 264, 64),       # three x tiles per row
    (1, 32, 2, 64, 81),        # largest D, D > W
    (1, 16, 2, 16, 1),         # D = 1
]


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", MMA32_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_corr_fwd_mma32(ops, oracle, shape, relu, monkeypatch, capfd):
    """DSSMD_CORR_FWD_VARIANT=3: the banded row GEMM on mma.sync.m16n8k8 with every fp32 operand split into tf32 hi + lo (three
    products per step) must meet the fp32 bar (1e-5) on post-ReLU features and on zero-mean ones (cancellation)."""
    B, C, H, W, D = shape
    L, R = feat((B, C, H, W), 31, relu), feat((B, C, H, W), 32, relu)
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "3")
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    out = ops.build_corr(L, R, D)
    torch.cuda.synchronize()
    assert "corr_fwd mma32 cfg" in capfd.readouterr().err
    ref = oracle.corr_fwd(L.cpu().numpy(), R.cpu().numpy(), D)
    assert rel_err(out.cpu().numpy(), ref) <= TOL_F32
    for d in range(min(D, W)):
        assert bool((out[:, d, :, :d] == 0).all())
    assert bool((out[:, W:] == 0).all())
    head = feat((B, 3, H, W), 33, False)
    assert torch.equal(ops.build_corr_cat(head, L, R, D), torch.cat((head, out), dim=1))


@pytest.mark.parametrize("bx", ["1", "0"])
def test_corr_mma32_cross_terms_on_bf16(ops, oracle, bx, monkeypatch, capfd):
    """The default mma.sync fp32 kernels compute hi * hi on tf32 and the two cross terms (hi * lo, lo * hi) on bf16 MMAs
    (DSSMD_MMA32_BF16X=0: all three on tf32).  Both must meet the fp32 bar on channels whose scales differ by six decades
    (the split is per element, so no channel may drown another), forward and both gradients."""
    B, C, H, W, D = 1, 64, 3, 160, 41
    sc = torch.logspace(-3, 3, C, device="cuda").view(1, C, 1, 1)
    L0, R0 = feat((B, C, H, W), 61, False) * sc, feat((B, C, H, W), 62, False) / sc
    g = feat((B, D, H, W), 63, False)
    monkeypatch.setenv("DSSMD_MMA32_BF16X", bx)
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "3")
    monkeypatch.setenv("DSSMD_CORR_BWD_VARIANT", "3")
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    L, R = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
    out = ops.build_corr(L, R, D)
    out.backward(g)
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    assert "corr_fwd mma32 cfg" in err and err.count("corr_bwd mma32 cfg") == 2, err
    assert rel_err(out.detach().cpu().numpy(), oracle.corr_fwd(L0.cpu().numpy(), R0.cpu().numpy(), D)) <= TOL_F32
    gL, gR = oracle.corr_bwd(g.cpu().numpy(), L0.cpu().numpy(), R0.cpu().numpy())
    assert rel_err(L.grad.cpu().numpy(), gL) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), gR) <= TOL_F32


@pytest.mark.parametrize("bad", [float("nan"), float("inf")])
def test_corr_mma32_nonfinite_inputs_stay_local(ops, bad, monkeypatch):
    """The tf32 split of the tensor-core kernels has no special case for inf / NaN: a non-finite feature must still make exactly the
    outputs that read it non-finite (through lo = v - hi) and leave every other output finite."""
    B, C, H, W, D = 1, 32, 4, 160, 24
    L, R = feat((B, C, H, W), 71, True), feat((B, C, H, W), 72, True)
    L[0, 5, 2, 77] = bad
    monkeypatch.setenv("DSSMD_CORR_FWD_VARIANT", "3")
    out = ops.build_corr(L, R, D)
    torch.cuda.synchronize()
    nf = ~torch.isfinite(out)
    expect = torch.zeros_like(nf)
    expect[0, :, 2, 77] = True           # out[d, y, x] reads L[c, y, x] for every d <= x
    assert bool((nf == expect).all())


WIDE_SHAPES = [
    (4, 128, 3, 120, 24),    # cfg5 rows: one tile of 120 pixels, two rows per block
    (1, 128, 40, 80, 24),    # cfg1 / cfg3 model shape (20 of 32 lanes)
    (2, 128, 5, 160, 24),    # KITTI rows: two x tiles of 80 pixels
    (1, 72, 7, 132, 17),     # ragged channel blocks (72 = 2 x 32 + 8), D < 24, tiles of 68 pixels
    (3, 32, 2, 128, 13),     # one channel block per row (ring depth limited to 2), full 128-pixel tile, smallest D
    (1, 20, 3, 8, 24),       # D > W, 16 channels per item
    (2, 40, 9, 300, 20),     # three x tiles, ragged channels
]


@pytest.mark.parametrize("cpw", [8, 4])
@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", WIDE_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_corr_bwd_wide(ops, oracle, shape, relu, cpw, monkeypatch, capfd):
    """corr_bwd_wide.cu (a lane owns 4 pixels x all D <= 24 disparities; decoupled warps on a ring of TMA stages) forced on every
    shape it accepts, at the fp32 bar against the oracle, bit-identical through the concat gradient and from run to run."""
    B, C, H, W, D = shape
    L0, R0 = feat((B, C, H, W), 51, relu), feat((B, C, H, W), 52, relu)
    g = feat((B, D, H, W), 53, False)
    monkeypatch.setenv("DSSMD_CORR_BWD_WIDE", "2")
    monkeypatch.setenv("DSSMD_CORR_BWD_WIDE_CPW", str(cpw))
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    L, R = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
    ops.build_corr(L, R, D).backward(g)
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    assert err.count("corr_bwd wide cfg") == 1, err
    gL, gR = oracle.corr_bwd(g.cpu().numpy(), L0.cpu().numpy(), R0.cpu().numpy())
    assert rel_err(L.grad.cpu().numpy(), gL) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), gR) <= TOL_F32
    head = feat((B, 3, H, W), 54, False).requires_grad_(True)
    L2, R2 = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
    gcat = torch.cat((feat((B, 3, H, W), 55, False), g), dim=1)
    ops.build_corr_cat(head, L2, R2, D).backward(gcat)
    assert torch.equal(L2.grad, L.grad) and torch.equal(R2.grad, R.grad)


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", MMA32_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_corr_bwd_mma32(ops, oracle, shape, relu, monkeypatch, capfd):
    """DSSMD_CORR_BWD_VARIANT=3: both gradients on mma.sync.m16n8k8 with 3xTF32 splitting of the operand rows and of the banded g
    operand, at the fp32 bar, plain and through the concat gradient."""
    B, C, H, W, D = shape
    L0, R0 = feat((B, C, H, W), 41, relu), feat((B, C, H, W), 42, relu)
    g = feat((B, D, H, W), 43, False)
    monkeypatch.setenv("DSSMD_CORR_BWD_VARIANT", "3")
    monkeypatch.setenv("DSSMD_CORR_DEBUG", "1")
    L, R = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
    ops.build_corr(L, R, D).backward(g)
    torch.cuda.synchronize()
    err = capfd.readouterr().err
    assert err.count("corr_bwd mma32 cfg") == 2, err
    gL, gR = oracle.corr_bwd(g.cpu().numpy(), L0.cpu().numpy(), R0.cpu().numpy())
    assert rel_err(L.grad.cpu().numpy(), gL) <= TOL_F32
    assert rel_err(R.grad.cpu().numpy(), gR) <= TOL_F32
    head = feat((B, 3, H, W), 44, False).requires_grad_(True)
    L2, R2 = L0.clone().requires_grad_(True), R0.clone().requires_grad_(True)
    gcat = torch.cat((feat((B, 3, H, W), 45, False), g), dim=1)
    ops.build_corr_cat(head, L2, R2, D).backward(gcat)
    assert torch.equal(L2.grad, L.grad) and torch.equal(R2.grad, R.grad)
