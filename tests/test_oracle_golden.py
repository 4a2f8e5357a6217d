"""Pins the CPU oracle against outputs of the upstream functions themselves.

tests/golden/*.npz were produced by oracle/gen_golden.py, which imports
build_corr / disp_warp / LRwarp_error from the upstream tree and runs them (and
their autograd backward) on CPU.  The upstream repo has no tests of its own for
this path, so these fixtures are the pin.
"""
import numpy as np
import pytest

from helpers import TOL_BF16, TOL_F32, load_cases, rel_err
from oracle import oracle as O

CORR = load_cases("corr")
WARP = load_cases("warp")
LRW = load_cases("lrwarp")


@pytest.mark.parametrize("name", sorted(CORR))
def test_corr_forward_backward_matches_reference(name):
    c = CORR[name]
    tol = TOL_BF16 if bool(c["bf16"]) else TOL_F32
    D = int(c["D"])
    out = O.corr_fwd(c["L"], c["R"], D)
    assert out.shape == c["out"].shape and out.flags["C_CONTIGUOUS"]
    assert rel_err(out, c["out"]) <= tol
    gL, gR = O.corr_bwd(c["g"], c["L"], c["R"])
    assert rel_err(gL, c["gL"]) <= tol
    assert rel_err(gR, c["gR"]) <= tol


@pytest.mark.parametrize("name", sorted(CORR))
def test_corr_f64_arbiter_agrees(name):
    c = CORR[name]
    if bool(c["bf16"]):
        pytest.skip("fp64 arbiter is for the fp32 cases")
    out64 = O.corr_fwd(c["L"].astype(np.float64), c["R"].astype(np.float64), int(c["D"]))
    assert rel_err(out64, c["out"]) <= TOL_F32


@pytest.mark.parametrize("name", sorted(WARP))
def test_warp_forward_backward_matches_reference(name):
    c = WARP[name]
    pad = str(c["pad"])
    warped, mask, bad = O.warp_fwd(c["img"], c["disp"], pad)
    assert not bad
    assert rel_err(warped, c["warped"]) <= TOL_F32
    assert np.array_equal(mask, c["mask"])  # 0/1 mask must be exact
    g_img, g_disp = O.warp_bwd(c["g"], c["img"], c["disp"], pad)
    assert rel_err(g_img, c["g_img"]) <= TOL_F32
    assert rel_err(g_disp, c["g_disp"]) <= TOL_F32


@pytest.mark.parametrize("name", sorted(LRW))
def test_lrwarp_error_matches_reference(name):
    c = LRW[name]
    err = O.lrwarp_error(c["imgL"], c["disp"], c["imgR"])
    assert err.shape == c["err"].shape
    assert rel_err(err, c["err"]) <= TOL_F32


def test_negative_or_nan_disparity_is_flagged():
    img = np.random.default_rng(0).random((1, 3, 4, 6), dtype=np.float32)
    disp = np.ones((1, 1, 4, 6), np.float32)
    assert not O.warp_fwd(img, disp)[2]
    d2 = disp.copy(); d2[0, 0, 2, 3] = -0.25
    assert O.warp_fwd(img, d2)[2]
    d3 = disp.copy(); d3[0, 0, 1, 1] = np.nan
    assert O.warp_fwd(img, d3)[2]


# ---- invariants of SURVEY.md Appendix B, checked on the oracle ----------------
def test_corr_invariants():
    rng = np.random.default_rng(1024)
    L = rng.standard_normal((2, 8, 3, 6)).astype(np.float32)
    R = rng.standard_normal((2, 8, 3, 6)).astype(np.float32)
    out = O.corr_fwd(L, R, 9)
    assert out.shape == (2, 9, 3, 6)
    for d in range(9):
        assert np.all(out[:, d, :, :d] == 0)          # zero left border
    assert np.all(out[:, 6:] == 0)                    # channels d >= W are zero
    assert rel_err(out[:, 0], (L * R).mean(1)) <= TOL_F32
    same = O.corr_fwd(L, L, 3)[:, 0]
    assert np.all(same >= 0)
    # shift test: R[..., x] = L[..., x+k]  =>  channel k == mean_c L^2 on x >= k
    k = 2
    Rs = np.zeros_like(L); Rs[..., :-k] = L[..., k:]
    outs = O.corr_fwd(L, Rs, 4)
    assert rel_err(outs[:, k, :, k:], (L[..., k:] ** 2).mean(1)) <= TOL_F32


def test_warp_invariants():
    rng = np.random.default_rng(7)
    img = rng.random((2, 3, 10, 24), dtype=np.float32)
    disp = (rng.random((2, 1, 10, 24), dtype=np.float32) * 5).astype(np.float32)
    warped, mask, _ = O.warp_fwd(img, disp)
    assert set(np.unique(mask)) <= {0.0, 1.0}
    assert np.all(mask[:, :, 0] == 0) and np.all(mask[:, :, -1] == 0)   # first/last rows invalid
    assert np.all(mask[:, 0] == mask[:, 1]) and np.all(mask[:, 0] == mask[:, 2])
    # disp = 0 is still a resample, not the identity
    w0, _, _ = O.warp_fwd(img, np.zeros_like(disp))
    assert np.abs(w0 - img).max() > 0.05


# ---- photometric L1 on the warp residual -------------
PHOTO = load_cases("photo")


@pytest.mark.parametrize("name", sorted(PHOTO))
def test_photometric_l1_matches_reference(name):
    """The fixture holds sum_s w_s * LRwarp_error(imgL, disp_s, imgR).abs().mean() and the autograd gradient of
    every disparity map, produced by upstream's LRwarp_error (oracle/gen_golden.py: gen_photo)."""
    c = PHOTO[name]
    total = 0.0
    for k in range(int(c["n"])):
        loss, g_disp = O.photometric_l1(c["imgL"], c[f"disp{k}"], c["imgR"])
        assert abs(loss - float(c["terms"][k])) <= TOL_F32 * abs(float(c["terms"][k]))
        w = float(c["weights"][k])
        assert rel_err(w * g_disp, c[f"g_disp{k}"]) <= TOL_F32
        total += w * loss
    assert abs(total - float(c["total"])) <= TOL_F32 * abs(float(c["total"]))


# ---- haze synthesis ----------------------------------
HAZE = load_cases("haze")


@pytest.mark.parametrize("name", sorted(HAZE))
def test_haze_chain_matches_reference(name):
    """Fixtures: upstream's convert_disp_to_depth -> depth2trans -> recover_haze_images (DeBug/inference.py) on the dtypes
    the data loaders hand over (oracle/gen_golden.py: gen_haze).  Result dtype follows upstream's type promotion."""
    c = HAZE[name]
    out = O.haze(c["clear"], c["disp"], True, c["baseline"], c["focal"], c["beta"], c["airlight"])
    for k in ("depth", "trans", "haze"):
        assert out[k].dtype == c[k].dtype and out[k].shape == c[k].shape, k
        assert rel_err(out[k], c[k]) <= (TOL_F32 if c[k].dtype == np.float32 else 1e-14), k
    # the three upstream entry points one by one (depth given instead of disparity)
    t = O.haze(None, c["depth"], False, None, None, c["beta"], None, want=("trans",))["trans"]
    assert rel_err(t, c["trans"]) <= (TOL_F32 if c["trans"].dtype == np.float32 else 1e-14)
    h = O.haze(c["clear"], c["depth"], False, None, None, c["beta"], c["airlight"], want=("haze",))["haze"]
    assert rel_err(h, c["haze"]) <= (TOL_F32 if c["haze"].dtype == np.float32 else 1e-14)


def test_haze_invariants():
    rng = np.random.default_rng(5)
    clear = rng.random((2, 3, 6, 8), dtype=np.float32)
    disp = (0.5 + rng.random((2, 1, 6, 8), dtype=np.float32) * 50).astype(np.float32)
    one = np.ones(2, np.float32)
    # beta = 0: transmission 1, the hazy image is the clear image
    o = O.haze(clear, disp, True, one * 0.1, one * 1050, one * 0, one * 0.9)
    assert np.all(o["trans"] == 1) and np.array_equal(o["haze"], clear)
    # transmission in (0, 1], hazy image between the clear image and the airlight
    o = O.haze(clear, disp, True, one * 0.1, one * 1050, one * 0.05, one * 0.9)
    assert np.all((o["trans"] > 0) & (o["trans"] <= 1))
    lo, hi = np.minimum(clear, 0.9), np.maximum(clear, 0.9)
    assert np.all((o["haze"] >= lo - 1e-6) & (o["haze"] <= hi + 1e-6))
    assert rel_err(o["depth"] * disp, np.full_like(disp, 105.0)) <= TOL_F32


# ---- ASM reconstruction loss -------------------------
ASM = load_cases("asm")


@pytest.mark.parametrize("name", sorted(ASM))
def test_recovered_clean_loss_matches_reference(name):
    """Fixtures: upstream's RecoveredCleanImagesLoss / V2 (losses/DSSMDLoss.py:81-145) and its autograd gradients w.r.t.
    the transmission map and the airlight (oracle/gen_golden.py: gen_asm)."""
    c = ASM[name]
    loss, g_t, g_a = O.recovered_clean_l1(c["trans"], c["airlight"], c["haze"], c["clean"])
    assert abs(loss - float(c["loss"])) <= TOL_F32 * abs(float(c["loss"]))
    assert rel_err(g_t, c["g_trans"]) <= TOL_F32
    assert rel_err(g_a.reshape(c["g_airlight"].shape), c["g_airlight"]) <= TOL_F32


def test_recovered_clean_loss_is_zero_for_the_true_parameters():
    rng = np.random.default_rng(3)
    clean = rng.random((2, 3, 6, 8), dtype=np.float32) * 0.9 + 0.05
    t = (0.3 + 0.6 * rng.random((2, 1, 6, 8), dtype=np.float32)).astype(np.float32)
    A = np.array([0.9, 0.8], np.float32)
    haze = clean * t + A.reshape(2, 1, 1, 1) * (1 - t)
    loss, _, _ = O.recovered_clean_l1(t, A, haze, clean)
    assert loss <= 1e-6


# ---- attention along the epipolar line -----------------
ATTN = load_cases("attn")


@pytest.mark.parametrize("name", sorted(ATTN))
def test_epipolar_attention_matches_reference(name):
    """Fixtures: upstream's STTM_Single (models/BidNet/STTM_Simple.py) run unmodified, q / k / v and the attention output
    captured with hooks, gradients from its autograd graph (oracle/gen_golden.py: gen_attn)."""
    c = ATTN[name]
    out, gq, gk, gv = O.epipolar_attention(c["q"], c["k"], c["v"], float(c["scale"]), c["g_out"])
    assert rel_err(out, c["out"]) <= TOL_F32
    assert rel_err(gq, c["g_q"]) <= TOL_F32
    assert rel_err(gk, c["g_k"]) <= TOL_F32
    assert rel_err(gv, c["g_v"]) <= TOL_F32
    assert np.array_equal(O.epipolar_attention(c["q"], c["k"], c["v"], float(c["scale"])), out)


def test_epipolar_attention_invariants():
    rng = np.random.default_rng(11)
    q = rng.standard_normal((1, 16, 2, 12)).astype(np.float32)
    k = rng.standard_normal((1, 16, 2, 12)).astype(np.float32)
    v = rng.standard_normal((1, 16, 2, 12)).astype(np.float32)
    # scale 0: uniform attention, the output is the row mean of v
    out = O.epipolar_attention(q, k, v, 0.0)
    assert rel_err(out, np.broadcast_to(v.mean(-1, keepdims=True), v.shape)) <= TOL_F32
    # the output is a convex combination of the row's values
    out = O.epipolar_attention(q, k, v, 0.25)
    assert np.all(out <= v.max(-1, keepdims=True) + 1e-6) and np.all(out >= v.min(-1, keepdims=True) - 1e-6)
    # rows are independent: changing row 1 of k leaves row 0 of the output alone
    k2 = k.copy(); k2[:, :, 1] += 1.0
    assert np.array_equal(O.epipolar_attention(q, k2, v, 0.25)[:, :, 0], out[:, :, 0])
