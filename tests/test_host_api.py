"""Host-side shim: signatures mirror upstream, errors follow upstream's conventions,
install() rebinds the names upstream's models resolve, and nothing falls back to CPU."""
import inspect
import sys
import types

import pytest
import torch

import dssmd_b200
from dssmd_b200 import cost_volume, disparity_warper, stereo_op, submoudles


def test_signatures_match_upstream():
    # upstream's signatures as literals, for boxes without an upstream checkout; tests/test_upstream_dropin.py compares
    # with inspect.signature of the real upstream functions where the tree is present
    assert str(inspect.signature(submoudles.build_corr)) == "(img_left, img_right, max_disp=40)"
    assert str(inspect.signature(disparity_warper.disp_warp)) == "(img, disp, padding_mode='border')"
    assert str(inspect.signature(disparity_warper.LRwarp_error)) == "(imgL, disp, imgR)"
    assert str(inspect.signature(disparity_warper.meshgrid)) == "(img, homogeneous=False)"
    assert str(inspect.signature(disparity_warper.normalize_coords)) == "(grid)"
    assert str(inspect.signature(cost_volume.CostVolume.__init__)) == "(self, max_disp, feature_similarity='correlation')"
    assert str(inspect.signature(cost_volume.CostVolume.forward)) == "(self, left_feature, right_feature)"
    for name in ("disp_warp", "LRwarp_error", "normalize_coords", "meshgrid", "print_tensor_shape"):
        assert hasattr(stereo_op, name)


def test_no_cpu_fallback():
    x = torch.rand(1, 4, 3, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dssmd_b200.build_corr(x, x, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dssmd_b200.disp_warp(torch.rand(1, 3, 4, 8), torch.rand(1, 1, 4, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dssmd_b200.LRwarp_error(torch.rand(1, 3, 4, 8), torch.rand(1, 1, 4, 8), torch.rand(1, 3, 4, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dssmd_b200.CostVolume(3)(x, x)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dssmd_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryNotBuiltError, match="no\\s+CPU/PyTorch fallback"):
        _lib.lib()


def test_pure_tensor_helpers_keep_upstream_semantics():
    img = torch.zeros(2, 3, 4, 5)
    grid = disparity_warper.meshgrid(img)
    assert grid.shape == (2, 2, 4, 5)
    assert torch.equal(grid[0, 0, 0], torch.arange(5.0)) and torch.equal(grid[1, 1, :, 0], torch.arange(4.0))
    assert disparity_warper.meshgrid(img, homogeneous=True).shape == (2, 3, 4, 5)
    norm = disparity_warper.normalize_coords(grid.clone())
    assert norm.shape == (2, 4, 5, 2)
    assert float(norm[..., 0].min()) == -1.0 and float(norm[..., 0].max()) == 1.0
    with pytest.raises(AssertionError):
        disparity_warper.normalize_coords(torch.zeros(1, 3, 4, 4))


def test_read_pfm_roundtrip(tmp_path):
    import numpy as np
    data = np.arange(12, dtype="<f4").reshape(3, 4)
    path = tmp_path / "d.pfm"
    with open(path, "wb") as f:
        f.write(b"Pf\n4 3\n-1.0\n")
        f.write(np.flipud(data).tobytes())
    out, scale = disparity_warper.read_pfm(str(path))
    assert scale == 1.0 and np.array_equal(out, data)


def _fake_upstream():
    """Minimal stand-in for the upstream package layout: defining modules plus a
    model module that star-imports build_corr (models/DSSMD.py:6-9)."""
    def mk(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def ref_build_corr(img_left, img_right, max_disp=40):
        return "upstream"

    def ref_disp_warp(img, disp, padding_mode='border'):
        return "upstream"

    def ref_lr(imgL, disp, imgR):
        return "upstream"

    class RefCV:
        pass

    mk("models"); mk("models.UtilsNet"); mk("utils")
    mk("models.UtilsNet.submoudles", build_corr=ref_build_corr)
    mk("models.UtilsNet.cost_volume", CostVolume=RefCV)
    mk("models.UtilsNet.stereo_op", disp_warp=ref_disp_warp, LRwarp_error=ref_lr)
    mk("utils.disparity_warper", disp_warp=ref_disp_warp, LRwarp_error=ref_lr)
    model = mk("models.DSSMD", build_corr=ref_build_corr, disp_warp=ref_disp_warp, CostVolume=RefCV, other=1)
    return model


def test_install_rebinds_defining_and_importing_modules():
    created = ["models", "models.UtilsNet", "utils", "models.UtilsNet.submoudles", "models.UtilsNet.cost_volume",
               "models.UtilsNet.stereo_op", "utils.disparity_warper", "models.DSSMD"]
    saved = {k: sys.modules.get(k) for k in created}
    try:
        model = _fake_upstream()
        patched = dssmd_b200.install(import_missing=False)
        assert ("models.UtilsNet.submoudles", "build_corr") in patched
        assert ("models.DSSMD", "build_corr") in patched
        assert sys.modules["models.UtilsNet.submoudles"].build_corr is dssmd_b200.build_corr
        assert model.build_corr is dssmd_b200.build_corr
        assert model.disp_warp is dssmd_b200.disp_warp
        assert model.CostVolume is dssmd_b200.CostVolume
        assert sys.modules["utils.disparity_warper"].LRwarp_error is dssmd_b200.LRwarp_error
        assert model.other == 1
        dssmd_b200.uninstall()
        assert model.build_corr(None, None) == "upstream"
    finally:
        dssmd_b200.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_haze_signatures_and_errors_follow_upstream():
    from dssmd_b200 import haze
    assert str(inspect.signature(haze.convert_disp_to_depth)) == "(baseline, focal_length, disp)"
    assert str(inspect.signature(haze.depth2trans)) == "(depth, beta)"
    assert str(inspect.signature(haze.recover_haze_images)) == "(clean_images, beta, A, depth)"
    disp = torch.rand(2, 1, 4, 8) + 1
    with pytest.raises(AssertionError):                       # upstream: assert baseline.shape[0] == b
        haze.convert_disp_to_depth(torch.ones(3), torch.ones(2), disp)
    with pytest.raises(AssertionError):
        haze.depth2trans(disp, torch.ones(1))
    with pytest.raises(AssertionError):
        haze.recover_haze_images(torch.rand(2, 3, 4, 8), torch.ones(2), torch.ones(5), disp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        haze.convert_disp_to_depth(torch.ones(2), torch.ones(2), disp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        haze.synthesize_haze(torch.rand(2, 3, 4, 8), disp, torch.ones(2), torch.ones(2), torch.ones(2), torch.ones(2))


def test_install_rebinds_haze_functions_in_trainers():
    names = ["DeBug", "DeBug.inference", "trainfiles", "trainfiles.dssmd_traner_new", "losses", "losses.DSSMDLoss"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        ups = {k: (lambda *a, _k=k, **kw: _k) for k in ("convert_disp_to_depth", "depth2trans", "recover_haze_images", "recover_depth")}
        up = ups["depth2trans"]
        for n in names:
            sys.modules[n] = types.ModuleType(n)
        for n in ("DeBug.inference", "trainfiles.dssmd_traner_new"):
            sys.modules[n].__dict__.update(ups)
        class UpLoss:
            loss_type = 'normal'
            def forward(self, transmission_map, airlight, haze_image, clean_image):
                return "upstream"
        up_forward = UpLoss.forward
        sys.modules["losses.DSSMDLoss"].RecoveredCleanImagesLoss = UpLoss
        sys.modules["trainfiles.dssmd_traner_new"].RecoveredCleanImagesLoss = UpLoss
        dssmd_b200.install(import_missing=False)
        tr = sys.modules["trainfiles.dssmd_traner_new"]
        # guarded patches: upstream's class keeps its identity, its forward takes the kernel where the kernel applies
        assert tr.RecoveredCleanImagesLoss is UpLoss and UpLoss.forward.__wrapped_upstream__ is up_forward
        cpu = [torch.rand(1, 1, 4, 8), torch.rand(1, 1), torch.rand(1, 3, 4, 8), torch.rand(1, 3, 4, 8)]
        assert UpLoss().forward(*cpu) == "upstream"                # CPU tensors: outside the kernel's limits -> upstream's code
        for n in ("convert_disp_to_depth", "depth2trans", "recover_haze_images"):
            assert getattr(tr, n).__kernel__ is getattr(dssmd_b200, n) and getattr(tr, n).__wrapped_upstream__ is ups[n]
        assert tr.depth2trans(torch.rand(2, 1, 4, 8), torch.ones(2)) == "depth2trans"   # CPU -> upstream's function
        assert tr.recover_depth is ups["recover_depth"]            # not on the path: untouched
        dssmd_b200.install(import_missing=False)                   # idempotent: no double wrapping
        assert UpLoss.forward.__wrapped_upstream__ is up_forward
        dssmd_b200.uninstall()
        assert tr.depth2trans is up and UpLoss.forward is up_forward
    finally:
        dssmd_b200.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_asm_loss_signature_and_no_cpu_fallback():
    from dssmd_b200 import losses
    assert str(inspect.signature(losses.RecoveredCleanImagesLoss.__init__)) == "(self, type='normal')"
    assert str(inspect.signature(losses.RecoveredCleanImagesLoss.forward)) == "(self, transmission_map, airlight, haze_image, clean_image)"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.RecoveredCleanImagesLoss()(torch.rand(1, 1, 4, 8), torch.rand(1, 1), torch.rand(1, 3, 4, 8), torch.rand(1, 3, 4, 8))


def test_build_corr_cat_signature_and_no_cpu_fallback():
    assert str(inspect.signature(submoudles.build_corr_cat)) == "(head, img_left, img_right, max_disp=40)"
    x = torch.rand(1, 4, 3, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dssmd_b200.build_corr_cat(x, x, x, 3)


def test_sttm_single_signature_state_dict_and_no_cpu_fallback():
    from dssmd_b200 import sttm
    assert str(inspect.signature(sttm.STTM_Single.__init__)) == "(self, dim, heads=8, dim_head=64, out_dim=256)"
    assert str(inspect.signature(sttm.STTM_Single.forward)) == "(self, left_feat, right_feat)"
    m = sttm.STTM_Single(dim=32, heads=1, dim_head=32, out_dim=32)
    keys = list(m.state_dict().keys())
    for k in ("q.weight", "k.weight", "v.weight", "norm.weight", "norm.bias", "to_out.0.weight", "to_out.1.running_mean", "to_out.3.weight"):
        assert k in keys
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 32, 2, 8), torch.randn(1, 32, 2, 8))
