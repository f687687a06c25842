"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz).

The reference has no tests of its own; the fixtures were produced by
tests/golden/make_golden.py importing /root/reference.  fp32 reference vs fp64 oracle:
tolerance 2e-5 relative (norm-wise), well inside the 1e-3 parity bar of BASELINE.json.
"""
import glob
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import dwt_oracle as O

TOL = 2e-5
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WHITEN = sorted(glob.glob(os.path.join(HERE, "w_*.npz")))


def f64(a):
    return np.asarray(a, dtype=np.float64)


@pytest.mark.parametrize("path", WHITEN, ids=[os.path.basename(p)[:-4] for p in WHITEN])
def test_whitening_matches_reference(path):
    z = np.load(path)
    gs = int(z["gs"])
    x1, x2, dy = f64(z["x1"]), f64(z["x2"]), f64(z["dy"])
    rm0, rv0 = f64(z["rm0"]).reshape(-1), f64(z["rv0"])
    y1, mean, w, rm1, rv1, _ = O.whiten_forward(x1, gs, running_mean=rm0, running_cov=rv0)
    assert rel_err(y1, z["y1"]) < TOL
    assert np.allclose(np.triu(w, 1), 0)                      # Cholesky basis, not ZCA
    assert rel_err(O.whiten_backward(x1, dy, mean, w), z["dx1"]) < 20 * TOL
    assert rel_err(rm1, z["rm1"].reshape(-1)) < TOL and rel_err(rv1, z["rv1"]) < TOL
    y2, _, _, rm2, rv2, _ = O.whiten_forward(x2, gs, running_mean=rm1, running_cov=rv1)
    assert rel_err(y2, z["y2"]) < TOL
    assert rel_err(rm2, z["rm2"].reshape(-1)) < TOL and rel_err(rv2, z["rv2"]) < TOL
    ye, _, we, rme, _, _ = O.whiten_forward(x1, gs, running_mean=rm2, running_cov=rv2, training=False)
    assert rel_err(ye, z["y_eval"]) < TOL
    assert rel_err(O.whiten_backward_eval(dy, we), z["dx_eval"]) < TOL
    assert np.array_equal(z["rm_eval"], z["rm2"]) and rme is rm2   # eval never touches the buffers
    # default buffers: zeros and an ALL-ONES matrix, not identity (whitening.py:23-24)
    c = x1.shape[1]
    gse = min(c, gs)
    _, _, _, rmd, rvd, _ = O.whiten_forward(x1, gs, running_mean=np.zeros(c), running_cov=np.ones((c // gse, gse, gse)))
    assert rel_err(rmd, z["rm_default1"].reshape(-1)) < TOL and rel_err(rvd, z["rv_default1"]) < TOL


@pytest.mark.parametrize("name", ["plain", "big", "k10", "n1", "ties"])
def test_mec_matches_reference(name):
    z = np.load(os.path.join(HERE, "mec.npz"))
    loss, gx, gy, _ = O.mec_loss(f64(z[name + "_x"]), f64(z[name + "_y"]))
    assert abs(loss - float(z[name + "_loss"])) < 1e-5 * max(1.0, abs(loss))
    assert rel_err(gx, z[name + "_gx"]) < TOL and rel_err(gy, z[name + "_gy"]) < TOL


@pytest.mark.parametrize("name", ["bn2d_affine", "bn2d_plain", "bn2d_hw4", "bn2d_cma", "bn1d_2", "bn1d_3", "bn3d"])
def test_bn_matches_reference(name):
    z = np.load(os.path.join(HERE, "bn.npz"))
    g = lambda k: f64(z[f"{name}_{k}"])
    w = g("weight") if f"{name}_weight" in z else None
    b = g("bias") if f"{name}_bias" in z else None
    mom = {"bn2d_cma": None, "bn3d": 0.3}.get(name, 0.1)
    f1 = 1.0 if mom is None else mom
    f2 = 0.5 if mom is None else mom
    y1, mean, invstd, rm1, rv1 = O.bn_forward(g("x1"), g("rm0"), g("rv0"), w, b, True, f1)
    assert rel_err(y1, g("y1")) < TOL and rel_err(rm1, g("rm1")) < TOL and rel_err(rv1, g("rv1")) < TOL
    dx, dw, db = O.bn_backward(g("x1"), g("dy"), mean, invstd, w)
    assert rel_err(dx, g("dx1")) < 10 * TOL
    if w is not None:
        assert rel_err(dw, g("dweight")) < TOL and rel_err(db, g("dbias")) < TOL
    _, _, _, rm2, rv2 = O.bn_forward(g("x2"), rm1, rv1, w, b, True, f2)
    assert rel_err(rm2, g("rm2")) < TOL and rel_err(rv2, g("rv2")) < TOL and int(z[f"{name}_nbt2"]) == 2
    ye, me, ie, _, _ = O.bn_forward(g("x1"), rm2, rv2, w, b, False)
    assert rel_err(ye, g("y_eval")) < TOL
    assert rel_err(O.bn_backward(g("x1"), g("dy"), me, ie, w, training=False)[0], g("dx_eval")) < TOL


def test_closed_form_backward_matches_autograd_fp64():
    """SURVEY §8a closed form vs torch autograd through the port, in fp64."""
    import torch
    from oracle import torch_port as P
    rng = np.random.default_rng(0)
    for (n, c, h, gs) in [(5, 8, 3, 4), (6, 16, 4, 16), (7, 6, 5, 1)]:
        x = rng.standard_normal((n, c, h, h)) @ np.eye(h) + 1.5
        x = x + 0.5 * np.roll(x, 1, axis=1)
        dy = rng.standard_normal(x.shape)
        m = P.WTransform2d(c, gs).double().train()
        xt = torch.tensor(x, requires_grad=True)
        (dx_ref,) = torch.autograd.grad(m(xt), xt, torch.tensor(dy))
        _, mean, w, *_ = O.whiten_forward(x, gs)
        assert rel_err(O.whiten_backward(x, dy, mean, w), dx_ref.numpy()) < 1e-11


# --------------------------------------------------------------------------- paired augmentation (§8f-4)
@pytest.mark.parametrize("case", ["ref", "strong", "tiny"])
def test_augmentation_oracle_is_bit_exact_vs_reference(case):
    """tests/golden/augment.npz = the reference's own _random_affine_augmentation / _gaussian_blur inside the
    torchvision pipeline; the numpy restatement (fixed-point cv2.warpAffine and all) must agree bit for bit."""
    from oracle import augment_oracle as A
    z = np.load(os.path.join(HERE, "augment.npz"))
    g = lambda k: z[f"{case}/{k}"]                                  # noqa: E731
    plain, aug = A.paired(g("images"), g("crop_plain"), g("crop_aug"), g("flip"), g("affine"), int(g("crop")))
    assert plain.dtype == np.float32 and aug.dtype == np.float32
    assert np.array_equal(plain, g("plain")) and np.array_equal(aug, g("aug"))


def test_augmentation_draws_have_the_reference_distributions():
    """Host logic of the product's draw_params: shapes, dtypes, ranges, moments (resnet50_dwt_mec_officehome.py:481-483,
    535-537: crop corners uniform, flip p = 0.5, affine = I + N(0, 0.1) with a zero translation column)."""
    import torch
    from dwt_b200.augment import draw_params
    p = draw_params(4096, 256, 224, np.random.default_rng(0))
    assert p["crop_plain"].dtype == torch.int32 and tuple(p["crop_aug"].shape) == (4096, 2)
    assert int(p["crop_plain"].min()) == 0 and int(p["crop_plain"].max()) == 32
    assert not torch.equal(p["crop_plain"], p["crop_aug"])           # the two views crop independently
    assert p["flip"].dtype == torch.uint8 and 0.45 < float(p["flip"].float().mean()) < 0.55
    a = p["affine"].numpy()
    assert a.dtype == np.float32 and a.shape == (4096, 2, 3) and np.all(a[:, :, 2] == 0)
    dev = a[:, :, :2] - np.eye(2, dtype=np.float32)
    assert abs(dev.mean()) < 5e-3 and 0.095 < dev.std() < 0.105
