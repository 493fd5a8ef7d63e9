"""Size-independent properties of the CPU oracle (adjointness, shift equivariance, A(1) = 1, constants
preserved by the resamplers), drawn with hypothesis over shapes, kernels and rates.  float64, so the bounds are tight; the
oracle is test infrastructure and these properties are what the full-size GPU tests rely on where no fixture exists."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as orc

KERNELS = ["Gaussian_R1", "Gaussian_R2", "Gaussian_R3", "Box_R2", "Box_R3", "Box_R4"]      # src/physics/kernels.py:3-10
FEW = settings(max_examples=12, deadline=None, derandomize=True)


def _rng(seed):
    return np.random.default_rng(seed)


@FEW
@given(name=st.sampled_from(KERNELS), h=st.integers(20, 40), w=st.integers(20, 40), seed=st.integers(0, 1000))
def test_circular_blur_adjoint_shift_and_dc(name, h, w, seed):
    k = orc.named_kernel(name)
    r = _rng(seed)
    x, v = r.standard_normal((2, 2, h, w)), r.standard_normal((2, 2, h, w))
    Ax, Atv = orc.blur_circular(x, k), orc.blur_circular(v, k, adjoint=True)
    assert abs((Ax * v).sum() - (x * Atv).sum()) < 1e-10 * np.abs(Ax * v).sum()               # <Ax, v> = <x, A^T v>
    sh = (int(r.integers(-h, h)), int(r.integers(-w, w)))
    assert np.abs(orc.blur_circular(np.roll(x, sh, (-2, -1)), k) - np.roll(Ax, sh, (-2, -1))).max() < 1e-12
    assert np.abs(orc.blur_circular(np.ones((1, 1, h, w)), k) - 1).max() < 1e-12                # taps sum to 1
    assert np.abs(Atv - orc.blur_circular(v, k)).max() < 1e-12                                 # symmetric kernels: A^T = A


@FEW
@given(name=st.sampled_from(KERNELS), padding=st.sampled_from(["valid", "circular", "replicate", "reflect", "zero"]),
       h=st.integers(22, 36), w=st.integers(22, 36), seed=st.integers(0, 1000))
def test_padded_blur_and_its_transpose_are_adjoint(name, padding, h, w, seed):
    f = np.ascontiguousarray(orc.named_kernel(name)[None, None].astype(np.float32))
    r = _rng(seed)
    x = r.standard_normal((1, 2, h, w)).astype(np.float32)
    y = orc.conv_v1(x, f, padding)
    v = r.standard_normal(y.shape).astype(np.float32)
    xt = orc.conv_transpose_v1(v, f, padding)
    assert xt.shape == x.shape
    lhs, rhs = (y.astype(np.float64) * v).sum(), (x.astype(np.float64) * xt).sum()
    assert abs(lhs - rhs) < 2e-4 * max(1.0, np.abs(y.astype(np.float64) * v).sum())


@FEW
@given(rate=st.sampled_from([2, 3, 4]), hb=st.integers(3, 9), wb=st.integers(3, 9), seed=st.integers(0, 1000))
def test_antialiased_decimation_adjoint_and_dc(rate, hb, wb, seed):
    h, w = hb * rate, wb * rate
    r = _rng(seed)
    x, v = r.standard_normal((1, 2, h, w)), r.standard_normal((1, 2, hb, wb))
    Ax, Atv = orc.down_aa(x, rate), orc.down_aa_vjp(v, rate, (h, w))
    assert Ax.shape == v.shape and Atv.shape == x.shape
    assert abs((Ax * v).sum() - (x * Atv).sum()) < 1e-10 * max(1.0, np.abs(Ax * v).sum())
    assert np.abs(orc.down_aa(np.ones((1, 1, h, w)), rate) - 1).max() < 1e-12                  # weights renormalised per output
    assert np.abs(orc.up_bicubic(np.ones((1, 1, hb, wb)), rate) - 1).max() < 1e-12


@FEW
@given(S=st.integers(12, 40), B=st.integers(1, 3), seed=st.integers(0, 1000))
def test_scale_transform_properties(S, B, seed):
    r = _rng(seed)
    x = r.standard_normal((B, 2, S, S))
    rate = r.choice([0.75, 0.5], size=B)
    center = 2 * r.random((B, 2)) - 1
    T = orc.scale_transform(x, rate, center)
    assert T.shape == x.shape
    assert np.abs(orc.scale_transform(np.full_like(x, 0.37), rate, center) - 0.37).max() < 1e-12   # cubic taps sum to 1
    v = r.standard_normal(x.shape)
    Ttv = orc.scale_transform_vjp(v, rate, center)
    assert abs((T * v).sum() - (x * Ttv).sum()) < 1e-10 * max(1.0, np.abs(T * v).sum())
    # rate 1, centre 0: the grid is 2j/S - 1 (not an identity): pixel j reads position j (S-1)/S, so the top-left sample is kept
    ident = orc.scale_transform(x, np.ones(B), np.zeros((B, 2)))
    assert np.abs(ident[..., 0, 0] - x[..., 0, 0]).max() < 1e-12
    # images of a batch never mix
    if B > 1:
        assert np.abs(orc.scale_transform(x[:1], rate[:1], center[:1]) - T[:1]).max() == 0.0


@FEW
@given(h=st.integers(9, 40), w=st.integers(9, 40), rate=st.sampled_from([0.75, 0.5]), aa=st.booleans())
def test_bicubic_resize_preserves_constants(h, w, rate, aa):
    y = orc.resize_bicubic(np.full((1, 2, h, w), 1.5), rate, aa)
    assert y.shape == (1, 2, int(np.floor(h * rate)), int(np.floor(w * rate)))
    assert np.abs(y - 1.5).max() < 1e-12


@FEW
@given(h=st.integers(5, 33), w=st.integers(5, 33), seed=st.integers(0, 1000))
def test_rotation_by_quarter_turns_is_exact(h, w, seed):
    x = _rng(seed).standard_normal((1, 2, h, w)).astype(np.float32)
    assert np.array_equal(orc.rotate_nearest(x, 180.0), x[..., ::-1, ::-1])
    if h == w:
        # quarter turns of a square image only move pixels (torchvision: positive angles are counter-clockwise)
        assert np.array_equal(orc.rotate_nearest(x, 90.0), np.rot90(x, 1, (-2, -1)))
        assert np.array_equal(orc.rotate_nearest(x, 270.0), np.rot90(x, -1, (-2, -1)))
    y = orc.rotate_nearest(x, 37.0)
    assert y.shape == x.shape and set(np.unique(y)) <= set(np.unique(x)) | {0.0}     # nearest neighbour: no new values


@FEW
@given(seed=st.integers(0, 1000), margin=st.integers(0, 3))
def test_sure_loss_pieces(seed, margin):
    """the reductions of SureGaussianLoss (src/losses/sure.py:35-76): mse over the interior, divergence estimate, constant"""
    r = _rng(seed)
    S = 12
    y1, y2, y, b = (r.standard_normal((2, 1, S, S)) for _ in range(4))
    tau, sigma2 = 1e-2, (5 / 255) ** 2
    loss, mse_v, div_v = orc.sure_loss(y1, y2, y, b, margin, margin, tau, sigma2, None)
    sl = (slice(None), slice(None), slice(margin, S - margin), slice(margin, S - margin)) if margin else (slice(None),) * 4
    assert abs(mse_v - ((y1 - y)[sl] ** 2).mean()) < 1e-12
    assert abs(div_v - (b[sl] * (y2 - y1)[sl] / tau).mean()) < 1e-9 * max(1.0, abs(div_v))      # mc_div (:7-32)
    assert abs(loss - (mse_v + 2 * sigma2 * div_v - sigma2 / 2)) < 1e-12             # sigma^2 / y.size(0), the reference's quirk (:66)
