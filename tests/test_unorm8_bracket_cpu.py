"""The property the UNORM8 blend's late start rests on (vkgs_b200/csrc/blend.cu, phase B), checked on the CPU:
one ROP step q -> rint(fma(s, a, q * (1 - a))) (engine.cc:281-291 blend equation on a B8G8R8A8_UNORM target,
render_pass.cc:15) is monotone non-decreasing in the destination value q, so for ANY run of splats F and any
destination q in [0, 255]:  F(0) <= F(q) <= F(255) - and where F(0) == F(255) the result does not depend on what lies
behind the run.  The magic-number rounding (x + 1.5 * 2^23) - 1.5 * 2^23 used on the device equals rintf on [0, 2^22)."""
import numpy as np

f32 = np.float32


def rop_step(q, s255, a):
    """q, a: float32 arrays; the device's operation order: om = 1 - a (rounded), t = q * om (rounded), fma(s255, a, t)
    (one rounding; emulated in float64: the product of two binary32 values is exact in binary64), rint."""
    om = (f32(1.0) - a).astype(f32)
    t = (q * om).astype(f32)
    x = (s255.astype(np.float64) * a.astype(np.float64) + t.astype(np.float64)).astype(f32)
    return np.rint(x).astype(f32)


def test_magic_rounding_equals_rint():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(200_000, dtype=f32) * f32(256.0), np.arange(0, 512, dtype=f32) * f32(0.5),
                        (np.arange(0, 256, dtype=f32) + f32(0.5)), rng.random(1000, dtype=f32) * f32(4.0e6)])
    magic = f32(12582912.0)
    y = ((x + magic).astype(f32) - magic).astype(f32)
    assert np.array_equal(y, np.rint(x).astype(f32))


def test_rop_step_is_monotone_in_destination():
    rng = np.random.default_rng(2)
    q = np.arange(256, dtype=f32)
    for _ in range(2000):
        a = f32(rng.random()) if rng.random() < 0.8 else f32(rng.choice([0.0, 1.0, 1e-4, 0.5, 0.999]))
        s = f32(rng.random() * 255.0)
        out = rop_step(q, np.full(256, s, f32), np.full(256, a, f32))
        assert np.all(np.diff(out) >= 0)
        assert out.min() >= 0 and out.max() <= 255


def test_bracket_encloses_every_destination_and_closes_behind_opaque_runs():
    rng = np.random.default_rng(3)
    closed = 0
    for trial in range(300):
        n = int(rng.integers(1, 60))
        # bimodal opacities like the synthetic scenes: mostly opaque cores, some faint layers
        a = np.where(rng.random(n) < 0.6, rng.random(n) * 0.5 + 0.5, rng.random(n) * 0.05).astype(f32)
        s = (rng.random(n) * 255.0).astype(f32)
        q = np.arange(256, dtype=f32)                    # every possible destination value at once
        for i in range(n):                               # back to front
            q = rop_step(q, np.full(256, s[i], f32), np.full(256, a[i], f32))
        assert np.all(np.diff(q) >= 0)                   # F is monotone: F(0) <= F(q) <= F(255)
        assert q[0] <= q.min() and q[255] >= q.max()
        if q[0] == q[255]:
            closed += 1
            assert np.all(q == q[0])                     # the bracket met: the run hides everything behind it exactly
    assert closed > 100


def test_faint_layers_can_keep_the_bracket_open():
    """Why the device checks instead of assuming: behind many faint layers an 8-bit destination 'sticks', so a
    transmittance threshold alone does not prove that what lies behind is hidden."""
    n = 4000
    a = np.full(n, 0.004, f32)                           # T = 0.996^4000 ~ 1e-7, far below any cut
    s = np.full(n, 128.0, f32)
    q = np.array([0.0, 255.0], f32)
    for i in range(n):
        q = rop_step(q, s[i:i + 1].repeat(2), a[i:i + 1].repeat(2))
    assert q[0] < q[1]                                   # fp32 blending would give 128 for both
