"""
Golden trajectories at the BENCHMARK shapes, from the REAL reference (/root/reference/accbpg) -> tests/golden/full_shape.npz

TEST INFRASTRUCTURE.  Run in the build container only (about 20 minutes of CPU):   python oracle/gen_golden_full.py
Only seeds and the reference's numerical outputs are stored (the instances are regenerated from the legacy NumPy seed
by the tests): F / gain / L_k arrays are a few kB.

  c2_*      D_opt_design(500, 50000, randseed=1)  (BASELINE.json configs[1], the bench instance)
            ABPG gamma=2 theta_eq=False 300 it; ABPG_gain gamma=2 150 it; D_opt_FW_away 1200 it; BPG line search 100 it
            and, next to ABPG_gain, the reference's own +-1 ulp noise curve (every operator output of the reference
            multiplied by 1 + k * 1.1e-16, k in {-1, 0, 1}): |F_perturbed - F| / |F| at every iteration
  m2000_*   D_opt_design(2000, 20000, randseed=2), ABPG_gain gamma=2 50 it     (the m of configs[4])
  kl_*      KL regression 2000 x 20000 + ShannonEntropySimplex, x* on the simplex (configs[2] family), ABPG_gain 100 it
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "full_shape.npz")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def noisy(f, h, rng):
    """The reference's f / h with every operator output perturbed by at most one ulp."""
    def nz(v):
        if np.ndim(v) == 0:
            return v * (1 + rng.randint(-1, 2) * 1.1e-16)
        return v * (1 + rng.randint(-1, 2, size=np.shape(v)) * 1.1e-16)

    class NF:
        def func_grad(self, x, flag=2):
            r = f.func_grad(x, flag)
            return (nz(r[0]), nz(r[1])) if flag == 2 else nz(r)

        def __call__(self, x):
            return self.func_grad(x, 0)

        def gradient(self, x):
            return self.func_grad(x, 1)

    class NH:
        def extra_Psi(self, x):
            return nz(h.extra_Psi(x))

        def divergence(self, x, y):
            return nz(h.divergence(x, y))

        def div_prox_map(self, y, g, L):
            return nz(h.div_prox_map(y, g, L))

        def prox_map(self, g, L):
            return nz(h.prox_map(g, L))
    return NF(), NH()


def kl_instance(m, n, seed):
    np.random.seed(seed)
    A = np.random.rand(m, n)
    A = A / A.sum(axis=0)
    xs = np.random.rand(n)
    xs /= xs.sum()
    b = np.dot(A, xs) * (1 + 0.01 * (np.random.rand(m) - 0.5))
    return A, b


def main():
    only = set(sys.argv[1:])
    ref = ref_loader.import_reference()
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    t0 = time.time()

    def want(tag):
        return not only or tag in only

    if want("c2"):
        f, h, L, x0 = ref.D_opt_design(500, 50000, randseed=1)
        r = quiet(ref.ABPG, f, h, L, x0, gamma=2, maxitrs=300, theta_eq=False, verbskip=1000)
        out["c2_abpg_F"], out["c2_abpg_G"] = r[1], r[2]
        print("c2 abpg", time.time() - t0, flush=True)
        r = quiet(ref.ABPG_gain, f, h, L, x0, gamma=2, maxitrs=150, verbskip=1000)
        out["c2_gain_F"], out["c2_gain_Gain"], out["c2_gain_Gdiv"] = r[1], r[2], r[3]
        print("c2 gain", time.time() - t0, flush=True)
        nf, nh = noisy(f, h, np.random.RandomState(0))
        rp = quiet(ref.ABPG_gain, nf, nh, L, x0, gamma=2, maxitrs=150, verbskip=1000)
        k = min(len(r[1]), len(rp[1]))
        out["c2_gain_noise_dF"] = np.abs(rp[1][:k] - r[1][:k]) / np.abs(r[1][:k])
        out["c2_gain_noise_Gain"] = rp[2][:k]
        print("c2 gain noise", time.time() - t0, flush=True)
        r = quiet(ref.BPG, f, h, L, x0, maxitrs=100, linesearch=True, ls_ratio=1.2, verbskip=1000)
        out["c2_bpg_F"], out["c2_bpg_Ls"] = r[1], r[2]
        nf, nh = noisy(f, h, np.random.RandomState(0))
        rp = quiet(ref.BPG, nf, nh, L, x0, maxitrs=100, linesearch=True, ls_ratio=1.2, verbskip=1000)
        out["c2_bpg_noise_dF"] = np.abs(rp[1] - r[1]) / np.abs(r[1])
        out["c2_bpg_noise_Ls"] = rp[2]
        print("c2 bpg", time.time() - t0, flush=True)
        r = quiet(ref.D_opt_FW_away, f.H, x0, 1e-12, 1200, verbskip=100000)
        out["c2_fwa_F"], out["c2_fwa_SP"], out["c2_fwa_SN"] = r[1], r[2], r[3]
        out["c2_fwa_support"] = np.nonzero(r[0] > 1e-8)[0]
        print("c2 fw-away", time.time() - t0, flush=True)
        np.savez_compressed(OUT, **out)
    if want("kl"):
        A, b = kl_instance(2000, 20000, 11)
        f, h = ref.KLdivRegression(A, b), ref.ShannonEntropySimplex()
        x0 = np.ones(20000) / 20000
        r = quiet(ref.ABPG_gain, f, h, 1.0, x0, gamma=2.0, maxitrs=100, verbskip=1000)
        out["kl_gain_F"], out["kl_gain_Gain"] = r[1], r[2]
        out["kl_b_sum"] = b.sum()
        out["kl_b"] = b          # A is bit-stable (legacy RNG); b went through a BLAS dot, so it is stored
        print("kl", time.time() - t0, flush=True)
        np.savez_compressed(OUT, **out)
    if want("m2000"):
        f, h, L, x0 = ref.D_opt_design(2000, 20000, randseed=2)
        r = quiet(ref.ABPG_gain, f, h, L, x0, gamma=2, maxitrs=50, verbskip=1000)
        out["m2000_gain_F"], out["m2000_gain_Gain"] = r[1], r[2]
        print("m2000", time.time() - t0, flush=True)
        np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(k, np.shape(v))
    print(os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
