"""Fit of the bf16-mode GELU used by rf_dw_tma.cu:  gelu(x) = x / (1 + exp(-x * P(min(x^2, 64)))), P of degree 4.

Iteratively re-weighted least squares towards the minimax fit of x*Phi(x) on |x| <= 8, then the error of an fp32
evaluation on |x| <= 30 against scipy's erfc.  Prints the coefficients pasted into gelu_erf2()."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erfc

np.seterr(all="ignore")
TMAX = 64.0


def phi(x):
    return 0.5 * erfc(-x / np.sqrt(2))


def model(c, x, dt=np.float64):
    x = x.astype(dt)
    t = np.minimum(x * x, dt(TMAX))
    p = dt(c[-1])
    for k in range(len(c) - 2, -1, -1):
        p = p * t + dt(c[k])
    return x / (dt(1) + np.exp2(-(x * p) * dt(1.4426950408889634)))


def main():
    x = np.linspace(-8, 8, 64001)
    g = x * phi(x)
    c = np.array([1.5954862968770487, 0.07317978805285182, -0.00033970554119043046, -4.944016282720699e-05,
                  1.7605811939406665e-06])
    w = np.ones_like(x)
    resid = lambda c: (model(c, x) - g) / np.maximum(np.abs(g), 2e-2)
    for _ in range(60):
        c = least_squares(lambda c: resid(c) * w, c, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
        e = np.abs(resid(c))
        w = w * (1 + 2 * e / e.max())
        w /= w.mean()
    print("coefficients:", c.tolist())
    xx = np.linspace(-30, 30, 600001)
    gg = xx * phi(xx)
    m = model(c.astype(np.float32), xx, np.float32).astype(np.float64)
    print("fp32 evaluation: max abs err %.3e, max err / max(|gelu|, 0.02) %.3e" %
          (np.abs(m - gg).max(), (np.abs(m - gg) / np.maximum(np.abs(gg), 2e-2)).max()))


if __name__ == "__main__":
    main()
