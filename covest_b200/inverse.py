"""Numeric inverses of increasing functions on the non-negative reals (covest/inverse.py)."""


def inverse_bs(f, precision=1e-8):
    """x with f(x) ~ y by bisection on [0, y]."""
    def f_inv(y):
        lo, hi = 0.0, float(y)
        prev, mid = 0.0, hi / 2
        while abs(mid - prev) > precision:
            if f(mid) < y:
                lo = mid
            else:
                hi = mid
            prev, mid = mid, (lo + hi) / 2
        return mid
    return f_inv


def inverse(f, delta=1e-8):
    """x with f(x) ~ y by Newton's method from y/2, forward-difference slope of width `delta`,
    stopping when the step falls below `delta` (inverse.py:20-40)."""
    def f_inv(y):
        def g(x):
            return f(x) - y

        def slope(x):
            return (g(x + delta) - g(x)) / delta

        x = float(y) / 2
        step = g(x) / slope(x)
        while abs(step) > delta:
            x -= step
            step = g(x) / slope(x)
        return x
    return f_inv
