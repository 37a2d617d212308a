"""numpy stand-in for ``numexpr.evaluate`` (numexpr is not installed in this image).

The reference uses numexpr only for the elementwise ``v * exp((A + B - 2 * C) / s)`` epilogue of the RBF
kernel (/root/reference/ital/gp.py:412,416,432,436).  Evaluating the same expression with numpy is
arithmetically equivalent up to libm's exp rounding.  Used only by tests/golden/make_golden.py.
"""
import numpy as np


def evaluate(expr, local_dict=None, global_dict=None):
    ns = {'exp': np.exp, 'log': np.log, 'sqrt': np.sqrt}
    ns.update(local_dict or {})
    return eval(expr, {'__builtins__': {}}, ns)
