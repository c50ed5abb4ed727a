"""The reference-side binding printed in INTEGRATION.md is executed as it stands (only the library path is made
absolute) and its gradients are compared with the oracle: the document cannot drift from the C ABI unnoticed."""
import os
import re
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import pipeline as op                  # noqa: E402  (the checker, never the product)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_namespace():
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    code = next(b for b in blocks if 'def predict_gradient_b200' in b)
    code = code.replace('"libedrgp_b200.so"', repr(os.path.join(ROOT, 'edrgp_b200', 'libedrgp_b200.so')))
    ns = {}
    exec(compile(code, 'INTEGRATION.md', 'exec'), ns)
    return ns


@pytest.mark.parametrize("n,d,m", [(700, 6, 20), (3000, 64, 512)])
def test_documented_ctypes_stub_reproduces_the_gradients(n, d, m):
    ns = _stub_namespace()
    w = op.make_workload(n, d, m, seed=n + m, k_true=2)
    yn = (w['y'] - w['y'].mean()) / w['y'].std()
    P, b, yy = op.inducing_stats_chunked(w['X'], yn, w['Z'], w['ell'], w['sf2'])
    sol = op.solve_from_stats(op.kuu(w['Z'], w['ell'], w['sf2']), P, b, yy, n, w['sf2'], w['noise'])
    G_ref = op.gradients_chunked(w['X'], w['Z'], w['ell'], w['sf2'], sol['alpha'], scale=w['y'].std())
    G = ns['predict_gradient_b200'](w['X'], w['Z'], w['ell'], w['sf2'], sol['alpha'], y_std=w['y'].std())
    assert G.shape == (n, d)
    assert np.max(np.abs(G - G_ref)) < 1e-8 * np.max(np.abs(G_ref))
