"""Replays the EDR cells of /root/reference/examples/BriefIntro.ipynb (the only GPy-produced numbers in the
reference tree) on the UNMODIFIED reference orchestration layer with the oracle standing in for GPy, and prints
the replayed values beside the notebook's prints -- including the two cells that are NOT reproduced.
``tests/test_gpy_known_answers.py`` asserts the reproduced ones.  Run here (needs /root/reference or oracle/_ref):

    python tests/golden/replay_notebook.py > profiles/r02_s4_gpy_known_answers.txt

(the sensitivity table and the GPU note at the end of that file were appended by hand: the same replay with
``fmin_l_bfgs_b``'s options changed, and the log of the ``-m gpu`` tests).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import build_ref                                         # noqa: E402
sys.path.append(build_ref.path())
import edrgp                                                         # noqa: E402
from edrgp.utils import SVDTransformer, discrepancy                  # noqa: E402
from oracle.estimator import GaussianProcessRegressor                # noqa: E402
from oracle import gpy_restatement as gpy                            # noqa: E402
import test_gpy_known_answers as ka                                  # noqa: E402

warnings.simplefilter('ignore')
np.set_printoptions(precision=6, suppress=True, linewidth=150)
X, B, y, y_sparse = ka._cells_21_and_32(edrgp)
GPR = GaussianProcessRegressor('RBF', [{'ARD': True}], normalizer=True)

edr = ka._reference_edr(edrgp, GPR).fit(X, y)
print("cell [29]  printed: Discrepancy = 0.135    replayed: %.6f" % discrepancy(B, edr.components_.T[:, :2]))
it = edrgp.EffectiveDimensionalityReduction(GPR, SVDTransformer(), n_components=2, step=1, normalize=False).fit(X, y)
print("cell [30]  printed: Discrepancy = 0.056    replayed: %.6f   (NOT reproduced: eight chained L-BFGS-B runs)"
      % discrepancy(B, it.components_.T))
edr = ka._reference_edr(edrgp, GPR).fit(X, y_sparse)
print("cell [34]  printed: Discrepancy = 0.061    replayed: %.6f"
      % discrepancy(ka.CELL_33_B_SPARSE, edr.components_.T[:, :2]))
ours = ka._rows_up_to_sign(edr.components_[:, :2], ka.CELL_35_COMPONENTS)
print("cell [35]  printed | replayed (row signs aligned) | difference")
for p, o in zip(ka.CELL_35_COMPONENTS, ours):
    print("   % .3f % .3f  |  % .6f % .6f  |  % .1e % .1e" % (p[0], p[1], o[0], o[1], o[0] - p[0], o[1] - p[1]))
print("   max |difference| = %.2e (the print keeps 3 decimals: 5e-4)" % np.max(np.abs(ours - ka.CELL_35_COMPONENTS)))

first = GaussianProcessRegressor('RBF', [{'ARD': True}], normalizer=True).fit(X, y_sparse)
sf2, ell, noise = ka._fitted_hyperparameters(first)
print("first-pass optimum: variance %.6g, noise %.6g, bound %.9f\n   lengthscales %s"
      % (sf2, noise, float(first.estimator_.log_likelihood()), np.array2string(ell, precision=5)))
sparse = gpy.SparseGPRegression(X, y_sparse[:, None], kernel=gpy.RBF(10, sf2, ell, ARD=True), Z=X.copy(),
                                normalizer=True)
sparse.noise_variance = noise
sparse.parameters_changed()
G = sparse.predictive_gradients(X)[0][:, :, 0]
print("variational sparse model at Z = X, same hyper-parameters: gradients vs the dense model %.2e (max, relative), "
      "bound %.9f" % (np.max(np.abs(G - edr._first_gradients_)) / np.max(np.abs(edr._first_gradients_)),
                      float(np.sum(sparse.log_likelihood()))))
from sklearn.decomposition import SparsePCA                          # noqa: E402
edr.refit(SparsePCA(n_components=2, alpha=2))
print("cell [37]  printed: rows 2 / 6 / 7 of refit_components_.T = -0.648 / -0.444 / -0.619, row 0 = (0, 1)   replayed "
      "(NOT reproduced: scikit-learn's SparsePCA changed):\n%s" % np.round(edr.refit_components_.T, 3))
