"""``_make_kernel`` of the estimator shell against the reference's rules (edrgp/gp_model/base.py:111-147):
names + options, a ready kernel object of this package, a ready ``GPy.kern`` object (duck-typed: GPy is not
installable here, so a stand-in class carrying GPy's module path and attribute names is used).  Host logic
only, no GPU."""
import numpy as np
import pytest

from edrgp_b200 import SparseGaussianProcessRegressor
from edrgp_b200 import model as emodel


def _shell(d, *args, **kw):
    est = SparseGaussianProcessRegressor(*args, **kw)
    est.n_features_ = d
    return est


class _Param(np.ndarray):
    """GPy's ``Param`` is an ndarray subclass; the shell must read it without GPy."""
    def __new__(cls, v):
        return np.atleast_1d(np.asarray(v, dtype=np.float64)).view(cls)


def _gpy_like(name, cls_name, module='GPy.kern.src.rbf', **attrs):
    kern = type(cls_name, (object,), {})()
    type(kern).__module__ = module
    kern.name = name
    for k, v in attrs.items():
        setattr(kern, k, v)
    return kern


def test_none_lets_the_model_pick_its_default():
    assert _shell(4)._make_kernel() is None


def test_names_and_options_like_the_reference():
    k = _shell(4, 'RBF', {'ARD': True, 'variance': 2.5})._make_kernel()
    assert isinstance(k, emodel.RBF) and k.ARD and k.input_dim == 4 and k.variance == 2.5
    assert k.lengthscale.shape == (4,)
    k = _shell(3, ['RBF'], [{'lengthscale': 0.5}])._make_kernel()
    assert not k.ARD and np.array_equal(k.lengthscale, [0.5])
    k = _shell(3, ['RBF'])._make_kernel()
    assert k.variance == 1.0 and np.array_equal(k.full_lengthscale(), np.ones(3))
    with pytest.raises(ValueError):
        _shell(3, ['RBF'], [{}, {}])._make_kernel()
    with pytest.raises(NotImplementedError):
        _shell(3, ['RBF', 'RBF'])._make_kernel()
    with pytest.raises(NotImplementedError):
        _shell(3, 'Matern52')._make_kernel()


def test_own_kernel_object_is_copied_not_shared():
    mine = emodel.RBF(3, 1.5, [1., 2., 3.], ARD=True)
    k = _shell(3, mine)._make_kernel()
    assert k is not mine and k.variance == 1.5 and np.array_equal(k.lengthscale, mine.lengthscale)
    k.lengthscale[0] = 9.
    assert mine.lengthscale[0] == 1.


def test_gpy_rbf_object_passes_through():
    gk = _gpy_like('rbf', 'RBF', input_dim=3, ARD=True, variance=_Param(0.7), lengthscale=_Param([1., 2., 4.]),
                   active_dims=np.arange(3))
    k = _shell(3, gk)._make_kernel()
    assert isinstance(k, emodel.RBF) and k.ARD and k.variance == 0.7
    assert np.array_equal(k.lengthscale, [1., 2., 4.]) and type(k.lengthscale) is np.ndarray
    iso = _gpy_like('rbf', 'RBF', input_dim=5, ARD=False, variance=_Param(2.), lengthscale=_Param(3.))
    k = _shell(5, iso)._make_kernel()
    assert not k.ARD and np.array_equal(k.full_lengthscale(), np.full(5, 3.))


def test_gpy_objects_outside_the_path_say_so():
    rbf = dict(input_dim=3, ARD=False, variance=_Param(1.), lengthscale=_Param(1.))
    add = _gpy_like('sum', 'Add', module='GPy.kern.src.add', input_dim=3,
                    parts=[_gpy_like('rbf', 'RBF', **rbf), _gpy_like('white', 'White', input_dim=3)])
    with pytest.raises(NotImplementedError):
        _shell(3, add)._make_kernel()
    with pytest.raises(NotImplementedError):
        _shell(3, _gpy_like('Mat32', 'Matern32', module='GPy.kern.src.stationary', **rbf))._make_kernel()
    with pytest.raises(NotImplementedError):
        _shell(3, _gpy_like('rbf', 'RBF', active_dims=np.array([0, 2]), **rbf))._make_kernel()
    with pytest.raises(ValueError):
        _shell(4, _gpy_like('rbf', 'RBF', **rbf))._make_kernel()


@pytest.mark.gpu
def test_gpy_rbf_object_fits_like_the_own_kernel():
    """End to end on the device: the duck-typed GPy object and the package's own holder give the same model."""
    from oracle import pipeline as op
    w = op.make_workload(400, 6, 16, seed=2, k_true=1)
    gk = _gpy_like('rbf', 'RBF', input_dim=6, ARD=True, variance=_Param(w['sf2']), lengthscale=_Param(w['ell']))
    fits = []
    for kern in (gk, emodel.RBF(6, w['sf2'], w['ell'], ARD=True)):
        est = SparseGaussianProcessRegressor(kernels=kern, Z=w['Z'], method='fixed', noise_var=w['noise'])
        fits.append(est.fit(w['X'], w['y']).predict_gradient(w['X'][:64]))
    assert np.array_equal(fits[0], fits[1])
