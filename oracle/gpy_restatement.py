"""NumPy FP64 restatement of the GPy arithmetic behind edr-gp's sparse-GP path.  TEST ORACLE.

GPy is a third-party dependency of the reference (``/root/reference/requirements.txt:3``,
``gpy>=1.8.4``, unpinned) that is absent from ``/root/reference`` and from this environment, so the
algorithm is restated here from GPy's published sources (GPy 1.9-1.13 agree on everything below
except the normaliser scaling of ``predictive_gradients``, see ``SparseGPRegression.
predictive_gradients``).  Every class names the GPy symbol it restates and the edr-gp call site
(file:line under ``/root/reference``) that reaches it.  Operation order follows GPy so that the
floating-point results are the ones the reference would produce: distance by GEMM expansion +
clip, jitter 1e-8 on Kuu, Cholesky chain for the Woodbury vector, per-dimension loop for
``gradients_X``.

PINNED at print precision (3 decimals) against the GPy-produced outputs kept in the reference's notebook
(``tests/test_gpy_known_answers.py``); PARITY UNPINNED at 1e-8: see ``oracle/__init__.py``.
"""
import numpy as np
from scipy import linalg as sla
from scipy import optimize as sopt

CONST_JITTER = 1e-8          # GPy VarDTC.const_jitter
_LIM_VAL = 36.0              # paramz.transformations._lim_val
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)
_EPS = np.finfo(np.float64).resolution


# --------------------------------------------------------------------------------------------
# GPy.util.linalg
# --------------------------------------------------------------------------------------------
def jitchol(A, maxtries=5):
    """GPy.util.linalg.jitchol: lower Cholesky, adding growing jitter on failure."""
    A = np.ascontiguousarray(A)
    try:
        return sla.cholesky(A, lower=True)
    except sla.LinAlgError:
        pass
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise sla.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        try:
            return sla.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
        except sla.LinAlgError:
            jitter *= 10
    raise sla.LinAlgError("not positive definite, even with jitter.")


def dtrtrs(A, B, lower=1, trans=0):
    """GPy.util.linalg.dtrtrs (LAPACK dtrtrs): solve op(A) X = B for triangular A."""
    return sla.solve_triangular(A, B, lower=bool(lower), trans=trans, check_finite=False)


def tdot(A):
    """GPy.util.linalg.tdot: A A^T (dsyrk)."""
    return A.dot(A.T)


def backsub_both_sides(L, X, transpose='left'):
    """GPy.util.linalg.backsub_both_sides: L^-T X L^-1 ('left') or L^-1 X L^-T."""
    if transpose == 'left':
        tmp = dtrtrs(L, X, lower=1, trans=1)
        return dtrtrs(L, tmp.T, lower=1, trans=1).T
    tmp = dtrtrs(L, X, lower=1, trans=0)
    return dtrtrs(L, tmp.T, lower=1, trans=0).T


def dpotri(L):
    """GPy.util.linalg.dpotri + symmetrify: inverse of L L^T from its lower factor."""
    n = L.shape[0]
    Li = dtrtrs(L, np.eye(n), lower=1)
    return Li.T.dot(Li)


def pdinv(A):
    """GPy.util.linalg.pdinv: (inverse, L, L^-1, logdet)."""
    L = jitchol(A)
    logdet = 2. * np.sum(np.log(np.diag(L)))
    Li = dtrtrs(L, np.eye(A.shape[0]), lower=1)
    Ai = Li.T.dot(Li)
    return Ai, L, Li, logdet


# --------------------------------------------------------------------------------------------
# GPy.util.normalizer.Standardize   (reached through normalizer=True,
# edrgp/gp_model/regression.py:127,157)
# --------------------------------------------------------------------------------------------
class Standardize(object):
    def scale_by(self, Y):
        Y = np.ma.masked_invalid(Y, copy=False)
        self.mean = np.asarray(Y.mean(0))
        self.std = np.asarray(Y.std(0))

    def normalize(self, Y):
        return (Y - self.mean) / self.std

    def inverse_mean(self, X):
        return (X * self.std) + self.mean

    def inverse_variance(self, var):
        return var * (self.std ** 2)


# --------------------------------------------------------------------------------------------
# GPy.kern.RBF (GPy.kern.src.stationary.Stationary + rbf.RBF)
# reached through edrgp/gp_model/base.py:111-147 ('RBF', [{'ARD': True}])
# --------------------------------------------------------------------------------------------
class RBF(object):
    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self.variance = float(variance)
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if self.ARD else 1)
        lengthscale = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)).copy()
        if self.ARD and lengthscale.size == 1:
            lengthscale = np.ones(self.input_dim) * lengthscale
        if not self.ARD:
            assert lengthscale.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        self.lengthscale = lengthscale

    def copy(self):
        return RBF(self.input_dim, self.variance, self.lengthscale.copy(), self.ARD)

    # -- Stationary._unscaled_dist / _scaled_dist
    @staticmethod
    def _unscaled_dist(X, X2=None):
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2. * tdot(X) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices(X.shape[0])] = 0.   # force diagonal to be zero
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        X1sq = np.sum(np.square(X), 1)
        X2sq = np.sum(np.square(X2), 1)
        r2 = -2. * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist(self, X, X2=None):
        if self.ARD:
            if X2 is not None:
                X2 = X2 / self.lengthscale
            return self._unscaled_dist(X / self.lengthscale, X2)
        return self._unscaled_dist(X, X2) / self.lengthscale

    def _inv_dist(self, X, X2=None):
        dist = self._scaled_dist(X, X2).copy()
        return 1. / np.where(dist != 0., dist, np.inf)

    # -- RBF.K_of_r / dK_dr
    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)

    def dK_dr(self, r):
        return -r * self.K_of_r(r)

    def K(self, X, X2=None):
        return self.K_of_r(self._scaled_dist(X, X2))

    def Kdiag(self, X):
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret

    # -- Stationary.gradients_X (the pure-NumPy branch, ``_gradients_X_pure``)
    def gradients_X(self, dL_dK, X, X2=None):
        invdist = self._inv_dist(X, X2)
        dL_dr = self.dK_dr(self._scaled_dist(X, X2)) * dL_dK
        tmp = invdist * dL_dr
        if X2 is None:
            tmp = tmp + tmp.T
            X2 = X
        grad = np.empty(X.shape, dtype=np.float64)
        for q in range(self.input_dim):
            np.sum(tmp * (X[:, q][:, None] - X2[:, q][None, :]), axis=1, out=grad[:, q])
        return grad / self.lengthscale ** 2

    def gradients_X_diag(self, dL_dKdiag, X):
        return np.zeros(X.shape)

    # -- Stationary.update_gradients_full / update_gradients_diag: return (dvariance, dlengthscale)
    def update_gradients_full(self, dL_dK, X, X2=None):
        dvar = np.sum(self.K(X, X2) * dL_dK) / self.variance
        dL_dr = self.dK_dr(self._scaled_dist(X, X2)) * dL_dK
        if self.ARD:
            tmp = dL_dr * self._inv_dist(X, X2)
            if X2 is None:
                X2 = X
            dlen = -np.array([np.sum(tmp * np.square(X[:, q:q + 1] - X2[:, q:q + 1].T))
                              for q in range(self.input_dim)]) / self.lengthscale ** 3
        else:
            r = self._scaled_dist(X, X2)
            dlen = np.atleast_1d(-np.sum(dL_dr * r) / self.lengthscale)
        return dvar, dlen

    def update_gradients_diag(self, dL_dKdiag, X):
        return np.sum(dL_dKdiag), np.zeros_like(self.lengthscale)


# --------------------------------------------------------------------------------------------
# GPy.inference.latent_function_inference.var_dtc.VarDTC.inference (homoscedastic Gaussian,
# certain inputs, no mean function) -- SURVEY.md section 8 row a6
# --------------------------------------------------------------------------------------------
class Posterior(object):
    def __init__(self, woodbury_inv, woodbury_vector, K, K_chol):
        self.woodbury_inv = woodbury_inv
        self.woodbury_vector = woodbury_vector
        self.K = K
        self.K_chol = K_chol


def vardtc_inference(kern, X, Z, noise_variance, Y, psi1=None):
    num_data, output_dim = Y.shape
    num_inducing = Z.shape[0]
    precision = 1. / np.fmax(np.atleast_1d(noise_variance).astype(np.float64), CONST_JITTER)
    precision = precision[:, None]                      # (1, 1), as in GPy
    VVT_factor = precision * Y
    trYYT = np.einsum('ij,ij->', Y, Y)

    Kmm = kern.K(Z).copy()
    Kmm[np.diag_indices(num_inducing)] += CONST_JITTER
    Lm = jitchol(Kmm)

    psi0 = kern.Kdiag(X)
    if psi1 is None:
        psi1 = kern.K(X, Z)
    tmp = psi1 * (np.sqrt(precision))
    tmp = dtrtrs(Lm, tmp.T, lower=1)
    A = tdot(tmp)

    B = np.eye(num_inducing) + A
    LB = jitchol(B)
    tmp = dtrtrs(Lm, psi1.T, lower=1, trans=0)
    _LBi_Lmi_psi1 = dtrtrs(LB, tmp, lower=1, trans=0)
    _LBi_Lmi_psi1Vf = np.dot(_LBi_Lmi_psi1, VVT_factor)
    tmp = dtrtrs(LB, _LBi_Lmi_psi1Vf, lower=1, trans=1)
    Cpsi1Vf = dtrtrs(Lm, tmp, lower=1, trans=1)

    delit = tdot(_LBi_Lmi_psi1Vf)
    data_fit = np.trace(delit)
    DBi_plus_BiPBi = backsub_both_sides(LB, output_dim * np.eye(num_inducing) + delit)
    delit = -0.5 * DBi_plus_BiPBi
    delit += -0.5 * B * output_dim
    delit += output_dim * np.eye(num_inducing)
    dL_dKmm = backsub_both_sides(Lm, delit)

    # _compute_dL_dpsi
    dL_dpsi0 = -0.5 * output_dim * (precision * np.ones([num_data, 1])).flatten()
    dL_dpsi1 = np.dot(VVT_factor, Cpsi1Vf.T)
    dL_dpsi2_beta = 0.5 * backsub_both_sides(Lm, output_dim * np.eye(num_inducing) - DBi_plus_BiPBi)
    dL_dpsi2 = precision * dL_dpsi2_beta
    dL_dpsi1 += 2. * np.dot(psi1, dL_dpsi2)

    # _compute_log_marginal_likelihood
    lik_1 = (-0.5 * num_data * output_dim * (np.log(2. * np.pi) - np.log(precision))
             - 0.5 * precision * trYYT)
    lik_2 = -0.5 * output_dim * (np.sum(precision * psi0) - np.trace(A))
    lik_3 = -output_dim * (np.sum(np.log(np.diag(LB))))
    lik_4 = 0.5 * data_fit
    log_marginal = lik_1 + lik_2 + lik_3 + lik_4          # (1, 1) array

    # _compute_dL_dR + Gaussian.exact_inference_gradients
    dL_dR = -0.5 * num_data * output_dim * precision + 0.5 * trYYT * precision ** 2
    dL_dR += 0.5 * output_dim * (psi0.sum() * precision ** 2 - np.trace(A) * precision)
    dL_dR += precision * (0.5 * np.sum(A * DBi_plus_BiPBi) - data_fit)
    dL_dthetaL = dL_dR.sum()

    grad_dict = {'dL_dKmm': dL_dKmm, 'dL_dKdiag': dL_dpsi0, 'dL_dKnm': dL_dpsi1,
                 'dL_dthetaL': dL_dthetaL}

    Bi = -dpotri(LB)
    Bi[np.diag_indices(num_inducing)] += 1
    woodbury_inv = backsub_both_sides(Lm, Bi)
    post = Posterior(woodbury_inv=woodbury_inv, woodbury_vector=Cpsi1Vf, K=Kmm, K_chol=Lm)
    post._A = A
    post._LB = LB
    return post, log_marginal, grad_dict


# --------------------------------------------------------------------------------------------
# paramz Logexp transform + L-BFGS-B driver (paramz.transformations.Logexp,
# paramz.optimization.opt_lbfgsb) -- SURVEY.md section 8 row a7
# --------------------------------------------------------------------------------------------
def logexp_f(x):
    return np.where(x > _LIM_VAL, x,
                    np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))


def logexp_finv(f):
    return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))


def logexp_gradfactor(f, df):
    return df * np.where(f > _LIM_VAL, 1., -np.expm1(-f))


class _Model(object):
    """Minimal paramz.Model: positive parameters under Logexp, L-BFGS-B, restarts."""

    _fail_count = 0
    _allowed_failures = 10

    def _objective_grads(self, x):
        try:
            self._set_optimizer_array(x)
            obj = -float(np.sum(self.log_likelihood()))
            grads = -self._transformed_gradients()
            self._fail_count = 0
        except (sla.LinAlgError, ZeroDivisionError, ValueError, FloatingPointError):
            if self._fail_count >= self._allowed_failures:
                raise
            self._fail_count += 1
            return np.inf, np.clip(np.zeros_like(x), -1e10, 1e10)
        return obj, np.clip(grads, -1e10, 1e10)

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, **kwargs):
        x0 = self._get_optimizer_array() if start is None else start
        if max_iters <= 0 or x0.size == 0:
            self._set_optimizer_array(x0)
            return self
        x_opt, f_opt, info = sopt.fmin_l_bfgs_b(self._objective_grads, x0,
                                                maxfun=max_iters, maxiter=max_iters)
        self._set_optimizer_array(x_opt)
        self.optimization_runs.append((f_opt, x_opt, info))
        return self

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=False, **kwargs):
        initial = self._get_optimizer_array().copy()
        first = len(self.optimization_runs)
        for i in range(num_restarts):
            try:
                if i > 0:
                    self._set_optimizer_array(np.random.normal(size=initial.size))
                self.optimize(**kwargs)
            except Exception:
                if not robust:
                    raise
        runs = self.optimization_runs[first:]
        if runs:
            best = int(np.argmin([r[0] for r in runs]))
            self._set_optimizer_array(runs[best][1])
        else:
            self._set_optimizer_array(initial)
        return self


# --------------------------------------------------------------------------------------------
# GPy.models.SparseGPRegression (+ core.SparseGP, core.GP) -- the object edr-gp stores as
# ``estimator_`` (edrgp/gp_model/regression.py:153-157; edrgp/gp_model/base.py:65-69)
# --------------------------------------------------------------------------------------------
class SparseGPRegression(_Model):
    def __init__(self, X, Y, kernel=None, Z=None, num_inducing=10, X_variance=None,
                 mean_function=None, normalizer=None):
        if X_variance is not None or mean_function is not None:
            raise NotImplementedError("uncertain inputs / mean functions are outside the path")
        num_data, input_dim = X.shape
        self.X = np.asarray(X, dtype=np.float64)
        if kernel is None:
            kernel = RBF(input_dim)
        self.kern = kernel
        if Z is None:
            i = np.random.permutation(num_data)[:min(num_inducing, num_data)]
            Z = self.X[i].copy()
        else:
            Z = np.array(Z, dtype=np.float64)
            assert Z.shape[1] == input_dim
        self.Z = Z
        self.noise_variance = 1.0                      # likelihoods.Gaussian() default
        if normalizer is True:
            self.normalizer = Standardize()
        elif normalizer is False or normalizer is None:
            # GPy: normalizer=None -> no normalisation (edr-gp's docstring says otherwise)
            self.normalizer = None
        else:
            self.normalizer = normalizer
        self.Y = np.asarray(Y, dtype=np.float64)
        if self.normalizer is not None:
            self.normalizer.scale_by(self.Y)
            self.Y_normalized = self.normalizer.normalize(self.Y)
        else:
            self.Y_normalized = self.Y
        self.optimization_runs = []
        self.fix_Z = False
        self.parameters_changed()

    # -- SparseGP.parameters_changed + _update_gradients
    def parameters_changed(self):
        self.posterior, self._log_marginal_likelihood, self.grad_dict = vardtc_inference(
            self.kern, self.X, self.Z, self.noise_variance, self.Y_normalized)
        g = self.grad_dict
        dvar, dlen = self.kern.update_gradients_diag(g['dL_dKdiag'], self.X)
        dv, dl = self.kern.update_gradients_full(g['dL_dKnm'], self.X, self.Z)
        dvar, dlen = dvar + dv, dlen + dl
        dv, dl = self.kern.update_gradients_full(g['dL_dKmm'], self.Z, None)
        self.grad_variance, self.grad_lengthscale = dvar + dv, dlen + dl
        self.grad_noise = g['dL_dthetaL']
        self.grad_Z = self.kern.gradients_X(g['dL_dKmm'], self.Z)
        self.grad_Z += self.kern.gradients_X(g['dL_dKnm'].T, self.Z, self.X)

    def log_likelihood(self):
        return self._log_marginal_likelihood

    # -- parameter vector in paramz order: Z, rbf.variance, rbf.lengthscale, noise variance
    def _positive(self):
        return np.concatenate([[self.kern.variance], self.kern.lengthscale, [self.noise_variance]])

    def _get_optimizer_array(self):
        pos = logexp_finv(self._positive())
        if self.fix_Z:
            return pos
        return np.concatenate([self.Z.ravel(), pos])

    def _set_optimizer_array(self, x):
        x = np.asarray(x, dtype=np.float64)
        nz = 0 if self.fix_Z else self.Z.size
        if nz:
            self.Z = x[:nz].reshape(self.Z.shape).copy()
        pos = logexp_f(x[nz:])
        self.kern.variance = float(pos[0])
        self.kern.lengthscale = pos[1:1 + self.kern.lengthscale.size].copy()
        self.noise_variance = float(pos[-1])
        self.parameters_changed()

    def _transformed_gradients(self):
        pos = self._positive()
        gpos = np.concatenate([[self.grad_variance], self.grad_lengthscale, [self.grad_noise]])
        gpos = logexp_gradfactor(pos, gpos)
        if self.fix_Z:
            return gpos
        return np.concatenate([self.grad_Z.ravel(), gpos])

    # -- GP.predict (Posterior._raw_predict + Gaussian.predictive_values + normaliser)
    def predict(self, Xnew):
        Kx = self.kern.K(self.Z, Xnew)
        mu = np.dot(Kx.T, self.posterior.woodbury_vector)
        Kxx = self.kern.Kdiag(Xnew)
        var = (Kxx - np.sum(np.dot(self.posterior.woodbury_inv.T, Kx) * Kx, 0))[:, None]
        var = np.clip(var, 1e-15, np.inf)
        var = var + self.noise_variance
        if self.normalizer is not None:
            mu = self.normalizer.inverse_mean(mu)
            var = self.normalizer.inverse_variance(var)
        return mu, var

    # -- GP.predictive_gradients: edr-gp keeps only [0][:, :, 0] (edrgp/gp_model/base.py:222).
    def predictive_gradients(self, Xnew, scale_by_normalizer=True):
        """Posterior-mean Jacobian (n, d, 1).  The variance gradient (second output of GPy, built
        from an n x n matrix) is discarded by edr-gp and is not restated (returns None).

        GPy >= 1.9.9 finishes with ``mean_jac = normalizer.inverse_mean(mean_jac) -
        normalizer.inverse_mean(0.)`` i.e. multiplies by std(y); earlier versions return the
        Jacobian of the *normalised* mean.  It is one global scalar: ``components_`` and every
        variance *ratio* are unchanged, ``subspace_variance_`` scales by std(y)^2.
        ``scale_by_normalizer`` selects the behaviour; default follows current GPy.
        """
        mean_jac = np.empty((Xnew.shape[0], Xnew.shape[1], 1))
        mean_jac[:, :, 0] = self.kern.gradients_X(self.posterior.woodbury_vector[:, 0:1].T,
                                                   Xnew, self.Z)
        if self.normalizer is not None and scale_by_normalizer:
            mean_jac = self.normalizer.inverse_mean(mean_jac) - self.normalizer.inverse_mean(0.)
        return mean_jac, None


# --------------------------------------------------------------------------------------------
# GPy.models.GPRegression (dense exact GP).  OUT of the hot path; restated only because the one
# reference test that touches the sparse path compares against it
# (edrgp/tests/test_edr.py:33-50).
# --------------------------------------------------------------------------------------------
class GPRegression(_Model):
    def __init__(self, X, Y, kernel=None, normalizer=None, noise_var=1.0):
        self.X = np.asarray(X, dtype=np.float64)
        self.kern = RBF(self.X.shape[1]) if kernel is None else kernel
        self.noise_variance = float(noise_var)
        self.normalizer = Standardize() if normalizer is True else (
            None if normalizer in (False, None) else normalizer)
        self.Y = np.asarray(Y, dtype=np.float64)
        if self.normalizer is not None:
            self.normalizer.scale_by(self.Y)
            self.Y_normalized = self.normalizer.normalize(self.Y)
        else:
            self.Y_normalized = self.Y
        self.optimization_runs = []
        self.parameters_changed()

    def parameters_changed(self):
        # ExactGaussianInference.inference
        Y = self.Y_normalized
        K = self.kern.K(self.X)
        Ky = K.copy()
        Ky[np.diag_indices(K.shape[0])] += self.noise_variance + 1e-8
        Wi, LW, LWi, W_logdet = pdinv(Ky)
        alpha = sla.cho_solve((LW, True), Y)
        self._log_marginal_likelihood = 0.5 * (-Y.size * np.log(2 * np.pi)
                                               - Y.shape[1] * W_logdet - np.sum(alpha * Y))
        dL_dK = 0.5 * (tdot(alpha) - Y.shape[1] * Wi)
        self.woodbury_vector = alpha
        self.grad_noise = np.diag(dL_dK).sum()
        self.grad_variance, self.grad_lengthscale = self.kern.update_gradients_full(dL_dK, self.X)

    def log_likelihood(self):
        return self._log_marginal_likelihood

    def _positive(self):
        return np.concatenate([[self.kern.variance], self.kern.lengthscale, [self.noise_variance]])

    def _get_optimizer_array(self):
        return logexp_finv(self._positive())

    def _set_optimizer_array(self, x):
        pos = logexp_f(np.asarray(x, dtype=np.float64))
        self.kern.variance = float(pos[0])
        self.kern.lengthscale = pos[1:1 + self.kern.lengthscale.size].copy()
        self.noise_variance = float(pos[-1])
        self.parameters_changed()

    def _transformed_gradients(self):
        gpos = np.concatenate([[self.grad_variance], self.grad_lengthscale, [self.grad_noise]])
        return logexp_gradfactor(self._positive(), gpos)

    def predict(self, Xnew):
        Kx = self.kern.K(self.X, Xnew)
        mu = np.dot(Kx.T, self.woodbury_vector)
        if self.normalizer is not None:
            mu = self.normalizer.inverse_mean(mu)
        return mu, None

    def predictive_gradients(self, Xnew, scale_by_normalizer=True):
        mean_jac = np.empty((Xnew.shape[0], Xnew.shape[1], 1))
        mean_jac[:, :, 0] = self.kern.gradients_X(self.woodbury_vector[:, 0:1].T, Xnew, self.X)
        if self.normalizer is not None and scale_by_normalizer:
            mean_jac = self.normalizer.inverse_mean(mean_jac) - self.normalizer.inverse_mean(0.)
        return mean_jac, None
