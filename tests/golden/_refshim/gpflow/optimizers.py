"""gpflow.optimizers.Scipy (gpflow/optimizers/scipy.py): pack the unconstrained variables into one float64 vector in the order given,
evaluate loss + gradient (torch autograd standing in for tf.GradientTape), call scipy.optimize.minimize(jac=True)."""
import numpy as _np
import scipy.optimize
import torch as _torch


class Scipy:
    def minimize(self, closure, variables, method='L-BFGS-B', step_callback=None, compile=True, allow_unused_variables=False, **scipy_kwargs):
        variables = tuple(variables)
        if not variables:
            raise ValueError('No trainable variables')

        def pack():
            return _np.concatenate([v.detach().numpy().reshape(-1) for v in variables])

        def unpack(x):
            k = 0
            with _torch.no_grad():
                for v in variables:
                    n = v.numel()
                    v.copy_(_torch.from_numpy(x[k:k + n].copy()).reshape(v.shape))
                    k += n

        def fun(x):
            unpack(x)
            for v in variables:
                v.grad = None
            loss = closure()
            grads = _torch.autograd.grad(loss, variables, allow_unused=allow_unused_variables)
            g = _np.concatenate([(_torch.zeros_like(v) if gr is None else gr).detach().numpy().reshape(-1) for v, gr in zip(variables, grads)])
            return float(loss.detach()), g

        result = scipy.optimize.minimize(fun, pack(), jac=True, method=method, **scipy_kwargs)
        unpack(result.x)
        return result
