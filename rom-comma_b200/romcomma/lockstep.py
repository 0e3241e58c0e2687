""" Lock-step fitting of independent small GPs.

The reference fits the folds of a repository one after another (romcomma/user/run.py:60-61) and, inside each fold, the L single-output GPs of a
variant model one after another (romcomma/gpr/models.py:359-361).  Each fit is a scipy L-BFGS-B run whose every iteration is ONE evaluation of
the log marginal likelihood and its gradient - at N ~ 2000 a chain of a few hundred latency-bound kernel launches that leaves most of a B200 idle
(cfg2: 2.5 ms per evaluation at 9 % of the FP64 tensor peak, plus a device-to-host synchronisation per evaluation).

Here the independent fits run as threads of one process - scipy's optimiser untouched, one thread per fit - and an ``EvaluationBroker`` collects
the evaluations they request: when every optimiser that is currently running waits for its next LML + gradient, the requests are evaluated
together, as ONE batched C-ABI call (rc_lml_grad_multi: per-problem inputs, outputs, sample counts and hyper-parameters) and ONE device-to-host
copy, and handed back.  Every problem is still computed on its own padded matrix by the same kernels in the same order, so a fit follows exactly
the trajectory it has when it runs alone; only requests of equal padded size share a launch.

    run_together([fit_a, fit_b, ...])     # callables; inside them gf.optimizers.Scipy().minimize(...) takes part in the lock step

``ROMCOMMA_B200_LOCKSTEP=0`` runs everything sequentially, as the reference does.
"""
from __future__ import annotations

import os
import threading
from contextlib import contextmanager
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from romcomma import _capi

_broker: Optional['EvaluationBroker'] = None
_broker_guard = threading.Lock()


def enabled() -> bool:
    return os.environ.get('ROMCOMMA_B200_LOCKSTEP', '1') != '0'


class _Request:
    __slots__ = ('model', 'X', 'Y', 'L', 'flags', 'ls', 'F', 'E', 'result', 'error')

    def __init__(self, model, X, Y, L, flags, ls, F, E):
        self.model, self.X, self.Y, self.L, self.flags, self.ls, self.F, self.E = model, X, Y, L, flags, ls, F, E
        self.result, self.error = None, None


class EvaluationBroker:
    """ Gathers the evaluation requests of the optimiser threads that are ``participating`` and runs them in batches."""

    def __init__(self):
        self._cond = threading.Condition()
        self._participants: set = set()
        self._pending: Dict[int, _Request] = {}
        self._plans: Dict[Tuple, Tuple[Any, List[Any]]] = {}
        self._streams: List[Any] = []
        self.batches: List[int] = []            # size of every batch that has been launched (diagnostics / tests)

    # -- who takes part ---------------------------------------------------------------------------------------------------------------
    def is_participant(self) -> bool:
        return threading.get_ident() in self._participants

    @contextmanager
    def participating(self):
        """ While inside, this thread is one of the optimisers the broker waits for before it launches a batch."""
        me = threading.get_ident()
        with self._cond:
            self._participants.add(me)
        try:
            yield self
        finally:
            with self._cond:
                self._participants.discard(me)
                self._fire_if_complete()           # the others may have been waiting for this thread only

    @contextmanager
    def stepping_aside(self):
        """ A participant that is about to wait for child threads (the outputs of a variant GP) must not be waited for itself."""
        me = threading.get_ident()
        with self._cond:
            was = me in self._participants
            self._participants.discard(me)
            if was:
                self._fire_if_complete()
        try:
            yield
        finally:
            if was:
                with self._cond:
                    self._participants.add(me)

    # -- evaluation -------------------------------------------------------------------------------------------------------------------
    def evaluate(self, model: Any, X: torch.Tensor, Y: torch.Tensor, L: int, flags: int, ls: np.ndarray, F: np.ndarray, E: np.ndarray) -> dict:
        """ Called by a participating thread: blocks until the batch this request joins has been evaluated; returns {lml, dF, dE, dls} on the host.
        X (N,M), Y (N,L): device tensors of the problem (unchanged between calls); ls (L,M), F, E (L,L): host arrays of this call."""
        request = _Request(model, X, Y, L, flags, np.array(ls, dtype=np.float64), np.array(F, dtype=np.float64).reshape(L, L),
                           np.array(E, dtype=np.float64).reshape(L, L))
        me = threading.get_ident()
        with self._cond:
            self._pending[me] = request
            self._fire_if_complete()
            while request.result is None and request.error is None:
                self._cond.wait()
        if request.error is not None:
            raise request.error
        return request.result

    def _fire_if_complete(self):
        """ (lock held) Launch when every participant has a request pending."""
        if not self._pending or any(p not in self._pending for p in self._participants):
            return
        requests = list(self._pending.values())
        self._pending.clear()
        groups: Dict[Tuple, List[_Request]] = {}
        for r in requests:
            n_pad = _capi.padded(r.L * r.X.shape[0])
            groups.setdefault((n_pad, r.X.shape[1], r.L, r.flags, r.X.device), []).append(r)
        # Launch every group first - each on a stream of its own, so that a lone odd-sized problem (the improper fold) overlaps the big batch -
        # then wait for them in turn: one device-to-host copy per group.
        launched = []
        for i, (key, group) in enumerate(groups.items()):
            group.sort(key=lambda r: id(r.model))
            try:
                launched.append((group, self._launch(key, group, i if len(groups) > 1 else None)))
            except BaseException as exc:           # hand the failure to every waiting optimiser of the group
                for r in group:
                    r.error = exc
        for group, (plan, stream) in launched:
            try:
                self._collect(group, plan, stream)
            except BaseException as exc:
                for r in group:
                    r.error = exc
        self._cond.notify_all()

    def _launch(self, key: Tuple, group: List[_Request], stream_index: Optional[int]):
        n_pad, M, L, flags, device = key
        # A cached plan holds the packed data of a COMPOSITION of problems, identified by the models' ids - which stay unique only while the models
        # are alive, so the cache entry keeps them alive (a freed model's id may be handed to the next one created).
        plan_key = key + tuple(id(r.model) for r in group)
        plan = self._plans.get(plan_key, (None, None))[0]
        if plan is None:
            if len(self._plans) >= 6:                # compositions change as optimisers finish: keep the workspaces of the latest few only
                self._plans.pop(next(iter(self._plans)))
            plan = _capi.LmlGradMultiPlan([r.X for r in group], [r.Y for r in group], L, flags)
            self._plans[plan_key] = (plan, [r.model for r in group])
        ls = _capi.dev(np.concatenate([r.ls.reshape(L, M) for r in group], axis=0), device)
        F, E = _capi.dev(np.stack([r.F for r in group]), device), _capi.dev(np.stack([r.E for r in group]), device)
        stream = None
        if stream_index is not None and device.type == 'cuda':
            while len(self._streams) <= stream_index:
                self._streams.append(torch.cuda.Stream(device))
            stream = self._streams[stream_index]
            stream.wait_stream(torch.cuda.current_stream(device))      # the hyper-parameter uploads above
            with torch.cuda.stream(stream):
                plan(ls, F, E)
        else:
            plan(ls, F, E)
        self.batches.append(len(group))
        return plan, stream

    @staticmethod
    def _collect(group: List[_Request], plan, stream):
        if stream is not None:
            torch.cuda.current_stream(stream.device).wait_stream(stream)
        out = plan.out.cpu().numpy()               # the one synchronisation of this group
        info = plan.info.cpu().numpy()
        for z, (r, res) in enumerate(zip(group, plan.unpack(out))):
            if int(info[z]) != 0:
                r.error = _capi.RomcommaB200Error('Cholesky decomposition was not successful. The input might not be valid.')
            else:
                r.result = res


def current() -> Optional[EvaluationBroker]:
    """ The broker of the running lock-step session if the calling thread takes part in it, else None."""
    broker = _broker
    return broker if broker is not None and broker.is_participant() else None


@contextmanager
def participating():
    """ Used by the optimiser driver around scipy.optimize.minimize: joins the running session, if there is one."""
    broker = _broker
    if broker is None:
        yield None
    else:
        with broker.participating():
            yield broker


def run_together(jobs: Sequence[Callable[[], Any]]) -> List[Any]:
    """ Run independent jobs (each fitting GPs through ``gf.optimizers.Scipy``) as threads whose LML evaluations are batched; returns their
    results in order and re-raises the first exception.  A single job, or ROMCOMMA_B200_LOCKSTEP=0, runs in the calling thread."""
    global _broker
    jobs = list(jobs)
    if len(jobs) <= 1 or not enabled():
        return [job() for job in jobs]
    with _broker_guard:
        owner = _broker is None
        if owner:
            _broker = EvaluationBroker()
        broker = _broker
    device = torch.cuda.current_device() if torch.cuda.is_available() else None
    results: List[Any] = [None] * len(jobs)
    errors: List[Optional[BaseException]] = [None] * len(jobs)

    def work(i: int):
        try:
            if device is not None:
                torch.cuda.set_device(device)          # the current device is per-thread state
            results[i] = jobs[i]()
        except BaseException as exc:
            errors[i] = exc

    try:
        with broker.stepping_aside():
            threads = [threading.Thread(target=work, args=(i,), name=f'romcomma-lockstep-{i}') for i in range(len(jobs))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
    finally:
        if owner:
            with _broker_guard:
                _broker = None
    for exc in errors:
        if exc is not None:
            raise exc
    return results
