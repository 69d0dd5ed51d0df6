"""Optimisers: host-side mirror of neuromancer/optimiser.py.

Global meta-parameters (lr, momentum, weight decay) are class-level like in the
reference (optimiser.py:19-55).  Each step is one fused kernel per parameter region
([regularised weights | biases]) of the flat device parameter buffer; the update
formulas are the reference's (Adam: optimiser.py:301-324; SGD: :146-160).
The reference additionally rotates three full parameter copies per step for
``repair()`` (optimiser.py:71-72, 103-118) -- kept here as an optional, off-by-default
snapshot because it is pure overhead on the hot path.
"""
import numpy as np
import torch

from .. import _lib
from .variables import VariableParam


class Optimiser(object):
    global_lr = VariableParam(value=1, name='lr')
    global_weight_decay = VariableParam(value=0, name='weight_decay')
    global_mom = VariableParam(value=0.9, name='mom')

    @classmethod
    def setlr(cls, val):
        cls.global_lr.set_value(val)

    @classmethod
    def setwd(cls, val):
        cls.global_weight_decay.set_value(val)

    @classmethod
    def setmom(cls, val):
        cls.global_mom.set_value(val)

    def __init__(self, model):
        self.model = model
        self.meta_params = dict(lr=self.global_lr, mom=self.global_mom, wd=self.global_weight_decay)
        self.last_exec_time = None
        self._state = None
        self._snapshot = None
        self.keep_history = False

    def set_opt_meta_params(self, value_dict):
        for k, v in value_dict.items():
            self.meta_params[k].set_value(v)

    def _hyper(self):
        return (float(np.asarray(self.global_lr.get_value()).reshape(-1)[0]),
                float(np.asarray(self.global_mom.get_value()).reshape(-1)[0]),
                float(np.asarray(self.global_weight_decay.get_value()).reshape(-1)[0]))

    def _regions(self, store):
        """(offset, count, weight-decay multiplier) per contiguous region of the flat buffer."""
        return list(store.regions)

    def _alloc(self, store, n):
        return [torch.zeros(store.total, dtype=torch.float32, device=store.device) for _ in range(n)]

    def step(self, store):
        raise NotImplementedError

    def _after_step(self, store):
        store.version += 1

    def repair(self):
        """Restore the parameters saved ``keep_history`` steps ago and clear the moments."""
        store = self.model._store
        if self._snapshot is not None:
            store.P.copy_(self._snapshot)
            store.version += 1
        if self._state is not None:
            for s in self._state:
                s.zero_()


def _slice_ptr(t, off):
    return _lib.C.c_void_p(t.data_ptr() + 4 * off)


class SGD(Optimiser):
    def step(self, store):
        h = _lib.get_handle()
        if self._state is None:
            self._state = self._alloc(store, 1)
        lr, mom, wd = self._hyper()
        if self.keep_history:
            self._snapshot = store.P.clone()
        for off, cnt, awd in self._regions(store):
            if cnt:
                h.call('e2_sgd_step', _slice_ptr(store.P, off), _slice_ptr(store.G, off),
                       _slice_ptr(self._state[0], off), cnt, lr, mom, wd, float(awd), h.stream())
        self._after_step(store)


class Adam(Optimiser):
    def __init__(self, model):
        super(Adam, self).__init__(model)
        self.beta2 = VariableParam(value=0.999, name='beta2')
        self.meta_params['beta2'] = self.beta2
        self.t = 0

    def step(self, store):
        h = _lib.get_handle()
        if self._state is None:
            self._state = self._alloc(store, 2)  # momentum, squared_accum
        lr, mom, wd = self._hyper()
        beta2 = float(np.asarray(self.beta2.get_value()).reshape(-1)[0])
        self.t += 1
        if self.keep_history:
            self._snapshot = store.P.clone()
        for off, cnt, awd in self._regions(store):
            if cnt:
                h.call('e2_adam_step', _slice_ptr(store.P, off), _slice_ptr(store.G, off),
                       _slice_ptr(self._state[0], off), _slice_ptr(self._state[1], off), cnt, lr, mom, beta2, wd,
                       float(awd), self.t, h.stream())
        self._after_step(store)

    # -- in-graph form (executor.Plan.train_step): hyper-parameters and step counter in device memory ----------
    fusable = True

    def dev_sync(self, store):
        """Make the device copies of (lr, mom, beta2, wd) and of the step counter match the host values.  Only
        touches the device when something changed since the last call (a schedule, set_opt_meta_params, ...)."""
        if self._state is None:
            self._state = self._alloc(store, 2)
        if getattr(self, '_hyper_dev', None) is None:
            self._hyper_dev = torch.zeros(8, dtype=torch.float32, device=store.device)
            self._t_dev = torch.zeros(1, dtype=torch.int32, device=store.device)
            self._hyper_host, self._t_host = None, None
        lr, mom, wd = self._hyper()
        cur = (lr, mom, float(np.asarray(self.beta2.get_value()).reshape(-1)[0]), wd)
        if cur != self._hyper_host:
            self._hyper_dev[:4].copy_(torch.tensor(cur, dtype=torch.float32))
            self._hyper_host = cur
        if self._t_host != self.t:
            self._t_dev.fill_(int(self.t))
            self._t_host = self.t

    def dev_prepare(self):
        """First launch of a fused step: t += 1 on the device, factor(t) refreshed."""
        h = _lib.get_handle()
        h.call('e2_adam_prepare', _lib.ptr(self._hyper_dev), _lib.ptr(self._t_dev), h.stream())

    def dev_step(self, store, off, cnt, apply_wd):
        if cnt <= 0:
            return
        h = _lib.get_handle()
        h.call('e2_adam_step_dev', _slice_ptr(store.P, off), _slice_ptr(store.G, off), _slice_ptr(self._state[0], off),
               _slice_ptr(self._state[1], off), cnt, _lib.ptr(self._hyper_dev), float(apply_wd), h.stream())

    def dev_after(self, store):
        """Host bookkeeping after a fused step has been submitted."""
        self.t += 1
        self._t_host = self.t
        self._after_step(store)
