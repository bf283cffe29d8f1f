"""Optimizers -- mirrors graphvite.optimizer (reference python/graphvite/optimizer.py and the
pybind classes in include/bind.h:757-999, whose C++ definitions are include/core/optimizer.h:42-319).
"""
from numpy import float32 as _f32

from . import _lib
from .base import auto

_SCHEDULES = {"constant": 0, "linear": 1}


class LRSchedule(object):
    """LRSchedule(type='constant') or LRSchedule(schedule_function)  (core/optimizer.h:42-85).

    schedule_function(batch_id, num_batch) returns a multiplicative factor for the learning rate.
    """

    def __init__(self, type="constant"):
        if callable(type):
            self.type = "custom"
            self.schedule_function = type
        else:
            if type not in _SCHEDULES:
                raise ValueError("Invalid schedule `%s`" % type)
            self.type = type
            if type == "linear":
                # single precision like LRSchedule::linear_schedule (core/optimizer.h:77-79)
                self.schedule_function = lambda batch_id, num_batch: float(
                    max(_f32(1) - _f32(batch_id) / _f32(num_batch), _f32(1e-4)))
            else:
                self.schedule_function = lambda batch_id, num_batch: 1

    def __call__(self, batch_id, num_batch):
        return self.schedule_function(batch_id, num_batch)

    def __repr__(self):
        return "lr schedule: %s" % self.type


def _as_schedule(schedule):
    return schedule if isinstance(schedule, LRSchedule) else LRSchedule(schedule)


class _OptimizerBase(object):
    """graphvite::Optimizer (core/optimizer.h:102-214)."""
    _type_id = -1
    type = "Default"
    momentum = alpha = beta1 = beta2 = epsilon = 0.0

    def __init__(self, lr=1e-4, weight_decay=0, schedule="linear"):
        self.lr = float(lr)
        self.init_lr = float(lr)
        self.weight_decay = float(weight_decay)
        self.schedule = _as_schedule(schedule)

    def _a_b(self):
        return 0.0, 0.0

    def _descriptor(self):
        """gv_optimizer_t; the returned object keeps the ctypes callback alive."""
        a, b = self._a_b()
        desc = _lib.OptimizerDesc()
        desc.type = self._type_id
        desc.lr = self.init_lr
        desc.weight_decay = self.weight_decay
        desc.a, desc.b, desc.epsilon = a, b, float(self.epsilon)
        if self.schedule.type == "custom":
            function = self.schedule.schedule_function
            desc.schedule = 2
            desc.schedule_fn = _lib.SCHEDULE_FN(lambda batch_id, num_batch, ctx: float(function(batch_id, num_batch)))
        else:
            desc.schedule = _SCHEDULES[self.schedule.type]
            desc.schedule_fn = _lib.SCHEDULE_FN()
        return desc

    def __repr__(self):  # Optimizer::info, core/optimizer.h:137-155
        lines = ["optimizer: %s" % self.type,
                 "learning rate: %g, %r" % (self.init_lr, self.schedule),
                 "weight decay: %g" % self.weight_decay]
        if self.type == "Momentum":
            lines.append("momentum: %g" % self.momentum)
        if self.type == "AdaGrad":
            lines.append("epsilon: %g" % self.epsilon)
        if self.type == "RMSprop":
            lines.append("alpha: %g, epsilon: %g" % (self.alpha, self.epsilon))
        if self.type == "Adam":
            lines.append("beta1: %g, beta2: %g, epsilon: %g" % (self.beta1, self.beta2, self.epsilon))
        return "\n".join(lines)


class _Default(_OptimizerBase):
    """Optimizer(auto) / Optimizer(lr): the solver picks its default type (core/optimizer.h:120-129)."""

    def __init__(self, lr=0.0):
        _OptimizerBase.__init__(self, lr, 0, "constant")


class SGD(_OptimizerBase):
    """SGD(lr=1e-4, weight_decay=0, schedule='linear')  (core/optimizer.h:272-277)"""
    _type_id, type = 0, "SGD"

    def __init__(self, lr=1e-4, weight_decay=0, schedule="linear"):
        _OptimizerBase.__init__(self, lr, weight_decay, schedule)


class Momentum(_OptimizerBase):
    """Momentum(lr=1e-4, weight_decay=0, momentum=0.999, schedule='linear')  (core/optimizer.h:279-287)"""
    _type_id, type = 1, "Momentum"

    def __init__(self, lr=1e-4, weight_decay=0, momentum=0.999, schedule="linear"):
        _OptimizerBase.__init__(self, lr, weight_decay, schedule)
        self.momentum = float(momentum)

    def _a_b(self):
        return self.momentum, 0.0


class AdaGrad(_OptimizerBase):
    """AdaGrad(lr=1e-4, weight_decay=0, epsilon=1e-10, schedule='linear')  (core/optimizer.h:289-297)"""
    _type_id, type = 2, "AdaGrad"

    def __init__(self, lr=1e-4, weight_decay=0, epsilon=1e-10, schedule="linear"):
        _OptimizerBase.__init__(self, lr, weight_decay, schedule)
        self.epsilon = float(epsilon)


class RMSprop(_OptimizerBase):
    """RMSprop(lr=1e-4, weight_decay=0, alpha=0.999, epsilon=1e-8, schedule='linear')  (core/optimizer.h:299-308)"""
    _type_id, type = 3, "RMSprop"

    def __init__(self, lr=1e-4, weight_decay=0, alpha=0.999, epsilon=1e-8, schedule="linear"):
        _OptimizerBase.__init__(self, lr, weight_decay, schedule)
        self.alpha = float(alpha)
        self.epsilon = float(epsilon)

    def _a_b(self):
        return self.alpha, 0.0


class Adam(_OptimizerBase):
    """Adam(lr=1e-4, weight_decay=0, beta1=0.999, beta2=0.99999, epsilon=1e-8, schedule='linear')
    (core/optimizer.h:310-319; no bias correction, like the reference)"""
    _type_id, type = 4, "Adam"

    def __init__(self, lr=1e-4, weight_decay=0, beta1=0.999, beta2=0.99999, epsilon=1e-8, schedule="linear"):
        _OptimizerBase.__init__(self, lr, weight_decay, schedule)
        self.beta1, self.beta2, self.epsilon = float(beta1), float(beta2), float(epsilon)

    def _a_b(self):
        return self.beta1, self.beta2


_TYPES = {"SGD": SGD, "Momentum": Momentum, "AdaGrad": AdaGrad, "RMSprop": RMSprop, "Adam": Adam}


class Optimizer(object):
    """Optimizer(type=auto, *args, **kwargs): create an optimizer of any type
    (python/graphvite/optimizer.py:29-44)."""

    def __new__(cls, type=auto, *args, **kwargs):
        if type == auto:
            return _Default()
        if isinstance(type, str):
            if type in _TYPES:
                return _TYPES[type](*args, **kwargs)
            raise ValueError("Unknown optimizer `%s`" % type)
        if isinstance(type, float):
            return _Default(type)
        raise ValueError("Unknown optimizer `%s`" % (type,))


def as_optimizer(value):
    """bind.h:793-794: int (auto) and float (learning rate) convert implicitly to Optimizer."""
    if isinstance(value, _OptimizerBase):
        return value
    if isinstance(value, bool):
        raise TypeError("optimizer must be an Optimizer, auto or a learning rate")
    if isinstance(value, int):
        if value != auto:
            raise ValueError("Only auto can be used for initializing a default optimizer. "
                             "Please use a float value if you want to specify the learning rate.")
        return _Default()
    if isinstance(value, float):
        return _Default(value)
    raise TypeError("optimizer must be an Optimizer, auto or a learning rate")


__all__ = ["Optimizer", "LRSchedule", "SGD", "Momentum", "AdaGrad", "RMSprop", "Adam"]
