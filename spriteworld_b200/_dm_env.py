"""dm_env surface used by this package.

If the real `dm_env` is installed it is used as is; otherwise a minimal equivalent of the
pieces Spriteworld relies on (environment.py:22,78,106,108; action_spaces.py:24;
pil_renderer.py:22) is provided, with the same TimeStep conventions:
FIRST: reward/discount None; MID: discount 1.0; LAST: discount 0.0.
"""
try:  # pragma: no cover - depends on the image
  import dm_env as _real
  from dm_env import specs  # noqa: F401
  Environment = _real.Environment
  TimeStep = _real.TimeStep
  StepType = _real.StepType
  restart = _real.restart
  transition = _real.transition
  termination = _real.termination
  HAVE_DM_ENV = True
except ImportError:
  import abc
  import enum
  import types
  from typing import Any, NamedTuple

  import numpy as np

  HAVE_DM_ENV = False

  class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2

    def first(self):
      return self is StepType.FIRST

    def mid(self):
      return self is StepType.MID

    def last(self):
      return self is StepType.LAST

  class TimeStep(NamedTuple):
    step_type: Any
    reward: Any
    discount: Any
    observation: Any

    def first(self):
      return self.step_type == StepType.FIRST

    def mid(self):
      return self.step_type == StepType.MID

    def last(self):
      return self.step_type == StepType.LAST

  class Environment(metaclass=abc.ABCMeta):

    @abc.abstractmethod
    def reset(self):
      """Starts a new episode; returns the FIRST TimeStep."""

    @abc.abstractmethod
    def step(self, action):
      """Advances one step."""

    @abc.abstractmethod
    def observation_spec(self):
      pass

    @abc.abstractmethod
    def action_spec(self):
      pass

    def reward_spec(self):
      return specs.Array(shape=(), dtype=float, name='reward')

    def discount_spec(self):
      return specs.BoundedArray(shape=(), dtype=float, minimum=0., maximum=1., name='discount')

    def close(self):
      pass

    def __enter__(self):
      return self

    def __exit__(self, *exc):
      self.close()

  def restart(observation):
    return TimeStep(StepType.FIRST, None, None, observation)

  def transition(reward, observation, discount=1.0):
    return TimeStep(StepType.MID, reward, discount, observation)

  def termination(reward, observation):
    return TimeStep(StepType.LAST, reward, 0.0, observation)

  class _Array(object):
    __slots__ = ('_shape', '_dtype', '_name')

    def __init__(self, shape, dtype, name=None):
      self._shape = tuple(int(d) for d in shape)
      self._dtype = np.dtype(dtype)
      self._name = name

    shape = property(lambda self: self._shape)
    dtype = property(lambda self: self._dtype)
    name = property(lambda self: self._name)

    def __repr__(self):
      return 'Array(shape=%r, dtype=%r, name=%r)' % (self._shape, self._dtype, self._name)

    def validate(self, value):
      value = np.asarray(value)
      if value.shape != self._shape:
        raise ValueError('shape %r != %r' % (value.shape, self._shape))
      if value.dtype != self._dtype:
        raise ValueError('dtype %r != %r' % (value.dtype, self._dtype))
      return value

    def generate_value(self):
      return np.zeros(self._shape, self._dtype)

  class _BoundedArray(_Array):
    __slots__ = ('_minimum', '_maximum')

    def __init__(self, shape, dtype, minimum, maximum, name=None):
      super().__init__(shape, dtype, name)
      self._minimum = np.array(minimum, dtype=self._dtype)
      self._maximum = np.array(maximum, dtype=self._dtype)
      self._minimum.setflags(write=False)
      self._maximum.setflags(write=False)

    minimum = property(lambda self: self._minimum)
    maximum = property(lambda self: self._maximum)

    def validate(self, value):
      value = super().validate(value)
      if (value < self._minimum).any() or (value > self._maximum).any():
        raise ValueError('value out of bounds')
      return value

    def generate_value(self):
      return (np.ones(self._shape, self._dtype) * self._dtype.type(self._minimum))

  class _DiscreteArray(_BoundedArray):
    __slots__ = ('_num_values',)

    def __init__(self, num_values, dtype=np.int32, name=None):
      if num_values <= 0:
        raise ValueError('num_values must be positive')
      super().__init__((), dtype, 0, num_values - 1, name)
      self._num_values = int(num_values)

    num_values = property(lambda self: self._num_values)

  specs = types.SimpleNamespace(Array=_Array, BoundedArray=_BoundedArray,
                                DiscreteArray=_DiscreteArray)
