"""Stand-in for dm_env: TimeStep / StepType / restart / transition / termination."""
import abc
import collections
import enum

from dm_env import specs  # noqa: F401


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


class TimeStep(
    collections.namedtuple('TimeStep',
                           ['step_type', 'reward', 'discount', 'observation'])):
  __slots__ = ()

  def first(self):
    return self.step_type == StepType.FIRST

  def mid(self):
    return self.step_type == StepType.MID

  def last(self):
    return self.step_type == StepType.LAST


class Environment(abc.ABC):

  @abc.abstractmethod
  def reset(self):
    pass

  @abc.abstractmethod
  def step(self, action):
    pass

  @abc.abstractmethod
  def observation_spec(self):
    pass

  @abc.abstractmethod
  def action_spec(self):
    pass

  def reward_spec(self):
    from dm_env import specs
    return specs.Array(shape=(), dtype=float, name='reward')

  def discount_spec(self):
    from dm_env import specs
    return specs.BoundedArray(shape=(), dtype=float, minimum=0., maximum=1., name='discount')

  def close(self):
    pass


def restart(observation):
  return TimeStep(StepType.FIRST, None, None, observation)


def transition(reward, observation, discount=1.0):
  return TimeStep(StepType.MID, reward, discount, observation)


def termination(reward, observation):
  return TimeStep(StepType.LAST, reward, 0.0, observation)
