"""Stand-in for dm_env.specs."""
import numpy as np


class Array(object):

  def __init__(self, shape, dtype, name=None):
    self.shape = tuple(int(d) for d in shape)
    self.dtype = np.dtype(dtype)
    self.name = name


class BoundedArray(Array):

  def __init__(self, shape, dtype, minimum, maximum, name=None):
    super(BoundedArray, self).__init__(shape, dtype, name)
    self.minimum = np.array(minimum, dtype=self.dtype)
    self.maximum = np.array(maximum, dtype=self.dtype)


class DiscreteArray(BoundedArray):

  def __init__(self, num_values, dtype=np.int32, name=None):
    super(DiscreteArray, self).__init__((), dtype, 0, num_values - 1, name)
    self.num_values = num_values
