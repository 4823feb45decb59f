"""Stand-in for dm_env.specs."""
import numpy as np


class Array(object):

  def __init__(self, shape, dtype, name=None):
    self.shape = tuple(int(d) for d in shape)
    self.dtype = np.dtype(dtype)
    self.name = name

  def validate(self, value):
    value = np.asarray(value)
    if value.shape != self.shape:
      raise ValueError('shape %r, expected %r' % (value.shape, self.shape))
    if value.dtype != self.dtype:
      raise ValueError('dtype %r, expected %r' % (value.dtype, self.dtype))
    return value

  def generate_value(self):
    return np.zeros(self.shape, self.dtype)


class BoundedArray(Array):

  def __init__(self, shape, dtype, minimum, maximum, name=None):
    super(BoundedArray, self).__init__(shape, dtype, name)
    self.minimum = np.array(minimum, dtype=self.dtype)
    self.maximum = np.array(maximum, dtype=self.dtype)

  def validate(self, value):
    value = super(BoundedArray, self).validate(value)
    if (value < self.minimum).any() or (value > self.maximum).any():
      raise ValueError('value out of bounds')
    return value

  def generate_value(self):
    return np.ones(self.shape, self.dtype) * self.dtype.type(self.minimum)


class DiscreteArray(BoundedArray):

  def __init__(self, num_values, dtype=np.int32, name=None):
    super(DiscreteArray, self).__init__((), dtype, 0, num_values - 1, name)
    self.num_values = num_values
