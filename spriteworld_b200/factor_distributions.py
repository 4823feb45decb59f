"""Distributions over sprite factors, with scalar and vectorised sampling.

Same algebra and constructor signatures as the reference's
`spriteworld/factor_distributions.py:49-411` (Continuous, Discrete, Mixture,
Intersection, Product, SetMinus, Selection; `sample(rng)`, `contains(spec)`, `keys`),
so the shipped configs build unchanged.  `sample()` draws from the random generator in
the same order and with the same calls as the reference (one `uniform` per Continuous,
one `choice` per Discrete/Mixture, rejection loops for the set operations), which keeps
single-env runs reproducible against it under a shared seed.

What is new is the batched form used when thousands of envs reset at once:
`sample_batch(n, rng)` returns a dict of typed NumPy columns of length n, and
`contains_batch(columns)` a boolean mask.  Column dtypes follow the scalar path
(float32 for `Continuous(dtype='float32')`, float64 for Python-float candidates, ...),
because comparisons and colour conversion depend on them.
"""
import abc

import numpy as np

_MAX_TRIES = int(1e5)


def _rng(rng):
  return np.random if rng is None else rng


def _pad(indent):
  return '  ' * indent


def _as_column(values):
  """Python objects -> the tightest homogeneous NumPy column (object array otherwise)."""
  kinds = set(type(v) for v in values)
  if kinds <= {float, np.float64}:
    return np.array(values, dtype=np.float64)
  if kinds <= {int, np.int64}:
    return np.array(values, dtype=np.int64)
  if len(kinds) == 1 and issubclass(next(iter(kinds)), np.generic):
    return np.array(values)
  out = np.empty(len(values), dtype=object)
  out[:] = values
  return out


class AbstractDistribution(abc.ABC):
  """Base class: a distribution over "specs" (dicts factor name -> value)."""

  @abc.abstractmethod
  def sample(self, rng=None):
    """One spec.  `rng` defaults to the global np.random."""

  @abc.abstractmethod
  def contains(self, spec):
    """Whether `spec` lies in the support."""

  @abc.abstractmethod
  def to_str(self, indent):
    """Indented, recursive description."""

  @property
  @abc.abstractmethod
  def keys(self):
    """Set of factor names of sampled specs."""

  @abc.abstractmethod
  def sample_batch(self, n, rng=None):
    """n specs at once: dict factor name -> column of length n."""

  @abc.abstractmethod
  def contains_batch(self, columns):
    """Boolean mask over the rows of `columns` (dict name -> column)."""

  def __str__(self):
    return self.to_str(indent=0)

  def _get_rng(self, rng=None):
    return _rng(rng)

  def _need(self, key, spec):
    if key not in spec:
      raise KeyError('key {} is not in spec {}, but must be to evaluate '
                     'containment.'.format(key, spec))


def _rejection_scalar(draw, accept, what):
  for _ in range(_MAX_TRIES):
    candidate = draw()
    if accept(candidate):
      return candidate
  raise ValueError('Maximum number of tried exceeded when trying to sample from {}.'.format(what))


def _rejection_batch(draw_batch, accept_batch, n, what):
  """Collects n accepted rows, in draw order."""
  kept = []
  have = 0
  need_more = n
  for _ in range(64):
    if need_more <= 0:
      break
    m = max(32, int(need_more * 1.5) + 8)
    cols = draw_batch(m)
    ok = np.asarray(accept_batch(cols), bool)
    if ok.any():
      kept.append({k: v[ok] for k, v in cols.items()})
      have += int(ok.sum())
      need_more = n - have
  if need_more > 0:
    raise ValueError('Maximum number of tried exceeded when trying to sample from {}.'.format(what))
  return {k: np.concatenate([c[k] for c in kept])[:n] for k in kept[0]}


class Continuous(AbstractDistribution):
  """Uniform on [minval, maxval)."""

  def __init__(self, key, minval, maxval, dtype='float32'):
    self.key, self.minval, self.maxval, self.dtype = key, minval, maxval, dtype

  def sample(self, rng=None):
    value = _rng(rng).uniform(low=self.minval, high=self.maxval)
    return {self.key: np.asarray(value).astype(self.dtype)[()]}

  def sample_batch(self, n, rng=None):
    return {self.key: _rng(rng).uniform(low=self.minval, high=self.maxval, size=n).astype(self.dtype)}

  def contains(self, spec):
    self._need(self.key, spec)
    return spec[self.key] >= self.minval and spec[self.key] < self.maxval

  def contains_batch(self, columns):
    self._need(self.key, columns)
    col = columns[self.key]
    if col.dtype == object:
      return np.array([v >= self.minval and v < self.maxval for v in col], bool)
    return (col >= self.minval) & (col < self.maxval)

  def to_str(self, indent):
    return _pad(indent) + '<Continuous: key={}, mival={}, maxval={}, dtype={}>'.format(
        self.key, self.minval, self.maxval, self.dtype)

  @property
  def keys(self):
    return {self.key}


class Discrete(AbstractDistribution):
  """Categorical over `candidates` (uniform unless `probs` is given)."""

  def __init__(self, key, candidates, probs=None):
    self.key, self.candidates, self.probs = key, candidates, probs

  def sample(self, rng=None):
    index = _rng(rng).choice(len(self.candidates), p=self.probs)
    return {self.key: self.candidates[index]}

  def sample_batch(self, n, rng=None):
    index = _rng(rng).choice(len(self.candidates), size=n, p=self.probs)
    if len(set(type(v) for v in self.candidates)) == 1:
      # homogeneous candidates: any sample of them types the same way, so type once and index
      return {self.key: _as_column(list(self.candidates))[index]}
    return {self.key: _as_column([self.candidates[i] for i in index])}

  def contains(self, spec):
    self._need(self.key, spec)
    return spec[self.key] in self.candidates

  def contains_batch(self, columns):
    self._need(self.key, columns)
    col = columns[self.key]
    if col.dtype == object and len(col) >= 64:
      # object columns (shape names): decide once per distinct value; values that do not
      # order against each other (mixed types) take the plain loop
      try:
        values, inverse = np.unique(col, return_inverse=True)
      except TypeError:
        values = None
      if values is not None:
        return np.array([v in self.candidates for v in values], bool)[inverse.reshape(-1)]
    return np.array([v in self.candidates for v in col], bool)

  def to_str(self, indent):
    return _pad(indent) + '<Discrete: key={}, candidates={}, probs={}>'.format(
        self.key, self.candidates, self.probs)

  @property
  def keys(self):
    return {self.key}


def _same_keys(components):
  first = components[0].keys
  for c in components[1:]:
    if c.keys != first:
      raise ValueError('All components must have the same key sets. However detected key '
                       'sets {} and {}'.format(first, c.keys))
  return first


def _nested_str(indent, name, parts, tail=''):
  body = ',\n'.join(p.to_str(indent + 2) for p in parts)
  return (_pad(indent) + '<' + name + ':\n' + _pad(indent + 1) + 'components=[\n' + body +
          ',\n' + _pad(indent + 1) + ']' + tail + '>')


class Mixture(AbstractDistribution):
  """Weighted mixture (not a union: overlaps are sampled more often)."""

  def __init__(self, components, probs=None):
    self.components = components
    n = len(components)
    self.probs = np.ones(n) / n if probs is None else np.array(probs)
    self._keys = _same_keys(components)

  def sample(self, rng=None):
    rng = _rng(rng)
    which = rng.choice(len(self.components), p=self.probs)
    return self.components[which].sample(rng=rng)

  def sample_batch(self, n, rng=None):
    rng = _rng(rng)
    which = rng.choice(len(self.components), size=n, p=self.probs)
    parts, where = [], []
    for i, comp in enumerate(self.components):
      rows = np.flatnonzero(which == i)
      if len(rows):
        parts.append(comp.sample_batch(len(rows), rng=rng))
        where.append(rows)
    order = np.argsort(np.concatenate(where), kind='stable')
    return {k: _concat_columns([p[k] for p in parts])[order] for k in parts[0]}

  def contains(self, spec):
    return any(c.contains(spec) for c in self.components)

  def contains_batch(self, columns):
    out = self.components[0].contains_batch(columns)
    for c in self.components[1:]:
      out = out | c.contains_batch(columns)
    return out

  def to_str(self, indent):
    return _nested_str(indent, 'Mixture', self.components,
                       ',\n' + _pad(indent + 1) + 'probs={}'.format(self.probs))

  @property
  def keys(self):
    return self._keys


def _concat_columns(cols):
  """Concatenates columns; mixed dtypes become an object column of NumPy/Python scalars
  (so that each value keeps the dtype the scalar path would have given it)."""
  if len({c.dtype for c in cols}) == 1:
    return np.concatenate(cols)
  out = np.empty(sum(len(c) for c in cols), dtype=object)
  i = 0
  for c in cols:
    for v in c:
      out[i] = v
      i += 1
  return out


class Intersection(AbstractDistribution):
  """Samples one component and rejects with the others."""

  def __init__(self, components, index_for_sampling=0):
    self.components = components
    self.index_for_sampling = index_for_sampling
    self._keys = _same_keys(components)

  def sample(self, rng=None):
    rng = _rng(rng)
    source = self.components[self.index_for_sampling]
    return _rejection_scalar(lambda: source.sample(rng=rng), self.contains, str(self))

  def sample_batch(self, n, rng=None):
    rng = _rng(rng)
    source = self.components[self.index_for_sampling]
    return _rejection_batch(lambda m: source.sample_batch(m, rng=rng), self.contains_batch, n,
                            str(self))

  def contains(self, spec):
    return all(c.contains(spec) for c in self.components)

  def contains_batch(self, columns):
    out = self.components[0].contains_batch(columns)
    for c in self.components[1:]:
      out = out & c.contains_batch(columns)
    return out

  def to_str(self, indent):
    return _nested_str(indent, 'Intersection', self.components,
                       ',\n' + _pad(indent + 1) +
                       'index_for_sampling={}'.format(self.index_for_sampling))

  @property
  def keys(self):
    return self._keys


class Product(AbstractDistribution):
  """Independent components over disjoint key sets."""

  def __init__(self, components):
    self.components = components
    keys, total = set(), 0
    for c in components:
      keys |= set(c.keys)
      total += len(c.keys)
    if len(keys) < total:
      raise ValueError('All components must have different keys, yet there are {} '
                       'overlapping keys.'.format(total - len(keys)))
    self._keys = keys

  def sample(self, rng=None):
    rng = _rng(rng)
    spec = {}
    for c in self.components:
      spec.update(c.sample(rng=rng))
    return spec

  def sample_batch(self, n, rng=None):
    rng = _rng(rng)
    cols = {}
    for c in self.components:
      cols.update(c.sample_batch(n, rng=rng))
    return cols

  def contains(self, spec):
    return all(c.contains(spec) for c in self.components)

  def contains_batch(self, columns):
    out = self.components[0].contains_batch(columns)
    for c in self.components[1:]:
      out = out & c.contains_batch(columns)
    return out

  def to_str(self, indent):
    return _nested_str(indent, 'Product', self.components)

  @property
  def keys(self):
    return self._keys


class _Filtered(AbstractDistribution):
  """base, thinned by a second distribution over a subset of its keys."""
  _label = None
  _second = None   # attribute name of the second distribution
  _keep_if_inside = None

  def _init(self, base, other, role):
    self.base = base
    self._keys = base.keys
    if not other.keys.issubset(self._keys):
      raise ValueError('Keys {} of {} is not a subset of keys {} of {} base distribution.'.format(
          other.keys, role, base.keys, self._label))
    self._other = other

  def _accept(self, spec):
    return self._other.contains(spec) == self._keep_if_inside

  def _accept_batch(self, columns):
    inside = self._other.contains_batch(columns)
    return inside if self._keep_if_inside else ~inside

  def sample(self, rng=None):
    rng = _rng(rng)
    return _rejection_scalar(lambda: self.base.sample(rng=rng), self._accept, str(self))

  def sample_batch(self, n, rng=None):
    rng = _rng(rng)
    return _rejection_batch(lambda m: self.base.sample_batch(m, rng=rng), self._accept_batch, n,
                            str(self))

  def contains(self, spec):
    return self.base.contains(spec) and self._accept(spec)

  def contains_batch(self, columns):
    return self.base.contains_batch(columns) & self._accept_batch(columns)

  def to_str(self, indent):
    return (_pad(indent) + '<' + self._label + ':\n' + _pad(indent + 1) + 'base=\n' +
            self.base.to_str(indent + 2) + ',\n' + _pad(indent + 1) + self._second + '=\n' +
            self._other.to_str(indent + 2) + '>')

  @property
  def keys(self):
    return self._keys


class SetMinus(_Filtered):
  """base without the region covered by hold_out (rejection sampling)."""
  _label, _second, _keep_if_inside = 'SetMinus', 'hold_out', False

  def __init__(self, base, hold_out):
    self._init(base, hold_out, 'hold_out')
    self.hold_out = hold_out


class Selection(_Filtered):
  """base restricted to the region accepted by `filtering` (rejection sampling)."""
  _label, _second, _keep_if_inside = 'Selection', 'filtering', True

  def __init__(self, base, filtering):
    self._init(base, filtering, 'filtering')
    self.filtering = filtering
