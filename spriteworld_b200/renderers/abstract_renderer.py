"""Renderer protocol (reference: renderers/abstract_renderer.py:27-47)."""
import abc


class AbstractRenderer(abc.ABC):

  @abc.abstractmethod
  def render(self, sprites=(), global_state=None):
    """Observation for one env from its sprites (back to front) and global state."""

  @abc.abstractmethod
  def observation_spec(self):
    """dm_env.specs.ArraySpec (or nested structure) of `render()`'s output."""
