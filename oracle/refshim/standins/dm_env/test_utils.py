"""Stand-in for dm_env.test_utils (dm_env is absent from this image).

EnvironmentTestMixin restates the checks of dm_env's reusable environment test: a fresh or
reset environment starts with a FIRST step without reward and discount, later steps carry a
reward and a discount that conform to the specs, observations conform to observation_spec(),
and the step after a LAST step is FIRST again.  Subclasses provide make_object_under_test().
"""


def _map(fn, spec):
  if isinstance(spec, dict):
    return {k: _map(fn, v) for k, v in spec.items()}
  if isinstance(spec, (list, tuple)):
    return type(spec)(_map(fn, v) for v in spec)
  return fn(spec)


class EnvironmentTestMixin(object):

  def setUp(self):
    super(EnvironmentTestMixin, self).setUp()
    self.environment = self.make_object_under_test()

  def tearDown(self):
    self.environment.close()
    super(EnvironmentTestMixin, self).tearDown()

  def make_object_under_test(self):
    raise NotImplementedError('make_object_under_test() must be provided by the test')

  def make_action(self):
    return _map(lambda s: s.generate_value(), self.environment.action_spec())

  def make_action_sequence(self):
    for _ in range(200):
      yield self.make_action()

  def reset_environment(self):
    step = self.environment.reset()
    self.assertValidStep(step)
    return step

  def step_environment(self, action=None):
    if action is None:
      action = self.make_action()
    step = self.environment.step(action)
    self.assertValidStep(step)
    return step

  # -- assertions ----------------------------------------------------------------------
  def assertConformsToSpec(self, value, spec):
    spec.validate(value)

  def assertValidObservation(self, observation):
    self.assertConformsToSpec(observation, self.environment.observation_spec())

  def assertValidReward(self, reward):
    self.assertConformsToSpec(reward, self.environment.reward_spec())

  def assertValidDiscount(self, discount):
    self.assertConformsToSpec(discount, self.environment.discount_spec())

  def assertValidStep(self, step):
    self.assertIn(int(step.step_type), (0, 1, 2))
    self.assertEqual([step.first(), step.mid(), step.last()].count(True), 1)
    if step.first():
      self.assertIsNone(step.reward)
      self.assertIsNone(step.discount)
    else:
      self.assertValidReward(step.reward)
      self.assertValidDiscount(step.discount)
    self.assertValidObservation(step.observation)

  # -- tests ---------------------------------------------------------------------------
  def test_reset(self):
    self.assertTrue(self.reset_environment().first())

  def test_step_on_fresh_environment(self):
    self.assertTrue(self.step_environment().first())
    self.assertFalse(self.step_environment().first())

  def test_step_after_reset(self):
    for _ in range(5):
      self.reset_environment()
      self.assertFalse(self.step_environment().first())

  def test_longlived_environment(self):
    step = self.reset_environment()
    for action in self.make_action_sequence():
      was_last = step.last()
      step = self.step_environment(action)
      if was_last:
        self.assertTrue(step.first())
