"""Example configurations (reference: spriteworld/configs/examples/); importing the package
imports both modules, as in the reference."""
from spriteworld_b200.configs.examples import goal_finding_clustering
from spriteworld_b200.configs.examples import goal_finding_embodied
