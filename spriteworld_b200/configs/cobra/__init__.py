"""The seven COBRA task configurations (reference: spriteworld/configs/cobra/).

Like the reference's package, importing it imports every configuration module, so that
`from spriteworld_b200.configs import cobra; cobra.sorting.get_config('train')` works.
"""
from spriteworld_b200.configs.cobra import clustering
from spriteworld_b200.configs.cobra import exploration
from spriteworld_b200.configs.cobra import goal_finding_more_distractors
from spriteworld_b200.configs.cobra import goal_finding_more_targets
from spriteworld_b200.configs.cobra import goal_finding_new_position
from spriteworld_b200.configs.cobra import goal_finding_new_shape
from spriteworld_b200.configs.cobra import sorting
