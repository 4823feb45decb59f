"""Example configurations (reference: spriteworld/configs/examples/)."""
