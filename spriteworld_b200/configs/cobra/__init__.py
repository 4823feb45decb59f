"""The seven COBRA task configurations (reference: spriteworld/configs/cobra/)."""
