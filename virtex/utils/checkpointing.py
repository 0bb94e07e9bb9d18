from virtex_b200.checkpointing import CheckpointManager  # noqa: F401
