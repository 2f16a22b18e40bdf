class TrainState:
    """Import-only stub."""
