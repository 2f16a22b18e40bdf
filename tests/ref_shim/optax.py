"""Import-only stub of ``optax``."""
