"""``jaxmarl`` 0.0.7 stand-in: only the base classes the reference derives from (test infrastructure)."""
