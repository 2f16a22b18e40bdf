"""Import-only stub of ``distrax`` (the fixtures supply actions; no sampling happens)."""
