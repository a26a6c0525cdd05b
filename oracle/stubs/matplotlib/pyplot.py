"""Import-only placeholder."""
