"""Test-infrastructure stub for pymeshlab."""
