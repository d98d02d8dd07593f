"""Re-authored entry points of the reference's src/ package (same public names)."""
